// gate_kernels.cuh -- closed-form gate kernels built ON THE DEVICE from the parameter vector, one thread per
// (parameter set, gate). Replaces the host-side per-gate builders of the reference
// (squander/src-cpp/gates/include/gate_kernel_templates.h; dispatch in U3.cpp:89-142, RY.cpp:27-105, CU.cpp:100-180,
// R.cpp, U2.cpp, RXX/RYY/RZZ.cpp) and Gates_block's sincos batch (Gates_block.cpp:269-279).
//
// Conventions copied from the reference: the stored parameter of a "theta" slot is theta/2 and sincos is taken of the
// stored value (U3.cpp:49-51, Gate.cpp:1430-1446); Adaptive == CRY with an identity activation (Adaptive.cpp:28-32,
// common/common.cpp:35-38); derivative kernels are the reference's explicit ones, including their zeroed entries
// (gate_kernel_templates.h:554-609).
#pragma once
#include "sq_types.cuh"
#include "../../include/sqgpu.h"

namespace sq {

struct Trig {
    double s[4], c[4];
};

__host__ __device__ __forceinline__ void k2(cplx* k, double r0, double i0, double r1, double i1, double r2, double i2, double r3,
                                   double i3) {
    k[0] = cmake(r0, i0);
    k[1] = cmake(r1, i1);
    k[2] = cmake(r2, i2);
    k[3] = cmake(r3, i3);
}

__host__ __device__ __forceinline__ void rot_phase(cplx* k, double sg, double cg) {  // multiply_2x2_by_phase :611-619
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = cmake(k[i].x * cg - k[i].y * sg, k[i].x * sg + k[i].y * cg);
}

// U3 body shared by U3 and CU. which: -1 forward, 0/1/2 derivative wrt theta/phi/lambda.
__host__ __device__ __forceinline__ void u3_body(cplx* k, int which, double st, double ct, double sp, double cp, double sl,
                                        double cl) {
    const double spl = sp * cl + cp * sl;
    const double cpl = cp * cl - sp * sl;
    if (which < 0) k2(k, ct, 0.0, -st * cl, -st * sl, st * cp, st * sp, ct * cpl, ct * spl);
    else if (which == 0) k2(k, -st, 0.0, -ct * cl, -ct * sl, ct * cp, ct * sp, -st * cpl, -st * spl);
    else if (which == 1) k2(k, 0.0, 0.0, 0.0, 0.0, -st * sp, st * cp, -ct * spl, ct * cpl);
    else k2(k, 0.0, 0.0, st * sl, -st * cl, 0.0, 0.0, -ct * spl, ct * cpl);
}

__host__ __device__ __forceinline__ void zero16(cplx* k) {
#pragma unroll
    for (int i = 0; i < 16; ++i) k[i] = czero();
}

// Writes the dim x dim kernel of `type` (which < 0) or its derivative wrt parameter `which`. Returns dim (0: unknown).
__host__ __device__ inline int build_gate_kernel(int type, const Trig& t, int which, cplx* k) {
    const double s0 = t.s[0], c0 = t.c[0], s1 = t.s[1], c1 = t.c[1], s2 = t.s[2], c2 = t.c[2], s3 = t.s[3], c3 = t.c[3];
    const double sq2 = 0.70710678118654752440;  // M_SQRT1_2
    const bool fwd = which < 0;
    switch (type) {
        case SQGPU_U3: u3_body(k, which, s0, c0, s1, c1, s2, c2); return 2;
        case SQGPU_CU:
            if (which == 3) {  // d/dgamma = i * kernel  (:647-656)
                u3_body(k, -1, s0, c0, s1, c1, s2, c2);
                rot_phase(k, s3, c3);
#pragma unroll
                for (int i = 0; i < 4; ++i) k[i] = cmake(-k[i].y, k[i].x);
            } else {
                u3_body(k, which, s0, c0, s1, c1, s2, c2);
                rot_phase(k, s3, c3);
            }
            return 2;
        case SQGPU_RX: case SQGPU_CRX:
            if (fwd) k2(k, c0, 0, 0, -s0, 0, -s0, c0, 0); else k2(k, -s0, 0, 0, -c0, 0, -c0, -s0, 0);
            return 2;
        case SQGPU_RY: case SQGPU_CRY: case SQGPU_ADAPTIVE:
            if (fwd) k2(k, c0, 0, -s0, 0, s0, 0, c0, 0); else k2(k, -s0, 0, -c0, 0, c0, 0, -s0, 0);
            return 2;
        case SQGPU_RZ: case SQGPU_CRZ:
            if (fwd) k2(k, c0, -s0, 0, 0, 0, 0, c0, s0); else k2(k, -s0, -c0, 0, 0, 0, 0, -s0, c0);
            return 2;
        case SQGPU_U1: case SQGPU_CP:
            if (fwd) k2(k, 1, 0, 0, 0, 0, 0, c0, s0); else k2(k, 0, 0, 0, 0, 0, 0, -s0, c0);
            return 2;
        case SQGPU_U2: {
            const double spl = s0 * c1 + c0 * s1, cpl = c0 * c1 - s0 * s1;
            if (fwd) k2(k, sq2, 0, -sq2 * c1, -sq2 * s1, sq2 * c0, sq2 * s0, sq2 * cpl, sq2 * spl);
            else if (which == 0) k2(k, 0, 0, 0, 0, -sq2 * s0, sq2 * c0, -sq2 * spl, sq2 * cpl);
            else k2(k, 0, 0, sq2 * s1, -sq2 * c1, 0, 0, -sq2 * spl, sq2 * cpl);
            return 2;
        }
        case SQGPU_R: case SQGPU_CR:
            if (fwd) k2(k, c0, 0, -s0 * s1, -s0 * c1, s0 * s1, -s0 * c1, c0, 0);
            else if (which == 0) k2(k, -s0, 0, -c0 * s1, -c0 * c1, c0 * s1, -c0 * c1, -s0, 0);
            else k2(k, 0, 0, -s0 * c1, s0 * s1, s0 * c1, s0 * s1, 0, 0);
            return 2;
        case SQGPU_X: case SQGPU_CNOT: case SQGPU_CCX: k2(k, 0, 0, 1, 0, 1, 0, 0, 0); return 2;
        case SQGPU_Y: k2(k, 0, 0, 0, -1, 0, 1, 0, 0); return 2;
        case SQGPU_Z: case SQGPU_CZ: k2(k, 1, 0, 0, 0, 0, 0, -1, 0); return 2;
        case SQGPU_H: case SQGPU_CH: k2(k, sq2, 0, sq2, 0, sq2, 0, -sq2, 0); return 2;
        case SQGPU_S: k2(k, 1, 0, 0, 0, 0, 0, 0, 1); return 2;
        case SQGPU_SDG: k2(k, 1, 0, 0, 0, 0, 0, 0, -1); return 2;
        case SQGPU_T: k2(k, 1, 0, 0, 0, 0, 0, sq2, sq2); return 2;
        case SQGPU_TDG: k2(k, 1, 0, 0, 0, 0, 0, sq2, -sq2); return 2;
        case SQGPU_SX: k2(k, .5, .5, .5, -.5, .5, -.5, .5, .5); return 2;
        case SQGPU_SXDG: k2(k, .5, -.5, .5, .5, .5, .5, .5, -.5); return 2;
        case SQGPU_RXX: {  // :669-695
            zero16(k);
            const cplx d = fwd ? cmake(c0, 0) : cmake(-s0, 0), o = fwd ? cmake(0, -s0) : cmake(0, -c0);
            k[0] = d; k[5] = d; k[10] = d; k[15] = d; k[3] = o; k[6] = o; k[9] = o; k[12] = o;
            return 4;
        }
        case SQGPU_RYY: {  // :704-730
            zero16(k);
            const cplx d = fwd ? cmake(c0, 0) : cmake(-s0, 0);
            const cplx op = fwd ? cmake(0, s0) : cmake(0, c0), om = fwd ? cmake(0, -s0) : cmake(0, -c0);
            k[0] = d; k[5] = d; k[10] = d; k[15] = d; k[3] = op; k[12] = op; k[6] = om; k[9] = om;
            return 4;
        }
        case SQGPU_RZZ: {  // :739-765
            zero16(k);
            const cplx m = fwd ? cmake(c0, -s0) : cmake(-s0, -c0), p = fwd ? cmake(c0, s0) : cmake(-s0, c0);
            k[0] = m; k[15] = m; k[5] = p; k[10] = p;
            return 4;
        }
        case SQGPU_SWAP: case SQGPU_CSWAP:
            zero16(k);
            k[0] = cmake(1, 0); k[6] = cmake(1, 0); k[9] = cmake(1, 0); k[15] = cmake(1, 0);
            return 4;
        case SQGPU_SYC:  // fSim(pi/2, pi/6), kernels/apply_dedicated_gate_kernel_to_input.cpp:582-640 (symmetric in its qubits)
            zero16(k);
            k[0] = cmake(1, 0); k[6] = cmake(0, -1); k[9] = cmake(0, -1); k[15] = cmake(0.86602540378443864676, -0.5);
            return 4;
        case SQGPU_CROT: {
            // local index = target bit | control bit << 1. control = 0: U3(theta, phi - pi/2, -phi + pi/2), control = 1: the
            // same with -theta (gate_kernel_templates.h:779-807; Gate.cpp:1568-1580 gives the "inverse" kernel to the
            // control = 1 rows). d/dtheta: theta -> theta + pi/2 on both branches (:831-842); d/dphi: U3(+-theta, phi, -phi)
            // with the diagonal zeroed (:866-877). Derivative kernels are NOT zero-filled: both branches carry them.
            zero16(k);
            double st = s0, ct = c0;
            if (which == 0) { st = c0; ct = -s0; }
            cplx b0[4], b1[4];
            if (which <= 0) {
                u3_body(b0, -1, st, ct, -c1, s1, c1, s1);
                u3_body(b1, -1, -st, ct, -c1, s1, c1, s1);
            } else {
                u3_body(b0, -1, st, ct, s1, c1, -s1, c1);
                u3_body(b1, -1, -st, ct, s1, c1, -s1, c1);
                b0[0] = b0[3] = b1[0] = b1[3] = czero();
            }
            k[0] = b0[0]; k[1] = b0[1]; k[4] = b0[2]; k[5] = b0[3];
            k[10] = b1[0]; k[11] = b1[1]; k[14] = b1[2]; k[15] = b1[3];
            return 4;
        }
        default: return 0;
    }
}

// ---- fused blocks ------------------------------------------------------------------------------------------------
// A block is a run of gates acting inside 1, 2 or 3 qubits (dim = 2, 4, 8). One WARP builds the tables of one
// (parameter set, block): matrices live in shared memory, every lane owns dim*dim/32 (at most 2) entries.

// element (r, c) of a member's kernel embedded into the block's basis (local index bit j <-> j-th ascending block qubit).
// `deriv`: derivative kernels of controlled gates are ZERO (not identity) where the control bit is 0 -- the reference's
// convention for derivative matrices (kernels/apply_kernel_to_input.cpp:93-97).
__device__ __forceinline__ cplx embed_elem(const DevMember& m, const cplx* k, bool deriv, int r, int c) {
    if (m.dim == 2) {
        const int tl = m.tl;
        if ((r & ~(1 << tl)) != (c & ~(1 << tl))) return czero();
        const int tr = (r >> tl) & 1, tc = (c >> tl) & 1;
        if (m.cl >= 0 && ((r >> m.cl) & 1) == 0) return (deriv || tr != tc) ? czero() : cmake(1.0, 0.0);
        return k[tr * 2 + tc];
    }
    const int lo = m.tl, hi = m.tl2, mask = (1 << lo) | (1 << hi);
    if ((r & ~mask) != (c & ~mask)) return czero();
    const int kr = ((r >> lo) & 1) | (((r >> hi) & 1) << 1), kc = ((c >> lo) & 1) | (((c >> hi) & 1) << 1);
    return k[kr * 4 + kc];
}

__device__ __forceinline__ void member_trig(const DevMember& m, const double* __restrict__ params, Trig& t) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        t.s[i] = 0.0;
        t.c[i] = 1.0;
        if (i < m.n_params) sincos(params[m.param_start + i], &t.s[i], &t.c[i]);
    }
}

__device__ __forceinline__ void member_kernel(const DevMember& m, const Trig& t, int which, const cplx* __restrict__ pool,
                                              cplx* k) {
    if (m.type == SQGPU_GENERAL) {
        for (int i = 0; i < m.dim * m.dim; ++i) k[i] = pool[m.pool_off + i];
        return;
    }
    build_gate_kernel(m.type, t, which, k);
}

// What fills a parameter's slot of the "derivative" tables. shift == 0: dK/dtheta_p. shift != 0 (sqgpu_cost_shifted_batched):
// the finite difference K(theta_p + shift) - K(theta_p). The trace functional is linear in each op's kernel, so the adjoint sweep
// and reduce_partials, unchanged, then return L(theta + shift e_p) - L(theta) for EVERY p from one sweep -- the shifted costs
// the COSINE engine asks for (optimization_engines/COSINE.cpp:255-291) without one forward pass per parameter. The embedding
// keeps the derivative convention (zero where a control bit is 0): identity - identity = 0 there.
__device__ __forceinline__ void slot_kernel(int type, const Trig& t, int p, double theta_p, double shift, cplx* k) {
    if (shift == 0.0) {
        build_gate_kernel(type, t, p, k);
        return;
    }
    Trig ts = t;
    sincos(theta_p + shift, &ts.s[p], &ts.c[p]);
    cplx k0[16];
    const int dim = build_gate_kernel(type, ts, -1, k);
    build_gate_kernel(type, t, -1, k0);
    for (int i = 0; i < dim * dim; ++i) k[i] = cmake(k[i].x - k0[i].x, k[i].y - k0[i].y);
}

// out = a * b (dim x dim, row-major) by one warp; out must not alias a or b
__device__ __forceinline__ void warp_mat_mul(const cplx* a, const cplx* b, cplx* out, int dim, int lane) {
    for (int i = lane; i < dim * dim; i += 32) {
        const int r = i / dim, c = i - r * dim;
        cplx acc = czero();
        for (int l = 0; l < dim; ++l) acc = cfma(a[r * dim + l], b[l * dim + c], acc);
        out[i] = acc;
    }
    __syncwarp();
}

// Block matrix M = E_{m-1} ... E_0 and, for every parameter p of member j, dM_p = (E_{m-1}..E_{j+1}) dE_{j,p} (E_{j-1}..E_0).
// This is the product rule the reference evaluates with full-size matrices (Gates_block::apply_derivate_to,
// Gates_block.cpp:1011-1150), restricted to the 2^k-dimensional space the run of gates acts on.
// Pass 1 walks the members forward and parks the running prefix in each parameter's output slot; pass 2 walks
// backward with the running suffix and finishes dM_p = suffix * dE * prefix in place.
// ws: 3 * 64 complex of shared memory owned by the warp.
__device__ inline void build_block_warp(const DevOp& op, const DevMember* __restrict__ members, const double* __restrict__ params,
                                        const cplx* __restrict__ pool, cplx* __restrict__ kdst, cplx* __restrict__ dkdst,
                                        int with_deriv, cplx* ws, int lane, double shift = 0.0) {
    const int dim = op.dim, d2 = dim * dim, nm = op.n_members;
    cplx* R = ws;        // running prefix / suffix
    cplx* E = ws + 64;   // embedded member
    cplx* T = ws + 128;  // product scratch
    cplx k[16];
    for (int i = lane; i < d2; i += 32) R[i] = (i / dim == i % dim) ? cmake(1.0, 0.0) : czero();
    __syncwarp();
    for (int j = 0; j < nm; ++j) {
        const DevMember m = members[op.member_off + j];
        if (with_deriv)
            for (int p = 0; p < m.n_params; ++p)
                for (int i = lane; i < d2; i += 32) dkdst[(size_t)(m.slot0 + p) * d2 + i] = R[i];
        Trig t;
        member_trig(m, params, t);
        member_kernel(m, t, -1, pool, k);
        for (int i = lane; i < d2; i += 32) E[i] = embed_elem(m, k, false, i / dim, i % dim);
        __syncwarp();
        warp_mat_mul(E, R, T, dim, lane);
        for (int i = lane; i < d2; i += 32) R[i] = T[i];
        __syncwarp();
    }
    for (int i = lane; i < d2; i += 32) kdst[i] = R[i];
    // with_deriv == 2: the derivative kernels are finished by build_block_derivs (one warp per member) instead of this warp
    // walking all members again
    if (with_deriv != 1 || op.n_params == 0) return;
    __syncwarp();
    for (int i = lane; i < d2; i += 32) R[i] = (i / dim == i % dim) ? cmake(1.0, 0.0) : czero();  // suffix
    __syncwarp();
    for (int j = nm - 1; j >= 0; --j) {
        const DevMember m = members[op.member_off + j];
        Trig t;
        member_trig(m, params, t);
        for (int p = 0; p < m.n_params; ++p) {
            cplx* dd = dkdst + (size_t)(m.slot0 + p) * d2;  // holds the prefix E_{j-1}..E_0
            slot_kernel(m.type, t, p, params[m.param_start + p], shift, k);  // (GENERAL members have no parameters)
            for (int i = lane; i < d2; i += 32) E[i] = embed_elem(m, k, true, i / dim, i % dim);
            __syncwarp();
            // T = dE * prefix (prefix read from global: written by this warp in pass 1)
            for (int i = lane; i < d2; i += 32) {
                const int r = i / dim, c = i - r * dim;
                cplx acc = czero();
                for (int l = 0; l < dim; ++l) acc = cfma(E[r * dim + l], dd[l * dim + c], acc);
                T[i] = acc;
            }
            __syncwarp();
            // dd = suffix * T
            for (int i = lane; i < d2; i += 32) {
                const int r = i / dim, c = i - r * dim;
                cplx acc = czero();
                for (int l = 0; l < dim; ++l) acc = cfma(R[r * dim + l], T[l * dim + c], acc);
                dd[i] = acc;
            }
            __syncwarp();
        }
        member_kernel(m, t, -1, pool, k);
        for (int i = lane; i < d2; i += 32) E[i] = embed_elem(m, k, false, i / dim, i % dim);
        __syncwarp();
        warp_mat_mul(R, E, T, dim, lane);
        for (int i = lane; i < d2; i += 32) R[i] = T[i];
        __syncwarp();
    }
}

// One warp per (parameter set b, op): fills the forward kernel table and the derivative kernel table.
// ktab[b * kern_total + op.kern_off + ...], dktab[b * dkern_total + op.dkern_off + slot * dim*dim + ...].

// Second half of the block tables: dM/dtheta_p = E_{nm-1} .. E_{j+1} dE_j(p) E_{j-1} .. E_0 for the parameters p of member j.
// build_block_warp left the prefix E_{j-1} .. E_0 in the parameter's slot; this warp rebuilds the suffix for ITS member (the
// same left-to-right products the single-warp version accumulated, so the tables are bit-identical) and finishes the slots.
// One warp per member: the critical path of a table build drops from ~2 P_block + 2 nm small matrix products to ~nm + 2 per
// parameter of one member -- it is what a single evaluation (BFGS) waits for before the executor starts.
__device__ inline void block_member_derivs(const DevOp& op, const DevMember* __restrict__ members, const double* __restrict__ params,
                                           const cplx* __restrict__ pool, cplx* __restrict__ dkdst, int j, cplx* ws, int lane, double shift) {
    const int dim = op.dim, d2 = dim * dim, nm = op.n_members;
    cplx* R = ws;        // suffix E_{nm-1} .. E_{j+1}
    cplx* E = ws + 64;   // embedded member
    cplx* T = ws + 128;  // product scratch
    cplx k[16];
    const DevMember m = members[op.member_off + j];
    if (m.n_params == 0) return;
    for (int i = lane; i < d2; i += 32) R[i] = (i / dim == i % dim) ? cmake(1.0, 0.0) : czero();
    __syncwarp();
    for (int jj = nm - 1; jj > j; --jj) {
        const DevMember mm = members[op.member_off + jj];
        Trig tt;
        member_trig(mm, params, tt);
        member_kernel(mm, tt, -1, pool, k);
        for (int i = lane; i < d2; i += 32) E[i] = embed_elem(mm, k, false, i / dim, i % dim);
        __syncwarp();
        warp_mat_mul(R, E, T, dim, lane);
        for (int i = lane; i < d2; i += 32) R[i] = T[i];
        __syncwarp();
    }
    Trig t;
    member_trig(m, params, t);
    for (int p = 0; p < m.n_params; ++p) {
        cplx* dd = dkdst + (size_t)(m.slot0 + p) * d2;  // holds the prefix E_{j-1}..E_0
        slot_kernel(m.type, t, p, params[m.param_start + p], shift, k);
        for (int i = lane; i < d2; i += 32) E[i] = embed_elem(m, k, true, i / dim, i % dim);
        __syncwarp();
        // T = dE * prefix (prefix read from global: written by build_kernel_tables)
        for (int i = lane; i < d2; i += 32) {
            const int r = i / dim, c = i - r * dim;
            cplx acc = czero();
            for (int l = 0; l < dim; ++l) acc = cfma(E[r * dim + l], dd[l * dim + c], acc);
            T[i] = acc;
        }
        __syncwarp();
        // dd = suffix * T
        for (int i = lane; i < d2; i += 32) {
            const int r = i / dim, c = i - r * dim;
            cplx acc = czero();
            for (int l = 0; l < dim; ++l) acc = cfma(R[r * dim + l], T[l * dim + c], acc);
            dd[i] = acc;
        }
        __syncwarp();
    }
}

static const int DERIV_WARPS = 8;
// one CTA per (parameter set, op), its warps deal the block's members
__global__ void __launch_bounds__(DERIV_WARPS * 32) build_block_derivs(const DevOp* __restrict__ ops, int n_ops, const DevMember* __restrict__ members,
                                                                       const double* __restrict__ params, int n_params, const cplx* __restrict__ pool,
                                                                       cplx* __restrict__ dktab, int dkern_total, double shift) {
    __shared__ cplx ws[DERIV_WARPS][192];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x / n_ops, kop = blockIdx.x - b * n_ops;
    const DevOp op = ops[kop];
    if (op.type != SQ_OP_BLOCK || op.n_params == 0 || op.kern_off < 0) return;
    const double* __restrict__ pb = params + (size_t)b * n_params;
    cplx* dkdst = dktab + (size_t)b * dkern_total + op.dkern_off;
    for (int j = warp; j < op.n_members; j += DERIV_WARPS) block_member_derivs(op, members, pb, pool, dkdst, j, ws[warp], lane, shift);
}

static const int TABLE_WARPS = 4;
__global__ void __launch_bounds__(TABLE_WARPS * 32) build_kernel_tables(
    const DevOp* __restrict__ ops, int n_ops, const DevMember* __restrict__ members, const double* __restrict__ params,
    int n_params, int batch, const cplx* __restrict__ pool, cplx* __restrict__ ktab, int kern_total,
    cplx* __restrict__ dktab, int dkern_total, int with_deriv, double shift) {
    __shared__ cplx ws[TABLE_WARPS][192];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long idx = (long long)blockIdx.x * TABLE_WARPS + warp;
    if (idx >= (long long)batch * n_ops) return;
    const int b = (int)(idx / n_ops);
    const DevOp op = ops[idx - (long long)b * n_ops];
    if (op.kern_off < 0) return;  // constant kernel (raw GENERAL) lives in the pool
    const double* __restrict__ pb = params + (size_t)b * n_params;
    cplx* kdst = ktab + (size_t)b * kern_total + op.kern_off;
    cplx* dkdst = dktab + (size_t)b * dkern_total + (op.dkern_off >= 0 ? op.dkern_off : 0);
    if (op.type == SQ_OP_BLOCK) {
        build_block_warp(op, members, pb, pool, kdst, dkdst, with_deriv, ws[warp], lane, shift);
        return;
    }
    if (lane != 0) return;
    Trig t;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        t.s[i] = 0.0;
        t.c[i] = 1.0;
        if (i < op.n_params) sincos(pb[op.param_start + i], &t.s[i], &t.c[i]);
    }
    cplx k[16];
    const int dim = build_gate_kernel(op.type, t, -1, k);
    // two-qubit kernels are built with kernel bit 0 = op.target; the op's local bit 0 is its LOWER qubit
    const bool flip = dim == 4 && op.target == op.q[1];
    auto src = [&](int i) {
        if (!flip) return i;
        const int r = i >> 2, c = i & 3;
        return ((((r & 1) << 1) | (r >> 1)) << 2) | (((c & 1) << 1) | (c >> 1));
    };
    for (int i = 0; i < dim * dim; ++i) kdst[i] = k[src(i)];
    if (with_deriv) {
        for (int p = 0; p < op.n_params; ++p) {
            slot_kernel(op.type, t, p, pb[op.param_start + p], shift, k);
            cplx* dd = dkdst + p * dim * dim;
            for (int i = 0; i < dim * dim; ++i) dd[i] = k[src(i)];
        }
    }
}

}  // namespace sq
