// exec_fused.cuh -- the fused circuit executor: one CTA keeps a tile of CT columns of the 2^n x C matrix in shared
// memory, runs the WHOLE device program on it, and emits only trace partials (cost), trace + W partials (adjoint
// gradient) or the transformed tile (apply / materialised derivative).
//
// It replaces, for one column tile, the reference's per-gate passes over the full matrix:
//   Gates_block::apply_to_inner forward loop        (gates/Gates_block.cpp:683-708)
//   apply_kernel_to_input row-pair update           (gates/kernels/apply_kernel_to_input.cpp:52-112)
//   apply_nqbit_kernel_to_matrix_input_impl         (gates/kernels/apply_large_kernel_to_input.cpp:123-213)
//   get_cost_function / get_trace* diagonals        (decomposition/N_Qubit_Decomposition_Cost_Function.cpp:73-664)
// and, for the gradient, the P materialised derivative matrices of Gates_block::apply_derivate_to
// (gates/Gates_block.cpp:1011-1150) by an adjoint sweep: with a_k the column after k ops and beta_k the row functional
// e_r^T G_{N-1}...G_{k+1}, the derivative of the trace term wrt a parameter of op k is sum_groups beta_k^T dK a_k.
// The executor accumulates W'_k[r][c] = sum_{groups, columns} beta_k[r] p_k[c] with p_k = K a_k the column AFTER the op
// (so the accumulation does not wait for the un-applied column); since a_k = K^dagger p_k, the wanted
// W_k[r][c] = sum beta_k[r] a_k[c] equals (W'_k conj(K))[r][c] and reduce_partials contracts W'_k with dK K^dagger
// (for fused blocks dK is the product-rule derivative of the block matrix, gate_kernels.cuh).
// Inactive (control = 0) pairs contribute nothing, which is exactly the reference's "zero rows in the derivative"
// convention (apply_kernel_to_input.cpp:93-97).
//
// The hot ops are the planner's fused blocks: dense 4x4 (two qubits) or 2x2 (one qubit) complex kernels without
// controls. Each thread keeps the kernel (and, in the backward sweep, the 16 W accumulators) in registers and owns
// whole amplitude groups, so a block costs one shared-memory round trip and one barrier instead of one per gate.
//
// Data layout in shared memory: element (row i, tile column c) at [phys(i) * CT + c], 16 B each, so a quarter-warp
// (8 lanes, one 128 B shared-memory wavefront) reads whole rows; phys() XOR-swizzles the low row bits so that rows that
// differ in the lane-varying bits land in different 16 B bank groups for every target qubit.
#pragma once
#include "sq_types.cuh"
#include "../../include/sqgpu.h"

namespace sq {

enum { MODE_COST = 0, MODE_GRAD = 1, MODE_APPLY = 2 };

struct OpTab;

struct ExecArgs {
    const cplx* in;          // input matrix (row-major, leading dimension ld_in)
    cplx* out;               // MODE_APPLY: output (may alias in)
    long long in_ystride;    // element stride of `in` per blockIdx.y (0: shared input)
    long long out_ystride;   // element stride of `out` per blockIdx.y
    int ld_in, ld_out;
    int rows, cols, n;       // rows = 2^n
    int ct, log_ct;          // tile width (columns), power of two
    int tiles, tiles_per_cta;
    const DevOp* ops;
    int n_ops;
    const cplx* ktab;        // [ysets][kern_total]
    int kern_total;
    const cplx* dktab;       // [ysets][dkern_total]
    int dkern_total;
    const cplx* pool;
    const struct OpTab* optabs;  // [ysets][n_ops] lookup tables of the DMMA block path (build_optabs)
    int k_shared;            // 1: every blockIdx.y uses kernel-table set 0 (materialised derivative: one parameter set)
    const int* deriv_op;     // MODE_APPLY: per blockIdx.y the op whose derivative kernel is applied (NULL: none)
    const int* deriv_slot;   //             and which of its derivative kernels
    int trace_offset;
    int n_trace_types;       // 1: main diagonal only, 2: + one-bit-flip sums, 3: + two-bit-flip sums
    double* tr_part;         // [y][chunks][6]
    cplx* w_part;            // [y][chunks][w_total]
    int w_total;
    int w_in_smem;           // accumulate W over the CTA's tiles in shared memory; else in the CTA's own (zero-initialised)
                             // slice of w_part with fire-and-forget reductions: one thread owns each address, so the
                             // summation order stays fixed
    const cplx* omega;       // [y][3] weights of the three trace types in the functional whose gradient is taken
    int dense_stage;         // complex elements of kernel staging for the generic dense path (0: none)
    int wmax;                // max dim*dim over parametric ops (complex), >= 4
};

static const int FUSED_THREADS = 512;

// ---- shared-memory swizzle -------------------------------------------------------------------------------------
// rows per 128 B wavefront: 8 / ct. The low log2(8/ct) bits of the physical row select the 16 B bank group set; they
// are XORed with images of the higher row bits so that rows differing in any of the lowest free bits do not collide.
template <int LOG_CT>
__device__ __forceinline__ int phys_row(int i) {
    if (LOG_CT >= 3) return i;
    if (LOG_CT == 2) return i ^ (__popc(i >> 1) & 1);
    if (LOG_CT == 1) {
        // image of row bit b: (1, 2, 3)[b % 3]; bits 0 and 1 map to themselves
        const unsigned m_lo = 0x6DB6DB6Cu;  // bits b >= 2 with b % 3 in {2, 0}
        const unsigned m_hi = 0x36DB6DB4u;  // bits b >= 2 with b % 3 in {2, 1}
        return i ^ ((__popc((unsigned)i & m_lo) & 1) | ((__popc((unsigned)i & m_hi) & 1) << 1));
    }
    return i ^ (((i >> 3) & 1) * 7);
}

// shared-memory element indices of the 4 rows of a two-qubit group: base | {0, b0, b1, b0|b1}. For CT = 4 the swizzle
// parity of (base | x) splits into parity(base) ^ parity(x), so one POPC per group suffices.
template <int LOG_CT>
__device__ __forceinline__ void group4_addr(int base, int b0, int b1, int c, int px0, int px1, int& e0, int& e1, int& e2, int& e3) {
    constexpr int CT = 1 << LOG_CT;
    if (LOG_CT == 2) {
        const int pb = __popc(base >> 1) & 1;
        e0 = ((base) ^ pb) * CT + c;
        e1 = ((base | b0) ^ (pb ^ px0)) * CT + c;
        e2 = ((base | b1) ^ (pb ^ px1)) * CT + c;
        e3 = ((base | b0 | b1) ^ (pb ^ px0 ^ px1)) * CT + c;
    } else {
        e0 = phys_row<LOG_CT>(base) * CT + c;
        e1 = phys_row<LOG_CT>(base | b0) * CT + c;
        e2 = phys_row<LOG_CT>(base | b1) * CT + c;
        e3 = phys_row<LOG_CT>(base | b0 | b1) * CT + c;
    }
}

// reduce 8 per-lane doubles over the warp; lanes with (lane & 3) == 0 end up holding the total of value
// index ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1) in v[0]   (9 double shuffles instead of 40)
__device__ __forceinline__ void warp_reduce8(double* v, int lane) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 16;
        const double send = up ? v[i] : v[i + 4];
        const double keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 8;
        const double send = up ? v[i] : v[i + 2];
        const double keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    {
        const bool up = lane & 4;
        const double send = up ? v[0] : v[1];
        const double keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 4);
    }
    v[0] += __shfl_xor_sync(full, v[0], 2);
    v[0] += __shfl_xor_sync(full, v[0], 1);
}

// per-warp W partial (complex w[n]) -> swarp slot; n = 4 or 16
template <int N>
__device__ __forceinline__ void warp_store_w(const cplx* w, double* slot, int lane) {
#pragma unroll
    for (int part = 0; part < N / 4; ++part) {
        double v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[2 * e] = w[part * 4 + e].x;
            v[2 * e + 1] = w[part * 4 + e].y;
        }
        warp_reduce8(v, lane);
        if ((lane & 3) == 0) slot[part * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = v[0];
    }
}

// ---- dense k-qubit blocks on the FP64 tensor cores ----------------------------------------------------------------
// A dense 2^k x 2^k complex kernel acting on a group of 2^k amplitudes is a real (2^(k+1)) x (2^(k+1)) matrix acting on
// the interleaved {re, im} vector:  Kreal[2r+a][2c+b] = { K.re if a == b;  -K.im if (a,b) = (0,1);  +K.im if (1,0) }.
// One warp transforms 8 (group, column) items per step with mma.sync.m8n8k4.f64 (DMMA): D[8 x 8 items] +=
// Kreal[8 x 4] * X[4 x 8 items], RT = 2^(k+1)/8 row tiles x KS = 2^(k+1)/4 k-steps. Replaces the reference's gather /
// dense matvec / scatter loop for 4-5 qubit kernels (apply_large_kernel_to_input.cpp:160-199,
// apply_large_kernel_to_input_AVX.cpp:90-136).
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

static const int DMMA_PAD = 4;  // doubles of row padding of the staged real kernel (conflict-free 8 x 4 fragment loads)

// forward application of a raw dense op (no controls) on the tile; requires nitems % 8 == 0. skr: staged real kernel
// [DIMR][DIMR + DMMA_PAD]; spat: row pattern of local index l. All warps of the CTA take part.
template <int LOG_CT, int KQ>
__device__ __forceinline__ void dense_dmma_forward(cplx* sa, const double* skr, const int* spat, const DevOp& op, int rows,
                                                   int tid, int nthr) {
    constexpr int CT = 1 << LOG_CT;
    constexpr int DIMR = 2 << KQ, RT = DIMR / 8, KS = DIMR / 4, LD = DIMR + DMMA_PAD;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    double* sad = reinterpret_cast<double*>(sa);
    const int nitems = (rows >> KQ) << LOG_CT;
    const int m = lane >> 2, kk = lane & 3;
    int q[KQ];
#pragma unroll
    for (int j = 0; j < KQ; ++j) q[j] = op.q[j];
    auto item_base = [&](int item, int& c) {
        c = item & (CT - 1);
        int base = item >> LOG_CT;
#pragma unroll
        for (int j = 0; j < KQ; ++j) base = insert_zero(base, q[j]);
        return base;
    };
    // A fragments: in registers when they fit (k <= 4: 32 doubles), else re-read from shared memory per DMMA
    constexpr bool A_IN_REGS = (RT * KS <= 32);
    double areg[A_IN_REGS ? RT * KS : 1];
    if (A_IN_REGS) {
#pragma unroll
        for (int rt = 0; rt < RT; ++rt)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) areg[rt * KS + ks] = skr[(8 * rt + m) * LD + 4 * ks + kk];
    }
    for (int b0 = warp * 8; b0 < nitems; b0 += nwarps * 8) {
        // B operand: lane holds real component 4*ks + kk of item b0 + m
        double bfrag[KS];
        {
            int c;
            const int base = item_base(b0 + m, c);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int comp = 4 * ks + kk;
                const int row = base | spat[comp >> 1];
                bfrag[ks] = sad[(phys_row<LOG_CT>(row) * CT + c) * 2 + (comp & 1)];
            }
        }
        double d[RT][2];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            d[rt][0] = 0.0;
            d[rt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double a = A_IN_REGS ? areg[rt * KS + ks] : skr[(8 * rt + m) * LD + 4 * ks + kk];
                dmma_m8n8k4(d[rt][0], d[rt][1], a, bfrag[ks]);
            }
        }
        __syncwarp();  // every lane has read its inputs of this 8-item batch before anyone overwrites them
        // D: lane holds output component 8*rt + m of items b0 + 2*kk and b0 + 2*kk + 1
        int c0, c1;
        const int base0 = item_base(b0 + 2 * kk, c0), base1 = item_base(b0 + 2 * kk + 1, c1);
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            const int comp = 8 * rt + m;
            const int pat = spat[comp >> 1];
            sad[(phys_row<LOG_CT>(base0 | pat) * CT + c0) * 2 + (comp & 1)] = d[rt][0];
            sad[(phys_row<LOG_CT>(base1 | pat) * CT + c1) * 2 + (comp & 1)] = d[rt][1];
        }
        __syncwarp();
    }
}

// ---- fused 2-/3-qubit blocks on the FP64 tensor cores -------------------------------------------------------------
// The block kernel K (4 x 4 or 8 x 8 complex) sits in shared memory (km); every lane derives the DMMA A-fragments of the
// real embeddings it needs: mode 0: K, 1: K^dagger, 2: K^T.
template <int DIM>
__device__ __forceinline__ double block_frag(const cplx* km, int x, int y, int mode) {
    const int r = x >> 1, a = x & 1, c = y >> 1, b = y & 1;
    const cplx e = (mode == 0) ? km[r * DIM + c] : km[c * DIM + r];
    const double im = (mode == 1) ? -e.y : e.y;
    return (a == b) ? e.x : (a ? im : -im);
}

// Addressing of the DMMA block path. phys_row() is GF(2)-linear (x ^ L(x >> s) with L linear), and so are the bit
// insertions that expand a group index into a row index; hence the shared-memory address of (item, component) splits into
//     address = B0(batch) ^ slot(lane, access)
// where B0 is computed once per 8-item batch and `slot` once per op and lane. Units: doubles (2 per complex element).
template <int LOG_CT, int KQ>
struct BlockGeom {
    static constexpr int CT = 1 << LOG_CT;
    static constexpr int LOGG = 3 - LOG_CT;  // log2(groups per 8-item batch)
    int q[3];   // block qubits, ascending (unused = 30)
    int Pq[3];  // address image of row bit q[j]
    int F[3];   // address image of the j-th lowest row bit that is NOT a block qubit (group-offset bits inside a batch)

    __device__ __forceinline__ void init(int q0, int q1, int q2) {
        q[0] = q0; q[1] = q1; q[2] = q2;
#pragma unroll
        for (int j = 0; j < 3; ++j) Pq[j] = (j < KQ) ? (phys_row<LOG_CT>(1 << q[j]) << (LOG_CT + 1)) : 0;
        int f = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            while (f == q0 || f == q1 || f == q2) ++f;
            F[j] = phys_row<LOG_CT>(1 << f) << (LOG_CT + 1);
            ++f;
        }
    }
    // B0 of the batch starting at item b0 (b0 % 8 == 0)
    __device__ __forceinline__ int batch_base(int b0) const {
        int base = b0 >> LOG_CT;
#pragma unroll
        for (int j = 0; j < KQ; ++j) base = insert_zero(base, q[j]);
        return phys_row<LOG_CT>(base) << (LOG_CT + 1);
    }
    // lane-constant part of the address of real component `comp` (= 2 * local amplitude + re/im) of batch item `item` (0..7)
    __device__ __forceinline__ int slot(int item, int comp) const {
        const int go = item >> LOG_CT, c = item & (CT - 1), amp = comp >> 1;
        int a = c * 2 + (comp & 1);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j < LOGG && ((go >> j) & 1)) a ^= F[j];
#pragma unroll
        for (int j = 0; j < KQ; ++j)
            if ((amp >> j) & 1) a ^= Pq[j];
        return a;
    }
};

// Per-op lookup table of the DMMA block path, built in shared memory by all threads while the PREVIOUS op runs
// (double-buffered): the A fragments of K, K^dagger, K^T in lane order and the lane-constant address slots.
struct OpTab {
    double frag[3][8][32];  // [mode][rt * KS + ks][lane]
    int slot[12][32];       // [0, KS): loads; [KS, KS + 2 RT): W' operands (h * RT + t); then stores (2 * rt + odd)
};

template <int KQ>
__device__ __forceinline__ int optab_entries() {
    constexpr int DIMR = 2 << KQ, RT = DIMR / 8, KS = DIMR / 4;
    return 3 * RT * KS * 32 + (KS + 4 * RT) * 32;
}

// which complex element of K entry e of the table needs (-1: none, a slot entry)
template <int KQ>
__device__ __forceinline__ int optab_kindex(int e) {
    constexpr int DIM = 1 << KQ, DIMR = 2 * DIM, RT = DIMR / 8, KS = DIMR / 4, NF = 3 * RT * KS * 32;
    if (e >= NF) return -1;
    const int mode = e / (RT * KS * 32), rem = e - mode * (RT * KS * 32), rtks = rem >> 5, lane = rem & 31;
    const int x = 8 * (rtks / KS) + (lane >> 2), y = 4 * (rtks % KS) + (lane & 3);
    return (mode == 0) ? (x >> 1) * DIM + (y >> 1) : (y >> 1) * DIM + (x >> 1);
}

template <int LOG_CT, int KQ>
__device__ __forceinline__ void optab_store(OpTab* T, int e, cplx kel, const BlockGeom<LOG_CT, KQ>& G) {
    constexpr int DIM = 1 << KQ, DIMR = 2 * DIM, RT = DIMR / 8, KS = DIMR / 4, NF = 3 * RT * KS * 32;
    if (e < NF) {
        const int mode = e / (RT * KS * 32), rem = e - mode * (RT * KS * 32), rtks = rem >> 5, lane = rem & 31;
        const int x = 8 * (rtks / KS) + (lane >> 2), y = 4 * (rtks % KS) + (lane & 3);
        const int a = x & 1, b = y & 1;
        const double im = (mode == 1) ? -kel.y : kel.y;
        T->frag[mode][rtks][lane] = (a == b) ? kel.x : (a ? im : -im);
        return;
    }
    const int r = e - NF, sidx = r >> 5, lane = r & 31, m = lane >> 2, kk = lane & 3;
    int v;
    if (sidx < KS) v = G.slot(m, 4 * sidx + kk);
    else if (sidx < KS + 2 * RT) {
        const int j = sidx - KS;
        v = G.slot(kk + 4 * (j / RT), 8 * (j % RT) + m);
    } else {
        const int j = sidx - KS - 2 * RT;
        v = G.slot(2 * kk + (j & 1), 8 * (j >> 1) + m);
    }
    T->slot[sidx][lane] = v;
}

// One CTA per (parameter set, op): fills the op's lookup table in global memory (ops the DMMA block path cannot take are
// skipped; their tables are never read).
template <int LOG_CT>
__device__ __forceinline__ void fill_optab(OpTab* T, const DevOp& op, const cplx* __restrict__ K) {
    if (op.dim == 8) {
        BlockGeom<LOG_CT, 3> G;
        G.init(op.q[0], op.q[1], op.q[2]);
        for (int e = threadIdx.x; e < optab_entries<3>(); e += blockDim.x) {
            const int ki = optab_kindex<3>(e);
            optab_store<LOG_CT, 3>(T, e, ki >= 0 ? K[ki] : czero(), G);
        }
    } else {
        BlockGeom<LOG_CT, 2> G;
        G.init(op.q[0], op.q[1], 30);
        for (int e = threadIdx.x; e < optab_entries<2>(); e += blockDim.x) {
            const int ki = optab_kindex<2>(e);
            optab_store<LOG_CT, 2>(T, e, ki >= 0 ? K[ki] : czero(), G);
        }
    }
}

__global__ void build_optabs(const DevOp* __restrict__ ops, int n_ops, const cplx* __restrict__ ktab, int kern_total,
                             const cplx* __restrict__ pool, int log_ct, OpTab* __restrict__ tabs) {
    const int b = blockIdx.x / n_ops, k = blockIdx.x - b * n_ops;
    const DevOp op = ops[k];
    if (op.ctrl_mask != 0 || (op.dim != 4 && op.dim != 8)) return;
    const cplx* K = op.kern_off >= 0 ? ktab + (size_t)b * kern_total + op.kern_off : pool + op.pool_off;
    OpTab* T = tabs + (size_t)b * n_ops + k;
    switch (log_ct) {
        case 0: fill_optab<0>(T, op, K); break;
        case 1: fill_optab<1>(T, op, K); break;
        case 2: fill_optab<2>(T, op, K); break;
        default: fill_optab<3>(T, op, K); break;
    }
}

// forward: x <- K x for every group of the tile. One warp handles 8 (group, column) items per step.
template <int LOG_CT, int KQ>
__device__ __forceinline__ void block_dmma_forward(cplx* sa, const OpTab* T, const BlockGeom<LOG_CT, KQ>& G, int rows, int tid,
                                                   int nthr) {
    constexpr int DIM = 1 << KQ, DIMR = 2 * DIM, RT = DIMR / 8, KS = DIMR / 4;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5, m = lane >> 2, kk = lane & 3;
    double* sad = reinterpret_cast<double*>(sa);
    double af[RT][KS];
    int sl_ld[KS], sl_st0[RT], sl_st1[RT];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) sl_ld[ks] = T->slot[ks][lane];
#pragma unroll
    for (int rt = 0; rt < RT; ++rt) {
        sl_st0[rt] = T->slot[KS + 2 * RT + 2 * rt][lane];
        sl_st1[rt] = T->slot[KS + 2 * RT + 2 * rt + 1][lane];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) af[rt][ks] = T->frag[0][rt * KS + ks][lane];
    }
    (void)m;
    (void)kk;
    const int nitems = (rows >> KQ) << LOG_CT;
    for (int b0 = warp * 8; b0 < nitems; b0 += nwarps * 8) {
        const int B0 = G.batch_base(b0);
        double xf[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) xf[ks] = sad[B0 ^ sl_ld[ks]];
        double d[RT][2];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            d[rt][0] = d[rt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) dmma_m8n8k4(d[rt][0], d[rt][1], af[rt][ks], xf[ks]);
        }
        __syncwarp();
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            sad[B0 ^ sl_st0[rt]] = d[rt][0];
            sad[B0 ^ sl_st1[rt]] = d[rt][1];
        }
        __syncwarp();
    }
}

// backward step of the adjoint sweep: a <- K^dagger p, beta <- K^T beta, W' += beta p^T (all on DMMA); the warp's W'
// (DIM x DIM complex) is written to wslot after the loop.
template <int LOG_CT, int KQ>
__device__ __forceinline__ void block_dmma_backward(cplx* sa, cplx* sb, const OpTab* T, const BlockGeom<LOG_CT, KQ>& G, int rows,
                                                    bool has_w, cplx* wslot, int tid, int nthr) {
    constexpr int DIM = 1 << KQ, DIMR = 2 * DIM, RT = DIMR / 8, KS = DIMR / 4;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5, m = lane >> 2, kk = lane & 3;
    double* sad = reinterpret_cast<double*>(sa);
    double* sbd = reinterpret_cast<double*>(sb);
    double adag[RT][KS], atr[RT][KS];
    int sl_ld[KS], sl_w[2][RT], sl_st0[RT], sl_st1[RT];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) sl_ld[ks] = T->slot[ks][lane];
#pragma unroll
    for (int rt = 0; rt < RT; ++rt) {
        sl_w[0][rt] = T->slot[KS + rt][lane];
        sl_w[1][rt] = T->slot[KS + RT + rt][lane];
        sl_st0[rt] = T->slot[KS + 2 * RT + 2 * rt][lane];
        sl_st1[rt] = T->slot[KS + 2 * RT + 2 * rt + 1][lane];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            adag[rt][ks] = T->frag[1][rt * KS + ks][lane];
            atr[rt][ks] = T->frag[2][rt * KS + ks][lane];
        }
    }
    double pacc[RT][RT][2];
#pragma unroll
    for (int tr = 0; tr < RT; ++tr)
#pragma unroll
        for (int tc = 0; tc < RT; ++tc) pacc[tr][tc][0] = pacc[tr][tc][1] = 0.0;
    const int nitems = (rows >> KQ) << LOG_CT;
    for (int b0 = warp * 8; b0 < nitems; b0 += nwarps * 8) {
        const int B0 = G.batch_base(b0);
        double pf[KS], bf[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            pf[ks] = sad[B0 ^ sl_ld[ks]];
            bf[ks] = sbd[B0 ^ sl_ld[ks]];
        }
        if (has_w) {  // W' operands: component 8t + m of items kk and kk + 4
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double pw[RT], bw[RT];
#pragma unroll
                for (int t = 0; t < RT; ++t) {
                    pw[t] = sad[B0 ^ sl_w[h][t]];
                    bw[t] = sbd[B0 ^ sl_w[h][t]];
                }
#pragma unroll
                for (int tr = 0; tr < RT; ++tr)
#pragma unroll
                    for (int tc = 0; tc < RT; ++tc) dmma_m8n8k4(pacc[tr][tc][0], pacc[tr][tc][1], bw[tr], pw[tc]);
            }
        }
        double da[RT][2], db[RT][2];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            da[rt][0] = da[rt][1] = db[rt][0] = db[rt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                dmma_m8n8k4(da[rt][0], da[rt][1], adag[rt][ks], pf[ks]);
                dmma_m8n8k4(db[rt][0], db[rt][1], atr[rt][ks], bf[ks]);
            }
        }
        __syncwarp();  // all reads of this batch are done
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            sad[B0 ^ sl_st0[rt]] = da[rt][0];
            sad[B0 ^ sl_st1[rt]] = da[rt][1];
            sbd[B0 ^ sl_st0[rt]] = db[rt][0];
            sbd[B0 ^ sl_st1[rt]] = db[rt][1];
        }
        __syncwarp();
    }
    if (has_w) {
        // lane (m, kk) holds P[8tr+m][8tc+2kk], P[8tr+m][8tc+2kk+1]; rows 2r (even m) and 2r+1 (odd m) combine to
        // W'[r][c] = (P[2r][2c] - P[2r+1][2c+1]) + i (P[2r][2c+1] + P[2r+1][2c]),  r = 4tr + m/2, c = 4tc + kk
#pragma unroll
        for (int tr = 0; tr < RT; ++tr)
#pragma unroll
            for (int tc = 0; tc < RT; ++tc) {
                const double o0 = __shfl_xor_sync(0xffffffffu, pacc[tr][tc][0], 4);
                const double o1 = __shfl_xor_sync(0xffffffffu, pacc[tr][tc][1], 4);
                if ((m & 1) == 0) wslot[(4 * tr + (m >> 1)) * DIM + 4 * tc + kk] = cmake(pacc[tr][tc][0] - o1, pacc[tr][tc][1] + o0);
            }
    }
}

// compact per-op record staged in shared memory (uniform reads, no global latency on the critical path)
struct SOp {
    int32_t dim;       // 2 / 4 / 8 for block ops
    int32_t q0, q1, q2;  // block ops: qubits ascending (unused entries 30)
    int32_t kern_off;  // kernel-table offset, -1: pool
    int32_t w_off;
    int32_t kind;      // 0: scalar 2x2 block, 2: DMMA 4x4 / 8x8 block, 1: generic path (controls, raw dense, derivative op)
    int32_t pool_lo, pool_hi;
};

static const int KM_ELEMS = 64;  // prefetched block kernel: up to 8 x 8 complex

template <int MODE, int LOG_CT>
__global__ void __launch_bounds__(FUSED_THREADS, 1) fused_exec(const ExecArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CT = 1 << LOG_CT;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int rows = A.rows;
    const int y = blockIdx.y;
    const int kset = A.k_shared ? 0 : y;
    const cplx* __restrict__ ktab = A.ktab + (size_t)kset * A.kern_total;
    const cplx* __restrict__ dktab = A.dktab ? A.dktab + (size_t)kset * A.dkern_total : nullptr;

    cplx* sa = reinterpret_cast<cplx*>(smem_raw);
    cplx* sb = sa + (size_t)rows * CT;                                   // MODE_GRAD only
    cplx* sk = (MODE == MODE_GRAD) ? sb + (size_t)rows * CT : sb;         // raw dense kernel staging
    cplx* skm = sk + A.dense_stage;                                       // [2][KM_ELEMS] prefetched block kernels
    OpTab* stab = reinterpret_cast<OpTab*>(skm + 2 * KM_ELEMS);           // [2] DMMA block lookup tables
    cplx* swarp = reinterpret_cast<cplx*>(stab + 2);                      // [2][nwarps][wmax]
    cplx* swacc = swarp + ((MODE == MODE_GRAD) ? 2 * nwarps * A.wmax : 0);  // [w_total] if w_in_smem
    double* sred = reinterpret_cast<double*>(swacc + ((MODE == MODE_GRAD && A.w_in_smem) ? A.w_total : 0));  // [nwarps][6]
    SOp* sops = reinterpret_cast<SOp*>(sred + nwarps * 6);               // [n_ops]

    const int chunk = blockIdx.x;
    const int nchunks = gridDim.x;
    const int deriv_op = (MODE == MODE_APPLY && A.deriv_op) ? A.deriv_op[y] : -1;
    const int deriv_slot = (MODE == MODE_APPLY && A.deriv_slot) ? A.deriv_slot[y] : 0;

    // ---- stage the op table ------------------------------------------------------------------------------------------
    for (int k = tid; k < A.n_ops; k += nthr) {
        const DevOp op = A.ops[k];
        SOp s;
        s.dim = op.dim;
        s.kern_off = op.kern_off;
        s.w_off = op.w_off;
        s.pool_lo = (int32_t)(op.pool_off & 0xffffffffLL);
        s.pool_hi = (int32_t)(op.pool_off >> 32);
        s.kind = 1;
        if (op.ctrl_mask == 0 && k != deriv_op) {
            if (op.dim == 2) s.kind = 0;
            else if ((op.dim == 4 || op.dim == 8) && ((((rows >> op.nq) << LOG_CT) & 7) == 0) && nthr >= 32) s.kind = 2;
        }
        s.q0 = op.dim == 2 ? op.target : op.q[0];
        s.q1 = op.dim == 2 ? 30 : op.q[1];
        s.q2 = op.dim == 8 ? op.q[2] : 30;
        sops[k] = s;
    }
    if (MODE == MODE_GRAD && A.w_in_smem) {
        for (int e = tid; e < A.w_total; e += nthr) swacc[e] = czero();
    }
    double tsum[6] = {0, 0, 0, 0, 0, 0};  // running trace sums of this CTA (thread 0)
    __syncthreads();

    // kernel element this thread prefetches for op k (threads 0..63)
    auto kernel_elem = [&](int k) -> cplx {
        const SOp s = sops[k];
        if (s.kind == 1 || tid >= s.dim * s.dim) return czero();
        const cplx* K = s.kern_off >= 0 ? ktab + s.kern_off : A.pool + (((long long)s.pool_hi << 32) | (unsigned)s.pool_lo);
        return K[tid];
    };

    // DMMA block table of op k: built per (parameter set, op) by build_optabs; copied global -> shared with cp.async while
    // the previous op computes (no registers held across the DMMA loops), into stab[k & 1]
    const OpTab* __restrict__ gtabs = A.optabs + (size_t)kset * A.n_ops;
    auto tab_prefetch = [&](int k) {
        if (sops[k].kind != 2) return;
        const char* src = reinterpret_cast<const char*>(gtabs + k);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(stab + (k & 1));
        for (int e = tid; e < (int)(sizeof(OpTab) / 16); e += nthr)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + e * 16), "l"(src + (size_t)e * 16));
    };
    auto tab_wait = [&]() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); };

    for (int ti = 0; ti < A.tiles_per_cta; ++ti) {
        const int tile = chunk * A.tiles_per_cta + ti;
        if (tile >= A.tiles) break;
        const int j0 = tile * CT;
        const int valid = min(CT, A.cols - j0);

        // ---- load the tile ------------------------------------------------------------------------------------
        {
            const cplx* __restrict__ src = A.in + (size_t)y * A.in_ystride + j0;
            for (int e = tid; e < rows * CT; e += nthr) {
                const int i = e >> LOG_CT, c = e & (CT - 1);
                cplx v = czero();
                if (c < valid) v = src[(size_t)i * A.ld_in + c];
                sa[phys_row<LOG_CT>(i) * CT + c] = v;
            }
            if (A.n_ops > 0 && tid < KM_ELEMS) skm[tid] = kernel_elem(0);
            if (A.n_ops > 0) {
                tab_prefetch(0);
                tab_wait();
            }
        }
        __syncthreads();

        // ---- forward sweep: op 0 first (Gates_block.cpp:683) ---------------------------------------------------
        for (int k = 0; k < A.n_ops; ++k) {
            const SOp s = sops[k];
            cplx next_elem = czero();
            const bool have_next = k + 1 < A.n_ops;
            if (have_next && tid < KM_ELEMS) next_elem = kernel_elem(k + 1);
            if (have_next) tab_prefetch(k + 1);
            const cplx* __restrict__ km = skm + (k & 1) * KM_ELEMS;
            if (s.kind == 2) {
                if (s.dim == 8) {
                    BlockGeom<LOG_CT, 3> G;
                    G.init(s.q0, s.q1, s.q2);
                    block_dmma_forward<LOG_CT, 3>(sa, stab + (k & 1), G, rows, tid, nthr);
                } else {
                    BlockGeom<LOG_CT, 2> G;
                    G.init(s.q0, s.q1, 30);
                    block_dmma_forward<LOG_CT, 2>(sa, stab + (k & 1), G, rows, tid, nthr);
                }
            } else if (s.kind == 0) {
                const cplx k00 = km[0], k01 = km[1], k10 = km[2], k11 = km[3];
                const int tbit = 1 << s.q0;
                const int nitems = (rows >> 1) << LOG_CT;
                for (int item = tid; item < nitems; item += nthr) {
                    const int c = item & (CT - 1);
                    const int i0 = insert_zero(item >> LOG_CT, s.q0);
                    const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                    const cplx a0 = sa[e0], a1 = sa[e1];
                    sa[e0] = cfma(k01, a1, cmul(k00, a0));
                    sa[e1] = cfma(k11, a1, cmul(k10, a0));
                }
            } else {
                // ---- generic path: controlled gates, raw dense kernels, derivative kernels ------------------------
                const DevOp& op = A.ops[k];
                const bool deriv = (MODE == MODE_APPLY) && (k == deriv_op);
                const int dim = op.dim;
                const cplx* __restrict__ K = deriv ? dktab + op.dkern_off + deriv_slot * dim * dim
                                                    : (op.kern_off >= 0 ? ktab + op.kern_off : A.pool + op.pool_off);
                if (dim == 2) {
                    const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
                    const int tbit = 1 << op.target;
                    const unsigned cm = op.ctrl_mask;
                    if (!deriv) {
                        const int f0 = op.fix[0], f1 = op.fix[1], f2 = op.fix[2];
                        const int nitems = (rows >> op.nfix) << LOG_CT;
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            const int i0 = insert_zero(insert_zero(insert_zero(item >> LOG_CT, f0), f1), f2) | cm;
                            const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                            const cplx a0 = sa[e0], a1 = sa[e1];
                            sa[e0] = cfma(k01, a1, cmul(k00, a0));
                            sa[e1] = cfma(k11, a1, cmul(k10, a0));
                        }
                    } else {  // derivative kernel: inactive pairs are zero-filled (apply_kernel_to_input.cpp:93-97)
                        const int nitems = (rows >> 1) << LOG_CT;
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            const int i0 = insert_zero(item >> LOG_CT, op.target);
                            const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                            if ((i0 & cm) == cm) {
                                const cplx a0 = sa[e0], a1 = sa[e1];
                                sa[e0] = cfma(k01, a1, cmul(k00, a0));
                                sa[e1] = cfma(k11, a1, cmul(k10, a0));
                            } else {
                                sa[e0] = czero();
                                sa[e1] = czero();
                            }
                        }
                    }
                } else {
                    // dense dim x dim kernel on ascending qubits (apply_large_kernel_to_input.cpp:160-199)
                    const int nq = op.nq;
                    const int nitems = (rows >> nq) << LOG_CT;
                    const bool use_dmma = !deriv && op.ctrl_mask == 0 && nq >= 3 && (nitems & 7) == 0;
                    if (use_dmma) {
                        // stage the real embedding of K (padded rows) and the local-index -> row-bit pattern
                        const int dimr = 2 * dim, ld = dimr + DMMA_PAD;
                        double* skr = reinterpret_cast<double*>(sk);
                        int* spat = reinterpret_cast<int*>(skr + dimr * ld);
                        for (int e = tid; e < dimr * dimr; e += nthr) {
                            const int r = e / dimr, cidx = e - r * dimr;
                            const cplx kv = K[(r >> 1) * dim + (cidx >> 1)];
                            const int a = r & 1, b = cidx & 1;
                            skr[r * ld + cidx] = (a == b) ? kv.x : (a ? kv.y : -kv.y);
                        }
                        for (int l = tid; l < dim; l += nthr) {
                            int r = 0;
                            for (int j = 0; j < nq; ++j) r |= ((l >> j) & 1) << op.q[j];
                            spat[l] = r;
                        }
                        __syncthreads();
                        if (nq == 3) dense_dmma_forward<LOG_CT, 3>(sa, skr, spat, op, rows, tid, nthr);
                        else if (nq == 4) dense_dmma_forward<LOG_CT, 4>(sa, skr, spat, op, rows, tid, nthr);
                        else dense_dmma_forward<LOG_CT, 5>(sa, skr, spat, op, rows, tid, nthr);
                    } else {
                        for (int e = tid; e < dim * dim; e += nthr) sk[e] = K[e];
                        __syncthreads();
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            int base = item >> LOG_CT;
                            for (int j = 0; j < nq; ++j) base = insert_zero(base, op.q[j]);
                            const bool active = (base & op.ctrl_mask) == op.ctrl_mask;
                            if (!active && !deriv) continue;
                            cplx v[32];
                            for (int l = 0; l < dim; ++l) {
                                int r = base;
                                for (int j = 0; j < nq; ++j) r |= ((l >> j) & 1) << op.q[j];
                                v[l] = active ? sa[phys_row<LOG_CT>(r) * CT + c] : czero();
                            }
                            for (int ro = 0; ro < dim; ++ro) {
                                cplx acc = czero();
                                if (active)
                                    for (int l = 0; l < dim; ++l) acc = cfma(sk[ro * dim + l], v[l], acc);
                                int r = base;
                                for (int j = 0; j < nq; ++j) r |= ((ro >> j) & 1) << op.q[j];
                                sa[phys_row<LOG_CT>(r) * CT + c] = acc;
                            }
                        }
                    }
                }
            }
            if (have_next && tid < KM_ELEMS) skm[((k + 1) & 1) * KM_ELEMS + tid] = next_elem;
            tab_wait();
            __syncthreads();
        }

        if (MODE == MODE_APPLY) {
            cplx* __restrict__ dst = A.out + (size_t)y * A.out_ystride + j0;
            for (int e = tid; e < rows * CT; e += nthr) {
                const int i = e >> LOG_CT, c = e & (CT - 1);
                if (c < valid) dst[(size_t)i * A.ld_out + c] = sa[phys_row<LOG_CT>(i) * CT + c];
            }
            __syncthreads();
            continue;
        }

        // ---- trace terms: sum_j M[(j+off) ^ mask, j] (N_Qubit_Decomposition_Cost_Function.cpp:137-160,191-404) ----
        {
            double t[6] = {0, 0, 0, 0, 0, 0};
            const int off = A.trace_offset;
            if (tid < valid) {
                const cplx v = sa[phys_row<LOG_CT>(j0 + tid + off) * CT + tid];
                t[0] = v.x;
                t[1] = v.y;
            }
            if (A.n_trace_types > 1) {
                for (int e = tid; e < A.n * CT; e += nthr) {
                    const int c = e & (CT - 1), qb = e >> LOG_CT;
                    if (c < valid) {
                        const cplx v = sa[phys_row<LOG_CT>((j0 + c + off) ^ (1 << qb)) * CT + c];
                        t[2] += v.x;
                        t[3] += v.y;
                    }
                }
            }
            if (A.n_trace_types > 2) {
                int e = 0;
                for (int q1 = 0; q1 < A.n - 1; ++q1)
                    for (int q2 = q1 + 1; q2 < A.n; ++q2)
                        for (int c = 0; c < valid; ++c, ++e)
                            if (e % nthr == tid) {
                                const cplx v = sa[phys_row<LOG_CT>((j0 + c + off) ^ ((1 << q1) | (1 << q2))) * CT + c];
                                t[4] += v.x;
                                t[5] += v.y;
                            }
            }
            const int nt = 2 * A.n_trace_types;
            for (int i = 0; i < nt; ++i)
                for (int sft = 16; sft > 0; sft >>= 1) t[i] += __shfl_xor_sync(0xffffffffu, t[i], sft);
            if (lane == 0)
                for (int i = 0; i < nt; ++i) sred[warp * 6 + i] = t[i];
            __syncthreads();
            if (tid == 0) {
                for (int i = 0; i < nt; ++i) {
                    double sum = 0;
                    for (int w = 0; w < nwarps; ++w) sum += sred[w * 6 + i];
                    tsum[i] += sum;
                }
            }
        }

        if (MODE == MODE_GRAD) {
            // ---- beta_N = sum_t omega_t * sum_{masks of type t} e_{(j+off)^mask} per column ---------------------
            for (int e = tid; e < rows * CT; e += nthr) sb[e] = czero();
            if (A.n_ops > 0 && tid < KM_ELEMS) skm[((A.n_ops - 1) & 1) * KM_ELEMS + tid] = kernel_elem(A.n_ops - 1);
            if (A.n_ops > 0) {
                tab_prefetch(A.n_ops - 1);
                tab_wait();
            }
            __syncthreads();
            {
                const int off = A.trace_offset;
                const cplx w0 = A.omega[(size_t)y * 3 + 0];
                if (tid < valid) sb[phys_row<LOG_CT>(j0 + tid + off) * CT + tid] = w0;
                if (A.n_trace_types > 1) {
                    const cplx w1 = A.omega[(size_t)y * 3 + 1];
                    for (int e = tid; e < A.n * CT; e += nthr) {
                        const int c = e & (CT - 1), qb = e >> LOG_CT;
                        if (c < valid) sb[phys_row<LOG_CT>((j0 + c + off) ^ (1 << qb)) * CT + c] = w1;
                    }
                }
                if (A.n_trace_types > 2) {
                    const cplx w2 = A.omega[(size_t)y * 3 + 2];
                    int e = 0;
                    for (int q1 = 0; q1 < A.n - 1; ++q1)
                        for (int q2 = q1 + 1; q2 < A.n; ++q2)
                            for (int c = 0; c < valid; ++c, ++e)
                                if (e % nthr == tid)
                                    sb[phys_row<LOG_CT>((j0 + c + off) ^ ((1 << q1) | (1 << q2))) * CT + c] = w2;
                }
            }
            __syncthreads();

            // ---- backward sweep ---------------------------------------------------------------------------------
            int buf = 0;
            for (int k = A.n_ops - 1; k >= 0; --k) {
                const SOp s = sops[k];
                cplx next_elem = czero();
                const bool have_next = k > 0;
                if (have_next && tid < KM_ELEMS) next_elem = kernel_elem(k - 1);
                if (have_next) tab_prefetch(k - 1);
                const bool has_w = s.w_off >= 0;
                cplx* wslot_c = swarp + (size_t)(buf * nwarps + warp) * A.wmax;
                double* wslot = reinterpret_cast<double*>(wslot_c);
                const cplx* __restrict__ km = skm + (k & 1) * KM_ELEMS;
                int wdim = s.dim;
                if (s.kind == 2) {
                    if (s.dim == 8) {
                        BlockGeom<LOG_CT, 3> G;
                        G.init(s.q0, s.q1, s.q2);
                        block_dmma_backward<LOG_CT, 3>(sa, sb, stab + (k & 1), G, rows, has_w, wslot_c, tid, nthr);
                    } else {
                        BlockGeom<LOG_CT, 2> G;
                        G.init(s.q0, s.q1, 30);
                        block_dmma_backward<LOG_CT, 2>(sa, sb, stab + (k & 1), G, rows, has_w, wslot_c, tid, nthr);
                    }
                } else if (s.kind == 0) {
                    const cplx k00 = km[0], k01 = km[1], k10 = km[2], k11 = km[3];
                    const int tbit = 1 << s.q0;
                    cplx W[4] = {czero(), czero(), czero(), czero()};
                    const int nitems = (rows >> 1) << LOG_CT;
                    for (int item = tid; item < nitems; item += nthr) {
                        const int c = item & (CT - 1);
                        const int i0 = insert_zero(item >> LOG_CT, s.q0);
                        const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                        const cplx p0 = sa[e0], p1 = sa[e1], b0 = sb[e0], b1 = sb[e1];
                        sa[e0] = cfmac(k10, p1, cfmac(k00, p0, czero()));
                        sa[e1] = cfmac(k11, p1, cfmac(k01, p0, czero()));
                        if (has_w) {
                            W[0] = cfma(b0, p0, W[0]);
                            W[1] = cfma(b0, p1, W[1]);
                            W[2] = cfma(b1, p0, W[2]);
                            W[3] = cfma(b1, p1, W[3]);
                        }
                        sb[e0] = cfma(k10, b1, cmul(k00, b0));
                        sb[e1] = cfma(k11, b1, cmul(k01, b0));
                    }
                    if (has_w) warp_store_w<4>(W, wslot, lane);
                } else {
                    const DevOp& op = A.ops[k];
                    const cplx* __restrict__ K = op.kern_off >= 0 ? ktab + op.kern_off : A.pool + op.pool_off;
                    wdim = op.dim;
                    if (op.dim == 2) {
                        const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
                        const int tbit = 1 << op.target;
                        const unsigned cm = op.ctrl_mask;
                        const int f0 = op.fix[0], f1 = op.fix[1], f2 = op.fix[2];
                        cplx W[4] = {czero(), czero(), czero(), czero()};
                        const int nitems = (rows >> op.nfix) << LOG_CT;
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            const int i0 = insert_zero(insert_zero(insert_zero(item >> LOG_CT, f0), f1), f2) | cm;
                            const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                            const cplx p0 = sa[e0], p1 = sa[e1], b0 = sb[e0], b1 = sb[e1];
                            sa[e0] = cfmac(k10, p1, cfmac(k00, p0, czero()));
                            sa[e1] = cfmac(k11, p1, cfmac(k01, p0, czero()));
                            if (has_w) {
                                W[0] = cfma(b0, p0, W[0]);
                                W[1] = cfma(b0, p1, W[1]);
                                W[2] = cfma(b1, p0, W[2]);
                                W[3] = cfma(b1, p1, W[3]);
                            }
                            sb[e0] = cfma(k10, b1, cmul(k00, b0));
                            sb[e1] = cfma(k11, b1, cmul(k01, b0));
                        }
                        if (has_w) warp_store_w<4>(W, wslot, lane);
                    } else {
                        // raw dense op (controlled two-target gates, GENERAL blocks, or a block too small for the tensor
                        // path): thread per (group, column), local arrays
                        const int dim = op.dim, nq = op.nq;
                        for (int e = tid; e < dim * dim; e += nthr) sk[e] = K[e];
                        __syncthreads();
                        const int nitems = (rows >> nq) << LOG_CT;
                        cplx wl[64];  // parametric dense ops have dim <= 8
                        if (has_w)
                            for (int e = 0; e < dim * dim; ++e) wl[e] = czero();
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            int base = item >> LOG_CT;
                            for (int j = 0; j < nq; ++j) base = insert_zero(base, op.q[j]);
                            if ((base & op.ctrl_mask) != op.ctrl_mask) continue;
                            cplx pv[32], bv[32];
                            int addr[32];
                            for (int l = 0; l < dim; ++l) {
                                int r = base;
                                for (int j = 0; j < nq; ++j) r |= ((l >> j) & 1) << op.q[j];
                                addr[l] = phys_row<LOG_CT>(r) * CT + c;
                                pv[l] = sa[addr[l]];
                                bv[l] = sb[addr[l]];
                            }
                            for (int ro = 0; ro < dim; ++ro) {
                                cplx acc = czero(), bacc = czero();
                                for (int l = 0; l < dim; ++l) {
                                    acc = cfmac(sk[l * dim + ro], pv[l], acc);   // (K^dagger p)[ro]
                                    bacc = cfma(sk[l * dim + ro], bv[l], bacc);  // (K^T beta)[ro]
                                }
                                sa[addr[ro]] = acc;
                                sb[addr[ro]] = bacc;
                                if (has_w)
                                    for (int r2 = 0; r2 < dim; ++r2) wl[r2 * dim + ro] = cfma(bv[r2], pv[ro], wl[r2 * dim + ro]);
                            }
                        }
                        if (has_w) {
                            for (int part = 0; part < dim * dim / 4; ++part) {
                                double v[8];
                                for (int e = 0; e < 4; ++e) {
                                    v[2 * e] = wl[part * 4 + e].x;
                                    v[2 * e + 1] = wl[part * 4 + e].y;
                                }
                                warp_reduce8(v, lane);
                                if ((lane & 3) == 0) wslot[part * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = v[0];
                            }
                        }
                    }
                }
                if (have_next && tid < KM_ELEMS) skm[((k - 1) & 1) * KM_ELEMS + tid] = next_elem;
                tab_wait();
                __syncthreads();
                if (has_w) {
                    const int nd = 2 * wdim * wdim;  // doubles
                    for (int e = tid; e < nd; e += nthr) {
                        double sum = 0;
                        for (int w = 0; w < nwarps; ++w)
                            sum += reinterpret_cast<const double*>(swarp + (size_t)(buf * nwarps + w) * A.wmax)[e];
                        if (A.w_in_smem) reinterpret_cast<double*>(swacc + s.w_off)[e] += sum;
                        else
                            atomicAdd(reinterpret_cast<double*>(A.w_part + ((size_t)y * nchunks + chunk) * A.w_total + s.w_off) + e, sum);
                    }
                    buf ^= 1;
                }
            }
            __syncthreads();
        }
    }

    if (MODE != MODE_APPLY) {
        if (tid == 0) {
            double* dst = A.tr_part + ((size_t)y * nchunks + chunk) * 6;
            for (int i = 0; i < 6; ++i) dst[i] = tsum[i];
        }
        if (MODE == MODE_GRAD && A.w_in_smem) {
            __syncthreads();
            cplx* dst = A.w_part + ((size_t)y * nchunks + chunk) * A.w_total;
            for (int e = tid; e < A.w_total; e += nthr) dst[e] = swacc[e];
        }
    }
}

}  // namespace sq
