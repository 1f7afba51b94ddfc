// exec_fused.cuh -- the fused circuit executor: one CTA keeps a tile of CT columns of the 2^n x C matrix in shared
// memory, runs the WHOLE gate program on it, and emits only trace partials (cost), trace + W partials (adjoint
// gradient) or the transformed tile (apply / materialised derivative).
//
// It replaces, for one column tile, the reference's per-gate passes over the full matrix:
//   Gates_block::apply_to_inner forward loop        (gates/Gates_block.cpp:683-708)
//   apply_kernel_to_input row-pair update           (gates/kernels/apply_kernel_to_input.cpp:52-112)
//   apply_nqbit_kernel_to_matrix_input_impl         (gates/kernels/apply_large_kernel_to_input.cpp:123-213)
//   get_cost_function / get_trace* diagonals        (decomposition/N_Qubit_Decomposition_Cost_Function.cpp:73-664)
// and, for the gradient, the P materialised derivative matrices of Gates_block::apply_derivate_to
// (gates/Gates_block.cpp:1011-1150) by an adjoint sweep: with a_k the column after k gates and beta_k the row functional
// e_r^T G_{N-1}...G_{k+1}, the derivative of the trace term wrt a parameter of gate k is sum_pairs beta_k^T dK a_k.
// The executor accumulates W_k[r][c] = sum_{active pairs, columns} beta_k[r] a_k[c]; the finalize kernel contracts W_k
// with the reference's derivative kernels dK. Inactive (control = 0) pairs contribute nothing, which is exactly the
// reference's "zero rows in the derivative" convention (apply_kernel_to_input.cpp:93-97).
//
// Data layout in shared memory: element (row i, tile column c) at [phys(i) * CT + c], 16 B each, so a quarter-warp
// (8 lanes, one 128 B shared-memory wavefront) reads whole rows; phys() XOR-swizzles the low row bits so that rows that
// differ in the lane-varying bits land in different 16 B bank groups for every target qubit.
#pragma once
#include "sq_types.cuh"
#include "../../include/sqgpu.h"

namespace sq {

enum { MODE_COST = 0, MODE_GRAD = 1, MODE_APPLY = 2 };

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
    int k_shared;            // 1: every blockIdx.y uses kernel-table set 0 (materialised derivative: one parameter set)
    const int* deriv_op;     // MODE_APPLY: per blockIdx.y the op whose derivative kernel is applied (NULL: none)
    const int* deriv_pidx;   //             and which of its parameters
    int trace_offset;
    int n_trace_types;       // 1: main diagonal only, 2: + one-bit-flip sums, 3: + two-bit-flip sums
    double* tr_part;         // [y][chunks][6]
    cplx* w_part;            // [y][chunks][w_total]
    int w_total;
    int w_in_smem;           // accumulate W over the CTA's tiles in shared memory (else tiles_per_cta must be 1)
    const cplx* omega;       // [y][3] weights of the three trace types in the functional whose gradient is taken
    int has_dense;           // program contains dim > 2 ops (reserves the 16 KB kernel staging buffer)
    int wmax;                // max dim*dim over parametric ops (complex), >= 4
};

static const int FUSED_THREADS = 512;
static const int DENSE_STAGE = 1024;  // complex elements (32 x 32)

// ---- shared-memory swizzle -------------------------------------------------------------------------------------
// rows per 128 B wavefront: 8 / ct. The low log2(8/ct) bits of the physical row select the 16 B bank group set.
template <int LOG_CT>
__device__ __forceinline__ int phys_row(int i) {
    if (LOG_CT >= 3) return i;
    if (LOG_CT == 2) return i ^ (__popc(i >> 1) & 1);
    if (LOG_CT == 1) {
        // images of row bit b >= 2: {3,1,2} for b % 3 == {2,0,1}; bits 0,1 map to themselves
        const unsigned lo = __popc(i & 0x36DB6DB4) & 1;  // bits b>=2 with b%3 in {2,0}
        const unsigned hi = __popc(i & 0x6DB6DB64 & ~0x3) & 1;  // placeholder, fixed below
        (void)hi;
        const unsigned m_lo = 0x6DB6DB6Cu;  // bits {2,3,5,6,8,9,...}: b%3 in {2,0}, b>=2
        const unsigned m_hi = 0x36DB6DB4u;  // bits {2,4,5,7,8,10,...}: b%3 in {2,1}, b>=2
        (void)lo;
        return i ^ ((__popc((unsigned)i & m_lo) & 1) | ((__popc((unsigned)i & m_hi) & 1) << 1));
    }
    return i ^ (((i >> 3) & 1) * 7);
}

__device__ __forceinline__ int phys_row_rt(int i, int log_ct) {
    switch (log_ct) {
        case 0: return phys_row<0>(i);
        case 1: return phys_row<1>(i);
        case 2: return phys_row<2>(i);
        default: return i;
    }
}

// expand a compact group index into a row index with zeros at the (ascending) fixed positions
__device__ __forceinline__ int expand_fixed(int g, const DevOp& op, int nfix, const int* fix) {
    int idx = g;
    for (int f = 0; f < nfix; ++f) idx = insert_zero(idx, fix[f]);
    return idx;
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

struct OpLocal {  // per-op values every thread needs, loaded once per op
    int dim, target, nfix, nq;
    unsigned ctrl_mask;
    int fix[6];
    int q[5];
};

__device__ __forceinline__ void load_op(const DevOp& op, OpLocal& o) {
    o.dim = op.dim;
    o.target = op.target;
    o.ctrl_mask = op.ctrl_mask;
    o.nq = op.nq;
    // fixed positions = target(s) and control bits, ascending
    unsigned m = op.ctrl_mask;
    if (op.dim == 2) m |= 1u << op.target;
    else
        for (int j = 0; j < op.nq; ++j) m |= 1u << op.q[j];
    o.nfix = 0;
    while (m) {
        const int p = __ffs(m) - 1;
        o.fix[o.nfix++] = p;
        m &= m - 1;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) o.q[j] = op.q[j];
}

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
    cplx* sk = (MODE == MODE_GRAD) ? sb + (size_t)rows * CT : sb;         // dense kernel staging
    cplx* swarp = sk + (A.has_dense ? DENSE_STAGE : 0);                   // [2][nwarps][wmax]
    cplx* swacc = swarp + ((MODE == MODE_GRAD) ? 2 * nwarps * A.wmax : 0);  // [w_total] if w_in_smem
    double* sred = reinterpret_cast<double*>(swacc + ((MODE == MODE_GRAD && A.w_in_smem) ? A.w_total : 0));  // [nwarps][6]

    const int chunk = blockIdx.x;
    const int nchunks = gridDim.x;
    const int deriv_op = (MODE == MODE_APPLY && A.deriv_op) ? A.deriv_op[y] : -1;
    const int deriv_p = (MODE == MODE_APPLY && A.deriv_pidx) ? A.deriv_pidx[y] : 0;

    if (MODE == MODE_GRAD && A.w_in_smem) {
        for (int e = tid; e < A.w_total; e += nthr) swacc[e] = czero();
    }
    double tsum[6] = {0, 0, 0, 0, 0, 0};  // running trace sums of this CTA (thread 0)

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
        }
        __syncthreads();

        // ---- forward sweep: gates[0] first (Gates_block.cpp:683) -----------------------------------------------
        for (int k = 0; k < A.n_ops; ++k) {
            const DevOp& op = A.ops[k];
            OpLocal o;
            load_op(op, o);
            const bool deriv = (MODE == MODE_APPLY) && (k == deriv_op);
            const cplx* __restrict__ K =
                deriv ? dktab + op.dkern_off + deriv_p * o.dim * o.dim
                      : (op.kern_off >= 0 ? ktab + op.kern_off : A.pool + op.pool_off);
            if (o.dim == 2) {
                const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
                const int tbit = 1 << o.target;
                if (!deriv) {
                    const int nitems = (rows >> o.nfix) << LOG_CT;
                    for (int item = tid; item < nitems; item += nthr) {
                        const int c = item & (CT - 1);
                        const int i0 = expand_fixed(item >> LOG_CT, op, o.nfix, o.fix) | o.ctrl_mask;
                        const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                        const cplx a0 = sa[e0], a1 = sa[e1];
                        sa[e0] = cfma(k01, a1, cmul(k00, a0));
                        sa[e1] = cfma(k11, a1, cmul(k10, a0));
                    }
                } else {  // derivative kernel: inactive pairs are zero-filled (apply_kernel_to_input.cpp:93-97)
                    const int nitems = (rows >> 1) << LOG_CT;
                    for (int item = tid; item < nitems; item += nthr) {
                        const int c = item & (CT - 1);
                        const int i0 = insert_zero(item >> LOG_CT, o.target);
                        const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                        if ((i0 & o.ctrl_mask) == o.ctrl_mask) {
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
                const int dim = o.dim;
                for (int e = tid; e < dim * dim; e += nthr) sk[e] = K[e];
                __syncthreads();
                unsigned qmask = 0;
                for (int j = 0; j < o.nq; ++j) qmask |= 1u << o.q[j];
                const int ngroups_all = rows >> o.nq;  // groups over non-target bits (control handled by predicate)
                const int nitems = ngroups_all << LOG_CT;
                for (int item = tid; item < nitems; item += nthr) {
                    const int c = item & (CT - 1);
                    int base = item >> LOG_CT;
                    for (int j = 0; j < o.nq; ++j) base = insert_zero(base, o.q[j]);
                    const bool active = (base & o.ctrl_mask) == o.ctrl_mask;
                    if (!active && !deriv) continue;
                    cplx v[32];
                    for (int l = 0; l < dim; ++l) {
                        int r = base;
                        for (int j = 0; j < o.nq; ++j) r |= ((l >> j) & 1) << o.q[j];
                        v[l] = active ? sa[phys_row<LOG_CT>(r) * CT + c] : czero();
                    }
                    for (int ro = 0; ro < dim; ++ro) {
                        cplx acc = czero();
                        if (active)
                            for (int l = 0; l < dim; ++l) acc = cfma(sk[ro * dim + l], v[l], acc);
                        int r = base;
                        for (int j = 0; j < o.nq; ++j) r |= ((ro >> j) & 1) << o.q[j];
                        sa[phys_row<LOG_CT>(r) * CT + c] = acc;
                    }
                }
            }
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
                for (int s = 16; s > 0; s >>= 1) t[i] += __shfl_xor_sync(0xffffffffu, t[i], s);
            if (lane == 0)
                for (int i = 0; i < nt; ++i) sred[warp * 6 + i] = t[i];
            __syncthreads();
            if (tid == 0) {
                for (int i = 0; i < nt; ++i) {
                    double s = 0;
                    for (int w = 0; w < nwarps; ++w) s += sred[w * 6 + i];
                    tsum[i] += s;
                }
            }
        }

        if (MODE == MODE_GRAD) {
            // ---- beta_N = sum_t omega_t * sum_{masks of type t} e_{(j+off)^mask} per column ---------------------
            for (int e = tid; e < rows * CT; e += nthr) sb[e] = czero();
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
                const DevOp& op = A.ops[k];
                OpLocal o;
                load_op(op, o);
                const cplx* __restrict__ K = op.kern_off >= 0 ? ktab + op.kern_off : A.pool + op.pool_off;
                const bool has_w = op.w_off >= 0;
                if (o.dim == 2) {
                    const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
                    const int tbit = 1 << o.target;
                    cplx w00 = czero(), w01 = czero(), w10 = czero(), w11 = czero();
                    const int nitems = (rows >> o.nfix) << LOG_CT;
                    for (int item = tid; item < nitems; item += nthr) {
                        const int c = item & (CT - 1);
                        const int i0 = expand_fixed(item >> LOG_CT, op, o.nfix, o.fix) | o.ctrl_mask;
                        const int e0 = phys_row<LOG_CT>(i0) * CT + c, e1 = phys_row<LOG_CT>(i0 | tbit) * CT + c;
                        const cplx p0 = sa[e0], p1 = sa[e1];  // column after the gate
                        const cplx b0 = sb[e0], b1 = sb[e1];  // row functional after the gate
                        // a_k = K^dagger a_{k+1}
                        const cplx a0 = cfmac(k10, p1, cfmac(k00, p0, czero()));
                        const cplx a1 = cfmac(k11, p1, cfmac(k01, p0, czero()));
                        sa[e0] = a0;
                        sa[e1] = a1;
                        if (has_w) {
                            w00 = cfma(b0, a0, w00);
                            w01 = cfma(b0, a1, w01);
                            w10 = cfma(b1, a0, w10);
                            w11 = cfma(b1, a1, w11);
                        }
                        // beta_{k-1} = K^T beta_k
                        sb[e0] = cfma(k10, b1, cmul(k00, b0));
                        sb[e1] = cfma(k11, b1, cmul(k01, b0));
                    }
                    if (has_w) {
                        double v[8] = {w00.x, w00.y, w01.x, w01.y, w10.x, w10.y, w11.x, w11.y};
                        warp_reduce8(v, lane);
                        if ((lane & 3) == 0) {
                            const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                            reinterpret_cast<double*>(swarp + (size_t)(buf * nwarps + warp) * A.wmax)[idx] = v[0];
                        }
                    }
                } else {
                    const int dim = o.dim;
                    for (int e = tid; e < dim * dim; e += nthr) sk[e] = K[e];
                    __syncthreads();
                    const int nitems = (rows >> o.nq) << LOG_CT;
                    // dense parametric ops are 4 x 4 (RXX/RYY/RZZ): 16 complex accumulators
                    cplx wl[16];
                    if (has_w)
                        for (int e = 0; e < 16; ++e) wl[e] = czero();
                    for (int item = tid; item < nitems; item += nthr) {
                        const int c = item & (CT - 1);
                        int base = item >> LOG_CT;
                        for (int j = 0; j < o.nq; ++j) base = insert_zero(base, o.q[j]);
                        if ((base & o.ctrl_mask) != o.ctrl_mask) continue;
                        cplx pv[32], bv[32];
                        int addr[32];
                        for (int l = 0; l < dim; ++l) {
                            int r = base;
                            for (int j = 0; j < o.nq; ++j) r |= ((l >> j) & 1) << o.q[j];
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
                            if (has_w && dim == 4) {
#pragma unroll
                                for (int r2 = 0; r2 < 4; ++r2) wl[r2 * 4 + ro] = cfma(bv[r2], acc, wl[r2 * 4 + ro]);
                            }
                        }
                    }
                    if (has_w) {
                        for (int part = 0; part < 4; ++part) {
                            double v[8];
                            for (int e = 0; e < 4; ++e) {
                                v[2 * e] = wl[part * 4 + e].x;
                                v[2 * e + 1] = wl[part * 4 + e].y;
                            }
                            warp_reduce8(v, lane);
                            if ((lane & 3) == 0) {
                                const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                                reinterpret_cast<double*>(swarp + (size_t)(buf * nwarps + warp) * A.wmax)[part * 8 + idx] = v[0];
                            }
                        }
                    }
                }
                __syncthreads();
                if (has_w) {
                    const int nd = 2 * o.dim * o.dim;  // doubles
                    if (tid < nd) {
                        double s = 0;
                        for (int w = 0; w < nwarps; ++w)
                            s += reinterpret_cast<const double*>(swarp + (size_t)(buf * nwarps + w) * A.wmax)[tid];
                        if (A.w_in_smem) reinterpret_cast<double*>(swacc + op.w_off)[tid] += s;
                        else
                            reinterpret_cast<double*>(A.w_part + ((size_t)y * nchunks + chunk) * A.w_total + op.w_off)[tid] = s;
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
