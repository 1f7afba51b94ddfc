// exec_stream.cuh -- one-gate-per-launch kernels over matrices / state vectors that live in HBM (or L2).
//
// These are the direct counterparts of the reference's per-gate CPU kernels and are HBM-bandwidth bound:
//   gate1q_stream  <- apply_kernel_to_input[_AVX[_parallel]]          (gates/kernels/apply_kernel_to_input.cpp:33-115,
//                     apply_kernel_to_input_AVX.cpp:665-1010), state-vector twin
//                     (apply_kernel_to_state_vector_input.cpp:33-229), and the dedicated CNOT/CZ/CH/CCX row kernels
//                     (apply_dedicated_gate_kernel_to_input.cpp:45-697) which are the same update with a 0/1 kernel
//   gatekq_stream  <- apply_nqbit_kernel_to_matrix_input_impl / apply_{2..5}qbit_kernel_to_state_vector_input
//                     (apply_large_kernel_to_input.cpp:123-213, apply_kernel_to_state_vector_input.cpp:244-617)
//   traces_stream  <- get_cost_function*, get_trace* (decomposition/N_Qubit_Decomposition_Cost_Function.cpp:73-664)
// Algorithmic traffic per launch: 32 B per touched amplitude (16 B read + 16 B write); a gate with c control bits
// touches rows * cols / 2^c amplitudes.
//
// They serve sqgpu_apply_gate (single-gate API / microbenchmarks) and every problem too large for the shared-memory
// executor (exec_fused.cuh).
#pragma once
#include "sq_types.cuh"

namespace sq {

struct StreamGate {
    cplx* data;
    long long ystride;  // elements per blockIdx.y
    int rows, cols, ld;
    int log_cols;       // log2(cols) if cols is a power of two, else -1
    int target;
    unsigned ctrl_mask;
    int nfix;
    int fix[6];
    const cplx* K;      // kernel (dim*dim complex)
    long long k_ystride;
    int nq;
    int q[5];
};

// every thread: UNROLL (row pair, column) items per grid-stride step, all loads issued before the first FMA so that
// 2 * UNROLL independent 16 B requests are in flight per thread. Consecutive threads walk along a row -> 16 B x 32
// coalesced accesses.
template <bool DERIV>
__global__ void __launch_bounds__(256) gate1q_stream(const StreamGate G) {
    constexpr int UNROLL = 4;
    const cplx* __restrict__ K = G.K + (size_t)blockIdx.y * G.k_ystride;
    const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
    cplx* __restrict__ d = G.data + (size_t)blockIdx.y * G.ystride;
    const int nfix = DERIV ? 1 : G.nfix;
    const long long nitems = (long long)(G.rows >> nfix) * G.cols;
    const size_t tstride = (size_t)(1 << G.target) * G.ld;
    const int f0 = DERIV ? G.target : G.fix[0], f1 = DERIV ? 30 : G.fix[1], f2 = DERIV ? 30 : G.fix[2];
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long item0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; item0 < nitems; item0 += step * UNROLL) {
        size_t off[UNROLL];
        bool ok[UNROLL], act[UNROLL];
        cplx a0[UNROLL], a1[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long item = item0 + u * step;
            ok[u] = item < nitems;
            int g = 0, j = 0;
            if (ok[u]) {
                if (G.log_cols >= 0) {
                    g = (int)(item >> G.log_cols);
                    j = (int)(item & (G.cols - 1));
                } else {
                    g = (int)(item / G.cols);
                    j = (int)(item - (long long)g * G.cols);
                }
            }
            int i0 = insert_zero(insert_zero(insert_zero(g, f0), f1), f2);
            if (!DERIV) i0 |= G.ctrl_mask;
            act[u] = !DERIV || (i0 & G.ctrl_mask) == G.ctrl_mask;
            off[u] = (size_t)i0 * G.ld + j;
            if (ok[u] && act[u]) {
                a0[u] = d[off[u]];
                a1[u] = d[off[u] + tstride];
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!ok[u]) continue;
            if (act[u]) {
                d[off[u]] = cfma(k01, a1[u], cmul(k00, a0[u]));
                d[off[u] + tstride] = cfma(k11, a1[u], cmul(k10, a0[u]));
            } else {
                d[off[u]] = czero();
                d[off[u] + tstride] = czero();
            }
        }
    }
}

// dense 2^KQ x 2^KQ kernel; the kernel matrix is staged in shared memory, the 2^KQ amplitudes of a group in registers
template <int KQ, bool DERIV>
__global__ void __launch_bounds__(128) gatekq_stream(const StreamGate G) {
    constexpr int DIM = 1 << KQ;
    __shared__ cplx sk[DIM * DIM];
    const cplx* __restrict__ K = G.K + (size_t)blockIdx.y * G.k_ystride;
    for (int e = threadIdx.x; e < DIM * DIM; e += blockDim.x) sk[e] = K[e];
    __syncthreads();
    cplx* __restrict__ d = G.data + (size_t)blockIdx.y * G.ystride;
    const long long nitems = (long long)(G.rows >> KQ) * G.cols;
    int pat[DIM];
#pragma unroll
    for (int l = 0; l < DIM; ++l) {
        int r = 0;
#pragma unroll
        for (int j = 0; j < KQ; ++j) r |= ((l >> j) & 1) << G.q[j];
        pat[l] = r;
    }
    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < nitems;
         item += (long long)gridDim.x * blockDim.x) {
        int g, j;
        if (G.log_cols >= 0) {
            g = (int)(item >> G.log_cols);
            j = (int)(item & (G.cols - 1));
        } else {
            g = (int)(item / G.cols);
            j = (int)(item - (long long)g * G.cols);
        }
        int base = g;
#pragma unroll
        for (int jq = 0; jq < KQ; ++jq) base = insert_zero(base, G.q[jq]);
        const bool active = (base & G.ctrl_mask) == G.ctrl_mask;
        if (!active && !DERIV) continue;
        cplx v[DIM];
#pragma unroll
        for (int l = 0; l < DIM; ++l) v[l] = d[(size_t)(base | pat[l]) * G.ld + j];
#pragma unroll
        for (int ro = 0; ro < DIM; ++ro) {
            cplx acc = czero();
            if (active) {
#pragma unroll
                for (int l = 0; l < DIM; ++l) acc = cfma(sk[ro * DIM + l], v[l], acc);
            }
            d[(size_t)(base | pat[ro]) * G.ld + j] = acc;
        }
    }
}

// traces[y][6] = {Re,Im} of sum_j M[(j+off)^mask, j] for mask classes {0}, {one bit}, {two bits}
__global__ void traces_stream(const cplx* __restrict__ data, long long ystride, int cols, int ld, int n,
                              int trace_offset, int n_trace_types, double* __restrict__ out, int out_stride) {
    __shared__ double sred[32 * 6];
    const cplx* __restrict__ d = data + (size_t)blockIdx.x * ystride;
    double t[6] = {0, 0, 0, 0, 0, 0};
    for (int j = threadIdx.x; j < cols; j += blockDim.x) {
        const int r = j + trace_offset;
        cplx v = d[(size_t)r * ld + j];
        t[0] += v.x;
        t[1] += v.y;
        if (n_trace_types > 1)
            for (int q = 0; q < n; ++q) {
                v = d[(size_t)(r ^ (1 << q)) * ld + j];
                t[2] += v.x;
                t[3] += v.y;
            }
        if (n_trace_types > 2)
            for (int q1 = 0; q1 < n - 1; ++q1)
                for (int q2 = q1 + 1; q2 < n; ++q2) {
                    v = d[(size_t)(r ^ ((1 << q1) | (1 << q2))) * ld + j];
                    t[4] += v.x;
                    t[5] += v.y;
                }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = 0; i < 6; ++i)
        for (int s = 16; s > 0; s >>= 1) t[i] += __shfl_xor_sync(0xffffffffu, t[i], s);
    if (lane == 0)
        for (int i = 0; i < 6; ++i) sred[warp * 6 + i] = t[i];
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0;
        for (int w = 0; w < nwarps; ++w) s += sred[w * 6 + threadIdx.x];
        out[(size_t)blockIdx.x * out_stride + threadIdx.x] = s;
    }
}

// y-batched replication of one matrix: dst[y] = src  (workspace setup of the streaming executor)
__global__ void replicate_matrix(const cplx* __restrict__ src, cplx* __restrict__ dst, long long n_elem) {
    cplx* __restrict__ d = dst + (size_t)blockIdx.y * n_elem;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += (long long)gridDim.x * blockDim.x)
        d[i] = src[i];
}

// dst[y][i][j] = src[i * ld_src + j0 + j]: the column chunk [j0, j0 + cw) of the resident matrix, once per parameter set
__global__ void copy_chunk(const cplx* __restrict__ src, int ld_src, int j0, int rows, int cw, cplx* __restrict__ dst) {
    cplx* __restrict__ d = dst + (size_t)blockIdx.y * rows * cw;
    const long long n = (long long)rows * cw;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / cw), j = (int)(e - (long long)i * cw);
        d[e] = src[(size_t)i * ld_src + j0 + j];
    }
}

// beta_N of the adjoint sweep for a column chunk (the chunk is zero-filled first):
// beta[y][(j0 + j + off) ^ mask][j] = omega[y][t] for every mask of trace type t < n_trace_types
__global__ void beta_init_stream(cplx* __restrict__ beta, int rows, int cw, int j0, int n, int trace_offset,
                                 int n_trace_types, const cplx* __restrict__ omega) {
    cplx* __restrict__ b = beta + (size_t)blockIdx.y * rows * cw;
    const cplx* w = omega + (size_t)blockIdx.y * 3;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cw; j += gridDim.x * blockDim.x) {
        const int r = j0 + j + trace_offset;
        b[(size_t)r * cw + j] = w[0];
        if (n_trace_types > 1)
            for (int q = 0; q < n; ++q) b[(size_t)(r ^ (1 << q)) * cw + j] = w[1];
        if (n_trace_types > 2)
            for (int q1 = 0; q1 < n - 1; ++q1)
                for (int q2 = q1 + 1; q2 < n; ++q2) b[(size_t)(r ^ ((1 << q1) | (1 << q2))) * cw + j] = w[2];
    }
}

}  // namespace sq
