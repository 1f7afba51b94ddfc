// vqe.cuh -- kernels of the state-vector (VQE) cost path and the streaming adjoint step.
//
//   csr_matvec_batched  <- mult(Matrix_sparse, Matrix&)                     (common/common.cpp:403-436)
//   expectation_batched <- Expectation_value_of_energy_real                 (variational_quantum_eigensolver/
//                                                                            Variational_Quantum_Eigensolver_Base.cpp:584-624)
//   adjoint1q_stream    :  one backward step of the adjoint gradient on matrices / state vectors in HBM. The reference
//                          materialises P derivative states (…Base.cpp:1131-1199 -> Gates_block::apply_derivate_to);
//                          here grad_p = 2 Re <d_p psi | H psi> = 2 Re sum_{r,c} dK_p[r][c] W[r][c] with
//                          W[r][c] = sum_pairs beta[r] a[c], beta_N = conj(H psi_N), beta_{k-1} = K^T beta_k,
//                          a_k = K^dagger a_{k+1}  -- the same recurrences as the unitary path (exec_fused.cuh).
#pragma once
#include "exec_stream.cuh"
#include "sq_types.cuh"

namespace sq {

// y[b][r] = sum_e values[e] * x[b][indices[e]]. G lanes share a row (G = 8 for short rows: a 20-qubit Heisenberg row holds
// 16 non-zeros on average, a whole warp per row would idle half its lanes), and every group applies the row to YB parameter
// sets at once so that indices / values are read once per YB states.
template <int G, int YB>
__global__ void csr_matvec_batched(int n_rows, int n_sets, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   const cplx* __restrict__ values, const cplx* __restrict__ x, cplx* __restrict__ yv,
                                   int conj_out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x & (G - 1);
    if (row >= n_rows) return;  // whole groups leave together (blockDim.x is a multiple of G, shuffles below are per group)
    const int y0 = blockIdx.y * YB;
    cplx acc[YB];
#pragma unroll
    for (int b = 0; b < YB; ++b) acc[b] = czero();
    const int e1 = indptr[row + 1];
    for (int e = indptr[row] + lane; e < e1; e += G) {
        const cplx v = values[e];
        const size_t col = (size_t)indices[e];
#pragma unroll
        for (int b = 0; b < YB; ++b)
            if (y0 + b < n_sets) acc[b] = cfma(v, x[(size_t)(y0 + b) * n_rows + col], acc[b]);
    }
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
#pragma unroll
    for (int b = 0; b < YB; ++b) {
        for (int s = G / 2; s > 0; s >>= 1) {
            acc[b].x += __shfl_xor_sync(gmask, acc[b].x, s);
            acc[b].y += __shfl_xor_sync(gmask, acc[b].y, s);
        }
        if (lane == 0 && y0 + b < n_sets) yv[(size_t)(y0 + b) * n_rows + row] = conj_out ? cmake(acc[b].x, -acc[b].y) : acc[b];
    }
}

// host launcher: group width from the average row length
inline void launch_csr_matvec(int n_rows, long long nnz, int n_sets, const int32_t* indptr, const int32_t* indices, const cplx* values,
                              const cplx* x, cplx* yv, int conj_out, cudaStream_t st) {
    constexpr int YB = 4;
    const long long avg = n_rows > 0 ? nnz / n_rows : 0;
    if (avg <= 24) {
        dim3 grid((unsigned)(((long long)n_rows * 8 + 255) / 256), (n_sets + YB - 1) / YB);
        csr_matvec_batched<8, YB><<<grid, 256, 0, st>>>(n_rows, n_sets, indptr, indices, values, x, yv, conj_out);
    } else {
        dim3 grid((unsigned)(((long long)n_rows * 32 + 255) / 256), (n_sets + YB - 1) / YB);
        csr_matvec_batched<32, YB><<<grid, 256, 0, st>>>(n_rows, n_sets, indptr, indices, values, x, yv, conj_out);
    }
}

// part[b][blockIdx.x] = sum_i Re(conj(left_i) * right_i)  (right may be stored conjugated: sign_im = -1)
__global__ void expectation_partial(int n_rows, const cplx* __restrict__ left, const cplx* __restrict__ right,
                                    double sign_im, double* __restrict__ part) {
    __shared__ double sred[32];
    const cplx* __restrict__ l = left + (size_t)blockIdx.y * n_rows;
    const cplx* __restrict__ r = right + (size_t)blockIdx.y * n_rows;
    double e = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (long long)gridDim.x * blockDim.x)
        e += l[i].x * r[i].x + sign_im * l[i].y * r[i].y;
    for (int s = 16; s > 0; s >>= 1) e += __shfl_xor_sync(0xffffffffu, e, s);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
        part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

__global__ void sum_partials(const double* __restrict__ part, int nparts, int width, double scale, double* __restrict__ out,
                             int out_stride) {
    // out[y*out_stride + w] = scale * sum_p part[(y*nparts + p)*width + w]
    const int y = blockIdx.x, w = threadIdx.x;
    if (w >= width) return;
    double s = 0;
    for (int p = 0; p < nparts; ++p) s += part[((size_t)y * nparts + p) * width + w];
    out[(size_t)y * out_stride + w] = scale * s;
}

// backward step for a 1-qubit (optionally controlled) gate: a <- K^dagger a, W += beta a^T, beta <- K^T beta.
// wpart[y][blockIdx.x][8] receives this block's W contribution.
__global__ void __launch_bounds__(256) adjoint1q_stream(const StreamGate G, cplx* __restrict__ beta, long long beta_ystride,
                                                        double* __restrict__ wpart, int want_w) {
    __shared__ double sred[8 * 8];
    const cplx* __restrict__ K = G.K + (size_t)blockIdx.y * G.k_ystride;
    const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
    cplx* __restrict__ d = G.data + (size_t)blockIdx.y * G.ystride;
    cplx* __restrict__ bt = beta + (size_t)blockIdx.y * beta_ystride;
    const long long nitems = (long long)(G.rows >> G.nfix) * G.cols;
    const int tbit = 1 << G.target;
    cplx w00 = czero(), w01 = czero(), w10 = czero(), w11 = czero();
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
        int i0 = g;
        for (int f = 0; f < G.nfix; ++f) i0 = insert_zero(i0, G.fix[f]);
        i0 |= G.ctrl_mask;
        const size_t o0 = (size_t)i0 * G.ld + j, o1 = (size_t)(i0 | tbit) * G.ld + j;
        const cplx p0 = d[o0], p1 = d[o1], b0 = bt[o0], b1 = bt[o1];
        const cplx a0 = cfmac(k10, p1, cfmac(k00, p0, czero()));
        const cplx a1 = cfmac(k11, p1, cfmac(k01, p0, czero()));
        d[o0] = a0;
        d[o1] = a1;
        if (want_w) {
            w00 = cfma(b0, p0, w00);  // W' = beta p^T (column after the gate), see reduce_partials
            w01 = cfma(b0, p1, w01);
            w10 = cfma(b1, p0, w10);
            w11 = cfma(b1, p1, w11);
        }
        bt[o0] = cfma(k10, b1, cmul(k00, b0));
        bt[o1] = cfma(k11, b1, cmul(k01, b0));
    }
    if (!want_w) return;
    double v[8] = {w00.x, w00.y, w01.x, w01.y, w10.x, w10.y, w11.x, w11.y};
    for (int i = 0; i < 8; ++i)
        for (int s = 16; s > 0; s >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int i = 0; i < 8; ++i) sred[warp * 8 + i] = v[i];
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w * 8 + threadIdx.x];
        wpart[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = s;
    }
}

// backward step for a dense 4 x 4 block on qubits q0 < q1 (no controls): same recurrences, 16 W entries.
// wpart[y][blockIdx.x][32] receives this block's W contribution.
__global__ void __launch_bounds__(128) adjoint2q_stream(const StreamGate G, cplx* __restrict__ beta, long long beta_ystride,
                                                        double* __restrict__ wpart, int want_w) {
    __shared__ double sred[4 * 32];
    const cplx* __restrict__ K = G.K + (size_t)blockIdx.y * G.k_ystride;
    cplx M[16], W[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        M[e] = K[e];
        W[e] = czero();
    }
    cplx* __restrict__ d = G.data + (size_t)blockIdx.y * G.ystride;
    cplx* __restrict__ bt = beta + (size_t)blockIdx.y * beta_ystride;
    const long long nitems = (long long)(G.rows >> 2) * G.cols;
    const int q0 = G.q[0], q1 = G.q[1];
    const size_t s0 = (size_t)(1 << q0) * G.ld, s1 = (size_t)(1 << q1) * G.ld;
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
        const int base = insert_zero(insert_zero(g, q0), q1);
        const size_t o0 = (size_t)base * G.ld + j;
        const size_t o[4] = {o0, o0 + s0, o0 + s1, o0 + s0 + s1};
        cplx p[4], b[4], a[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            p[l] = d[o[l]];
            b[l] = bt[o[l]];
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            a[cc] = cfmac(M[12 + cc], p[3], cfmac(M[8 + cc], p[2], cfmac(M[4 + cc], p[1], cfmac(M[cc], p[0], czero()))));
#pragma unroll
        for (int l = 0; l < 4; ++l) d[o[l]] = a[l];
        if (want_w) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) W[r * 4 + cc] = cfma(b[r], p[cc], W[r * 4 + cc]);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            bt[o[cc]] = cfma(M[12 + cc], b[3], cfma(M[8 + cc], b[2], cfma(M[4 + cc], b[1], cmul(M[cc], b[0]))));
    }
    if (!want_w) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        double re = W[e].x, im = W[e].y;
        for (int s = 16; s > 0; s >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, s);
            im += __shfl_xor_sync(0xffffffffu, im, s);
        }
        if (lane == 0) {
            sred[warp * 32 + 2 * e] = re;
            sred[warp * 32 + 2 * e + 1] = im;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w * 32 + threadIdx.x];
        wpart[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 32 + threadIdx.x] = s;
    }
}

// grad[y][p] = scale * Re(dL_p) from the traces buffer reduce_partials wrote
__global__ void grad_from_traces(const double* __restrict__ traces, int n_params, double scale, double* __restrict__ grad) {
    const int y = blockIdx.x;
    for (int p = threadIdx.x; p < n_params; p += blockDim.x)
        grad[(size_t)y * n_params + p] = scale * traces[(((size_t)y * (1 + n_params) + 1 + p) * 3) * 2];
}

}  // namespace sq
