// reduce.cuh -- deterministic reduction of the executor's per-CTA partials and the cost / gradient formulas.
//
//   reduce_partials   : tr_part[y][chunk][6], w_part[y][chunk][w_total]  ->  traces[y][1+P][3][2]
//                       k = 0: the three trace terms of C(theta) U; k = 1+p: dL_p = sum_{r,c} dK_p[r][c] W[r][c], the
//                       derivative of the functional L = sum_t omega_t T_t (stored in slot t = 0). The executor stores
//                       W' = sum beta p^T (p = column after the op), W = W' conj(K), hence
//                       dL_p = sum_{r,r'} W'[r][r'] (dK_p K^dagger)[r][r'].
//   cost_from_traces  : calculate_cost_function (decomposition/Optimization_Interface.cpp:677-735) and the gradient
//                       component formulas (Optimization_Interface.cpp:1397-1458) on (possibly rank-summed) traces.
//   make_omega        : weights of the second pass for the Hilbert-Schmidt-with-corrections variants.
#pragma once
#include "sq_types.cuh"
#include "../../include/sqgpu.h"

namespace sq {

// traces layout helpers
__host__ __device__ __forceinline__ size_t tr_index(int y, int k, int n_k, int t) { return (((size_t)y * n_k + k) * 3 + t) * 2; }

// w_part[y][0][e] <- sum over chunks of w_part[y][chunk][e] in a fixed order: one coalesced pass over the partials instead of
// one strided gather per parameter in reduce_partials. A block folds 32 consecutive elements; its 8 warps each sum every 8th
// chunk (512 B per warp load), then warp 0 adds the 8 group sums in ascending order -- a single parameter set (BFGS: batch 1
// with 512 chunks) keeps w_total / 32 blocks busy instead of w_total / 256 threads walking 512 chunks each.
__global__ void __launch_bounds__(256) fold_w_chunks(cplx* __restrict__ w_part, int nchunks, int w_total) {
    __shared__ cplx sgrp[8][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    cplx acc = czero();
    if (e < w_total) {
        const cplx* base = w_part + (size_t)blockIdx.y * nchunks * w_total + e;
        for (int ch = grp; ch < nchunks; ch += 8) acc = cadd(acc, base[(size_t)ch * w_total]);
    }
    sgrp[grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && e < w_total) {
        cplx t = sgrp[0][lane];
#pragma unroll
        for (int g = 1; g < 8; ++g) t = cadd(t, sgrp[g][lane]);
        w_part[(size_t)blockIdx.y * nchunks * w_total + e] = t;
    }
}
static inline unsigned fold_grid_x(int w_total) { return (unsigned)((w_total + 31) / 32); }

// traces[y][0] <- sum of the trace partials; traces[y][1 + p] <- sum_{r, r2} (dK_p K^dagger)[r][r2] W'[r][r2] per parameter.
// grid = (parameter sets, parameter blocks); ONE WARP per parameter: its lanes take the dim^2 entries (coalesced reads of the
// derivative kernel, the kernel and W'), then a shuffle reduction in a fixed order. (Round 1 ran one thread per parameter in
// one block per set: 0.7 ms for the 1290 parameters of C3 at batch 1 -- 40 % of a single evaluation's latency.)
// w_folded: the W partials of chunk 0 already hold the sum over chunks (fold_w_chunks)
__global__ void reduce_partials(const double* __restrict__ tr_part, int nchunks, const cplx* __restrict__ w_part,
                                int w_total, const DevOp* __restrict__ ops, const int* __restrict__ param_op,
                                const int* __restrict__ param_slot, const cplx* __restrict__ dktab, int dkern_total,
                                const cplx* __restrict__ ktab, int kern_total, int n_params, int with_grad,
                                double* __restrict__ traces, int w_folded = 0, int w_slices = 0) {
    if (w_slices <= 0) w_slices = nchunks;  // slices of w_part per parameter set (chunks x warps with per-warp slices)
    const int y = blockIdx.x;
    const int n_k = 1 + (with_grad ? n_params : 0);
    if (blockIdx.y == 0 && threadIdx.x < 6) {
        double s = 0;
        for (int ch = 0; ch < nchunks; ++ch) s += tr_part[((size_t)y * nchunks + ch) * 6 + threadIdx.x];
        traces[tr_index(y, 0, n_k, 0) + threadIdx.x] = s;
    }
    if (!with_grad) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int wch = w_folded ? 1 : w_slices;
    for (int p = blockIdx.y * wpb + warp; p < n_params; p += gridDim.y * wpb) {
        const DevOp op = ops[param_op[p]];
        const int dim = op.dim, d2 = dim * dim;
        const cplx* dk = dktab + (size_t)y * dkern_total + op.dkern_off + param_slot[p] * d2;
        const cplx* kk = ktab + (size_t)y * kern_total + op.kern_off;  // parametric ops always have a table kernel
        cplx acc = czero();
        for (int e = lane; e < d2; e += 32) {
            const int r = e / dim, r2 = e - r * dim;
            cplx w = czero();
            for (int ch = 0; ch < wch; ++ch) w = cadd(w, w_part[((size_t)y * w_slices + ch) * w_total + op.w_off + e]);
            cplx dkk = czero();  // (dK K^dagger)[r][r2] = sum_c dK[r][c] conj(K[r2][c])
            for (int c = 0; c < dim; ++c) dkk = cfmac(kk[r2 * dim + c], dk[r * dim + c], dkk);
            acc = cfma(dkk, w, acc);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, sft);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, sft);
        }
        if (lane == 0) {
            double* dst = traces + tr_index(y, 1 + p, n_k, 0);
            dst[0] = acc.x;
            dst[1] = acc.y;
            dst[2] = dst[3] = dst[4] = dst[5] = 0.0;
        }
    }
}
// parameter blocks per set: four warps (= four parameters at a time) per block, at most 64 blocks
static inline unsigned reduce_grid_y(int n_params) { return (unsigned)std::max(1, std::min(64, (n_params + 3) / 4)); }

struct CostCfg {
    int variant;
    double prev, c1, c2;
};

__device__ __forceinline__ double cost_formula(const CostCfg& c, const double* t, double n) {
    const double sp = sqrt(c.prev);
    switch (c.variant) {
        case SQGPU_FROBENIUS_NORM: return 1.0 - t[0] / n;
        case SQGPU_FROBENIUS_NORM_CORRECTION1: return (1.0 - t[0] / n) - sp * (t[2] / n) * c.c1;
        case SQGPU_FROBENIUS_NORM_CORRECTION2: return (1.0 - t[0] / n) - sp * ((t[2] / n) * c.c1 + (t[4] / n) * c.c2);
        case SQGPU_HILBERT_SCHMIDT_TEST: { const double d = 1.0 / n; return 1.0 - d * d * (t[0] * t[0] + t[1] * t[1]); }
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1: {
            const double d = 1.0 / n;
            return 1 - d * d * (t[0] * t[0] + t[1] * t[1] + sp * c.c1 * (t[2] * t[2] + t[3] * t[3]));
        }
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2: {
            const double d = 1.0 / n;
            return 1 - d * d * (t[0] * t[0] + t[1] * t[1] + sp * (c.c1 * (t[2] * t[2] + t[3] * t[3]) + c.c2 * (t[4] * t[4] + t[5] * t[5])));
        }
        case SQGPU_INFIDELITY: return 1.0 - ((t[0] * t[0] + t[1] * t[1]) / n + 1) / (n + 1);
        case SQGPU_SUM_OF_SQUARES: return t[0];  // slot 0 carries sum |M_ij - delta_ij|^2 (Cost_Function.cpp:443-457)
        default: return nan("");
    }
}

// dl = {Re, Im} of dL_p with the omega of `make_omega` / the variant defaults
__device__ __forceinline__ double grad_formula(const CostCfg& c, const double* t, const double* dl, double n) {
    switch (c.variant) {
        case SQGPU_FROBENIUS_NORM:
        case SQGPU_FROBENIUS_NORM_CORRECTION1:
        case SQGPU_FROBENIUS_NORM_CORRECTION2: return (1.0 - dl[0] / n) - 1.0;
        case SQGPU_HILBERT_SCHMIDT_TEST: {
            const double d = 1.0 / n;
            return -2.0 * d * d * t[0] * dl[0] - 2.0 * d * d * t[1] * dl[1];
        }
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1:
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2: {
            const double d = 1.0 / n;
            return -2.0 * d * d * dl[0];
        }
        case SQGPU_INFIDELITY: return -2.0 / n / (n + 1) * t[0] * dl[0] - 2.0 / n / (n + 1) * t[1] * dl[1];
        case SQGPU_SUM_OF_SQUARES: return dl[0];  // real_trace_conj_dot(Upartial, dM_p), Optimization_Interface.cpp:1432-1434
        default: return nan("");
    }
}

// cost at theta + shift e_p from the traces of theta and dl = L(theta + shift e_p) - L(theta) (sqgpu_cost_shifted_batched): the
// Frobenius variants are affine in L, the |trace|^2 variants take the shifted complex trace
__device__ __forceinline__ double shifted_formula(const CostCfg& c, const double* t, const double* dl, double n) {
    switch (c.variant) {
        case SQGPU_FROBENIUS_NORM:
        case SQGPU_FROBENIUS_NORM_CORRECTION1:
        case SQGPU_FROBENIUS_NORM_CORRECTION2: return cost_formula(c, t, n) - dl[0] / n;
        case SQGPU_HILBERT_SCHMIDT_TEST:
        case SQGPU_INFIDELITY: {
            const double ts[6] = {t[0] + dl[0], t[1] + dl[1], 0, 0, 0, 0};
            return cost_formula(c, ts, n);
        }
        default: return nan("");
    }
}

__global__ void cost_from_traces(const double* __restrict__ traces, int n_params, int with_grad, int cols_total,
                                 CostCfg cfg, double* __restrict__ cost, double* __restrict__ grad) {
    const int y = blockIdx.x;
    const int n_k = 1 + (with_grad ? n_params : 0);
    const double* t = traces + tr_index(y, 0, n_k, 0);
    const double n = (double)cols_total;
    if (threadIdx.x == 0 && cost) cost[y] = cost_formula(cfg, t, n);
    if (!with_grad || !grad) return;
    for (int p = threadIdx.x; p < n_params; p += blockDim.x)
        grad[(size_t)y * n_params + p] = with_grad == 2 ? shifted_formula(cfg, t, traces + tr_index(y, 1 + p, n_k, 0), n)
                                                        : grad_formula(cfg, t, traces + tr_index(y, 1 + p, n_k, 0), n);
}

// omega[y][3]: weights of the trace types in L. Variant defaults need no data; the HS-with-corrections variants take
// w_t * conj(T_t) from the traces of a cost-only pre-pass (Optimization_Interface.cpp:1414-1434).
__global__ void make_omega(const double* __restrict__ traces, int n_k, CostCfg cfg, int batch, cplx* __restrict__ omega) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= batch) return;
    const double sp = sqrt(cfg.prev);
    cplx w0 = cmake(1.0, 0.0), w1 = czero(), w2 = czero();
    switch (cfg.variant) {
        case SQGPU_FROBENIUS_NORM_CORRECTION1: w1 = cmake(sp * cfg.c1, 0); break;
        case SQGPU_FROBENIUS_NORM_CORRECTION2: w1 = cmake(sp * cfg.c1, 0); w2 = cmake(sp * cfg.c2, 0); break;
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1:
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2: {
            const double* t = traces + tr_index(y, 0, n_k, 0);
            w0 = cmake(t[0], -t[1]);
            w1 = cmake(sp * cfg.c1 * t[2], -sp * cfg.c1 * t[3]);
            if (cfg.variant == SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2) w2 = cmake(sp * cfg.c2 * t[4], -sp * cfg.c2 * t[5]);
            break;
        }
        default: break;
    }
    omega[(size_t)y * 3 + 0] = w0;
    omega[(size_t)y * 3 + 1] = w1;
    omega[(size_t)y * 3 + 2] = w2;
}

}  // namespace sq
