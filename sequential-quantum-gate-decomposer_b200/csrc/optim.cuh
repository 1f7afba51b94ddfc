// optim.cuh -- device-resident optimizer inner loops; included at the end of sqgpu.cu.
//
// The reference's optimizers evaluate one point at a time and come back to the host between evaluations
// (optimization_engines/ADAM.cpp:199-330: optimization_problem_combined, then Adam::update on the host; common/BFGS_Powell.cpp:
// 70-200: one optimization_problem_combined per trial step length of the line search). With the cost path on the GPU that
// host round trip (H2D of the parameters, launch, synchronise, D2H of cost + gradient, update) is what is left per step for
// small circuits. Two entry points remove it:
//
//   sqgpu_adam_init / sqgpu_adam_steps / sqgpu_adam_get   `batch` independent ADAM trajectories; parameters, moments and the
//       optimizer's scalar state stay in HBM, every step is  cost+gradient (the executor)  ->  adam_update  on the device, the
//       whole loop is enqueued without a host synchronisation (one CUDA graph replay per step after the first). The update is
//       Adam::update (common/Adam.cpp:120-262) in its sequential semantics, operation by operation (no FMA contraction), so the
//       trajectory equals the host-driven one bit for bit (oracle/sq_oracle.c: sqo_adam_update is the host mirror the tests use).
//   sqgpu_line_search_batched   k trial step lengths of a line search as ONE batch: theta_j = x + alpha_j d is formed on the
//       device (2 P + k doubles cross the bus instead of k P), cost and the directional derivative g_j . d come back.
#pragma once

namespace sq {

struct AdamCfg {
    double eta, beta1, beta2, epsilon;
};

// scalar state of one trajectory: the members of class Adam (common/include/Adam.h) that Adam::update reads and writes
struct AdamState {
    double beta1_t, beta2_t, f0_mean, decreasing_test, f0_prev, best_cost;
    int f0_idx, decreasing_idx, iter_t, status;
    double f0_vec[100];
    int decreasing_vec[20];
};

__global__ void adam_reset_kernel(AdamState* st, int batch) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= batch) return;
    AdamState& s = st[y];  // Adam::reset, common/Adam.cpp:84-108
    s.beta1_t = 1.0;
    s.beta2_t = 1.0;
    s.f0_mean = 0.0;
    s.decreasing_test = -1.0;
    s.f0_prev = 1.7976931348623157e308;  // DBL_MAX
    s.best_cost = 1.7976931348623157e308;
    s.f0_idx = s.decreasing_idx = s.iter_t = s.status = 0;
    for (int i = 0; i < 100; ++i) s.f0_vec[i] = 0.0;
    for (int i = 0; i < 20; ++i) s.decreasing_vec[i] = -1;
}

// One block per trajectory. scratch[y][2][P]: the running bias-correction products (written by thread 0).
__global__ void adam_update_kernel(double* __restrict__ theta, const double* __restrict__ grad, const double* __restrict__ cost,
                                   double* __restrict__ mom, double* __restrict__ var, AdamState* __restrict__ states,
                                   double* __restrict__ scratch, double* __restrict__ best_theta, double* __restrict__ cost_hist,
                                   int P, AdamCfg cfg) {
    const int y = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    double* th = theta + (size_t)y * P;
    const double* g = grad + (size_t)y * P;
    double* m = mom + (size_t)y * P;
    double* v = var + (size_t)y * P;
    double* b1 = scratch + (size_t)y * 2 * P;
    double* b2 = b1 + P;
    AdamState& s = states[y];
    const double f0 = cost[y];
    __shared__ int s_barren, s_better;
    if (tid == 0) {
        if (cost_hist) cost_hist[y] = f0;
        // ADAM.cpp:219-222: the best point seen so far is the one that produced f0 (before the update)
        s_better = f0 < s.best_cost ? 1 : 0;
        if (s_better) s.best_cost = f0;
        // ---- Adam::update, common/Adam.cpp:139-208: local-minimum statistics, decreasing test, barren-plateau test ----------
        s.f0_mean = __dadd_rn(s.f0_mean, __ddiv_rn(__dsub_rn(f0, s.f0_vec[s.f0_idx]), 100.0));
        s.f0_vec[s.f0_idx] = f0;
        s.f0_idx = (s.f0_idx + 1) % 100;
        double var_f0 = 0.0;
        for (int i = 0; i < 100; ++i) {
            const double d = __dsub_rn(s.f0_vec[i], s.f0_mean);
            var_f0 = __dadd_rn(var_f0, __dmul_rn(d, d));
        }
        var_f0 = __ddiv_rn(__dsqrt_rn(var_f0), 100.0);
        if (f0 < s.f0_prev) {
            if (s.decreasing_vec[s.decreasing_idx] != 1) s.decreasing_test = __dadd_rn(s.decreasing_test, __ddiv_rn(2.0, 20.0));
            s.decreasing_vec[s.decreasing_idx] = 1;
        } else {
            if (s.decreasing_vec[s.decreasing_idx] == 1) s.decreasing_test = __dsub_rn(s.decreasing_test, __ddiv_rn(2.0, 20.0));
            s.decreasing_vec[s.decreasing_idx] = -1;
        }
        s.decreasing_idx = (s.decreasing_idx + 1) % 20;
        s.f0_prev = f0;
        double grad_var = 0.0;
        for (int i = 0; i < P; ++i) grad_var = __dadd_rn(grad_var, v[i]);
        s_barren = (grad_var < cfg.epsilon && s.decreasing_test > 0.7) ? 1 : 0;
        // the bias-correction products advance ONCE PER PARAMETER inside the reference's loop (Adam.cpp:226-229: beta1_t and
        // beta2_t are members updated in the loop body): a sequential scan; once both have underflowed to 0 they stay there
        double t1 = s.beta1_t, t2 = s.beta2_t;
        if (t1 == 0.0 && t2 == 0.0) {
            b1[0] = -1.0;  // marker: every product is 0
        } else {
            for (int i = 0; i < P; ++i) {
                t1 = __dmul_rn(t1, cfg.beta1);
                t2 = __dmul_rn(t2, cfg.beta2);
                b1[i] = t1;
                b2[i] = t2;
            }
            s.beta1_t = t1;
            s.beta2_t = t2;
        }
        s.iter_t += 1;
        s.status = (fabs(__dsub_rn(s.f0_mean, f0)) < 1e-6 && s.decreasing_test <= 0.7 && __ddiv_rn(var_f0, s.f0_mean) < 1e-6) ? 1 : 0;
    }
    __syncthreads();
    const bool all_zero = b1[0] == -1.0;
    const double eps = s_barren ? __ddiv_rn(cfg.epsilon, 100.0) : cfg.epsilon;
    const double omb1 = __dsub_rn(1.0, cfg.beta1), omb2 = __dsub_rn(1.0, cfg.beta2);
    for (int i = tid; i < P; i += nthr) {
        const double gi = g[i], thi = th[i];
        if (s_better) best_theta[(size_t)y * P + i] = thi;
        const double mi = __dadd_rn(__dmul_rn(cfg.beta1, m[i]), __dmul_rn(omb1, gi));                    // Adam.cpp:222
        const double vi = __dadd_rn(__dmul_rn(cfg.beta2, v[i]), __dmul_rn(__dmul_rn(omb2, gi), gi));     // Adam.cpp:223
        m[i] = mi;
        v[i] = vi;
        const double mom_bias_corr = __ddiv_rn(mi, __dsub_rn(1.0, all_zero ? 0.0 : b1[i]));
        const double var_bias_corr = __ddiv_rn(vi, __dsub_rn(1.0, all_zero ? 0.0 : b2[i]));
        th[i] = __dsub_rn(thi, __ddiv_rn(__dmul_rn(cfg.eta, mom_bias_corr), __dadd_rn(__dsqrt_rn(var_bias_corr), eps)));
    }
}

// theta[j][i] = x[i] + alpha[j] * d[i]
__global__ void line_points_kernel(const double* __restrict__ x, const double* __restrict__ d, const double* __restrict__ alpha,
                                   double* __restrict__ theta, int P) {
    const int j = blockIdx.y;
    const double a = alpha[j];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x)
        theta[(size_t)j * P + i] = __dadd_rn(x[i], __dmul_rn(a, d[i]));
}

// dphi[j] = sum_i grad[j][i] * d[i], fixed summation order (one block per point, strided partial sums folded by thread 0)
__global__ void directional_kernel(const double* __restrict__ grad, const double* __restrict__ d, double* __restrict__ dphi, int P) {
    __shared__ double part[256];
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) acc += grad[(size_t)j * P + i] * d[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int t = 0; t < blockDim.x; ++t) sum += part[t];
        dphi[j] = sum;
    }
}

}  // namespace sq

struct AdamRun {
    int batch = 0, P = 0;
    sq::AdamCfg cfg{1e-3, 0.68, 0.8, 1e-4};  // Adam::Adam(), common/Adam.cpp:31-36
    DevBuf theta, mom, var, grad, cost, states, scratch, best_theta, hist;
    cudaGraphExec_t graph = nullptr;  // one captured step (cost+gradient executor launches + update)
    void release() {
        DevBuf* bufs[] = {&theta, &mom, &var, &grad, &cost, &states, &scratch, &best_theta, &hist};
        for (DevBuf* b : bufs) b->release();
        if (graph) cudaGraphExecDestroy(graph);
        graph = nullptr;
    }
};

namespace {

void release_adam(sqgpu_ctx* c) {
    if (!c->adam) return;
    c->adam->release();
    delete c->adam;
    c->adam = nullptr;
}

int adam_enqueue_step(sqgpu_ctx* c, AdamRun* a, double* d_hist_slot, cudaStream_t st) {
    int rc = eval_dev(c, a->theta.as<double>(), a->batch, true, a->cost.as<double>(), a->grad.as<double>(), st);
    if (rc) return rc;
    sq::adam_update_kernel<<<a->batch, 256, 0, st>>>(a->theta.as<double>(), a->grad.as<double>(), a->cost.as<double>(), a->mom.as<double>(),
                                                    a->var.as<double>(), a->states.as<sq::AdamState>(), a->scratch.as<double>(),
                                                    a->best_theta.as<double>(), d_hist_slot, a->P, a->cfg);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return SQGPU_OK;
}

}  // namespace

extern "C" {

int sqgpu_adam_init(sqgpu_handle_t c, const double* theta0, int batch, double eta, double beta1, double beta2, double epsilon) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 1 || batch > 65535 || (!theta0 && c->n_params > 0)) return fail(SQGPU_ERR_INVALID, "bad arguments");
    if (!(eta > 0) || !(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1) || !(epsilon > 0)) return fail(SQGPU_ERR_INVALID, "bad ADAM hyperparameters");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (!c->adam) c->adam = new AdamRun();
    AdamRun* a = c->adam;
    if (a->graph) {
        cudaGraphExecDestroy(a->graph);
        a->graph = nullptr;
    }
    a->batch = batch;
    a->P = c->n_params;
    a->cfg = sq::AdamCfg{eta, beta1, beta2, epsilon};
    const size_t np = std::max<size_t>(1, (size_t)batch * a->P) * sizeof(double);
    if ((rc = a->theta.ensure(np)) || (rc = a->mom.ensure(np)) || (rc = a->var.ensure(np)) || (rc = a->grad.ensure(np)) ||
        (rc = a->best_theta.ensure(np)) || (rc = a->scratch.ensure(2 * np)) || (rc = a->cost.ensure((size_t)batch * sizeof(double))) ||
        (rc = a->states.ensure((size_t)batch * sizeof(sq::AdamState))))
        return rc;
    if (a->P) CUDA_TRY(cudaMemcpyAsync(a->theta.p, theta0, (size_t)batch * a->P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(a->mom.p, 0, np, c->stream));  // initialize_moment_and_variance, Adam.cpp:113-118
    CUDA_TRY(cudaMemsetAsync(a->var.p, 0, np, c->stream));
    if (a->P) CUDA_TRY(cudaMemcpyAsync(a->best_theta.p, a->theta.p, (size_t)batch * a->P * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    sq::adam_reset_kernel<<<(batch + 127) / 128, 128, 0, c->stream>>>(a->states.as<sq::AdamState>(), batch);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_adam_steps(sqgpu_handle_t c, int n_steps, double* cost_history) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (n_steps < 0) return fail(SQGPU_ERR_INVALID, "negative number of steps");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    if (!c->adam || c->adam->batch == 0) return fail(SQGPU_ERR_STATE, "call sqgpu_adam_init first");
    AdamRun* a = c->adam;
    if (a->P != c->n_params) return fail(SQGPU_ERR_STATE, "the circuit changed since sqgpu_adam_init");
    if (n_steps == 0) return SQGPU_OK;
    CallScope cs(c, c->stream);
    int rc;
    if ((rc = a->hist.ensure((size_t)n_steps * a->batch * sizeof(double)))) return rc;
    double* hist = a->hist.as<double>();
    cudaStream_t st = c->stream;
    int done = 0;
    // the first step runs eagerly (it may grow workspaces); from the second on one CUDA graph replay per step: the graph leaves
    // the step's costs in a->cost, a small device copy moves them to the step's row of the history.
    if ((rc = adam_enqueue_step(c, a, hist, st))) return rc;
    done = 1;
    if (n_steps > 1 && !c->opt.no_graph) {
        if (!a->graph) {
            cudaGraph_t g = nullptr;
            c->capturing = true;
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                rc = adam_enqueue_step(c, a, nullptr, st);
                e = cudaStreamEndCapture(st, &g);
            }
            c->capturing = false;
            if (rc == SQGPU_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&a->graph, g, 0);
            if (g) cudaGraphDestroy(g);
            if (rc != SQGPU_OK || e != cudaSuccess) {  // capture not possible for this configuration: plain launches
                cudaGetLastError();
                if (a->graph) cudaGraphExecDestroy(a->graph);
                a->graph = nullptr;
            }
        }
    }
    for (; done < n_steps; ++done) {
        if (a->graph) {
            CUDA_TRY(cudaGraphLaunch(a->graph, st));
            CUDA_TRY(cudaMemcpyAsync(hist + (size_t)done * a->batch, a->cost.p, (size_t)a->batch * sizeof(double), cudaMemcpyDeviceToDevice, st));
            c->launches += 8;
        } else if ((rc = adam_enqueue_step(c, a, hist + (size_t)done * a->batch, st))) {
            return rc;
        }
    }
    if (cost_history) CUDA_TRY(cudaMemcpyAsync(cost_history, hist, (size_t)n_steps * a->batch * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return SQGPU_OK;
}

int sqgpu_adam_get(sqgpu_handle_t c, double* theta, double* best_cost, double* best_theta, int* status) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    if (!c->adam || c->adam->batch == 0) return fail(SQGPU_ERR_STATE, "call sqgpu_adam_init first");
    AdamRun* a = c->adam;
    CallScope cs(c, c->stream);
    const size_t np = (size_t)a->batch * a->P * sizeof(double);
    if (theta && np) CUDA_TRY(cudaMemcpyAsync(theta, a->theta.p, np, cudaMemcpyDeviceToHost, c->stream));
    if (best_theta && np) CUDA_TRY(cudaMemcpyAsync(best_theta, a->best_theta.p, np, cudaMemcpyDeviceToHost, c->stream));
    std::vector<sq::AdamState> hs;
    if (best_cost || status) {
        hs.resize(a->batch);
        CUDA_TRY(cudaMemcpyAsync(hs.data(), a->states.p, (size_t)a->batch * sizeof(sq::AdamState), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int y = 0; y < a->batch && !hs.empty(); ++y) {
        if (best_cost) best_cost[y] = hs[y].best_cost;
        if (status) status[y] = hs[y].status;
    }
    return SQGPU_OK;
}

int sqgpu_line_search_batched(sqgpu_handle_t c, const double* x, const double* dir, const double* alphas, int k, double* cost,
                              double* dphi) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (k < 0 || k > 65535) return fail(SQGPU_ERR_INVALID, "bad number of trial points");
    if (k == 0) return SQGPU_OK;
    if (!x || !dir || !alphas || !cost) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc = check_ready(c, true);
    if (rc) return rc;
    const int P = c->n_params;
    if (P == 0) return fail(SQGPU_ERR_INVALID, "the circuit has no parameters");
    const bool with_grad = dphi != nullptr;
    // layout of wParams: [theta k x P | x P | dir P | alphas k | dphi k]
    if ((rc = c->wParams.ensure(((size_t)k * P + 2 * P + 2 * k) * sizeof(double)))) return rc;
    if ((rc = c->wCost.ensure((size_t)k * sizeof(double)))) return rc;
    if (with_grad && (rc = c->wGrad.ensure((size_t)k * P * sizeof(double)))) return rc;
    double* d_theta = c->wParams.as<double>();
    double* d_x = d_theta + (size_t)k * P;
    double* d_dir = d_x + P;
    double* d_alpha = d_dir + P;
    double* d_dphi = d_alpha + k;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(d_x, x, P * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_dir, dir, P * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_alpha, alphas, k * sizeof(double), cudaMemcpyHostToDevice, st));
    sq::line_points_kernel<<<dim3((P + 255) / 256, k), 256, 0, st>>>(d_x, d_dir, d_alpha, d_theta, P);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    if ((rc = eval_dev(c, d_theta, k, with_grad, c->wCost.as<double>(), with_grad ? c->wGrad.as<double>() : nullptr, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(cost, c->wCost.p, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (with_grad) {
        sq::directional_kernel<<<k, 256, 0, st>>>(c->wGrad.as<double>(), d_dir, d_dphi, P);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(dphi, d_dphi, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return SQGPU_OK;
}

}  // extern "C"
