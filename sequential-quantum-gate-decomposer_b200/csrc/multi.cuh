// multi.cuh -- several GPUs behind ONE handle; included at the end of sqgpu.cu.
//
// Replaces the reference's own multi-accelerator / MPI split of the batched cost path
// (Optimization_Interface::optimization_problem_batched_DFE, decomposition/Optimization_Interface.cpp:806-832 and the
// MPI_Allgather variant :962-1004): a C++ host that asked for accelerator_num = G calls sqgpu_create_multi once and then the
// same entry points as on a single device. One host thread drives all devices: every per-device piece of work is enqueued
// asynchronously on that device's stream, the host only waits at the end.
//
// Two sharding axes:
//   batch    parameter vectors are independent (Optimization_Interface.cpp:1009-1025): device d evaluates a contiguous slice of
//            the batch on its own copy of the matrix; results land in the caller's arrays at the slice offsets -- within one
//            process that IS the gather, no collective is needed.
//   columns  left multiplication never mixes columns (kernels/apply_kernel_to_input.cpp:69-89): device d holds
//            U[:, b_d : e_d) and evaluates the raw trace terms with the shard's row offset; ONE ncclAllReduce (sum, fp64) of the
//            trace buffer [B x (1 + P) x 3 x 2] per evaluation -- issued on the devices' compute streams, between the executor
//            and the cost formulas -- gives every device the full traces (the Hilbert-Schmidt variants need the full complex
//            trace before their non-linear formulas, Optimization_Interface.cpp:1414-1419). The Hilbert-Schmidt correction
//            variants need the summed traces of the circuit itself for the weights of their gradient functional: a second,
//            small all-reduce of [B x 3 x 2] in front of the gradient pass.
// NCCL is bound at run time (dlopen of libnccl.so.2: a process that already carries an NCCL, e.g. through torch, keeps using
// that copy); a group of one device never touches it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    int load() {
        if (lib) return SQGPU_OK;
        void* l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!l) l = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!l) return fail(SQGPU_ERR_UNSUPPORTED, "column sharding over several devices needs NCCL: %s", dlerror());
#define SQ_NCCL_SYM(field, name)                                         \
    field = reinterpret_cast<decltype(field)>(dlsym(l, name));          \
    if (!field) return fail(SQGPU_ERR_UNSUPPORTED, "NCCL symbol %s not found", name);
        SQ_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        SQ_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        SQ_NCCL_SYM(AllReduce, "ncclAllReduce")
        SQ_NCCL_SYM(GroupStart, "ncclGroupStart")
        SQ_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        SQ_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SQ_NCCL_SYM
        lib = l;
        return SQGPU_OK;
    }
};

static NcclApi g_nccl;

#define NCCL_TRY(expr)                                                                                         \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess) return fail(SQGPU_ERR_CUDA, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

struct MultiGpu {
    int n = 0;
    int requested_mode = SQGPU_SHARD_AUTO;
    int mode = SQGPU_SHARD_BATCH;  // resolved when the matrix is uploaded
    std::vector<sqgpu_ctx*> dev;
    std::vector<ncclComm_t> comms;
    bool have_comms = false;
    int rows = 0, cols_total = 0;
    std::vector<int> col_begin, col_end;
    // the front handle keeps the configuration so that it can be re-applied when the sharding mode changes
    int variant = SQGPU_FROBENIUS_NORM, trace_offset = 0;
    double prev = 1.0, c1 = 1.0 / 1.7, c2 = 0.5;
};

namespace {

void shard_range(int total, int part, int parts, int* b, int* e) {
    const int base = total / parts, rem = total % parts;
    *b = part * base + std::min(part, rem);
    *e = *b + base + (part < rem ? 1 : 0);
}

int multi_init_comms(MultiGpu* m) {
    if (m->have_comms || m->n == 1) return SQGPU_OK;
    int rc = g_nccl.load();
    if (rc) return rc;
    std::vector<int> ids(m->n);
    for (int d = 0; d < m->n; ++d) ids[d] = m->dev[d]->device;
    m->comms.assign(m->n, nullptr);
    NCCL_TRY(g_nccl.CommInitAll(m->comms.data(), m->n, ids.data()));
    m->have_comms = true;
    return SQGPU_OK;
}

// in-place sum over the devices of buf_d[0 .. count) (device d's buffer), on the devices' compute streams
int multi_allreduce(MultiGpu* m, const std::vector<double*>& bufs, size_t count) {
    if (m->n == 1 || count == 0) return SQGPU_OK;
    int rc = multi_init_comms(m);
    if (rc) return rc;
    NCCL_TRY(g_nccl.GroupStart());
    for (int d = 0; d < m->n; ++d) {
        ncclResult_t r = g_nccl.AllReduce(bufs[d], bufs[d], count, ncclDouble, ncclSum, m->comms[d], m->dev[d]->stream);
        if (r != ncclSuccess) {
            g_nccl.GroupEnd();
            return fail(SQGPU_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
        }
    }
    NCCL_TRY(g_nccl.GroupEnd());
    for (int d = 0; d < m->n; ++d) m->dev[d]->launches++;
    return SQGPU_OK;
}

int multi_sync_all(MultiGpu* m) {
    for (int d = 0; d < m->n; ++d) {
        DeviceGuard g(m->dev[d]->device);
        CUDA_TRY(cudaStreamSynchronize(m->dev[d]->stream));
    }
    return SQGPU_OK;
}

int multi_apply_cost(MultiGpu* m) {
    for (int d = 0; d < m->n; ++d) {
        int rc = set_cost_checked(m->dev[d], m->variant, m->trace_offset, m->prev, m->c1, m->c2);
        if (rc) return rc;
    }
    return SQGPU_OK;
}

int multi_upload(sqgpu_ctx* front, const double* data, int rows, int cols, int stride) {
    MultiGpu* m = front->multi;
    int mode = m->requested_mode;
    if (mode == SQGPU_SHARD_AUTO)  // tall matrices shard by columns (one column tile per CTA leaves few parameter sets in
                                   // flight); small ones and state vectors by parameter vectors
        mode = (cols >= 2048 && cols >= m->n) ? SQGPU_SHARD_COLUMNS : SQGPU_SHARD_BATCH;
    if (mode == SQGPU_SHARD_COLUMNS && cols < m->n) return fail(SQGPU_ERR_INVALID, "column sharding: %d columns over %d devices", cols, m->n);
    m->mode = mode;
    m->rows = rows;
    m->cols_total = cols;
    m->col_begin.assign(m->n, 0);
    m->col_end.assign(m->n, cols);
    for (int d = 0; d < m->n; ++d) {
        sqgpu_ctx* c = m->dev[d];
        int b = 0, e = cols;
        if (mode == SQGPU_SHARD_COLUMNS) shard_range(cols, d, m->n, &b, &e);
        m->col_begin[d] = b;
        m->col_end[d] = e;
        int rc = sqgpu_upload_matrix(c, data + 2 * (size_t)b, rows, e - b, stride);
        if (rc) return rc;
        c->shard_offset = (mode == SQGPU_SHARD_COLUMNS) ? b : 0;
        c->shard_cols_total = cols;
    }
    front->rows = rows;
    front->cols = cols;
    return SQGPU_OK;
}

// cost (and gradient) for a batch of host parameter vectors on a multi-device handle
int multi_eval(sqgpu_ctx* front, const double* params, int batch, bool with_grad, double* cost, double* grad) {
    MultiGpu* m = front->multi;
    if (batch < 0) return fail(SQGPU_ERR_INVALID, "negative batch");
    if (batch == 0) return SQGPU_OK;
    const int P = front->n_params;
    if ((!params && P > 0) || !cost || (with_grad && !grad && P > 0)) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    int rc;
    if (m->mode == SQGPU_SHARD_BATCH) {
        std::vector<int> b0(m->n), b1(m->n);
        for (int d = 0; d < m->n; ++d) {  // enqueue everything first ...
            shard_range(batch, d, m->n, &b0[d], &b1[d]);
            const int nb = b1[d] - b0[d];
            if (nb == 0) continue;
            sqgpu_ctx* c = m->dev[d];
            DeviceGuard g(c->device);
            if ((rc = check_ready(c, true))) return rc;
            const size_t np = (size_t)nb * P;
            if ((rc = c->wParams.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
            if ((rc = c->wCost.ensure((size_t)nb * sizeof(double)))) return rc;
            if (with_grad && (rc = c->wGrad.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
            if (np) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params + (size_t)b0[d] * P, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            if ((rc = eval_dev(c, c->wParams.as<double>(), nb, with_grad, c->wCost.as<double>(), with_grad ? c->wGrad.as<double>() : nullptr, c->stream))) return rc;
        }
        for (int d = 0; d < m->n; ++d) {  // ... then collect: the devices are all running by now
            const int nb = b1[d] - b0[d];
            if (nb == 0) continue;
            sqgpu_ctx* c = m->dev[d];
            DeviceGuard g(c->device);
            CUDA_TRY(cudaMemcpyAsync(cost + b0[d], c->wCost.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if (with_grad && P) CUDA_TRY(cudaMemcpyAsync(grad + (size_t)b0[d] * P, c->wGrad.p, (size_t)nb * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        return multi_sync_all(m);
    }
    // ---- columns ------------------------------------------------------------------------------------------------------
    if (!variant_supported(m->variant)) return fail(SQGPU_ERR_UNSUPPORTED, "cost function variant %d is not supported on the device path", m->variant);
    const bool hs_corr = m->variant == SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1 || m->variant == SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2;
    const int n_k = 1 + (with_grad ? P : 0);
    const size_t np = (size_t)batch * P;
    std::vector<double*> tr(m->n), tr0(m->n);
    for (int d = 0; d < m->n; ++d) {
        sqgpu_ctx* c = m->dev[d];
        DeviceGuard g(c->device);
        if ((rc = check_ready(c, true))) return rc;
        if ((rc = c->wParams.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
        if ((rc = c->wTraces.ensure((size_t)batch * n_k * 6 * sizeof(double)))) return rc;
        if (np) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        tr[d] = c->wTraces.as<double>();
    }
    if (with_grad && hs_corr) {  // traces of the circuit itself, summed over the shards, for the weights of the functional
        for (int d = 0; d < m->n; ++d) {
            sqgpu_ctx* c = m->dev[d];
            DeviceGuard g(c->device);
            if ((rc = c->wMat.ensure((size_t)batch * 6 * sizeof(double)))) return rc;
            tr0[d] = c->wMat.as<double>();
            if ((rc = traces_dev(c, c->wParams.as<double>(), batch, false, tr0[d], c->stream, false))) return rc;
        }
        if ((rc = multi_allreduce(m, tr0, (size_t)batch * 6))) return rc;
    }
    for (int d = 0; d < m->n; ++d) {
        sqgpu_ctx* c = m->dev[d];
        DeviceGuard g(c->device);
        if ((rc = traces_dev(c, c->wParams.as<double>(), batch, with_grad, tr[d], c->stream, false, (with_grad && hs_corr) ? tr0[d] : nullptr))) return rc;
    }
    if ((rc = multi_allreduce(m, tr, (size_t)batch * n_k * 6))) return rc;
    {   // every device holds the full traces now; device 0 applies the cost formulas and answers
        sqgpu_ctx* c = m->dev[0];
        DeviceGuard g(c->device);
        if ((rc = c->wCost.ensure((size_t)batch * sizeof(double)))) return rc;
        if (with_grad && (rc = c->wGrad.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
        if ((rc = cost_from_traces_dev(c, tr[0], batch, with_grad, m->cols_total, c->wCost.as<double>(), with_grad ? c->wGrad.as<double>() : nullptr, c->stream))) return rc;
        CUDA_TRY(cudaMemcpyAsync(cost, c->wCost.p, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (with_grad && np) CUDA_TRY(cudaMemcpyAsync(grad, c->wGrad.p, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    return multi_sync_all(m);
}

// VQE: parameter sets over the devices (a 2^n state is never split)
int multi_vqe(sqgpu_ctx* front, const double* params, int batch, bool with_grad, double* energy, double* grad) {
    MultiGpu* m = front->multi;
    if (batch < 0) return fail(SQGPU_ERR_INVALID, "negative batch");
    if (batch == 0) return SQGPU_OK;
    const int P = front->n_params;
    if ((!params && P > 0) || !energy || (with_grad && !grad && P > 0)) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    if (m->mode != SQGPU_SHARD_BATCH) return fail(SQGPU_ERR_INVALID, "the VQE path shards parameter sets: use SQGPU_SHARD_BATCH (or AUTO with a state vector)");
    int rc;
    std::vector<int> b0(m->n), b1(m->n);
    for (int d = 0; d < m->n; ++d) {
        shard_range(batch, d, m->n, &b0[d], &b1[d]);
        const int nb = b1[d] - b0[d];
        if (nb == 0) continue;
        sqgpu_ctx* c = m->dev[d];
        DeviceGuard g(c->device);
        const size_t np = (size_t)nb * P;
        if ((rc = c->wParams.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
        if ((rc = c->wCost.ensure((size_t)nb * sizeof(double)))) return rc;
        if (with_grad && (rc = c->wGrad.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
        if (np) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params + (size_t)b0[d] * P, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if ((rc = vqe_dev(c, c->wParams.as<double>(), nb, with_grad, c->wCost.as<double>(), with_grad ? c->wGrad.as<double>() : nullptr, c->stream))) return rc;
    }
    for (int d = 0; d < m->n; ++d) {
        const int nb = b1[d] - b0[d];
        if (nb == 0) continue;
        sqgpu_ctx* c = m->dev[d];
        DeviceGuard g(c->device);
        CUDA_TRY(cudaMemcpyAsync(energy + b0[d], c->wCost.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (with_grad && P) CUDA_TRY(cudaMemcpyAsync(grad + (size_t)b0[d] * P, c->wGrad.p, (size_t)nb * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    return multi_sync_all(m);
}

int multi_destroy(sqgpu_ctx* front) {
    MultiGpu* m = front->multi;
    if (m->have_comms)
        for (ncclComm_t cm : m->comms)
            if (cm) g_nccl.CommDestroy(cm);
    for (sqgpu_ctx* c : m->dev) sqgpu_destroy(c);
    delete m;
    front->multi = nullptr;
    delete front;
    return SQGPU_OK;
}

int multi_set_circuit(sqgpu_ctx* front, const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool,
                      int64_t pool_len) {
    for (sqgpu_ctx* c : front->multi->dev) {
        const int rc = sqgpu_set_circuit(c, gates, n_gates, n_params, qbit_num, matrix_pool, pool_len);
        if (rc) return rc;
    }
    front->n_params = n_params;
    front->qbit_num = qbit_num;
    front->n_gates = n_gates;
    front->circuit_set = true;
    return SQGPU_OK;
}

int multi_set_cost(sqgpu_ctx* front, int variant, int trace_offset, double prev, double c1, double c2) {
    MultiGpu* m = front->multi;
    if (!variant_supported(variant)) return fail(SQGPU_ERR_UNSUPPORTED, "cost function variant %d is not supported on the device path", variant);
    m->variant = variant;
    m->trace_offset = trace_offset;
    m->prev = prev;
    m->c1 = c1;
    m->c2 = c2;
    return multi_apply_cost(m);
}

int multi_set_option(sqgpu_ctx* front, const char* name, int64_t value) {
    for (sqgpu_ctx* c : front->multi->dev) {
        const int rc = option_set(c->opt, name, value);
        if (rc) return rc;
    }
    return SQGPU_OK;
}

int multi_set_hamiltonian(sqgpu_ctx* front, int n_rows, int64_t nnz, const int32_t* indptr, const int32_t* indices, const double* values) {
    for (sqgpu_ctx* c : front->multi->dev) {
        const int rc = sqgpu_set_hamiltonian_csr(c, n_rows, nnz, indptr, indices, values);
        if (rc) return rc;
    }
    return SQGPU_OK;
}

long long multi_launches(sqgpu_ctx* front) {
    long long n = 0;
    for (sqgpu_ctx* c : front->multi->dev) n += c->launches;
    return n;
}

sqgpu_ctx* multi_first(sqgpu_ctx* front) { return front->multi->dev[0]; }

}  // namespace

extern "C" {

int sqgpu_create_multi(int n_devices, const int* devices, int mode, sqgpu_handle_t* out) {
    if (!out) return fail(SQGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return fail(SQGPU_ERR_INVALID, "n_devices should be between 1 and 64, got %d", n_devices);
    if (mode != SQGPU_SHARD_AUTO && mode != SQGPU_SHARD_BATCH && mode != SQGPU_SHARD_COLUMNS) return fail(SQGPU_ERR_INVALID, "unknown sharding mode %d", mode);
    for (int d = 0; d < n_devices; ++d)
        for (int e = 0; e < d; ++e)
            if (devices && devices[d] == devices[e]) return fail(SQGPU_ERR_INVALID, "device %d listed twice", devices[d]);
    MultiGpu* m = new MultiGpu();
    m->n = n_devices;
    m->requested_mode = mode;
    for (int d = 0; d < n_devices; ++d) {
        sqgpu_handle_t h = nullptr;
        const int rc = sqgpu_create(devices ? devices[d] : d, &h);
        if (rc) {
            for (sqgpu_ctx* c : m->dev) sqgpu_destroy(c);
            delete m;
            return rc;
        }
        m->dev.push_back(h);
    }
    sqgpu_ctx* front = new sqgpu_ctx();  // a front object: no stream, no device memory of its own
    front->device = m->dev[0]->device;
    front->sm_count = m->dev[0]->sm_count;
    front->multi = m;
    *out = front;
    return SQGPU_OK;
}

int sqgpu_multi_info(sqgpu_handle_t c, int* n_devices, int* mode) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (n_devices) *n_devices = c->multi ? c->multi->n : 1;
    if (mode) *mode = c->multi ? c->multi->mode : SQGPU_SHARD_BATCH;
    return SQGPU_OK;
}

// Column sharding driven from OUTSIDE the library (one process per GPU, torch.distributed / MPI for the exchange): the
// resident matrix of this handle is U[:, col_begin : col_begin + cols) of a matrix with cols_total columns.
int sqgpu_set_shard(sqgpu_handle_t c, int col_begin, int cols_total) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (c->multi) return fail(SQGPU_ERR_INVALID, "a multi-device handle shards by itself");
    if (col_begin < 0 || cols_total < 0) return fail(SQGPU_ERR_INVALID, "bad shard arguments");
    std::lock_guard<std::mutex> lk(c->mtx);
    c->shard_offset = col_begin;
    c->shard_cols_total = cols_total;
    return SQGPU_OK;
}

// gradient traces of a column shard for the Hilbert-Schmidt correction variants: d_global_traces0 [batch][3][2] are the traces of
// the circuit itself (sqgpu_traces_batched_dev with with_grad = 0) already summed over all shards
int sqgpu_grad_traces_with_global_dev(sqgpu_handle_t c, const double* d_params, int batch, const double* d_global_traces0,
                                      double* d_traces, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (c->multi) return fail(SQGPU_ERR_UNSUPPORTED, "not available on a multi-device handle");
    if (batch < 0 || (batch > 0 && (!d_traces || !d_global_traces0))) return fail(SQGPU_ERR_INVALID, "bad arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return traces_dev(c, d_params, batch, true, d_traces, (cudaStream_t)stream, false, d_global_traces0);
}

}  // extern "C"
