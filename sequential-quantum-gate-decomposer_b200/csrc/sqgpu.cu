// sqgpu.cu -- context, circuit lowering and the extern "C" entry points declared in include/sqgpu.h.
//
// Host side of the B200 engine: owns device memory (grow-only workspaces, the resident matrix of
// sqgpu_upload_matrix == the reference's load2LMEM, common/common_DFE.cpp:129-132), lowers the gate descriptors to
// the device program once per sqgpu_set_circuit, and launches
//   build_kernel_tables -> fused_exec<COST|GRAD|APPLY> (or the streaming kernels) -> reduce_partials -> cost_from_traces
// on one stream without host round trips. No CPU evaluation path exists in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sqgpu.h"
#include "exec_fused.cuh"
#include "exec_stream.cuh"
#include "gate_kernels.cuh"
#include "reduce.cuh"
#include "sq_types.cuh"
#include "vqe.cuh"

using namespace sq;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(_e == cudaErrorMemoryAllocation ? SQGPU_ERR_NOMEM : SQGPU_ERR_CUDA, "%s failed: %s", \
                        #expr, cudaGetErrorString(_e));                                                     \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return SQGPU_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(SQGPU_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return SQGPU_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct KernelTimer {  // CUDA-event timing of the library's dominant kernels on the launching stream, one ring per kernel name
    static const int RING = 64, NAMES = 8;
    struct Ring {
        std::string name;
        cudaEvent_t e0[RING], e1[RING];
        int n = 0;
        bool init = false;
    } rings[NAMES];
    Ring* cur = nullptr;  // ring of the last time_begin
    Ring* get(const char* name) {
        int free_slot = -1;
        for (int i = 0; i < NAMES; ++i) {
            if (rings[i].init && rings[i].name == name) return &rings[i];
            if (!rings[i].init && free_slot < 0) free_slot = i;
        }
        if (free_slot < 0) free_slot = 0;  // more names than rings: recycle the first
        Ring& r = rings[free_slot];
        if (!r.init) {
            for (int i = 0; i < RING; ++i) {
                cudaEventCreate(&r.e0[i]);
                cudaEventCreate(&r.e1[i]);
            }
            r.init = true;
        }
        if (r.name != name) {
            r.name = name;
            r.n = 0;
        }
        return &r;
    }
    void destroy() {
        for (auto& r : rings) {
            if (!r.init) continue;
            for (int i = 0; i < RING; ++i) {
                cudaEventDestroy(r.e0[i]);
                cudaEventDestroy(r.e1[i]);
            }
            r.init = false;
        }
    }
};

// Per-handle switches (sqgpu_set_option). Defaults are the product configuration; the others exist for the parity tests
// (both executor paths are covered) and for kernel experiments. They replace the environment variables of round 1: a
// setenv in some host thread can no longer change the numerics path of a running handle.
struct Options {
    int no_fuse = 0;           // every gate is its own op
    int max_fuse_qubits = 3;   // widest fused block of plan3 (2 or 3)
    int fuse_consecutive = 0;  // runs of consecutive gates only (no commuting reorder)
    int force_stream = 0;      // cost / gradient through the one-op-per-launch streaming executor
    int vqe_stream = 0;        // VQE through the streaming path instead of the windowed executor
    int window = 11;           // window width of the state-vector segment planner
    int split = 0;             // CTAs per SM the fused executor aims for (0: default of the call site)
    int split_force = 0;
    int threads = 0;           // cap of the fused executor's CTA size (0: none)
    int ctas_per_sm = 48;      // grid granularity of the cost / gradient executor
    int verbose = 0;
    int no_graph = 0;          // device-resident optimizer loops: plain launches instead of one CUDA graph replay per step
    int tall_window = 1;       // matrices whose column does not fit shared memory: windowed executor (0: streaming fallback)
    int async_tiles = 1;       // windowed executor (SQ_WIN_BULK builds): double-buffered tiles in the forward segments
    int cluster = 1;           // thread-block-cluster executor: 0 off; 1 (default) where it is the measured winner (a column that
                               // fits one CTA only once per SM) or the only fused option; 2: also instead of the windowed executor
    int split_tables = 1;      // derivative kernel tables of fused blocks by one warp per member (build_block_derivs): 0 off, 1 for small
                               // batches (the single evaluation a BFGS line search waits for), 2 always; bit-identical tables either way
    int const_fuse_qubits = 4; // constant sub-circuits are multiplied out on the host into dense blocks of up to this many qubits (0: off)
};

typedef int Options::*OptionField;
struct OptionName { const char* name; OptionField field; long long lo, hi; };
const OptionName kOptionNames[] = {
    {"no_fuse", &Options::no_fuse, 0, 1},
    {"max_fuse_qubits", &Options::max_fuse_qubits, 2, 3},
    {"fuse_consecutive", &Options::fuse_consecutive, 0, 1},
    {"force_stream", &Options::force_stream, 0, 1},
    {"vqe_stream", &Options::vqe_stream, 0, 1},
    {"window", &Options::window, 1, 30},
    {"split", &Options::split, 0, 16},
    {"split_force", &Options::split_force, 0, 1},
    {"threads", &Options::threads, 0, 1024},
    {"ctas_per_sm", &Options::ctas_per_sm, 1, 1024},
    {"verbose", &Options::verbose, 0, 1},
    {"no_graph", &Options::no_graph, 0, 1},
    {"tall_window", &Options::tall_window, 0, 1},
    {"async_tiles", &Options::async_tiles, 0, 1},
    {"const_fuse_qubits", &Options::const_fuse_qubits, 0, 5},
    {"cluster", &Options::cluster, 0, 2},
    {"split_tables", &Options::split_tables, 0, 2},
};

int option_set(Options& o, const char* name, long long value) {
    if (!name) return fail(SQGPU_ERR_INVALID, "option name is NULL");
    for (const auto& on : kOptionNames)
        if (!strcmp(on.name, name)) {
            if (value < on.lo || value > on.hi) return fail(SQGPU_ERR_INVALID, "option %s: value %lld outside [%lld, %lld]", name, value, on.lo, on.hi);
            o.*(on.field) = (int)value;
            return SQGPU_OK;
        }
    return fail(SQGPU_ERR_INVALID, "unknown option '%s'", name);
}

// "name=value,name=value" (sqgpu_plan_stats_opt)
int options_parse(Options& o, const char* text) {
    if (!text) return SQGPU_OK;
    std::string t(text);
    size_t pos = 0;
    while (pos < t.size()) {
        size_t end = t.find(',', pos);
        if (end == std::string::npos) end = t.size();
        const std::string item = t.substr(pos, end - pos);
        pos = end + 1;
        if (item.empty()) continue;
        const size_t eq = item.find('=');
        if (eq == std::string::npos) return fail(SQGPU_ERR_INVALID, "option '%s': expected name=value", item.c_str());
        char* endp = nullptr;
        const long long v = strtoll(item.c_str() + eq + 1, &endp, 10);
        if (!endp || *endp) return fail(SQGPU_ERR_INVALID, "option '%s': value is not an integer", item.c_str());
        int rc = option_set(o, item.substr(0, eq).c_str(), v);
        if (rc) return rc;
    }
    return SQGPU_OK;
}

}  // namespace

// A lowered device program. Two are kept per circuit: `plan2` fuses gate runs inside at most two qubits (4 x 4 blocks;
// used by the streaming / apply / VQE paths, whose kernels stop at two-qubit blocks) and `plan3` inside at most three
// qubits (8 x 8 blocks on the FP64 tensor cores; used by the shared-memory executor).
struct Plan {
    std::vector<DevOp> ops;          // device program (fused blocks + raw ops)
    std::vector<DevMember> members;  // gates inside the fused blocks
    std::vector<int> param_op;       // parameter -> op that owns it
    std::vector<int> param_slot;     // parameter -> index of its derivative kernel inside that op
    DevBuf dOps, dMembers, dParamOp;
    DevBuf wKtab, wDKtab, wOpTab;    // per-parameter-set kernel tables and DMMA block lookup tables (workspace)
    DevBuf wDenseTab, wDenseTab5;    // fragment tables of the raw dense 3-/4- and 5-qubit ops (constant kernels): per tile width
    int n_dense = 0, n_dense5 = 0, dense_logct = -1;
    int n_ops = 0, kern_total = 0, dkern_total = 0, w_total = 0, wmax = 4;
    int dense_stage = 0;             // complex elements of kernel staging the executor's generic dense path needs
    bool has_dense = false;          // raw ops wider than one qubit (GENERAL blocks, controlled two-target gates): DNS kernels
};

struct sqgpu_ctx {
    int device = 0;
    int sm_count = 148;
    int smem_optin = 0;
    int smem_per_sm = 0;
    std::mutex mtx;
    cudaStream_t stream = nullptr;

    // resident matrix (sqgpu_upload_matrix)
    DevBuf U;
    int rows = 0, cols = 0;

    // circuit
    Plan plan2, plan3;
    // windowed state-vector executor (VQE): plan3's ops reordered into segments whose joint support fits `win_w` qubits,
    // qubit indices rewritten to positions inside the segment's window
    Plan planW;
    // cluster executor (n >= 12): plan3 with RESPLIT ops inserted and qubits rewritten to row-bit positions, for clusters of
    // 2^rho CTAs (index rho - 1); cl_fin[rho - 1][q] = where logical qubit q sits after the forward sweep
    static const int MAX_RHO = 3;
    Plan planC[MAX_RHO];
    bool cl_ok[MAX_RHO] = {false, false, false};
    signed char cl_fin[MAX_RHO][32];
    int cl_resplits[MAX_RHO] = {0, 0, 0};
    int use_rho = 0;                 // cluster size (log2) of the evaluation in progress; 0: single-CTA tiles
    struct Segment { int begin, end; unsigned wmask; };
    std::vector<Segment> segs;
    int win_w = 0;
    Plan* P = &plan2;                // plan the helpers below operate on (set by the entry point, under the mutex)
    DevBuf dPool;
    int n_params = 0, qbit_num = 0, n_gates = 0, n_const_fused = 0;
    bool all_unitary = true, circuit_set = false;
    double table_shift = 0.0;  // != 0 while sqgpu_cost_shifted_batched runs: the derivative tables hold K(theta_p + shift) - K(theta_p)
    // geometry of the last gradient reduction of the resident fused executor: its W' partials stay valid, so another shift of
    // the same parameter sets needs new tables and a second reduce_partials only, not a second sweep
    struct { bool valid = false; int batch = 0, chunks = 0, w_slices = 0; const void* plan = nullptr; } replay;
    std::vector<cplx> pool;

    // cost configuration
    CostCfg cfg{SQGPU_FROBENIUS_NORM, 1.0, 1.0 / 1.7, 0.5};
    int trace_offset = 0;
    // column sharding (sqgpu_set_shard / multi-device handles): the resident matrix is U[:, shard_offset : shard_offset + cols]
    // of a matrix with shard_cols_total columns; diagonal element j of the shard sits in row j + shard_offset for EVERY variant
    int shard_offset = 0, shard_cols_total = 0;
    struct MultiGpu* multi = nullptr;  // non-null: this handle is the front of a multi-device group (multi.cuh)
    struct AdamRun* adam = nullptr;    // device-resident ADAM trajectories (optim.cuh)
    bool capturing = false;            // a CUDA graph capture is in progress on the handle's stream: no event timers

    // workspaces
    DevBuf wParams, wTrPart, wWPart, wTraces, wOmega, wCost, wGrad, wMat, wDerivIdx, wTraces0;

    // Hamiltonian (VQE)
    DevBuf hIndptr, hIndices, hValues;
    int h_rows = 0;
    long long h_nnz = 0;

    long long launches = 0;
    KernelTimer timer;
    Options opt;
    double last_flops[2] = {0, 0};  // FP64 flops the last cost / gradient executor launch issues: tensor (DMMA), scalar (DFMA)
    int last_shape[6] = {-1, 0, 0, 0, 0, 0};  // fused executor launch of the last evaluation: log_ct, threads, chunks, tiles per CTA, smem, cluster size

    // Every entry point that enqueues work records `last_done` on its stream when it returns; the next entry point makes its
    // own stream wait for it first, so calls on DIFFERENT streams cannot race on the handle's shared workspaces (the mutex
    // only serialises the host side), and the blocking uploads wait for it before they overwrite device data.
    cudaEvent_t last_done = nullptr;
    bool last_valid = false;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// Orders this call after the previous one on the handle, whatever streams the two use (see sqgpu_ctx::last_done).
struct CallScope {
    sqgpu_ctx* c;
    cudaStream_t st;
    CallScope(sqgpu_ctx* c_, cudaStream_t st_) : c(c_), st(st_) {
        if (c->last_valid) cudaStreamWaitEvent(st, c->last_done, 0);
    }
    ~CallScope() {
        if (c->last_done && cudaEventRecord(c->last_done, st) == cudaSuccess) c->last_valid = true;
    }
};

// before a blocking upload overwrites device data: everything enqueued on the handle so far has finished
void wait_idle(sqgpu_ctx* c) {
    if (c->last_valid) cudaEventSynchronize(c->last_done);
    if (c->stream) cudaStreamSynchronize(c->stream);
}

// restores the handle's current plan when a helper that switches it returns (also on its error paths)
struct PlanScope {
    sqgpu_ctx* c;
    Plan* saved;
    explicit PlanScope(sqgpu_ctx* c_) : c(c_), saved(c_->P) {}
    ~PlanScope() { c->P = saved; }
};

// 16 independent DFMA chains per thread: measures the FP64 pipe, nothing else
__global__ void fp64_fma_burn(double* out, int iters, double m) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 tensor-core (DMMA m8n8k4) burn: 4 independent accumulator fragments per warp
__global__ void fp64_dmma_burn(double* out, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - threadIdx.x * 1e-4;
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// both pipes at once: does DMMA run beside DFMA?
__global__ void fp64_mixed_burn(double* out, int iters, double m) {
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - threadIdx.x * 1e-4;
    double c[2][2] = {{0, 0}, {0, 0}};
    double f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fma(f[i], m, 1e-9);
    }
    double s = c[0][0] + c[0][1] + c[1][0] + c[1][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int n_trace_types_of(int variant) {
    switch (variant) {
        case SQGPU_FROBENIUS_NORM_CORRECTION1:
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1: return 2;
        case SQGPU_FROBENIUS_NORM_CORRECTION2:
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2: return 3;
        default: return 1;
    }
}

bool variant_supported(int v) {
    return v == SQGPU_FROBENIUS_NORM || v == SQGPU_FROBENIUS_NORM_CORRECTION1 || v == SQGPU_FROBENIUS_NORM_CORRECTION2 ||
           v == SQGPU_HILBERT_SCHMIDT_TEST || v == SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1 ||
           v == SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2 || v == SQGPU_INFIDELITY || v == SQGPU_SUM_OF_SQUARES;
}

// the trace offset only enters the Frobenius-family cost functions (get_cost_function*, :73-404); get_trace* ignore it
int effective_offset(const sqgpu_ctx* c) {
    return (c->cfg.variant <= SQGPU_FROBENIUS_NORM_CORRECTION2 ? c->trace_offset : 0) + c->shard_offset;
}

int param_count_of(int type) {
    switch (type) {
        case SQGPU_U3: return 3;
        case SQGPU_CU: return 4;
        case SQGPU_U2: case SQGPU_R: case SQGPU_CR: case SQGPU_CROT: return 2;
        case SQGPU_RX: case SQGPU_RY: case SQGPU_RZ: case SQGPU_U1: case SQGPU_CRY: case SQGPU_CRX: case SQGPU_CRZ:
        case SQGPU_CP: case SQGPU_ADAPTIVE: case SQGPU_RXX: case SQGPU_RYY: case SQGPU_RZZ: return 1;
        case SQGPU_GENERAL: case SQGPU_CZ: case SQGPU_CNOT: case SQGPU_CH: case SQGPU_X: case SQGPU_Y: case SQGPU_Z:
        case SQGPU_H: case SQGPU_S: case SQGPU_SDG: case SQGPU_T: case SQGPU_TDG: case SQGPU_SX: case SQGPU_SXDG:
        case SQGPU_CCX: case SQGPU_SWAP: case SQGPU_CSWAP: case SQGPU_SYC: return 0;
        default: return -1;
    }
}

// fixed row-index bits of an op's group enumeration: its targets and controls, ascending; unused slots = 30
void fill_fix(DevOp& op) {
    unsigned m = op.ctrl_mask;
    if (op.dim == 2) m |= 1u << op.target;
    else
        for (int j = 0; j < op.nq; ++j) m |= 1u << op.q[j];
    op.nfix = 0;
    for (int b = 0; b < 30; ++b)
        if ((m >> b) & 1) {
            if (op.nfix < 6) op.fix[op.nfix] = b;
            op.nfix++;
        }
    for (int f = op.nfix; f < 6; ++f) op.fix[f] = 30;
}

unsigned support_mask(const DevOp& op) {
    unsigned m = op.ctrl_mask;
    if (op.dim == 2) m |= 1u << op.target;
    else
        for (int j = 0; j < op.nq; ++j) m |= 1u << op.q[j];
    return m;
}

int popcount32(unsigned v) {
    int c = 0;
    for (; v; v &= v - 1) ++c;
    return c;
}

// Windowed plan for state vectors: the ops of plan3 are scheduled into SEGMENTS whose joint support fits
// win_w qubits. Inside a segment the executor keeps a 2^win_w x CT tile of the state in shared memory -- rows = the window
// qubits, columns = values of the other bits -- and runs all of the segment's ops on it, so the state makes one HBM round
// trip per segment instead of one per gate (the reference streams it once per gate,
// apply_kernel_to_state_vector_input.cpp:33-229). Ops on disjoint qubits commute: an op may be pulled forward when no
// earlier unscheduled op shares a qubit with it. Qubit indices are rewritten to positions inside the window.
int build_window_plan(sqgpu_ctx* c, bool upload) {
    const Plan& src = c->plan3;
    Plan& dst = c->planW;
    const int N = (int)src.ops.size(), n = c->qbit_num;
    const int w = std::max(1, std::min(n, c->opt.window));
    c->win_w = w;
    c->segs.clear();
    std::vector<unsigned> sup(N);
    for (int k = 0; k < N; ++k) sup[k] = support_mask(src.ops[k]);
    // Scheduling. A segment is defined by its window W (w qubits): it takes, in program order, every unscheduled op that
    // lies inside W and shares no qubit with an earlier op left behind. Two schedules are built and the shorter one kept:
    //   (a) first fit: W grows with the ops as they come (pads with the highest free qubits);
    //   (b) best of several candidate windows per segment -- every contiguous range of w qubits (nearest-neighbour
    //       ansatz circuits: HEA n = 20 needs 15 instead of 20 segments at w = 10), the first-fit set, and first-fit sets
    //       seeded with each of the first few ready ops -- scored by the number of ops they admit.
    auto pad = [&](unsigned S) {
        for (int q = n - 1; q >= 0 && popcount32(S) < w; --q) S |= 1u << q;
        return S;
    };
    auto admitted = [&](unsigned W, const std::vector<char>& done, std::vector<int>* got) {
        unsigned blocked = 0;
        int cnt = 0;
        for (int k = 0; k < N; ++k) {
            if (done[k]) continue;
            if ((sup[k] & blocked) == 0 && (sup[k] & ~W) == 0) {
                ++cnt;
                if (got) got->push_back(k);
            } else {
                blocked |= sup[k];
            }
        }
        return cnt;
    };
    auto first_fit_set = [&](const std::vector<char>& done, int seed) {
        unsigned S = seed >= 0 ? sup[seed] : 0u, blocked = 0;
        for (int k = 0; k < N; ++k) {
            if (done[k]) continue;
            if ((sup[k] & blocked) == 0 && popcount32(S | sup[k]) <= w) S |= sup[k];
            else blocked |= sup[k];
        }
        return pad(S);
    };
    auto schedule = [&](bool candidates, std::vector<int>& order, std::vector<sqgpu_ctx::Segment>& segs) {
        std::vector<char> done(N, 0);
        int remaining = N;
        order.clear();
        segs.clear();
        while (remaining > 0) {
            unsigned bestW = first_fit_set(done, -1);
            int best = admitted(bestW, done, nullptr);
            if (candidates) {
                std::vector<unsigned> cand;
                for (int a = 0; a + w <= n; ++a) cand.push_back(((w >= 32 ? 0u : (1u << w)) - 1u) << a);
                unsigned blocked = 0;
                int seeds = 0;
                for (int k = 0; k < N && seeds < 8; ++k) {
                    if (done[k]) continue;
                    if ((sup[k] & blocked) == 0) {
                        cand.push_back(first_fit_set(done, k));
                        ++seeds;
                    }
                    blocked |= sup[k];
                }
                for (unsigned W : cand) {
                    const int cnt = admitted(W, done, nullptr);
                    if (cnt > best) {
                        best = cnt;
                        bestW = W;
                    }
                }
            }
            if (best <= 0) return false;
            sqgpu_ctx::Segment sg;
            sg.begin = (int)order.size();
            std::vector<int> got;
            admitted(bestW, done, &got);
            for (int k : got) {
                done[k] = 1;
                order.push_back(k);
            }
            remaining -= (int)got.size();
            sg.end = (int)order.size();
            sg.wmask = bestW;
            segs.push_back(sg);
        }
        return true;
    };
    std::vector<int> order, order_b, newidx(N, -1);
    std::vector<sqgpu_ctx::Segment> segs_b;
    if (!schedule(false, order, c->segs)) return fail(SQGPU_ERR_UNSUPPORTED, "window planner made no progress (op support wider than the window)");
    if (schedule(true, order_b, segs_b) && segs_b.size() < c->segs.size()) {
        order.swap(order_b);
        c->segs.swap(segs_b);
    }
    for (int i = 0; i < N; ++i) newidx[order[i]] = i;
    dst.ops.clear();
    for (const auto& sg : c->segs)
        for (int i = sg.begin; i < sg.end; ++i) {
            DevOp op = src.ops[order[i]];
            auto loc = [&](int q) { return popcount32(sg.wmask & ((1u << q) - 1u)); };
            if (op.dim == 2) op.target = loc(op.target);
            else {
                const bool hi_first = op.target == op.q[1];
                for (int j = 0; j < op.nq; ++j) op.q[j] = loc(op.q[j]);
                op.target = hi_first ? op.q[1] : op.q[0];
            }
            unsigned cm = 0;
            for (int q = 0; q < 30; ++q)
                if ((op.ctrl_mask >> q) & 1) cm |= 1u << loc(q);
            op.ctrl_mask = cm;
            fill_fix(op);
            dst.ops.push_back(op);
        }
    dst.members = src.members;
    dst.param_slot = src.param_slot;
    dst.param_op = src.param_op;
    for (auto& po : dst.param_op)
        if (po >= 0) po = newidx[po];
    dst.n_ops = N;
    dst.kern_total = src.kern_total;
    dst.dkern_total = src.dkern_total;
    dst.w_total = src.w_total;
    dst.wmax = src.wmax;
    dst.dense_stage = src.dense_stage;
    dst.n_dense = src.n_dense;
    dst.n_dense5 = src.n_dense5;
    dst.dense_logct = -1;
    dst.has_dense = src.has_dense || c->win_w <= 6;
    if (!upload) return SQGPU_OK;  // planning only (sqgpu_plan_stats)
    int rc;
    const size_t np1 = std::max<size_t>(dst.param_op.size(), 1);
    if ((rc = dst.dOps.ensure(std::max<size_t>(1, dst.ops.size()) * sizeof(DevOp)))) return rc;
    if ((rc = dst.dMembers.ensure(std::max<size_t>(1, dst.members.size()) * sizeof(DevMember)))) return rc;
    if ((rc = dst.dParamOp.ensure(2 * np1 * sizeof(int)))) return rc;
    if (!dst.ops.empty()) CUDA_TRY(cudaMemcpy(dst.dOps.p, dst.ops.data(), dst.ops.size() * sizeof(DevOp), cudaMemcpyHostToDevice));
    if (!dst.members.empty()) CUDA_TRY(cudaMemcpy(dst.dMembers.p, dst.members.data(), dst.members.size() * sizeof(DevMember), cudaMemcpyHostToDevice));
    if (!dst.param_op.empty()) {
        CUDA_TRY(cudaMemcpy(dst.dParamOp.p, dst.param_op.data(), dst.param_op.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dst.dParamOp.as<int>() + np1, dst.param_slot.data(), dst.param_slot.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return SQGPU_OK;
}

// Cluster plan (n >= 12; SURVEY a9/a10 for columns that do not fit ONE CTA's shared memory): plan3 for a thread-block cluster
// of 2^rho CTAs that share a column tile. rho SPLIT qubits select the CTA, the other n - rho qubits the row inside it; an op
// may only touch local qubits. The sweep starts with the top rho qubits split (CTA r loads rows [r 2^(n-rho), ...)). Before an
// op that touches a split qubit s, a RESPLIT op exchanges s with a local qubit t that the op does not touch -- the one whose
// next use lies farthest ahead (Belady) -- through distributed shared memory; from then on t selects the CTA and s sits on
// t's old row bit. Every op's qubits are rewritten to the row-bit positions they have when it runs (kernel bit order is kept,
// so a block's positions need not ascend: BlockGeom sorts them where it inserts zero bits). The backward sweep runs the same
// list in reverse (a RESPLIT is its own inverse). Not available with raw dense / two-target ops (their generic paths want
// ascending positions): returns false and the caller keeps the windowed executor.
bool build_cluster_plan(sqgpu_ctx* c, int rho, bool upload, int* rc_out) {
    *rc_out = SQGPU_OK;
    const Plan& src = c->plan3;
    Plan& dst = c->planC[rho - 1];
    const int N = (int)src.ops.size(), n = c->qbit_num, L = n - rho;
    c->cl_ok[rho - 1] = false;
    if (L < 6) return false;
    for (const DevOp& op : src.ops)
        if (op.dim > 2 && op.type != SQ_OP_BLOCK) return false;
    // next use of every qubit at or after op k
    std::vector<int> nxt((size_t)(N + 1) * n, N + n);
    for (int k = N - 1; k >= 0; --k) {
        const unsigned sup = support_mask(src.ops[k]);
        for (int q = 0; q < n; ++q) nxt[(size_t)k * n + q] = ((sup >> q) & 1) ? k : nxt[(size_t)(k + 1) * n + q];
    }
    std::vector<int> pos(n), inv_local(L), rank_q(rho);
    for (int q = 0; q < n; ++q) pos[q] = q < L ? q : 32 + (q - L);
    for (int q = 0; q < L; ++q) inv_local[q] = q;
    for (int i = 0; i < rho; ++i) rank_q[i] = L + i;
    dst.ops.clear();
    std::vector<int> newidx(N, -1);
    int n_resplit = 0;
    for (int k = 0; k < N; ++k) {
        DevOp op = src.ops[k];
        const unsigned sup = support_mask(op);
        for (int i = 0; i < rho; ++i) {
            const int s = rank_q[i];
            if (!((sup >> s) & 1)) continue;
            int best_j = -1, best_next = -1;
            for (int j = 0; j < L; ++j) {
                const int t = inv_local[j];
                if ((sup >> t) & 1) continue;
                const int nu = nxt[(size_t)k * n + t];
                if (nu > best_next) { best_next = nu; best_j = j; }
            }
            if (best_j < 0) return false;
            const int t = inv_local[best_j];
            DevOp r;
            memset(&r, 0, sizeof(r));
            r.type = SQ_OP_RESPLIT;
            r.kern_off = r.dkern_off = r.w_off = -1;
            r.member_off = -1;
            r.target = best_j;
            r.nq = i;
            for (int f = 0; f < 6; ++f) r.fix[f] = 30;
            dst.ops.push_back(r);
            ++n_resplit;
            rank_q[i] = t;
            inv_local[best_j] = s;
            pos[t] = 32 + i;
            pos[s] = best_j;
        }
        if (op.dim == 2) op.target = pos[op.target];
        else {
            const bool hi_first = op.target == op.q[1];
            for (int j = 0; j < op.nq; ++j) op.q[j] = pos[op.q[j]];
            op.target = hi_first ? op.q[1] : op.q[0];
        }
        unsigned cm = 0;
        for (int q = 0; q < n; ++q)
            if ((op.ctrl_mask >> q) & 1) cm |= 1u << pos[q];
        op.ctrl_mask = cm;
        fill_fix(op);
        newidx[k] = (int)dst.ops.size();
        dst.ops.push_back(op);
    }
    for (int q = 0; q < 32; ++q) c->cl_fin[rho - 1][q] = (signed char)(q < n ? pos[q] : 0);
    c->cl_resplits[rho - 1] = n_resplit;
    dst.members = src.members;
    dst.param_slot = src.param_slot;
    dst.param_op = src.param_op;
    for (auto& po : dst.param_op)
        if (po >= 0) po = newidx[po];
    dst.n_ops = (int)dst.ops.size();
    dst.kern_total = src.kern_total;
    dst.dkern_total = src.dkern_total;
    dst.w_total = src.w_total;
    dst.wmax = src.wmax;
    dst.dense_stage = src.dense_stage;
    dst.n_dense = dst.n_dense5 = 0;
    dst.dense_logct = -1;
    dst.has_dense = false;
    c->cl_ok[rho - 1] = true;
    if (!upload) return true;
    int rc;
    const size_t np1 = std::max<size_t>(dst.param_op.size(), 1);
    if ((rc = dst.dOps.ensure(std::max<size_t>(1, dst.ops.size()) * sizeof(DevOp))) || (rc = dst.dMembers.ensure(std::max<size_t>(1, dst.members.size()) * sizeof(DevMember))) ||
        (rc = dst.dParamOp.ensure(2 * np1 * sizeof(int)))) {
        *rc_out = rc;
        return false;
    }
    cudaError_t e = cudaSuccess;
    if (!dst.ops.empty()) e = cudaMemcpy(dst.dOps.p, dst.ops.data(), dst.ops.size() * sizeof(DevOp), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !dst.members.empty()) e = cudaMemcpy(dst.dMembers.p, dst.members.data(), dst.members.size() * sizeof(DevMember), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !dst.param_op.empty()) {
        e = cudaMemcpy(dst.dParamOp.p, dst.param_op.data(), dst.param_op.size() * sizeof(int), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dst.dParamOp.as<int>() + np1, dst.param_slot.data(), dst.param_slot.size() * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        *rc_out = fail(SQGPU_ERR_CUDA, "cluster plan upload failed: %s", cudaGetErrorString(e));
        c->cl_ok[rho - 1] = false;
        return false;
    }
    return true;
}

// Lower one descriptor to a DevOp (offsets are assigned by the caller). Returns 0 or a status.
int lower_gate(const sqgpu_gate_desc& g, int qbit_num, const double* pool, int64_t pool_len, DevOp* out, bool* unitary) {
    DevOp op;
    memset(&op, 0, sizeof(op));
    op.type = g.type;
    op.kern_off = -1;
    op.dkern_off = -1;
    op.w_off = -1;
    const int np = param_count_of(g.type);
    if (np < 0) return fail(SQGPU_ERR_UNSUPPORTED, "gate type %d is not supported on the device path", g.type);
    if (g.n_params != np) return fail(SQGPU_ERR_INVALID, "gate type %d takes %d parameters, descriptor says %d", g.type, np, g.n_params);
    op.n_params = np;
    op.param_start = g.param_start;
    auto qok = [&](int q) { return q >= 0 && q < qbit_num; };
    *unitary = true;
    if (g.type == SQGPU_GENERAL) {
        const int k = g.n_qubits;
        if (k < 1 || k > SQGPU_MAX_GENERAL_QUBITS) return fail(SQGPU_ERR_INVALID, "GENERAL gate: 1..5 qubits supported, got %d", k);
        const int dim = 1 << k;
        if (!pool || g.matrix_off < 0 || g.matrix_off + (int64_t)dim * dim > pool_len) return fail(SQGPU_ERR_INVALID, "GENERAL gate: kernel outside the matrix pool");
        for (int j = 0; j < k; ++j) {
            if (!qok(g.qubits[j])) return fail(SQGPU_ERR_INVALID, "GENERAL gate: qubit %d out of range", g.qubits[j]);
            if (j && g.qubits[j] <= g.qubits[j - 1]) return fail(SQGPU_ERR_INVALID, "GENERAL gate: qubits must be ascending");
        }
        op.pool_off = g.matrix_off;
        op.member_off = -1;
        if (k == 1) {
            op.dim = 2;
            op.target = g.qubits[0];
        } else {
            op.dim = dim;
            op.nq = k;
            for (int j = 0; j < k; ++j) op.q[j] = g.qubits[j];
        }
        // K K^dagger == I ? (the adjoint sweep un-applies gates with K^dagger)
        const double* K = pool + 2 * g.matrix_off;
        double worst = 0;
        for (int r = 0; r < dim; ++r)
            for (int c = 0; c < dim; ++c) {
                double re = 0, im = 0;
                for (int l = 0; l < dim; ++l) {
                    const double ar = K[2 * (r * dim + l)], ai = K[2 * (r * dim + l) + 1];
                    const double br = K[2 * (c * dim + l)], bi = -K[2 * (c * dim + l) + 1];
                    re += ar * br - ai * bi;
                    im += ar * bi + ai * br;
                }
                worst = std::max(worst, std::max(std::fabs(re - (r == c ? 1.0 : 0.0)), std::fabs(im)));
            }
        *unitary = worst < 1e-9;
        fill_fix(op);
        *out = op;
        return SQGPU_OK;
    }
    if (g.type == SQGPU_SYC || g.type == SQGPU_CROT) {
        // two-qubit kernels over (target, control): a dense 4 x 4 op on the sorted pair; op.target keeps the qubit that is
        // kernel bit 0 (build_kernel_tables flips the kernel when that is the higher qubit; block members carry tl / tl2)
        if (!qok(g.target) || !qok(g.control) || g.control == g.target) return fail(SQGPU_ERR_INVALID, "gate type %d: bad target / control qubits %d / %d", g.type, g.target, g.control);
        op.dim = 4;
        op.nq = 2;
        op.q[0] = std::min(g.target, g.control);
        op.q[1] = std::max(g.target, g.control);
        op.target = g.target;
        op.member_off = -1;
        fill_fix(op);
        *out = op;
        return SQGPU_OK;
    }
    const bool two_target = g.type == SQGPU_RXX || g.type == SQGPU_RYY || g.type == SQGPU_RZZ || g.type == SQGPU_SWAP || g.type == SQGPU_CSWAP;
    if (!qok(g.target)) return fail(SQGPU_ERR_INVALID, "gate type %d: target qubit %d out of range", g.type, g.target);
    unsigned cm = 0;
    const bool needs_ctrl = g.type == SQGPU_CNOT || g.type == SQGPU_CZ || g.type == SQGPU_CH || g.type == SQGPU_CU ||
                            g.type == SQGPU_CRY || g.type == SQGPU_CRX || g.type == SQGPU_CRZ || g.type == SQGPU_CP ||
                            g.type == SQGPU_CR || g.type == SQGPU_ADAPTIVE || g.type == SQGPU_CCX || g.type == SQGPU_CSWAP;
    if (needs_ctrl) {
        if (!qok(g.control) || g.control == g.target) return fail(SQGPU_ERR_INVALID, "gate type %d: bad control qubit %d", g.type, g.control);
        cm |= 1u << g.control;
        if (g.type == SQGPU_CCX) {
            if (!qok(g.control2) || g.control2 == g.target || g.control2 == g.control) return fail(SQGPU_ERR_INVALID, "CCX: bad second control qubit %d", g.control2);
            cm |= 1u << g.control2;
        }
    } else if (g.control >= 0) {
        return fail(SQGPU_ERR_INVALID, "gate type %d takes no control qubit", g.type);
    }
    op.ctrl_mask = cm;
    if (two_target) {
        if (!qok(g.target2) || g.target2 == g.target || ((cm >> g.target2) & 1)) return fail(SQGPU_ERR_INVALID, "gate type %d: bad second target qubit %d", g.type, g.target2);
        op.dim = 4;
        op.nq = 2;
        op.q[0] = std::min(g.target, g.target2);
        op.q[1] = std::max(g.target, g.target2);
        op.target = op.q[0];
    } else {
        op.dim = 2;
        op.target = g.target;
    }
    op.member_off = -1;
    fill_fix(op);
    *out = op;
    return SQGPU_OK;
}

// ---- launch planning for the fused executor ---------------------------------------------------------------------

struct FusedPlan {
    bool ok = false;
    int log_ct = 0, threads = 32;
    int tiles = 0, tiles_per_cta = 1, chunks = 1;
    bool w_in_smem = false;
    int rho = 0;            // cluster executor: log2(CTAs per cluster), rows per CTA = rows >> rho
    bool dbuf = false;      // window forward segments: two tile buffers
    bool w_direct = false;  // W' partials: one global slice per (parameter set, CTA, warp), see SQ_W_DIRECT
    int w_slices = 1;       // slices of w_part per parameter set = chunks * (w_direct ? warps per CTA : 1)
    size_t smem = 0;
};

size_t fused_smem(int mode, int rows, int ct, int threads, int dense_stage, int wmax, int w_total, bool w_in_smem, int n_ops, bool window = false,
                  bool dbuf = false) {
    const bool has_b = mode == MODE_GRAD || mode == MODE_BWD;
    const bool w_direct = SQ_W_DIRECT && has_b && !w_in_smem;
    size_t s = (size_t)rows * ct * sizeof(cplx) * ((has_b || (dbuf && mode == MODE_APPLY)) ? 2 : 1);  // a + beta, or two tile buffers
    s += (size_t)dense_stage * sizeof(cplx);
    s += TAB_RING * sizeof(OpTabS);          // ring of DMMA block lookup tables of one sweep direction
    s += TAB_RING * sizeof(unsigned long long);  // ... and their mbarriers
    s += (size_t)n_ops * sizeof(SOp);        // staged op table
    const int nwarps = threads / 32;
    if (has_b) {
        if (!w_direct) s += (size_t)2 * nwarps * wmax * sizeof(cplx);
        if (w_in_smem) s += (size_t)w_total * sizeof(cplx);
    }
    s += (size_t)(nwarps + 1) * 6 * sizeof(double);  // trace partials per warp + the CTA's running sums
    if (window || mode == MODE_BWD) s += (size_t)rows * ct * sizeof(int);  // window mode: offset in the state of every tile slot
    return s;
}

FusedPlan plan_fused(const sqgpu_ctx* c, int mode, int rows, int cols, int ysets, int default_split = 2, bool window = false, int rho = 0) {
    FusedPlan p;
    p.rho = rho;  // cluster executor: `rows` is what ONE CTA holds, chunks counts CTAs (clusters x 2^rho)
    // window forward segments: two tile buffers, the next tile streams in while the current one is computed (option async_tiles)
    const bool dbuf = (SQ_WIN_BULK != 0) && window && mode == MODE_APPLY && c->opt.async_tiles != 0;
    // test hook (option force_stream): cost / gradient evaluations go down the chunked streaming executor
    if (c->opt.force_stream && (mode == MODE_COST || mode == MODE_GRAD)) return p;
    const size_t budget = (size_t)c->smem_optin;
    int max_log = rho > 0 ? 1 : 3;  // (the cluster executor is instantiated for one- and two-column tiles)
    while ((1 << max_log) > cols && max_log > 0) --max_log;  // no wider than the matrix (cols = 1: state vector)
    auto threads_for = [&](int ct) {
        const int items = (rows / 4) * ct;  // groups of a two-qubit block
        return std::min(FUSED_THREADS, std::max(64, (items + 31) / 32 * 32));  // >= 64: the 8 x 8 block kernel prefetch
    };
    // widest tile that fits; the W accumulator lives in shared memory when it fits beside that tile, otherwise in the
    // CTA's own slice of w_part in global memory
    int pick = -1;
    bool pick_wsm = false;
    for (int lc = max_log; lc >= 0 && pick < 0; --lc) {
        const int ct = 1 << lc;
        const bool can_wsm = mode == MODE_GRAD && c->P->w_total > 0;  // MODE_BWD accumulates over several launches: global
        if (can_wsm && fused_smem(mode, rows, ct, threads_for(ct), c->P->dense_stage, c->P->wmax, c->P->w_total, true, c->P->n_ops, window, dbuf) <= budget) {
            pick = lc;
            pick_wsm = true;
        } else if (fused_smem(mode, rows, ct, threads_for(ct), c->P->dense_stage, c->P->wmax, c->P->w_total, false, c->P->n_ops, window, dbuf) <= budget) {
            pick = lc;
            pick_wsm = false;
        }
    }
    if (pick < 0) return p;
    int pick_threads = threads_for(1 << pick);
    // Columns are independent, so two half-width CTAs per SM do the work of one: while one CTA sits in the barrier /
    // table prologue between two ops, the other keeps the FP64 tensor pipe busy. Taken when both fit in the SM.
    {
        const int split = c->opt.split > 0 ? c->opt.split : default_split;
        const bool fc = c->opt.split_force != 0;
        for (int want = split; want >= 2; want /= 2) {  // the most CTAs per SM that fit, then fewer
            int lc = pick, thr = pick_threads, ways = 1;
            while (ways < want && lc > 0 && thr >= 128 && (thr >= 256 || want > 4)) {
                --lc;
                thr /= 2;
                ways *= 2;
            }
            if (ways < 2) break;
            const size_t sm = fused_smem(mode, rows, 1 << lc, thr, c->P->dense_stage, c->P->wmax, c->P->w_total, false, c->P->n_ops, window, dbuf);
            if (fc || ((sm + 1024) * ways <= (size_t)c->smem_per_sm && thr * ways * 128 <= 65536)) {
                pick = lc;
                pick_threads = thr;
                pick_wsm = false;
                break;
            }
        }
    }
    if (c->opt.threads >= 64) pick_threads = std::min(pick_threads, c->opt.threads);  // experiment hook: CTA size
    {
        const int lc = pick, ct = 1 << lc;
        p.ok = true;
        p.log_ct = lc;
        p.threads = pick_threads;
        p.w_in_smem = pick_wsm;
        p.w_direct = SQ_W_DIRECT && (mode == MODE_GRAD || mode == MODE_BWD) && !pick_wsm;
        p.smem = fused_smem(mode, rows, ct, p.threads, c->P->dense_stage, c->P->wmax, c->P->w_total, pick_wsm, c->P->n_ops, window, dbuf);
        p.dbuf = dbuf;
        p.tiles = (cols + ct - 1) / ct;
        if (mode == MODE_APPLY && !window) {
            p.tiles_per_cta = 1;
            p.chunks = p.tiles;
        } else {
            // Grid granularity: chunks CTAs per parameter set, each with tiles_per_cta tiles. The target is ctas_per_sm CTAs per
            // SM (window segments: a few long-lived ones, so that the per-CTA prologue is paid once per ~10 tiles; the windowed
            // backward pass keeps one W' slice per CTA alive across all segment launches); around that target the chunk count
            // with the shortest schedule is taken -- waves x tiles per CTA, the CTAs of a wave running side by side on
            // `slots` = SMs x resident CTAs -- so that the last wave is full (C5 backward: 19 chunks of 14 tiles need 9 waves for
            // 8.2 waves of work; 37 chunks of 7 tiles fill 16 waves exactly).
            const int per_sm = mode == MODE_BWD ? std::min(c->opt.ctas_per_sm, 8) : (mode == MODE_APPLY ? std::min(c->opt.ctas_per_sm, 16) : c->opt.ctas_per_sm);
            const int R = 1 << rho;  // the unit of the schedule is a cluster of R CTAs
            const int want_ctas = std::max(1, c->sm_count * std::max(1, per_sm) / R);
            const int target = std::min(p.tiles, std::max(1, (want_ctas + ysets - 1) / ysets));
            const int resident = std::max(1, std::min((int)((size_t)c->smem_per_sm / (p.smem + 1024)), 65536 / (128 * std::max(p.threads, 32))));  // shared memory, 128 registers per thread
            const long long slots = std::max<long long>(1, (long long)c->sm_count * resident / R);
            long long best_cost = -1;
            int best_chunks = target;
            for (int ch = std::max(1, target / 2); ch <= std::min(p.tiles, 2 * target + 1); ++ch) {
                const int tpc = (p.tiles + ch - 1) / ch;
                const int used = (p.tiles + tpc - 1) / tpc;  // chunks that actually get tiles
                const long long waves = ((long long)used * ysets + slots - 1) / slots;
                const long long cost = waves * tpc;
                // ties: the count nearest to the target
                if (best_cost < 0 || cost < best_cost || (cost == best_cost && std::abs(used - target) < std::abs(best_chunks - target))) {
                    best_cost = cost;
                    best_chunks = used;
                }
            }
            p.tiles_per_cta = (p.tiles + best_chunks - 1) / best_chunks;
            p.chunks = ((p.tiles + p.tiles_per_cta - 1) / p.tiles_per_cta) * R;
        }
        p.w_slices = p.chunks * (p.w_direct ? p.threads / 32 : 1);
    }
    return p;
}

template <int MODE>
cudaError_t launch_fused_mode(const ExecArgs& a, const FusedPlan& p, int ysets, cudaStream_t st) {
    dim3 grid(p.chunks, ysets);
    cudaError_t e = cudaSuccess;
    // two flavours of every kernel: with the raw-dense-op paths (circuits with GENERAL blocks, controlled two-target gates,
    // materialised block derivatives, tiny tiles) and without (see fused_exec, DNS)
    const bool dns = a.dns != 0;
#define SQ_LAUNCH(LC)                                                                                          \
    case LC:                                                                                                   \
        if (dns) {                                                                                             \
            e = cudaFuncSetAttribute(fused_exec<MODE, LC, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem); \
            if (e != cudaSuccess) return e;                                                                    \
            fused_exec<MODE, LC, false, true><<<grid, p.threads, p.smem, st>>>(a);                             \
        } else {                                                                                               \
            e = cudaFuncSetAttribute(fused_exec<MODE, LC, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem); \
            if (e != cudaSuccess) return e;                                                                    \
            fused_exec<MODE, LC, false, false><<<grid, p.threads, p.smem, st>>>(a);                            \
        }                                                                                                      \
        break;
    if (p.rho > 0) {
        // cluster executor: 2^rho consecutive CTAs in x form a thread-block cluster (distributed shared memory)
        if constexpr (MODE == MODE_COST || MODE == MODE_GRAD) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid;
            cfg.blockDim = dim3(p.threads);
            cfg.dynamicSmemBytes = p.smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1u << p.rho;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
#define SQ_LAUNCH_CL(LC)                                                                                                  \
    case LC:                                                                                                              \
        e = cudaFuncSetAttribute(fused_exec<MODE, LC, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);   \
        if (e != cudaSuccess) return e;                                                                                   \
        e = cudaLaunchKernelEx(&cfg, fused_exec<MODE, LC, true, false>, a);                                                      \
        if (e != cudaSuccess) return e;                                                                                   \
        break;
            switch (p.log_ct) {
                SQ_LAUNCH_CL(0)
                SQ_LAUNCH_CL(1)
                default: return cudaErrorInvalidValue;
            }
#undef SQ_LAUNCH_CL
            return cudaGetLastError();
        } else {
            return cudaErrorInvalidValue;
        }
    }
    switch (p.log_ct) {
        SQ_LAUNCH(0)
        SQ_LAUNCH(1)
        SQ_LAUNCH(2)
        SQ_LAUNCH(3)
    }
#undef SQ_LAUNCH
    return cudaGetLastError();
}

// launch plan of the cost / gradient executor for the evaluation in progress (single-CTA tiles or the cluster executor)
FusedPlan plan_eval(const sqgpu_ctx* c, int mode, int ysets) {
    if (c->use_rho > 0) return plan_fused(c, mode, c->rows >> c->use_rho, c->cols, ysets, 2, false, c->use_rho);
    return plan_fused(c, mode, c->rows, c->cols, ysets);
}

// Cluster size (log2) for a cost / gradient evaluation: the smallest cluster whose per-CTA share of the column leaves two CTAs
// per SM, else the smallest that fits at all; 0: no cluster plan for this circuit or nothing fits. Temporarily points c->P at
// the candidate plans; the caller sets c->P afterwards.
int pick_cluster_rho(sqgpu_ctx* c, bool grad, int batch) {
    if (c->cfg.variant == SQGPU_SUM_OF_SQUARES || c->qbit_num < 12) return 0;
    const int mode = grad ? MODE_GRAD : MODE_COST;
    PlanScope keep(c);
    int first_fit = 0;
    for (int rho = 1; rho <= sqgpu_ctx::MAX_RHO; ++rho) {
        if (!c->cl_ok[rho - 1]) continue;
        c->P = &c->planC[rho - 1];
        const FusedPlan p = plan_fused(c, mode, c->rows >> rho, c->cols, batch, 2, false, rho);
        if (!p.ok) continue;
        if (!first_fit) first_fit = rho;
        if ((p.smem + 1024) * 2 <= (size_t)c->smem_per_sm) return rho;
    }
    return first_fit;
}

// ---- building blocks (all enqueue on `st`, no host sync) --------------------------------------------------------

static const int SPLIT_TABLES_MAX_BATCH = 8;
int run_tables(sqgpu_ctx* c, const double* d_params, int batch, bool with_deriv, cudaStream_t st) {
    int rc;
    if ((rc = c->P->wKtab.ensure(std::max<size_t>(1, (size_t)batch * c->P->kern_total) * sizeof(cplx)))) return rc;
    if ((rc = c->P->wDKtab.ensure(std::max<size_t>(1, (size_t)batch * c->P->dkern_total) * sizeof(cplx)))) return rc;
    const long long total = (long long)batch * c->P->n_ops;
    if (total == 0) return SQGPU_OK;
    // small batches: the table build is on the critical path of the evaluation, the member-parallel second kernel shortens it;
    // large batches fill the device with one warp per (set, op) and the extra suffix products would only add work
    const bool split = with_deriv && c->n_params > 0 && !c->P->members.empty() &&
                       (c->opt.split_tables == 2 || (c->opt.split_tables == 1 && batch <= SPLIT_TABLES_MAX_BATCH));
    // one warp per (parameter set, op)
    build_kernel_tables<<<(unsigned)((total + TABLE_WARPS - 1) / TABLE_WARPS), TABLE_WARPS * 32, 0, st>>>(
        c->P->dOps.as<DevOp>(), c->P->n_ops, c->P->dMembers.as<DevMember>(), d_params, c->n_params, batch, c->dPool.as<cplx>(),
        c->P->wKtab.as<cplx>(), c->P->kern_total, c->P->wDKtab.as<cplx>(), c->P->dkern_total, with_deriv ? (split ? 2 : 1) : 0, c->table_shift);
    c->launches++;
    if (split) {
        // derivative kernels of the fused blocks: the prefixes are in place, one warp per member finishes them
        const long long ctas = (long long)batch * c->P->n_ops;
        build_block_derivs<<<(unsigned)ctas, DERIV_WARPS * 32, 0, st>>>(c->P->dOps.as<DevOp>(), c->P->n_ops, c->P->dMembers.as<DevMember>(), d_params,
                                                                          c->n_params, c->dPool.as<cplx>(), c->P->wDKtab.as<cplx>(), c->P->dkern_total, c->table_shift);
        c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return SQGPU_OK;
}

// lookup tables of the DMMA block path for `ysets` parameter sets (after run_tables), for tile width 2^log_ct
int run_optabs(sqgpu_ctx* c, int ysets, int log_ct, cudaStream_t st) {
    if (c->P->n_ops == 0 || ysets == 0) return SQGPU_OK;
    int rc;
    if ((rc = c->P->wOpTab.ensure((size_t)ysets * c->P->n_ops * sizeof(OpTab)))) return rc;
    build_optabs<<<ysets * c->P->n_ops, 128, 0, st>>>(c->P->dOps.as<DevOp>(), c->P->n_ops, c->P->wKtab.as<cplx>(), c->P->kern_total,
                                                       c->dPool.as<cplx>(), log_ct, c->P->wOpTab.as<OpTab>());
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return SQGPU_OK;
}

// fragment tables of the raw dense 3-/4-qubit ops: constants of (circuit, tile width), rebuilt only when the width changes
int run_dense_tabs(sqgpu_ctx* c, int log_ct, cudaStream_t st) {
    Plan* P = c->P;
    if ((P->n_dense == 0 && P->n_dense5 == 0) || P->dense_logct == log_ct) return SQGPU_OK;
    int rc;
    if ((rc = P->wDenseTab.ensure(std::max<size_t>(1, (size_t)3 * P->n_dense) * sizeof(DenseTab)))) return rc;  // K, K^dagger, K^T per op
    if ((rc = P->wDenseTab5.ensure(std::max<size_t>(1, (size_t)P->n_dense5) * sizeof(DenseTab5)))) return rc;
    const DevOp* ops = P->dOps.as<DevOp>();
    const cplx* pool = c->dPool.as<cplx>();
    DenseTab* tabs = P->wDenseTab.as<DenseTab>();
    DenseTab5* tabs5 = P->wDenseTab5.as<DenseTab5>();
    switch (log_ct) {
        case 0: build_dense_tabs<0><<<P->n_ops, 128, 0, st>>>(ops, P->n_ops, pool, tabs, tabs5); break;
        case 1: build_dense_tabs<1><<<P->n_ops, 128, 0, st>>>(ops, P->n_ops, pool, tabs, tabs5); break;
        case 2: build_dense_tabs<2><<<P->n_ops, 128, 0, st>>>(ops, P->n_ops, pool, tabs, tabs5); break;
        default: build_dense_tabs<3><<<P->n_ops, 128, 0, st>>>(ops, P->n_ops, pool, tabs, tabs5); break;
    }
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    P->dense_logct = log_ct;
    return SQGPU_OK;
}

void fill_common_args(const sqgpu_ctx* c, const FusedPlan& p, ExecArgs& a, int rows, int cols) {
    memset(&a, 0, sizeof(a));
    a.rows = rows;
    a.cols = cols;
    a.n = c->qbit_num;
    a.ct = 1 << p.log_ct;
    a.log_ct = p.log_ct;
    a.tiles = p.tiles;
    a.tiles_per_cta = p.tiles_per_cta;
    a.ops = c->P->dOps.as<DevOp>();
    a.n_ops = c->P->n_ops;
    a.ktab = c->P->wKtab.as<cplx>();
    a.kern_total = c->P->kern_total;
    a.dktab = c->P->wDKtab.as<cplx>();
    a.dkern_total = c->P->dkern_total;
    a.pool = c->dPool.as<cplx>();
    a.optabs = c->P->wOpTab.as<OpTab>();
    a.dense_tabs = (c->P->n_dense > 0 && c->P->dense_logct == p.log_ct) ? c->P->wDenseTab.as<DenseTab>() : nullptr;
    a.dense_tabs5 = (c->P->n_dense5 > 0 && c->P->dense_logct == p.log_ct) ? c->P->wDenseTab5.as<DenseTab5>() : nullptr;
    a.dense_stage = c->P->dense_stage;
    a.wmax = c->P->wmax;
    a.w_total = c->P->w_total;
    a.w_in_smem = p.w_in_smem ? 1 : 0;
    a.w_direct = p.w_direct ? 1 : 0;
    a.dbuf = p.dbuf ? 1 : 0;
    a.rho = p.rho;
    a.dns = c->P->has_dense ? 1 : 0;
    if (p.rho > 0) memcpy(a.fin_pos, c->cl_fin[p.rho - 1], sizeof(a.fin_pos));
}

void time_begin(sqgpu_ctx* c, const char* name, cudaStream_t st) {
    if (c->capturing) return;
    KernelTimer::Ring* r = c->timer.cur = c->timer.get(name);
    cudaEventRecord(r->e0[r->n % KernelTimer::RING], st);
}
void time_end(sqgpu_ctx* c, cudaStream_t st) {
    KernelTimer::Ring* r = c->timer.cur;
    if (!r || c->capturing) return;
    cudaEventRecord(r->e1[r->n % KernelTimer::RING], st);
    r->n++;
}

int run_exec_streaming(sqgpu_ctx* c, int batch, bool grad, const cplx* d_omega, double* d_traces, cudaStream_t st);
int run_exec_tall_window(sqgpu_ctx* c, int batch, bool grad, const cplx* d_omega, double* d_traces, cudaStream_t st);
bool tall_window_fits(sqgpu_ctx* c, bool grad);
void exec_flops(const Plan& P, int rows, int cols, int log_ct, bool grad, int ysets, double* tensor, double* scalar);
StreamGate make_stream_gate(const DevOp& op, cplx* data, long long ystride, int rows, int cols, int ld, const cplx* K, long long k_ystride);
int launch_stream_gate(sqgpu_ctx* c, const DevOp& op, bool deriv, cplx* data, long long ystride, int ysets, int rows, int cols,
                       int ld, const cplx* K, long long k_ystride, cudaStream_t st);

// one executor pass over the resident matrix for `batch` parameter sets whose kernel tables are already built:
// fills wTrPart (and wWPart), then reduces into d_traces[batch][1+P or 1][3][2].
int run_exec_resident(sqgpu_ctx* c, int batch, bool grad, const cplx* d_omega, double* d_traces, cudaStream_t st) {
    const int mode = grad ? MODE_GRAD : MODE_COST;
    FusedPlan p = plan_eval(c, mode, batch);
    if (!p.ok)  // column too tall for shared memory: windowed executor (planW) or one op per launch (plan2)
        return c->P == &c->planW ? run_exec_tall_window(c, batch, grad, d_omega, d_traces, st) : run_exec_streaming(c, batch, grad, d_omega, d_traces, st);
    int rc;
    if ((rc = run_optabs(c, batch, p.log_ct, st))) return rc;
    if ((rc = run_dense_tabs(c, p.log_ct, st))) return rc;
    if ((rc = c->wTrPart.ensure((size_t)batch * p.chunks * 6 * sizeof(double)))) return rc;
    if (grad && (rc = c->wWPart.ensure(std::max<size_t>(1, (size_t)batch * p.w_slices * c->P->w_total) * sizeof(cplx)))) return rc;
    ExecArgs a;
    fill_common_args(c, p, a, c->rows >> p.rho, c->cols);
    a.in = c->U.as<cplx>();
    a.in_ystride = 0;
    a.ld_in = c->cols;
    a.trace_offset = effective_offset(c);
    a.n_trace_types = n_trace_types_of(c->cfg.variant);
    a.sum_sq = c->cfg.variant == SQGPU_SUM_OF_SQUARES ? 1 : 0;
    a.tr_part = c->wTrPart.as<double>();
    a.w_part = c->wWPart.as<cplx>();
    a.omega = d_omega;
    if (grad && !p.w_in_smem && c->P->w_total > 0)
        CUDA_TRY(cudaMemsetAsync(c->wWPart.p, 0, (size_t)batch * p.w_slices * c->P->w_total * sizeof(cplx), st));
    c->last_shape[0] = p.log_ct; c->last_shape[1] = p.threads; c->last_shape[2] = p.chunks; c->last_shape[3] = p.tiles_per_cta;
    c->last_shape[4] = (int)p.smem; c->last_shape[5] = 1 << p.rho;
    exec_flops(*c->P, c->rows, c->cols, p.log_ct, grad, batch, &c->last_flops[0], &c->last_flops[1]);
    time_begin(c, grad ? "fused_exec<GRAD>" : "fused_exec<COST>", st);
    cudaError_t e = grad ? launch_fused_mode<MODE_GRAD>(a, p, batch, st) : launch_fused_mode<MODE_COST>(a, p, batch, st);
    time_end(c, st);
    c->launches++;
    if (e != cudaSuccess) return fail(SQGPU_ERR_CUDA, "fused_exec launch failed: %s", cudaGetErrorString(e));
    if (grad && p.w_slices > 1 && c->P->w_total > 0) {
        fold_w_chunks<<<dim3(fold_grid_x(c->P->w_total), batch), 256, 0, st>>>(c->wWPart.as<cplx>(), p.w_slices, c->P->w_total);
        c->launches++;
    }
    reduce_partials<<<dim3(batch, reduce_grid_y(c->n_params)), 128, 0, st>>>(c->wTrPart.as<double>(), p.chunks, c->wWPart.as<cplx>(), c->P->w_total,
                                           c->P->dOps.as<DevOp>(), c->P->dParamOp.as<int>(), c->P->dParamOp.as<int>() + std::max(c->n_params, 1),
                                           c->P->wDKtab.as<cplx>(), c->P->dkern_total, c->P->wKtab.as<cplx>(), c->P->kern_total, c->n_params, grad ? 1 : 0, d_traces, 1,
                                           p.w_slices);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    c->replay.valid = grad;
    c->replay.batch = batch;
    c->replay.chunks = p.chunks;
    c->replay.w_slices = p.w_slices;
    c->replay.plan = c->P;
    return SQGPU_OK;
}

// FP64 flops the fused executor ISSUES for one launch over `ysets` parameter sets (a model of its instruction stream, to be
// checked against ncu's sm__ops_path_tensor_src_fp64): a DMMA m8n8k4 is 512 flops; per batch of 8 (group, column) items a
// 3-qubit block issues 6 DMMA forward and 12 (+ 6 for W') backward in the 3M formulation (8 and 16 (+ 8) in the real
// embedding, SQ_BLOCK_3M = 0), a 2-qubit block 2 and 4 (+ 2); the scalar paths issue
// 4 complex multiply-adds (32 flops) per row pair forward and 8 (+ 4 for W) backward; raw dense 2^k kernels (2^k/4)^2 * 2
// DMMA per batch (3-5 qubits) forward.
void exec_flops(const Plan& P, int rows, int cols, int log_ct, bool grad, int ysets, double* tensor, double* scalar) {
    double t = 0, sc = 0;
    const double colsd = (double)cols;
    for (const DevOp& op : P.ops) {
        if (op.type == SQ_OP_RESPLIT) continue;  // data exchange between the CTAs of a cluster: no arithmetic
        const bool has_w = op.w_off >= 0;
        if (op.dim == 2) {
            const double pairs = (double)(rows >> (op.ctrl_mask ? op.nfix : 1)) * colsd;
            sc += pairs * (32.0 + (grad ? 64.0 + (has_w ? 32.0 : 0.0) : 0.0));
            continue;
        }
        const double items = (double)(rows >> op.nq) * colsd;
        const bool dmma_block = op.ctrl_mask == 0 && (op.dim == 4 || op.dim == 8) && ((((rows >> op.nq) << log_ct) & 7) == 0);
        if (dmma_block) {
            double per8;  // DMMA per batch of 8 items
            if (op.dim == 8 && SQ_BLOCK_3M) {
                // three real products per complex product: 6 DMMA forward, 12 (+ 6 for W') backward, and per lane and batch
                // 2 DADD forward, 4 (+ 4) backward for the operand sums
                per8 = 6.0 + (grad ? 12.0 + (has_w ? 6.0 : 0.0) : 0.0);
                sc += items / 8.0 * 32.0 * (2.0 + (grad ? 4.0 + (has_w ? 4.0 : 0.0) : 0.0));  // the operand sums u + v
            } else {
                per8 = op.dim == 8 ? (8.0 + (grad ? 16.0 + (has_w ? 8.0 : 0.0) : 0.0)) : (2.0 + (grad ? 4.0 + (has_w ? 2.0 : 0.0) : 0.0));
            }
            t += items / 8.0 * per8 * 512.0;
        } else if (op.ctrl_mask == 0 && op.nq >= 3 && op.type == SQGPU_GENERAL && (!grad || op.dtab > 0) && ((((rows >> op.nq) << log_ct) & 7) == 0)) {
            const double nt = op.dim / 4.0;
            const double passes = grad ? 3.0 : 1.0;  // forward, and in the adjoint sweep K^dagger a and K^T beta (constant kernels: no W')
            if ((op.nq <= 4 && SQ_DENSE_3M) || (op.nq == 5 && SQ_DENSE5_3M)) {  // three-product form: 3 x (dim / 8) x (dim / 4) DMMA and dim / 4 DADD per batch
                t += passes * items / 8.0 * (nt * nt * 1.5) * 512.0;
                sc += passes * items / 8.0 * 32.0 * nt;
            } else {
                t += passes * items / 8.0 * (nt * nt * 2.0) * 512.0;
            }
        } else {
            const double act = items / (double)(1 << popcount32(op.ctrl_mask));
            sc += act * 8.0 * op.dim * op.dim * (grad ? 3.0 : 1.0);
        }
    }
    *tensor = t * ysets;
    *scalar = sc * ysets;
}

int check_ready(const sqgpu_ctx* c, bool need_matrix) {
    if (!c->circuit_set) return fail(SQGPU_ERR_STATE, "no circuit set (call sqgpu_set_circuit first)");
    if (need_matrix) {
        if (!c->U.p) return fail(SQGPU_ERR_STATE, "no matrix uploaded (call sqgpu_upload_matrix first)");
        if (c->rows != (1 << c->qbit_num))
            return fail(SQGPU_ERR_INVALID, "Wrong matrix size: the circuit has %d qubits, the matrix %d rows", c->qbit_num, c->rows);
    }
    return SQGPU_OK;
}

// max parameter sets per executor launch so that the W partials stay below ~1.5 GiB
int batch_slice(const sqgpu_ctx* c, int batch, bool grad) {
    {
        FusedPlan pf = plan_eval(c, grad ? MODE_GRAD : MODE_COST, batch);
        if (!pf.ok) return std::min(batch, 32);  // streaming fallback: bound the replicated chunk workspace
    }
    if (!grad || c->P->w_total == 0) return std::min(batch, 65535);
    // a smaller slice is planned with more chunks per parameter set (the grid is filled either way): iterate to a fixed point
    const size_t lim = (size_t)(SQ_W_DIRECT ? 8192 : 1536) << 20;  // B200: 180 GB of HBM
    int slice = std::min(batch, 65535);
    for (int it = 0; it < 16; ++it) {
        const FusedPlan p = plan_eval(c, MODE_GRAD, slice);
        const size_t per = (size_t)std::max(1, p.w_slices) * c->P->w_total * sizeof(cplx);
        const int fit = (int)std::max<size_t>(1, std::min<size_t>((size_t)slice, lim / std::max<size_t>(per, 1)));
        if (fit >= slice) break;
        slice = fit;
    }
    return slice;
}

// traces for a batch (device pointers). Layout d_traces[batch][n_k][3][2], n_k = 1 + (with_grad ? P : 0).
// d_global_traces0 (optional, [batch][3][2]): the traces of the circuit itself already summed over all column shards -- the
// Hilbert-Schmidt correction variants take the weights of their gradient functional from them (omega_t = w_t conj(T_t))
int traces_dev(sqgpu_ctx* c, const double* d_params, int batch, bool with_grad, double* d_traces, cudaStream_t st,
               bool allow_two_pass, const double* d_global_traces0 = nullptr) {
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (batch <= 0) return SQGPU_OK;
    // three-qubit blocks for the shared-memory executor, two-qubit blocks for the streaming fallback
    c->P = &c->plan3;
    c->use_rho = 0;
    {
        // Which executor (measured, profiles/README_r2.md): the single-CTA one wherever a column fits twice per SM; where it
        // fits only once per SM (n = 12 gradient) a cluster of two CTAs with half the rows each (+2...6 %); where it does not
        // fit at all (gradient n >= 13, cost n >= 14) the windowed executor, which beats the clusters from two adaptive levels
        // on (n = 13: 180 against 172 evals/s, n = 14: 299 against 235) -- clusters there only on request (option cluster = 2)
        // or when the circuit has no window plan; last the one-op-per-launch streaming kernels.
        const FusedPlan p0 = plan_fused(c, with_grad ? MODE_GRAD : MODE_COST, c->rows, c->cols, batch);
        int rho = 0;
        if (c->opt.cluster) {
            const bool one_per_sm = p0.ok && (p0.smem + 1024) * 2 > (size_t)c->smem_per_sm;
            if (p0.ok ? (one_per_sm && c->qbit_num >= 12) : (c->opt.cluster == 2 || !tall_window_fits(c, with_grad))) rho = pick_cluster_rho(c, with_grad, batch);
        }
        if (rho > 0) {  // thread-block clusters share the column over distributed shared memory
            c->P = &c->planC[rho - 1];
            c->use_rho = rho;
        } else if (!p0.ok) {
            c->P = tall_window_fits(c, with_grad) ? &c->planW : &c->plan2;
        }
    }
    if (c->cols + effective_offset(c) > c->rows) return fail(SQGPU_ERR_INVALID, "trace_offset %d + cols %d exceeds rows %d", effective_offset(c), c->cols, c->rows);
    if (with_grad && !c->all_unitary) return fail(SQGPU_ERR_UNSUPPORTED, "gradient with a non-unitary GENERAL gate is not supported (the adjoint sweep needs K^-1 = K^dagger)");
    const bool hs_corr = c->cfg.variant == SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1 || c->cfg.variant == SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2;
    if (with_grad && hs_corr && !allow_two_pass && !d_global_traces0)
        return fail(SQGPU_ERR_UNSUPPORTED, "raw gradient traces for the Hilbert-Schmidt correction variants need the globally summed traces first");
    const int n_k = 1 + (with_grad ? c->n_params : 0);
    const int slice = batch_slice(c, batch, with_grad);
    for (int b0 = 0; b0 < batch; b0 += slice) {
        const int nb = std::min(slice, batch - b0);
        const double* dp = d_params + (size_t)b0 * c->n_params;
        double* dt = d_traces + (size_t)b0 * n_k * 6;
        if ((rc = run_tables(c, dp, nb, with_grad, st))) return rc;
        if (!with_grad) {
            if ((rc = run_exec_resident(c, nb, false, nullptr, dt, st))) return rc;
            continue;
        }
        if ((rc = c->wOmega.ensure((size_t)nb * 3 * sizeof(cplx)))) return rc;
        const double* tr_for_omega = nullptr;
        int nk_for_omega = 1;
        if (hs_corr && d_global_traces0) {
            tr_for_omega = d_global_traces0 + (size_t)b0 * 6;
        } else if (hs_corr) {  // pass 1: traces of the circuit itself, pass 2 uses omega_t = w_t conj(T_t)
            if ((rc = c->wTraces0.ensure((size_t)nb * 6 * sizeof(double)))) return rc;
            if ((rc = run_exec_resident(c, nb, false, nullptr, c->wTraces0.as<double>(), st))) return rc;
            tr_for_omega = c->wTraces0.as<double>();
        }
        make_omega<<<(nb + 127) / 128, 128, 0, st>>>(tr_for_omega, nk_for_omega, c->cfg, nb, c->wOmega.as<cplx>());
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        if ((rc = run_exec_resident(c, nb, true, c->wOmega.as<cplx>(), dt, st))) return rc;
    }
    return SQGPU_OK;
}

int cost_from_traces_dev(sqgpu_ctx* c, const double* d_traces, int batch, bool with_grad, int cols_total, double* d_cost,
                         double* d_grad, cudaStream_t st, bool shifted = false) {
    if (batch <= 0) return SQGPU_OK;
    if (!variant_supported(c->cfg.variant)) return fail(SQGPU_ERR_UNSUPPORTED, "cost function variant %d is not supported on the device path", c->cfg.variant);
    for (int b0 = 0; b0 < batch; b0 += 65535) {
        const int nb = std::min(65535, batch - b0);
        const int n_k = 1 + (with_grad ? c->n_params : 0);
        cost_from_traces<<<nb, 128, 0, st>>>(d_traces + (size_t)b0 * n_k * 6, c->n_params, with_grad ? (shifted ? 2 : 1) : 0, cols_total, c->cfg,
                                             d_cost ? d_cost + b0 : nullptr, d_grad ? d_grad + (size_t)b0 * c->n_params : nullptr);
        c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return SQGPU_OK;
}

int eval_dev(sqgpu_ctx* c, const double* d_params, int batch, bool with_grad, double* d_cost, double* d_grad, cudaStream_t st) {
    if (!variant_supported(c->cfg.variant)) return fail(SQGPU_ERR_UNSUPPORTED, "cost function variant %d is not supported on the device path", c->cfg.variant);
    int rc = check_ready(c, true);
    if (rc) return rc;
    const int n_k = 1 + (with_grad ? c->n_params : 0);
    if ((rc = c->wTraces.ensure(std::max<size_t>(1, (size_t)batch * n_k * 6) * sizeof(double)))) return rc;
    if ((rc = traces_dev(c, d_params, batch, with_grad, c->wTraces.as<double>(), st, true))) return rc;
    return cost_from_traces_dev(c, c->wTraces.as<double>(), batch, with_grad, c->cols, d_cost, d_grad, st);  // a single, unsharded handle
}

// cost at theta + shifts[s] e_p for every parameter p of every set, from ONE adjoint sweep per set (see slot_kernel, gate_kernels.cuh).
// The sweep does not depend on the shift: further shifts re-build the tables and repeat reduce_partials on the W' partials of
// the first one (resident fused executor, batch in one slice; otherwise one sweep per shift).
int shifted_eval_dev(sqgpu_ctx* c, const double* d_params, int batch, const double* shifts, int n_shifts, double* d_cost0, double* d_shifted,
                     cudaStream_t st) {
    const int v = c->cfg.variant;
    if (v != SQGPU_FROBENIUS_NORM && v != SQGPU_FROBENIUS_NORM_CORRECTION1 && v != SQGPU_FROBENIUS_NORM_CORRECTION2 &&
        v != SQGPU_HILBERT_SCHMIDT_TEST && v != SQGPU_INFIDELITY)
        return fail(SQGPU_ERR_UNSUPPORTED, "shifted costs from one sweep need a cost that is a function of one linear trace functional (variants 0, 1, 2, 3, 9), not variant %d", v);
    for (int s = 0; s < n_shifts; ++s)
        if (shifts[s] == 0.0) return fail(SQGPU_ERR_INVALID, "shift must not be 0");
    int rc = check_ready(c, true);
    if (rc) return rc;
    const int n_k = 1 + c->n_params;
    if ((rc = c->wTraces.ensure(std::max<size_t>(1, (size_t)batch * n_k * 6) * sizeof(double)))) return rc;
    c->replay.valid = false;
    for (int s = 0; s < n_shifts && !rc; ++s) {
        c->table_shift = shifts[s];
        if (s > 0 && c->replay.valid && c->replay.batch == batch && c->replay.plan == c->P && batch_slice(c, batch, true) >= batch) {
            if (!(rc = run_tables(c, d_params, batch, true, st))) {
                reduce_partials<<<dim3(batch, reduce_grid_y(c->n_params)), 128, 0, st>>>(
                    c->wTrPart.as<double>(), c->replay.chunks, c->wWPart.as<cplx>(), c->P->w_total, c->P->dOps.as<DevOp>(), c->P->dParamOp.as<int>(),
                    c->P->dParamOp.as<int>() + std::max(c->n_params, 1), c->P->wDKtab.as<cplx>(), c->P->dkern_total, c->P->wKtab.as<cplx>(),
                    c->P->kern_total, c->n_params, 1, c->wTraces.as<double>(), 1, c->replay.w_slices);
                c->launches++;
                if (cudaGetLastError() != cudaSuccess) rc = fail(SQGPU_ERR_CUDA, "reduce_partials launch failed");
            }
        } else {
            c->replay.valid = false;
            rc = traces_dev(c, d_params, batch, true, c->wTraces.as<double>(), st, true);
            if (batch_slice(c, batch, true) < batch) c->replay.valid = false;  // the partials hold the last slice only
        }
        c->table_shift = 0.0;
        if (!rc) rc = cost_from_traces_dev(c, c->wTraces.as<double>(), batch, true, c->cols, d_cost0, d_shifted + (size_t)s * batch * c->n_params, st, true);
    }
    c->table_shift = 0.0;
    c->replay.valid = false;
    return rc;
}

// ---- apply paths -------------------------------------------------------------------------------------------------

StreamGate make_stream_gate(const DevOp& op, cplx* data, long long ystride, int rows, int cols, int ld, const cplx* K, long long k_ystride) {
    StreamGate g;
    memset(&g, 0, sizeof(g));
    g.data = data;
    g.ystride = ystride;
    g.rows = rows;
    g.cols = cols;
    g.ld = ld;
    g.log_cols = -1;
    for (int l = 0; l < 31; ++l)
        if ((1 << l) == cols) g.log_cols = l;
    g.target = op.target;
    g.ctrl_mask = op.ctrl_mask;
    g.nfix = op.nfix;
    for (int f = 0; f < 6; ++f) g.fix[f] = op.fix[f];
    g.K = K;
    g.k_ystride = k_ystride;
    g.nq = op.nq;
    for (int j = 0; j < 5; ++j) g.q[j] = op.q[j];
    return g;
}

// one gate on a device matrix with the streaming kernels
int launch_stream_gate(sqgpu_ctx* c, const DevOp& op, bool deriv, cplx* data, long long ystride, int ysets, int rows, int cols,
                       int ld, const cplx* K, long long k_ystride, cudaStream_t st) {
    StreamGate g = make_stream_gate(op, data, ystride, rows, cols, ld, K, k_ystride);
    if (op.dim == 2) {
        const long long items = (long long)(rows >> (deriv ? 1 : g.nfix)) * cols;
        const int thr = 256;
        // 4 items per thread and grid-stride step; cap the grid at 8 CTAs per SM worth of waves x 4
        const unsigned blocks = (unsigned)std::min<long long>((items + 4 * thr - 1) / (4 * thr), (long long)c->sm_count * 32);
        dim3 grid(std::max(1u, blocks), ysets);
        if (deriv) gate1q_stream<true><<<grid, thr, 0, st>>>(g);
        else gate1q_stream<false><<<grid, thr, 0, st>>>(g);
    } else {
        const long long items = (long long)(rows >> op.nq) * cols;
        const int thr = 128;
        const unsigned blocks = (unsigned)std::min<long long>((items + thr - 1) / thr, (long long)c->sm_count * 32);
        dim3 grid(std::max(1u, blocks), ysets);
#define SQ_KQ(KQ)                                                            \
    case KQ:                                                                 \
        if (deriv) gatekq_stream<KQ, true><<<grid, thr, 0, st>>>(g);         \
        else gatekq_stream<KQ, false><<<grid, thr, 0, st>>>(g);              \
        break;
        switch (op.nq) {
            SQ_KQ(2)
            SQ_KQ(3)
            SQ_KQ(4)
            SQ_KQ(5)
            default: return fail(SQGPU_ERR_INVALID, "dense gate on %d qubits", op.nq);
        }
#undef SQ_KQ
    }
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return SQGPU_OK;
}

// apply the whole program (kernel-table set 0) to `ysets` device matrices; deriv_op[y] (host) >= 0 selects the op whose
// derivative kernel (parameter deriv_p[y]) replaces the forward one in matrix y.
int apply_program_dev(sqgpu_ctx* c, const cplx* d_in, long long in_ystride, cplx* d_out, long long out_ystride, int ysets,
                      int rows, int cols, const std::vector<int>* deriv_op, const std::vector<int>* deriv_p, cudaStream_t st) {
    int rc;
    FusedPlan p = plan_fused(c, MODE_APPLY, rows, cols, ysets);
    if (p.ok) {
        if ((rc = run_optabs(c, 1, p.log_ct, st))) return rc;
        if ((rc = run_dense_tabs(c, p.log_ct, st))) return rc;
        const int* d_dop = nullptr;
        const int* d_dp = nullptr;
        if (deriv_op) {
            if ((rc = c->wDerivIdx.ensure((size_t)2 * ysets * sizeof(int)))) return rc;
            CUDA_TRY(cudaMemcpyAsync(c->wDerivIdx.p, deriv_op->data(), ysets * sizeof(int), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(c->wDerivIdx.as<int>() + ysets, deriv_p->data(), ysets * sizeof(int), cudaMemcpyHostToDevice, st));
            d_dop = c->wDerivIdx.as<int>();
            d_dp = c->wDerivIdx.as<int>() + ysets;
        }
        ExecArgs a;
        fill_common_args(c, p, a, rows, cols);
        a.in = d_in;
        a.out = d_out;
        a.in_ystride = in_ystride;
        a.out_ystride = out_ystride;
        a.ld_in = cols;
        a.ld_out = cols;
        a.k_shared = 1;
        a.deriv_op = d_dop;
        a.deriv_slot = d_dp;
        if (d_dop) a.dns = 1;  // the derivative kernel of a fused block goes down the generic dense path
        for (int y0 = 0; y0 < ysets; y0 += 65535) {
            const int ny = std::min(65535, ysets - y0);
            ExecArgs b = a;
            b.in = d_in + (size_t)y0 * in_ystride;
            b.out = d_out + (size_t)y0 * out_ystride;
            if (d_dop) {
                b.deriv_op = d_dop + y0;
                b.deriv_slot = d_dp + y0;
            }
            time_begin(c, "fused_exec<APPLY>", st);
            cudaError_t e = launch_fused_mode<MODE_APPLY>(b, p, ny, st);
            time_end(c, st);
            c->launches++;
            if (e != cudaSuccess) return fail(SQGPU_ERR_CUDA, "fused_exec<APPLY> launch failed: %s", cudaGetErrorString(e));
        }
        return SQGPU_OK;
    }
    // streaming route: matrices too tall for shared memory. out[y] <- in[y] first, then gate by gate in place.
    const long long n_elem = (long long)rows * cols;
    for (int y = 0; y < ysets; ++y) {
        const cplx* src = d_in + (size_t)y * in_ystride;
        cplx* dst = d_out + (size_t)y * out_ystride;
        if (src != dst) CUDA_TRY(cudaMemcpyAsync(dst, src, n_elem * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    }
    for (int k = 0; k < c->P->n_ops; ++k) {
        const DevOp& op = c->P->ops[k];
        const cplx* Kf = op.kern_off >= 0 ? c->P->wKtab.as<cplx>() + op.kern_off : c->dPool.as<cplx>() + op.pool_off;
        if (!deriv_op) {
            if ((rc = launch_stream_gate(c, op, false, d_out, out_ystride, ysets, rows, cols, cols, Kf, 0, st))) return rc;
            continue;
        }
        // matrices whose derivative op is k use the derivative kernel; batch the rest in contiguous runs
        int y = 0;
        while (y < ysets) {
            const bool d = (*deriv_op)[y] == k;
            int y1 = y + 1;
            if (!d)
                while (y1 < ysets && (*deriv_op)[y1] != k) ++y1;
            const cplx* K = d ? c->P->wDKtab.as<cplx>() + op.dkern_off + (size_t)(*deriv_p)[y] * op.dim * op.dim : Kf;
            if ((rc = launch_stream_gate(c, op, d, d_out + (size_t)y * out_ystride, out_ystride, y1 - y, rows, cols, cols, K, 0, st))) return rc;
            y = y1;
        }
    }
    return SQGPU_OK;
}

// State vectors too long for one shared-memory tile (n >= 14): the windowed executor -- one pass over the state per
// SEGMENT of the window plan instead of one per gate (build_window_plan). Parameters are in c->wParams. Returns 1 when no
// window plan fits (caller falls back to one op per launch).
int apply_window_dev(sqgpu_ctx* c, cplx* d_inout, int rows, cudaStream_t st) {
    PlanScope keep(c);
    c->P = &c->planW;
    const int w = c->win_w, wr = 1 << w, wc = rows >> w;
    const FusedPlan pf = plan_fused(c, MODE_APPLY, wr, wc, 1, 4, true);
    if (!pf.ok || c->segs.empty()) return 1;
    int rc;
    if ((rc = run_tables(c, c->wParams.as<double>(), 1, false, st))) return rc;
    if ((rc = run_optabs(c, 1, pf.log_ct, st))) return rc;
    if ((rc = run_dense_tabs(c, pf.log_ct, st))) return rc;
    time_begin(c, "fused_exec<WINDOW_FWD>", st);
    for (const auto& sg : c->segs) {
        ExecArgs a;
        fill_common_args(c, pf, a, wr, wc);
        a.n = w;
        a.in = d_inout;
        a.out = d_inout;
        a.wmask = sg.wmask;
        a.ops += sg.begin;
        a.n_ops = sg.end - sg.begin;
        a.optabs += sg.begin;
        a.optab_stride = c->P->n_ops;
        a.k_shared = 1;
        cudaError_t e = launch_fused_mode<MODE_APPLY>(a, pf, 1, st);
        c->launches++;
        if (e != cudaSuccess) return fail(SQGPU_ERR_CUDA, "fused_exec launch failed: %s", cudaGetErrorString(e));
    }
    time_end(c, st);
    return SQGPU_OK;
}

// Matrices whose column does not fit shared memory (gradient n >= 13, cost n >= 14), first choice: the WINDOWED executor
// of the state-vector path. A chunk of 2^m columns of the matrix, stored row-major [2^n][2^m], is a vector over n + m index
// bits whose upper n bits are the qubits; the segments of the window plan (build_window_plan) apply to it with their masks
// shifted by m, the tile's columns being the values of all other bits (the non-window qubits AND the matrix columns). The
// chunk makes one HBM round trip per SEGMENT (fused_exec<MODE_APPLY>, then fused_exec<MODE_BWD> on (a, beta) in reverse)
// instead of one per gate as in run_exec_streaming below, and the ops are the 3-qubit DMMA blocks of plan3.
// tall_window_fits: the decision (before the kernel tables are built for the plan); c->P is left untouched.
bool tall_window_fits(sqgpu_ctx* c, bool grad) {
    if (!c->opt.tall_window || c->segs.empty() || c->cfg.variant == SQGPU_SUM_OF_SQUARES) return false;
    if (c->win_w >= c->qbit_num) return false;  // a single tile would hold the column: the resident executor's case
    PlanScope keep(c);
    c->P = &c->planW;
    const int wr = 1 << c->win_w, wc = 1 << (c->qbit_num - c->win_w);
    if (!plan_fused(c, MODE_APPLY, wr, wc, 1, 4, true).ok) return false;
    if (grad && !plan_fused(c, MODE_BWD, wr, wc, 1, 1).ok) return false;
    return true;
}

int run_exec_tall_window(sqgpu_ctx* c, int batch, bool grad, const cplx* d_omega, double* d_traces, cudaStream_t st) {
    const int rows = c->rows, cols = c->cols, n = c->qbit_num, w = c->win_w, wr = 1 << w;
    // chunk width: a power of two that divides cols (every chunk has the same shape), at most 64 MiB of column data per set
    int m = 0;
    while (((cols >> (m + 1)) << (m + 1)) == cols && ((size_t)rows << (m + 1)) * sizeof(cplx) <= ((size_t)64 << 20)) ++m;
    const int cw = 1 << m, nchunks = cols / cw;
    const size_t ce = (size_t)rows * cw;
    const int wc = (int)(ce >> w);
    const FusedPlan pf = plan_fused(c, MODE_APPLY, wr, wc, batch, 4, true);
    const FusedPlan pb = grad ? plan_fused(c, MODE_BWD, wr, wc, batch, 2) : pf;
    if (!pf.ok || !pb.ok) return fail(SQGPU_ERR_STATE, "windowed executor: no shared-memory plan for a %d-qubit window", w);
    int rc;
    if ((rc = c->wMat.ensure((size_t)(grad ? 2 : 1) * batch * ce * sizeof(cplx)))) return rc;
    if ((rc = c->wTrPart.ensure((size_t)batch * nchunks * 6 * sizeof(double)))) return rc;
    if (grad && (rc = c->wWPart.ensure(std::max<size_t>(1, (size_t)batch * pb.w_slices * c->P->w_total) * sizeof(cplx)))) return rc;
    cplx* A = c->wMat.as<cplx>();
    cplx* Bt = A + (size_t)batch * ce;
    double* tr_part = c->wTrPart.as<double>();
    const int ntt = n_trace_types_of(c->cfg.variant);
    const int off = effective_offset(c);
    auto seg_args = [&](const FusedPlan& p, const sqgpu_ctx::Segment& sg, ExecArgs& a) {
        fill_common_args(c, p, a, wr, wc);
        a.n = w;
        a.in = A;
        a.out = A;
        a.in_ystride = a.out_ystride = (long long)ce;
        a.wmask = sg.wmask << m;
        a.ops += sg.begin;
        a.n_ops = sg.end - sg.begin;
        a.optabs += sg.begin;
        a.optab_stride = c->P->n_ops;
    };
    if ((rc = run_optabs(c, batch, pf.log_ct, st))) return rc;
    if ((rc = run_dense_tabs(c, pf.log_ct, st))) return rc;
    int tab_log_ct = pf.log_ct;
    if (grad && c->P->w_total > 0) CUDA_TRY(cudaMemsetAsync(c->wWPart.p, 0, (size_t)batch * pb.w_slices * c->P->w_total * sizeof(cplx), st));
    c->last_shape[0] = pb.log_ct; c->last_shape[1] = pb.threads; c->last_shape[2] = pb.chunks; c->last_shape[3] = pb.tiles_per_cta;
    c->last_shape[4] = (int)pb.smem; c->last_shape[5] = 1;
    exec_flops(*c->P, rows, cols, pb.log_ct, grad, batch, &c->last_flops[0], &c->last_flops[1]);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int j0 = ch * cw;
        copy_chunk<<<dim3(std::min(c->sm_count * 8, std::max(1, (int)(ce / 256))), batch), 256, 0, st>>>(c->U.as<cplx>(), cols, j0, rows, cw, A);
        c->launches++;
        if (tab_log_ct != pf.log_ct) {
            if ((rc = run_optabs(c, batch, pf.log_ct, st))) return rc;
            if ((rc = run_dense_tabs(c, pf.log_ct, st))) return rc;
            tab_log_ct = pf.log_ct;
        }
        time_begin(c, "fused_exec<WINDOW_FWD>", st);
        for (const auto& sg : c->segs) {
            ExecArgs a;
            seg_args(pf, sg, a);
            cudaError_t e = launch_fused_mode<MODE_APPLY>(a, pf, batch, st);
            c->launches++;
            if (e != cudaSuccess) return fail(SQGPU_ERR_CUDA, "fused_exec launch failed: %s", cudaGetErrorString(e));
        }
        time_end(c, st);
        traces_stream<<<batch, 256, 0, st>>>(A, (long long)ce, cw, cw, n, off + j0, ntt, tr_part + (size_t)ch * 6, nchunks * 6);
        c->launches++;
        if (!grad) continue;
        CUDA_TRY(cudaMemsetAsync(Bt, 0, (size_t)batch * ce * sizeof(cplx), st));
        beta_init_stream<<<dim3((cw + 127) / 128, batch), 128, 0, st>>>(Bt, rows, cw, j0, n, off, ntt, d_omega);
        c->launches++;
        if (tab_log_ct != pb.log_ct) {
            if ((rc = run_optabs(c, batch, pb.log_ct, st))) return rc;
            if ((rc = run_dense_tabs(c, pb.log_ct, st))) return rc;
            tab_log_ct = pb.log_ct;
        }
        time_begin(c, "fused_exec<WINDOW_BWD>", st);
        for (int si = (int)c->segs.size() - 1; si >= 0; --si) {
            ExecArgs a;
            seg_args(pb, c->segs[si], a);
            a.beta = Bt;
            a.w_part = c->wWPart.as<cplx>();  // the CTAs' slices accumulate over segments and chunks (one writer per address per launch)
            cudaError_t e = launch_fused_mode<MODE_BWD>(a, pb, batch, st);
            c->launches++;
            if (e != cudaSuccess) return fail(SQGPU_ERR_CUDA, "fused_exec launch failed: %s", cudaGetErrorString(e));
        }
        time_end(c, st);
    }
    CUDA_TRY(cudaGetLastError());
    if (grad && pb.w_slices > 1 && c->P->w_total > 0) {
        fold_w_chunks<<<dim3(fold_grid_x(c->P->w_total), batch), 256, 0, st>>>(c->wWPart.as<cplx>(), pb.w_slices, c->P->w_total);
        c->launches++;
    }
    reduce_partials<<<dim3(batch, reduce_grid_y(c->n_params)), 128, 0, st>>>(tr_part, nchunks, c->wWPart.as<cplx>(), c->P->w_total, c->P->dOps.as<DevOp>(), c->P->dParamOp.as<int>(),
                                           c->P->dParamOp.as<int>() + std::max(c->n_params, 1), c->P->wDKtab.as<cplx>(), c->P->dkern_total,
                                           c->P->wKtab.as<cplx>(), c->P->kern_total, c->n_params, grad ? 1 : 0, d_traces, 1, pb.w_slices);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return SQGPU_OK;
}

// Fallback for matrices whose columns do not fit shared memory (n >= 13 gradient, n >= 14 cost): column chunks of the
// resident matrix are replicated per parameter set into an L2-sized workspace and the program runs one op per launch
// with the streaming kernels; partials have the same layout as the fused executor's, one "chunk" per column chunk.
int run_exec_streaming(sqgpu_ctx* c, int batch, bool grad, const cplx* d_omega, double* d_traces, cudaStream_t st) {
    const int rows = c->rows, cols = c->cols;
    if (c->cfg.variant == SQGPU_SUM_OF_SQUARES) return fail(SQGPU_ERR_UNSUPPORTED, "SUM_OF_SQUARES is not implemented on the streaming executor (n = %d)", c->qbit_num);
    if (grad)
        for (const DevOp& op : c->P->ops) {
            const bool ok = op.dim == 2 || (op.dim == 4 && op.nq == 2 && op.ctrl_mask == 0);
            if (!ok) return fail(SQGPU_ERR_UNSUPPORTED, "streaming gradient with 3+ qubit dense or controlled two-target gates is not implemented (n = %d)", c->qbit_num);
        }
    // chunk width: ~16 MiB of column data per parameter set, at least 1 column
    int cw = (int)std::max<long long>(1, std::min<long long>(cols, ((long long)16 << 20) / ((long long)rows * (long long)sizeof(cplx))));
    const int nchunks = (cols + cw - 1) / cw;
    const size_t chunk_elems = (size_t)rows * cw;
    const int nblk = std::min(c->sm_count * 8, std::max(1, (int)(chunk_elems / 1024)));
    int rc;
    if ((rc = c->wMat.ensure((size_t)(grad ? 2 : 1) * batch * chunk_elems * sizeof(cplx)))) return rc;
    if ((rc = c->wTrPart.ensure(((size_t)batch * nchunks * 6 + (size_t)batch * nblk * 32) * sizeof(double)))) return rc;
    if (grad && (rc = c->wWPart.ensure(std::max<size_t>(1, (size_t)batch * nchunks * c->P->w_total) * sizeof(cplx)))) return rc;
    cplx* A = c->wMat.as<cplx>();
    cplx* Bt = A + (size_t)batch * chunk_elems;
    double* tr_part = c->wTrPart.as<double>();
    double* wscratch = tr_part + (size_t)batch * nchunks * 6;
    const int ntt = n_trace_types_of(c->cfg.variant);
    const int off = effective_offset(c);
    time_begin(c, "gate_stream (chunked)", st);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int j0 = ch * cw, w = std::min(cw, cols - j0);
        const size_t ce = (size_t)rows * w;
        copy_chunk<<<dim3(std::min(c->sm_count * 8, std::max(1, (int)(ce / 256))), batch), 256, 0, st>>>(c->U.as<cplx>(), cols, j0, rows, w, A);
        c->launches++;
        for (int k = 0; k < c->P->n_ops; ++k) {
            const DevOp& op = c->P->ops[k];
            const cplx* K = op.kern_off >= 0 ? c->P->wKtab.as<cplx>() + op.kern_off : c->dPool.as<cplx>() + op.pool_off;
            const long long kst = op.kern_off >= 0 ? c->P->kern_total : 0;
            if ((rc = launch_stream_gate(c, op, false, A, (long long)ce, batch, rows, w, w, K, kst, st))) return rc;
        }
        // per-chunk traces land at tr_part[y][ch][6]: launch with out pointing at chunk ch and stride nchunks*6
        traces_stream<<<batch, 256, 0, st>>>(A, (long long)ce, w, w, c->qbit_num, off + j0, ntt, tr_part + (size_t)ch * 6, nchunks * 6);
        c->launches++;
        if (!grad) continue;
        CUDA_TRY(cudaMemsetAsync(Bt, 0, (size_t)batch * ce * sizeof(cplx), st));
        beta_init_stream<<<dim3((w + 127) / 128, batch), 128, 0, st>>>(Bt, rows, w, j0, c->qbit_num, off, ntt, d_omega);
        c->launches++;
        for (int k = c->P->n_ops - 1; k >= 0; --k) {
            const DevOp& op = c->P->ops[k];
            const cplx* K = op.kern_off >= 0 ? c->P->wKtab.as<cplx>() + op.kern_off : c->dPool.as<cplx>() + op.pool_off;
            const long long kst = op.kern_off >= 0 ? c->P->kern_total : 0;
            StreamGate g = make_stream_gate(op, A, (long long)ce, rows, w, w, K, kst);
            const int want_w = op.n_params > 0 ? 1 : 0;
            int blocks, width;
            if (op.dim == 2) {
                const long long items = (long long)(rows >> g.nfix) * w;
                blocks = (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, nblk));
                width = 8;
                adjoint1q_stream<<<dim3(blocks, batch), 256, 0, st>>>(g, Bt, (long long)ce, wscratch, want_w);
            } else {
                const long long items = (long long)(rows >> 2) * w;
                blocks = (int)std::max<long long>(1, std::min<long long>((items + 127) / 128, nblk));
                width = 32;
                adjoint2q_stream<<<dim3(blocks, batch), 128, 0, st>>>(g, Bt, (long long)ce, wscratch, want_w);
            }
            c->launches++;
            if (want_w) {
                sum_partials<<<batch, 32, 0, st>>>(wscratch, blocks, width, 1.0,
                                                   reinterpret_cast<double*>(c->wWPart.as<cplx>() + (size_t)ch * c->P->w_total + op.w_off),
                                                   2 * nchunks * c->P->w_total);
                c->launches++;
            }
        }
    }
    time_end(c, st);
    CUDA_TRY(cudaGetLastError());
    if (grad && nchunks > 1 && c->P->w_total > 0) {
        fold_w_chunks<<<dim3(fold_grid_x(c->P->w_total), batch), 256, 0, st>>>(c->wWPart.as<cplx>(), nchunks, c->P->w_total);
        c->launches++;
    }
    reduce_partials<<<dim3(batch, reduce_grid_y(c->n_params)), 128, 0, st>>>(tr_part, nchunks, c->wWPart.as<cplx>(), c->P->w_total, c->P->dOps.as<DevOp>(), c->P->dParamOp.as<int>(),
                                           c->P->dParamOp.as<int>() + std::max(c->n_params, 1), c->P->wDKtab.as<cplx>(), c->P->dkern_total,
                                           c->P->wKtab.as<cplx>(), c->P->kern_total, c->n_params, grad ? 1 : 0, d_traces, 1);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return SQGPU_OK;
}

int set_cost_checked(sqgpu_ctx* c, int variant, int trace_offset, double prev, double c1, double c2) {
    if (!variant_supported(variant)) return fail(SQGPU_ERR_UNSUPPORTED, "cost function variant %d is not supported on the device path", variant);
    if (trace_offset < 0) return fail(SQGPU_ERR_INVALID, "negative trace offset");
    if (prev < 0) return fail(SQGPU_ERR_INVALID, "negative previous cost function value");
    c->cfg.variant = variant;
    c->cfg.prev = prev;
    c->cfg.c1 = c1;
    c->cfg.c2 = c2;
    c->trace_offset = trace_offset;
    return SQGPU_OK;
}

}  // namespace

namespace {
void release_adam(sqgpu_ctx* c);
int multi_upload(sqgpu_ctx* front, const double* data, int rows, int cols, int stride);
int multi_eval(sqgpu_ctx* front, const double* params, int batch, bool with_grad, double* cost, double* grad);
int multi_vqe(sqgpu_ctx* front, const double* params, int batch, bool with_grad, double* energy, double* grad);
int multi_apply_cost(struct MultiGpu* m);
int multi_destroy(sqgpu_ctx* front);
int multi_set_circuit(sqgpu_ctx* front, const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool, int64_t pool_len);
int multi_set_cost(sqgpu_ctx* front, int variant, int trace_offset, double prev, double c1, double c2);
int multi_set_option(sqgpu_ctx* front, const char* name, int64_t value);
int multi_set_hamiltonian(sqgpu_ctx* front, int n_rows, int64_t nnz, const int32_t* indptr, const int32_t* indices, const double* values);
long long multi_launches(sqgpu_ctx* front);
sqgpu_ctx* multi_first(sqgpu_ctx* front);
}  // namespace

#define SQ_NOT_ON_MULTI(c) \
    if ((c) && (c)->multi) return fail(SQGPU_ERR_UNSUPPORTED, "%s is not available on a multi-device handle", __func__)

// =====================================================================================================================
// extern "C" entry points
// =====================================================================================================================

extern "C" {

const char* sqgpu_last_error(void) { return g_last_error.c_str(); }

int sqgpu_abi_version(void) { return SQGPU_ABI_VERSION; }

int sqgpu_device_count(int* count) {
    if (!count) return fail(SQGPU_ERR_INVALID, "count is NULL");
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SQGPU_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major >= 10) ++ok;
    }
    *count = ok;
    return SQGPU_OK;
}

int sqgpu_create(int device, sqgpu_handle_t* out) {
    if (!out) return fail(SQGPU_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SQGPU_ERR_NO_DEVICE, "no CUDA device available (%s); the sqgpu engine has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(SQGPU_ERR_NO_DEVICE, "device %d out of range (%d visible)", device, n);
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major < 10) return fail(SQGPU_ERR_NO_DEVICE, "device %d has compute capability %d.x; this library is built for sm_100a only", device, major);
    DeviceGuard guard(device);
    sqgpu_ctx* c = new sqgpu_ctx();
    c->device = device;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&c->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return fail(SQGPU_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    e = cudaEventCreateWithFlags(&c->last_done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaStreamDestroy(c->stream);
        delete c;
        return fail(SQGPU_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
    }
    *out = c;
    return SQGPU_OK;
}

int sqgpu_destroy(sqgpu_handle_t c) {
    if (!c) return SQGPU_OK;
    if (c->multi) return multi_destroy(c);
    {
        DeviceGuard guard(c->device);
        std::lock_guard<std::mutex> lk(c->mtx);
        wait_idle(c);
        DevBuf* bufs[] = {&c->U, &c->plan2.dOps, &c->plan2.dMembers, &c->plan2.dParamOp, &c->plan2.wKtab, &c->plan2.wDKtab, &c->plan2.wOpTab, &c->plan3.dOps, &c->plan3.dMembers, &c->plan3.dParamOp, &c->plan3.wKtab, &c->plan3.wDKtab, &c->plan3.wOpTab, &c->planW.dOps, &c->planW.dMembers, &c->planW.dParamOp, &c->planW.wKtab, &c->planW.wDKtab, &c->planW.wOpTab, &c->plan2.wDenseTab, &c->plan3.wDenseTab, &c->planW.wDenseTab, &c->plan2.wDenseTab5, &c->plan3.wDenseTab5, &c->planW.wDenseTab5, &c->dPool, &c->wParams, &c->wTrPart, &c->wWPart,
                          &c->wTraces, &c->wOmega, &c->wCost, &c->wGrad, &c->wMat, &c->wDerivIdx, &c->wTraces0,
                          &c->hIndptr, &c->hIndices, &c->hValues};
        for (DevBuf* b : bufs) b->release();
        release_adam(c);
        c->timer.destroy();
        if (c->last_done) cudaEventDestroy(c->last_done);
        cudaStreamDestroy(c->stream);
    }
    delete c;
    return SQGPU_OK;
}

int sqgpu_upload_matrix(sqgpu_handle_t c, const double* data, int rows, int cols, int stride) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (!data || rows <= 0 || cols <= 0 || stride < cols) return fail(SQGPU_ERR_INVALID, "bad matrix arguments");
    if (rows & (rows - 1)) return fail(SQGPU_ERR_INVALID, "rows must be a power of two, got %d", rows);
    if (cols > rows) return fail(SQGPU_ERR_INVALID, "cols (%d) cannot exceed rows (%d)", cols, rows);
    if (c->multi) {
        std::lock_guard<std::mutex> lk(c->mtx);
        return multi_upload(c, data, rows, cols, stride);
    }
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    wait_idle(c);
    int rc = c->U.ensure((size_t)rows * cols * sizeof(cplx));
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(c->U.p, (size_t)cols * sizeof(cplx), data, (size_t)stride * sizeof(cplx), (size_t)cols * sizeof(cplx),
                               rows, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->rows = rows;
    c->cols = cols;
    return SQGPU_OK;
}

// lowering + planning (host only) and, with `upload`, the transfer of the three plans to the device
// N3, constant sub-circuits (the reference multiplies gates out into <= 5-qubit kernels in
// Gates_block::apply_to's fusion rule, gates/Gates_block.cpp:632-681, and in squander/partitioning/partition.py:50-93): gates
// WITHOUT parameters whose joint support fits `max_q` qubits are multiplied out ON THE HOST, once per sqgpu_set_circuit, into
// one constant dense kernel that the executor runs as a raw GENERAL op on the tensor cores (dense_dmma_forward2 / 5). The
// grouping is the same dependency-aware first fit as the block planner's: a group starts at the first constant gate not yet
// placed and takes, in program order, every later constant gate that fits its qubit set and shares no qubit with a gate that
// was passed over. A group is only replaced when that pays by the flop model: by itself it would need MORE 3-qubit blocks
// than the dense kernel costs (a 2^k x 2^k kernel costs 2^(k-3) block units per amplitude), so the parametric layers of the
// decomposition circuits, whose CNOTs ride along in neighbouring blocks for free, are left alone.
// `pool` is the matrix pool of the circuit; new kernels are appended to it.
static void host_apply_const_gate(const DevOp& r, const std::vector<cplx>& pool, const int* qs, int nqs, std::vector<cplx>& M) {
    const int D = 1 << nqs;
    auto local_bit = [&](int q) { for (int j = 0; j < nqs; ++j) if (qs[j] == q) return j; return -1; };
    const int dim = r.dim, nq = r.dim == 2 ? 1 : r.nq;
    std::vector<cplx> K((size_t)std::max(dim * dim, 16));
    if (r.type == SQGPU_GENERAL) {
        for (int e = 0; e < dim * dim; ++e) K[e] = pool[(size_t)r.pool_off + e];
    } else {
        Trig t;
        memset(&t, 0, sizeof(t));
        build_gate_kernel(r.type, t, -1, K.data());
    }
    int pos[5] = {0, 0, 0, 0, 0};
    if (r.dim == 2) pos[0] = local_bit(r.target);
    else
        for (int j = 0; j < nq; ++j) pos[j] = local_bit(r.q[j]);
    unsigned tmask = 0, cmask = 0;
    for (int j = 0; j < nq; ++j) tmask |= 1u << pos[j];
    for (int q = 0; q < 30; ++q)
        if ((r.ctrl_mask >> q) & 1) cmask |= 1u << local_bit(q);
    std::vector<cplx> v(dim), o(dim);
    for (int col = 0; col < D; ++col)
        for (int base = 0; base < D; ++base) {
            if ((base & tmask) != 0 || ((unsigned)base & cmask) != cmask) continue;
            for (int l = 0; l < dim; ++l) {
                int row = base;
                for (int j = 0; j < nq; ++j) row |= ((l >> j) & 1) << pos[j];
                v[l] = M[(size_t)row * D + col];
            }
            for (int lo = 0; lo < dim; ++lo) {
                double re = 0, im = 0;
                for (int l = 0; l < dim; ++l) {
                    const cplx k = K[lo * dim + l];
                    re += k.x * v[l].x - k.y * v[l].y;
                    im += k.x * v[l].y + k.y * v[l].x;
                }
                o[lo] = cmake(re, im);
            }
            for (int l = 0; l < dim; ++l) {
                int row = base;
                for (int j = 0; j < nq; ++j) row |= ((l >> j) & 1) << pos[j];
                M[(size_t)row * D + col] = o[l];
            }
        }
}

static int fuse_constant_runs(const std::vector<DevOp>& raw, int qbit_num, int max_q, std::vector<cplx>& pool, std::vector<DevOp>& out) {
    const int N = (int)raw.size();
    out.clear();
    int n_fused = 0;
    auto is_const = [&](const DevOp& r) { return r.n_params == 0 && popcount32(support_mask(r)) <= max_q; };
    const unsigned all_qubits = qbit_num >= 32 ? 0xffffffffu : ((1u << qbit_num) - 1u);
    std::vector<char> placed(std::max(N, 1), 0);
    for (int first = 0; first < N; ++first) {
        if (placed[first]) continue;
        placed[first] = 1;
        if (max_q < 4 || !is_const(raw[first])) {
            out.push_back(raw[first]);
            continue;
        }
        std::vector<int> group(1, first);
        unsigned support = support_mask(raw[first]), blocked = 0;
        const int scan_end = std::min(N, first + 8192);
        for (int i = first + 1; i < scan_end; ++i) {
            if (placed[i]) continue;
            const unsigned sup = support_mask(raw[i]);
            if (is_const(raw[i]) && (sup & blocked) == 0 && popcount32(support | sup) <= max_q) {
                group.push_back(i);
                support |= sup;
            } else {
                blocked |= sup;
                if ((blocked & all_qubits) == all_qubits) break;
            }
        }
        const int nqs = popcount32(support);
        // 3-qubit blocks the group would need by itself (the block planner's first fit, restricted to the group)
        int blocks3 = 0;
        {
            std::vector<char> done(group.size(), 0);
            for (size_t a = 0; a < group.size(); ++a) {
                if (done[a]) continue;
                unsigned bs = 0, blk = 0;
                for (size_t b = a; b < group.size(); ++b) {
                    if (done[b]) continue;
                    const unsigned sup = support_mask(raw[group[b]]);
                    if ((sup & blk) == 0 && popcount32(bs | sup) <= 3) {
                        bs |= sup;
                        done[b] = 1;
                    } else {
                        blk |= sup;
                    }
                }
                ++blocks3;
            }
        }
        if (nqs < 4 || blocks3 <= (1 << (nqs - 3))) {  // not worth a dense kernel: leave the gates to the block planner
            out.push_back(raw[first]);
            continue;
        }
        int qs[5] = {0, 0, 0, 0, 0}, nq = 0;
        for (int q = 0; q < 30; ++q)
            if ((support >> q) & 1) qs[nq++] = q;
        const int D = 1 << nqs;
        std::vector<cplx> M((size_t)D * D, cmake(0, 0));
        for (int i = 0; i < D; ++i) M[(size_t)i * D + i] = cmake(1, 0);
        for (int gi : group) {
            host_apply_const_gate(raw[gi], pool, qs, nqs, M);
            placed[gi] = 1;
        }
        DevOp op;
        memset(&op, 0, sizeof(op));
        op.type = SQGPU_GENERAL;
        op.kern_off = op.dkern_off = op.w_off = -1;
        op.member_off = -1;
        op.dim = D;
        op.nq = nqs;
        for (int j = 0; j < nqs; ++j) op.q[j] = qs[j];
        op.pool_off = (int64_t)pool.size();
        pool.insert(pool.end(), M.begin(), M.end());
        fill_fix(op);
        out.push_back(op);
        ++n_fused;
    }
    return n_fused;
}

static int set_circuit_impl(sqgpu_ctx* c, const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num,
                            const double* matrix_pool, int64_t pool_len, bool upload) {
    if (n_gates < 0 || n_params < 0 || qbit_num < 1 || qbit_num > 30) return fail(SQGPU_ERR_INVALID, "bad circuit arguments");
    if (n_gates > 0 && !gates) return fail(SQGPU_ERR_INVALID, "gates is NULL");
    // 1. lower every descriptor to a raw op (validation happens here)
    std::vector<DevOp> raw(n_gates);
    bool all_unitary = true;
    std::vector<char> param_used(std::max(n_params, 1), 0);
    for (int i = 0; i < n_gates; ++i) {
        bool unitary = true;
        int rc = lower_gate(gates[i], qbit_num, matrix_pool, pool_len, &raw[i], &unitary);
        if (rc) return rc;
        all_unitary = all_unitary && unitary;
        const DevOp& op = raw[i];
        if (op.n_params > 0) {
            if (op.param_start < 0 || op.param_start + op.n_params > n_params)
                return fail(SQGPU_ERR_INVALID, "gate %d: parameters [%d, %d) outside the parameter vector of length %d", i,
                            op.param_start, op.param_start + op.n_params, n_params);
            for (int p = 0; p < op.n_params; ++p) {
                if (param_used[op.param_start + p]) return fail(SQGPU_ERR_INVALID, "parameter %d is used by two gates", op.param_start + p);
                param_used[op.param_start + p] = 1;
            }
        }
    }
    for (int p = 0; p < n_params; ++p)
        if (!param_used[p]) return fail(SQGPU_ERR_INVALID, "parameter %d is not used by any gate", p);

    // 2. plan: fuse runs of consecutive gates whose joint support is at most `max_q` qubits into one dense block
    //    (the device-side analogue of Gates_block's <=2-qubit fusion rule, Gates_block.cpp:632-681, applied to the
    //    flattened circuit and extended to the gradient by the product rule in build_block_warp)
    const bool fuse = !c->opt.no_fuse;
    // plan2 (streaming fallback) keeps the gates as they are; plan3 (shared-memory / windowed executor) sees constant
    // sub-circuits multiplied out into dense kernels (fuse_constant_runs)
    std::vector<cplx> pool_ext;
    if (matrix_pool && pool_len > 0) pool_ext.assign(reinterpret_cast<const cplx*>(matrix_pool), reinterpret_cast<const cplx*>(matrix_pool) + pool_len);
    std::vector<DevOp> raw_fused;
    int n_const_fused = 0;
    if (fuse && c->opt.const_fuse_qubits >= 4) n_const_fused = fuse_constant_runs(raw, qbit_num, c->opt.const_fuse_qubits, pool_ext, raw_fused);
    const std::vector<DevOp> raw_plain = raw;
    auto build_plan = [&](int max_q, const std::vector<DevOp>& raw, Plan& out) {
        const int n_gates = (int)raw.size();
        std::vector<DevOp> ops;
        std::vector<DevMember> members;
        std::vector<int> param_op(std::max(n_params, 1), -1), param_slot(std::max(n_params, 1), 0);
        int kern_total = 0, dkern_total = 0, w_total = 0, wmax = 4;
        int dense_stage = 0, n_dense = 0, n_dense5 = 0;
        std::vector<int> pend;
        unsigned pend_support = 0;
        auto finish_op = [&](DevOp& op) {
            const int d2 = op.dim * op.dim;
            if (!(op.type == SQGPU_GENERAL)) {
                op.kern_off = kern_total;
                kern_total += d2;
            }
            if (op.dim > 2) {
                // generic dense path: dim^2 complex; raw 4-5 qubit kernels on the tensor cores: padded real embedding + patterns
                int need = op.dim * op.dim;
                if (op.dim > 16) need = (2 * op.dim) * (2 * op.dim + 4) / 2 + op.dim;
                else if (op.type == SQGPU_GENERAL && op.dim >= 8) need = (int)(sizeof(DenseTab) / sizeof(cplx)) + 1;  // DMMA fragment table
                if (op.type == SQGPU_GENERAL && op.ctrl_mask == 0 && (op.nq == 3 || op.nq == 4)) op.dtab = ++n_dense;
                if (op.type == SQGPU_GENERAL && op.ctrl_mask == 0 && op.nq == 5) op.dtab = -(++n_dense5);
                dense_stage = std::max(dense_stage, need);
            }
            if (op.n_params > 0) {
                op.dkern_off = dkern_total;
                dkern_total += d2 * op.n_params;
                op.w_off = w_total;
                w_total += d2;
                wmax = std::max(wmax, d2);
            }
            ops.push_back(op);
        };
        auto flush = [&]() {
            if (pend.empty()) return;
            DevOp b;
            memset(&b, 0, sizeof(b));
            b.type = SQ_OP_BLOCK;
            b.kern_off = b.dkern_off = b.w_off = -1;
            int qs[3] = {0, 0, 0}, nqs = 0;
            for (int q = 0; q < 30; ++q)
                if ((pend_support >> q) & 1) qs[nqs++] = q;
            auto local_bit = [&](int q) { for (int j = 0; j < nqs; ++j) if (qs[j] == q) return j; return 0; };
            b.dim = 1 << nqs;
            if (nqs == 1) {
                b.target = qs[0];
            } else {
                b.nq = nqs;
                for (int j = 0; j < nqs; ++j) b.q[j] = qs[j];
            }
            b.member_off = (int)members.size();
            b.n_members = (int)pend.size();
            int slot = 0;
            for (int gi : pend) {
                const DevOp& r = raw[gi];
                DevMember m;
                memset(&m, 0, sizeof(m));
                m.type = r.type;
                m.dim = r.dim;
                m.cl = -1;
                if (r.dim == 2) {
                    m.tl = local_bit(r.target);
                    if (r.ctrl_mask)
                        for (int q = 0; q < 30; ++q)
                            if ((r.ctrl_mask >> q) & 1) m.cl = local_bit(q);
                } else {
                    const bool hi_first = r.target == r.q[1];  // kernel bit 0 sits on the higher qubit (CROT with target > control)
                    m.tl = local_bit(hi_first ? r.q[1] : r.q[0]);
                    m.tl2 = local_bit(hi_first ? r.q[0] : r.q[1]);
                }
                m.param_start = r.param_start;
                m.n_params = r.n_params;
                m.slot0 = slot;
                m.pool_off = r.pool_off;
                for (int p = 0; p < r.n_params; ++p) {
                    param_op[r.param_start + p] = (int)ops.size();
                    param_slot[r.param_start + p] = slot + p;
                }
                slot += r.n_params;
                members.push_back(m);
            }
            b.n_params = slot;
            fill_fix(b);
            finish_op(b);
            pend.clear();
            pend_support = 0;
        };
        // Dependency-aware first fit: a block starts at the first gate not yet placed and then takes, in program order,
        // every later gate that still fits its qubit set and shares no qubit with a gate that was passed over (such gates
        // commute with everything in between, so pulling them forward is a valid reordering; inside a block the program
        // order is kept). The all-pairs adaptive structure closes its triangles this way -- (0,1), (0,2) and the later
        // (1,2) become one 8x8 block -- n = 10, L = 4: 84 ops instead of the 100 of consecutive-run fusion, 17 % fewer
        // flops per amplitude. Option fuse_consecutive restores runs of consecutive gates only.
        const bool consecutive_only = c->opt.fuse_consecutive != 0;
        const unsigned all_qubits = qbit_num >= 32 ? 0xffffffffu : ((1u << qbit_num) - 1u);
        auto fusable_op = [&](const DevOp& r) {
            return fuse && ((r.dim == 2 && popcount32(r.ctrl_mask) <= 1) || (r.dim == 4 && r.ctrl_mask == 0));
        };
        std::vector<char> placed(std::max(n_gates, 1), 0);
        for (int first = 0; first < n_gates; ++first) {
            if (placed[first]) continue;
            const DevOp& r0 = raw[first];
            if (!fusable_op(r0)) {
                DevOp op = r0;
                for (int p = 0; p < op.n_params; ++p) {
                    param_op[op.param_start + p] = (int)ops.size();
                    param_slot[op.param_start + p] = p;
                }
                finish_op(op);
                placed[first] = 1;
                continue;
            }
            unsigned blocked = 0;
            const int scan_end = std::min(n_gates, first + 8192);  // bounds the planner's work on very long circuits
            for (int i = first; i < scan_end; ++i) {
                if (placed[i]) continue;
                const DevOp& r = raw[i];
                const unsigned sup = support_mask(r);
                if (fusable_op(r) && (sup & blocked) == 0 && popcount32(pend_support | sup) <= max_q && (int)pend.size() < SQ_MAX_MEMBERS) {
                    pend.push_back(i);
                    pend_support |= sup;
                    placed[i] = 1;
                } else {
                    if (consecutive_only) break;
                    blocked |= sup;
                    if ((blocked & all_qubits) == all_qubits) break;
                }
            }
            flush();
        }
        flush();
        if (upload) {
            int rc;
            const size_t np1 = std::max(n_params, 1);
            if ((rc = out.dOps.ensure(std::max<size_t>(1, ops.size()) * sizeof(DevOp)))) return rc;
            if ((rc = out.dMembers.ensure(std::max<size_t>(1, members.size()) * sizeof(DevMember)))) return rc;
            if ((rc = out.dParamOp.ensure(2 * np1 * sizeof(int)))) return rc;
            if (!ops.empty()) CUDA_TRY(cudaMemcpy(out.dOps.p, ops.data(), ops.size() * sizeof(DevOp), cudaMemcpyHostToDevice));
            if (!members.empty()) CUDA_TRY(cudaMemcpy(out.dMembers.p, members.data(), members.size() * sizeof(DevMember), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(out.dParamOp.p, param_op.data(), np1 * sizeof(int), cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(out.dParamOp.as<int>() + np1, param_slot.data(), np1 * sizeof(int), cudaMemcpyHostToDevice));
        }
        out.ops.swap(ops);
        out.members.swap(members);
        out.param_op.swap(param_op);
        out.param_slot.swap(param_slot);
        out.n_ops = (int)out.ops.size();
        out.kern_total = kern_total;
        out.dkern_total = dkern_total;
        out.w_total = w_total;
        out.wmax = wmax;
        out.dense_stage = dense_stage;
        out.n_dense = n_dense;
        out.n_dense5 = n_dense5;
        out.dense_logct = -1;
        out.has_dense = false;
        for (const DevOp& o : out.ops)
            if (o.dim > 2 && o.type != SQ_OP_BLOCK) out.has_dense = true;
        // a fused block too small for the tensor path (fewer than 8 (group, column) items per tile) takes the generic dense path:
        // only circuits of up to ~5 qubits; they run the DNS kernels as well
        if (qbit_num <= 6) out.has_dense = true;
        return (int)SQGPU_OK;
    };

    int rc;
    if (upload) {
        if ((rc = c->dPool.ensure(std::max<size_t>(1, pool_ext.size()) * sizeof(cplx)))) return rc;
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (!pool_ext.empty()) CUDA_TRY(cudaMemcpy(c->dPool.p, pool_ext.data(), pool_ext.size() * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    c->circuit_set = false;
    if ((rc = build_plan(2, raw_plain, c->plan2))) return rc;
    const int max_q3 = c->opt.max_fuse_qubits == 2 ? 2 : 3;
    if ((rc = build_plan(max_q3, n_const_fused > 0 ? raw_fused : raw_plain, c->plan3))) return rc;
    c->n_const_fused = n_const_fused;
    c->qbit_num = qbit_num;
    if ((rc = build_window_plan(c, upload))) return rc;
    for (int rho = 1; rho <= sqgpu_ctx::MAX_RHO; ++rho) {
        c->cl_ok[rho - 1] = false;
        if (qbit_num < 12 || !fuse) continue;
        int rc2 = SQGPU_OK;
        build_cluster_plan(c, rho, upload, &rc2);
        if (rc2) return rc2;
    }
    c->P = &c->plan2;
    c->n_gates = n_gates;
    c->n_params = n_params;
    c->qbit_num = qbit_num;
    c->all_unitary = all_unitary;
    c->circuit_set = upload;
    return SQGPU_OK;
}

int sqgpu_set_circuit(sqgpu_handle_t c, const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num,
                      const double* matrix_pool, int64_t pool_len) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (c->multi) {
        std::lock_guard<std::mutex> lk(c->mtx);
        return multi_set_circuit(c, gates, n_gates, n_params, qbit_num, matrix_pool, pool_len);
    }
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    wait_idle(c);
    return set_circuit_impl(c, gates, n_gates, n_params, qbit_num, matrix_pool, pool_len, true);
}

int sqgpu_plan_stats(const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool,
                     int64_t pool_len, int64_t* stats, int n_stats) {
    return sqgpu_plan_stats_opt(gates, n_gates, n_params, qbit_num, matrix_pool, pool_len, nullptr, stats, n_stats);
}

int sqgpu_plan_stats_opt(const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool,
                         int64_t pool_len, const char* options, int64_t* stats, int n_stats) {
    if (!stats || n_stats < 0) return fail(SQGPU_ERR_INVALID, "NULL stats");
    sqgpu_ctx* tmp = new sqgpu_ctx();  // never touches a device: no stream, no allocations
    int rc = options_parse(tmp->opt, options);
    if (rc == SQGPU_OK) rc = set_circuit_impl(tmp, gates, n_gates, n_params, qbit_num, matrix_pool, pool_len, false);
    if (rc == SQGPU_OK) {
        int max_seg = 0;
        for (const auto& sg : tmp->segs) max_seg = std::max(max_seg, sg.end - sg.begin);
        const int64_t v[SQGPU_PLAN_STATS] = {tmp->plan2.n_ops, tmp->plan3.n_ops, (int64_t)tmp->segs.size(), max_seg, tmp->win_w,
                                             tmp->plan3.kern_total, tmp->plan3.dkern_total, tmp->plan3.w_total,
                                             tmp->plan3.n_dense + tmp->plan3.n_dense5, (int64_t)tmp->plan3.members.size()};
        for (int i = 0; i < n_stats && i < SQGPU_PLAN_STATS; ++i) stats[i] = v[i];
    }
    delete tmp;
    return rc;
}

// ops[i*8 .. i*8+8) = {dim, q0, q1, q2, q3, q4, n_params, n_members} of op i of the chosen plan (2: <=2-qubit blocks, 3: <=3-qubit
// blocks, 0: the window plan in segment order); unused qubit slots are -1. Returns the number of ops through *n_ops.
int sqgpu_plan_ops(const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool,
                   int64_t pool_len, const char* options, int which, int32_t* ops, int cap, int* n_ops) {
    if (!n_ops) return fail(SQGPU_ERR_INVALID, "NULL n_ops");
    sqgpu_ctx* tmp = new sqgpu_ctx();
    int rc = options_parse(tmp->opt, options);
    if (rc == SQGPU_OK) rc = set_circuit_impl(tmp, gates, n_gates, n_params, qbit_num, matrix_pool, pool_len, false);
    if (rc == SQGPU_OK) {
        const int rho = which - 10;  // 11, 12, 13: the cluster plans for 2, 4, 8 CTAs
        if (rho >= 1 && rho <= sqgpu_ctx::MAX_RHO && !tmp->cl_ok[rho - 1]) {
            *n_ops = 0;  // no cluster plan for this circuit (fewer than 12 qubits, raw dense ops, ...)
            delete tmp;
            return SQGPU_OK;
        }
        const Plan& P = (rho >= 1 && rho <= sqgpu_ctx::MAX_RHO) ? tmp->planC[rho - 1] : (which == 2 ? tmp->plan2 : (which == 3 ? tmp->plan3 : tmp->planW));
        *n_ops = (int)P.ops.size();
        for (int i = 0; i < (int)P.ops.size() && i < cap; ++i) {
            const DevOp& op = P.ops[i];
            int32_t* o = ops + (size_t)i * 8;
            o[0] = op.dim;
            for (int j = 0; j < 5; ++j) o[1 + j] = -1;
            if (op.type == SQ_OP_RESPLIT) {  // dim 0: {local row bit, cluster-rank bit}
                o[1] = op.target;
                o[2] = op.nq;
            } else if (op.dim == 2) o[1] = op.target;
            else
                for (int j = 0; j < op.nq && j < 5; ++j) o[1 + j] = op.q[j];
            o[6] = op.n_params;
            o[7] = op.n_members;
        }
    }
    delete tmp;
    return rc;
}

int sqgpu_set_cost(sqgpu_handle_t c, int variant, int trace_offset, double prev_cost_fnv_val, double correction1_scale,
                   double correction2_scale) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(c->mtx);
    if (c->multi) return multi_set_cost(c, variant, trace_offset, prev_cost_fnv_val, correction1_scale, correction2_scale);
    return set_cost_checked(c, variant, trace_offset, prev_cost_fnv_val, correction1_scale, correction2_scale);
}

// ---- device-resident entry points ----------------------------------------------------------------------------------

int sqgpu_cost_batched_dev(sqgpu_handle_t c, const double* d_params, int batch, double* d_cost, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0 || (batch > 0 && (!d_params && c->n_params > 0)) || (batch > 0 && !d_cost)) return fail(SQGPU_ERR_INVALID, "bad arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return eval_dev(c, d_params, batch, false, d_cost, nullptr, (cudaStream_t)stream);
}

int sqgpu_cost_grad_batched_dev(sqgpu_handle_t c, const double* d_params, int batch, double* d_cost, double* d_grad, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0 || (batch > 0 && (!d_params && c->n_params > 0)) || (batch > 0 && (!d_cost || (!d_grad && c->n_params > 0)))) return fail(SQGPU_ERR_INVALID, "bad arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return eval_dev(c, d_params, batch, true, d_cost, d_grad, (cudaStream_t)stream);
}

int sqgpu_cost_shifted_batched_dev(sqgpu_handle_t c, const double* d_params, int batch, const double* shifts, int n_shifts, double* d_cost, double* d_shifted,
                                   void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0 || n_shifts < 1 || !shifts || (batch > 0 && (!d_params && c->n_params > 0)) || (batch > 0 && (!d_cost || (!d_shifted && c->n_params > 0)))) return fail(SQGPU_ERR_INVALID, "bad arguments");
    if (batch == 0) return SQGPU_OK;
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return shifted_eval_dev(c, d_params, batch, shifts, n_shifts, d_cost, d_shifted, (cudaStream_t)stream);
}

int sqgpu_cost_shifted_batched(sqgpu_handle_t c, const double* params, int batch, const double* shifts, int n_shifts, double* cost, double* shifted) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0 || n_shifts < 1 || !shifts) return fail(SQGPU_ERR_INVALID, "negative batch or no shifts");
    if (batch == 0) return SQGPU_OK;
    if ((!params && c->n_params > 0) || !cost || (!shifted && c->n_params > 0)) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc = check_ready(c, true);
    if (rc) return rc;
    const size_t np = (size_t)batch * c->n_params;
    if ((rc = c->wParams.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
    if ((rc = c->wCost.ensure((size_t)batch * sizeof(double)))) return rc;
    if ((rc = c->wGrad.ensure(std::max<size_t>(1, np * n_shifts) * sizeof(double)))) return rc;
    if (np) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = shifted_eval_dev(c, c->wParams.as<double>(), batch, shifts, n_shifts, c->wCost.as<double>(), c->wGrad.as<double>(), c->stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(cost, c->wCost.p, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (np) CUDA_TRY(cudaMemcpyAsync(shifted, c->wGrad.p, np * n_shifts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_traces_batched_dev(sqgpu_handle_t c, const double* d_params, int batch, int with_grad, double* d_traces, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0 || (batch > 0 && !d_traces)) return fail(SQGPU_ERR_INVALID, "bad arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return traces_dev(c, d_params, batch, with_grad != 0, d_traces, (cudaStream_t)stream, false);
}

int sqgpu_cost_from_traces_dev(sqgpu_handle_t c, const double* d_traces, int batch, int with_grad, int cols_total,
                               double* d_cost, double* d_grad, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0 || cols_total <= 0 || (batch > 0 && (!d_traces || !d_cost))) return fail(SQGPU_ERR_INVALID, "bad arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    if (!c->circuit_set) return fail(SQGPU_ERR_STATE, "no circuit set");
    return cost_from_traces_dev(c, d_traces, batch, with_grad != 0, cols_total, d_cost, d_grad, (cudaStream_t)stream);
}

// ---- host-buffer entry points (what the reference's hooks call) ------------------------------------------------------

static int host_eval(sqgpu_handle_t c, const double* params, int batch, bool with_grad, double* cost, double* grad) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (c->multi) {
        std::lock_guard<std::mutex> lk(c->mtx);
        return multi_eval(c, params, batch, with_grad, cost, grad);
    }
    if (batch < 0) return fail(SQGPU_ERR_INVALID, "negative batch");
    if (batch == 0) return SQGPU_OK;
    if ((!params && c->n_params > 0) || !cost || (with_grad && !grad && c->n_params > 0)) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc = check_ready(c, true);
    if (rc) return rc;
    const size_t np = (size_t)batch * c->n_params;
    if ((rc = c->wParams.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
    if ((rc = c->wCost.ensure((size_t)batch * sizeof(double)))) return rc;
    if (with_grad && (rc = c->wGrad.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
    if (np) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = eval_dev(c, c->wParams.as<double>(), batch, with_grad, c->wCost.as<double>(), with_grad ? c->wGrad.as<double>() : nullptr, c->stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(cost, c->wCost.p, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (with_grad && np) CUDA_TRY(cudaMemcpyAsync(grad, c->wGrad.p, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_cost_batched(sqgpu_handle_t c, const double* params, int batch, double* cost) {
    return host_eval(c, params, batch, false, cost, nullptr);
}

int sqgpu_cost_grad_batched(sqgpu_handle_t c, const double* params, int batch, double* cost, double* grad) {
    return host_eval(c, params, batch, true, cost, grad);
}

int sqgpu_traces_batched(sqgpu_handle_t c, const double* params, int batch, int with_grad, double* traces) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0) return fail(SQGPU_ERR_INVALID, "negative batch");
    if (batch == 0) return SQGPU_OK;
    if ((!params && c->n_params > 0) || !traces) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc = check_ready(c, true);
    if (rc) return rc;
    const size_t np = (size_t)batch * c->n_params;
    const int n_k = 1 + (with_grad ? c->n_params : 0);
    if ((rc = c->wParams.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
    if ((rc = c->wTraces.ensure((size_t)batch * n_k * 6 * sizeof(double)))) return rc;
    if (np) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = traces_dev(c, c->wParams.as<double>(), batch, with_grad != 0, c->wTraces.as<double>(), c->stream, false))) return rc;
    CUDA_TRY(cudaMemcpyAsync(traces, c->wTraces.p, (size_t)batch * n_k * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_cost_from_traces(sqgpu_handle_t c, const double* traces, int batch, int with_grad, int cols_total, double* cost, double* grad) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (batch < 0 || cols_total <= 0) return fail(SQGPU_ERR_INVALID, "bad arguments");
    if (batch == 0) return SQGPU_OK;
    if (!traces || !cost || (with_grad && !grad && c->n_params > 0)) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    if (!c->circuit_set) return fail(SQGPU_ERR_STATE, "no circuit set");
    int rc;
    const size_t np = (size_t)batch * c->n_params;
    const int n_k = 1 + (with_grad ? c->n_params : 0);
    if ((rc = c->wTraces.ensure((size_t)batch * n_k * 6 * sizeof(double)))) return rc;
    if ((rc = c->wCost.ensure((size_t)batch * sizeof(double)))) return rc;
    if (with_grad && (rc = c->wGrad.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->wTraces.p, traces, (size_t)batch * n_k * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = cost_from_traces_dev(c, c->wTraces.as<double>(), batch, with_grad != 0, cols_total, c->wCost.as<double>(),
                                   with_grad ? c->wGrad.as<double>() : nullptr, c->stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(cost, c->wCost.p, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (with_grad && np) CUDA_TRY(cudaMemcpyAsync(grad, c->wGrad.p, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_apply(sqgpu_handle_t c, const double* params, double* inout, int rows, int cols, int stride) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (!inout || rows <= 0 || cols <= 0 || stride < cols) return fail(SQGPU_ERR_INVALID, "bad matrix arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    c->P = &c->plan2;  // apply paths: two-qubit blocks (the streaming kernels stop there)
    int rc = check_ready(c, false);
    if (rc) return rc;
    if (rows != (1 << c->qbit_num)) return fail(SQGPU_ERR_INVALID, "Wrong input size in Gates_block gate apply: %d rows for %d qubits", rows, c->qbit_num);
    if (!params && c->n_params > 0) return fail(SQGPU_ERR_INVALID, "params is NULL");
    const size_t bytes = (size_t)rows * cols * sizeof(cplx);
    if ((rc = c->wMat.ensure(bytes))) return rc;
    if ((rc = c->wParams.ensure(std::max<size_t>(1, c->n_params) * sizeof(double)))) return rc;
    if (c->n_params) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params, c->n_params * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpy2DAsync(c->wMat.p, (size_t)cols * sizeof(cplx), inout, (size_t)stride * sizeof(cplx), (size_t)cols * sizeof(cplx), rows, cudaMemcpyHostToDevice, c->stream));
    rc = 1;
    if (cols == 1 && !plan_fused(c, MODE_APPLY, rows, 1, 1).ok) rc = apply_window_dev(c, c->wMat.as<cplx>(), rows, c->stream);
    if (rc == 1) {
        c->P = &c->plan2;
        if ((rc = run_tables(c, c->wParams.as<double>(), 1, false, c->stream))) return rc;
        if ((rc = apply_program_dev(c, c->wMat.as<cplx>(), 0, c->wMat.as<cplx>(), 0, 1, rows, cols, nullptr, nullptr, c->stream))) return rc;
    } else if (rc) {
        return rc;
    }
    CUDA_TRY(cudaMemcpy2DAsync(inout, (size_t)stride * sizeof(cplx), c->wMat.p, (size_t)cols * sizeof(cplx), (size_t)cols * sizeof(cplx), rows, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_apply_derivative(sqgpu_handle_t c, const double* params, const double* in, int rows, int cols, int stride, double* out) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (!in || rows <= 0 || cols <= 0 || stride < cols) return fail(SQGPU_ERR_INVALID, "bad matrix arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    c->P = &c->plan2;  // apply paths: two-qubit blocks (the streaming kernels stop there)
    int rc = check_ready(c, false);
    if (rc) return rc;
    if (rows != (1 << c->qbit_num)) return fail(SQGPU_ERR_INVALID, "Wrong input size in Gates_block gate apply: %d rows for %d qubits", rows, c->qbit_num);
    const int P = c->n_params;
    if (P == 0) return SQGPU_OK;
    if (!params || !out) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    const size_t n_elem = (size_t)rows * cols;
    // derivative matrices are produced in slices of at most ~1 GiB
    const int slice = (int)std::max<size_t>(1, std::min<size_t>((size_t)P, ((size_t)1 << 30) / (n_elem * sizeof(cplx))));
    if ((rc = c->wMat.ensure((size_t)(slice + 1) * n_elem * sizeof(cplx)))) return rc;
    if ((rc = c->wParams.ensure((size_t)P * sizeof(double)))) return rc;
    cplx* d_in = c->wMat.as<cplx>();
    cplx* d_out = d_in + n_elem;
    CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpy2DAsync(d_in, (size_t)cols * sizeof(cplx), in, (size_t)stride * sizeof(cplx), (size_t)cols * sizeof(cplx), rows, cudaMemcpyHostToDevice, c->stream));
    if ((rc = run_tables(c, c->wParams.as<double>(), 1, true, c->stream))) return rc;
    for (int p0 = 0; p0 < P; p0 += slice) {
        const int np = std::min(slice, P - p0);
        std::vector<int> dop(np), dp(np);
        for (int i = 0; i < np; ++i) {
            dop[i] = c->P->param_op[p0 + i];
            dp[i] = c->P->param_slot[p0 + i];
        }
        if ((rc = apply_program_dev(c, d_in, 0, d_out, (long long)n_elem, np, rows, cols, &dop, &dp, c->stream))) return rc;
        CUDA_TRY(cudaMemcpyAsync(out + 2 * (size_t)p0 * n_elem, d_out, (size_t)np * n_elem * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return SQGPU_OK;
}

// shared by the host and device single-gate entry points: d_inout is a device matrix
static int apply_gate_on_device(sqgpu_ctx* c, const sqgpu_gate_desc* gate, const double* gate_params, const double* matrix_pool,
                                int deriv_param, cplx* d_inout, int rows, int cols, int ld, cudaStream_t st) {
    int n = 0;
    while ((1 << n) < rows) ++n;
    if ((1 << n) != rows) return fail(SQGPU_ERR_INVALID, "rows must be a power of two");
    sqgpu_gate_desc g = *gate;
    g.param_start = 0;
    int64_t pool_len = 0;
    if (g.type == SQGPU_GENERAL) {
        if (g.n_qubits < 1 || g.n_qubits > 5) return fail(SQGPU_ERR_INVALID, "GENERAL gate: 1..5 qubits supported");
        pool_len = g.matrix_off + ((int64_t)1 << (2 * g.n_qubits));
    }
    DevOp op;
    bool unitary;
    int rc = lower_gate(g, n, matrix_pool, pool_len, &op, &unitary);
    if (rc) return rc;
    if (deriv_param >= op.n_params) return fail(SQGPU_ERR_INVALID, "derivative parameter %d out of range", deriv_param);
    if (deriv_param >= 0 && op.n_params == 0) return fail(SQGPU_ERR_INVALID, "gate has no parameters");
    if (op.n_params > 0 && !gate_params) return fail(SQGPU_ERR_INVALID, "gate_params is NULL");
    const int d2 = op.dim * op.dim;
    // scratch: [op][params(4)][ktab <= 1024][dktab 4 * 16] in one small device buffer (16-byte aligned sections)
    const size_t off_par = 128, off_k = 256, off_dk = off_k + 1024 * sizeof(cplx);
    static_assert(sizeof(DevOp) <= 128, "DevOp does not fit its scratch slot");
    if ((rc = c->wDerivIdx.ensure(off_dk + 4 * 16 * sizeof(cplx)))) return rc;
    char* base = c->wDerivIdx.as<char>();
    const cplx* K;
    if (op.type == SQGPU_GENERAL) {
        CUDA_TRY(cudaMemcpyAsync(base + off_k, matrix_pool + 2 * g.matrix_off, d2 * sizeof(cplx), cudaMemcpyHostToDevice, st));
        K = reinterpret_cast<cplx*>(base + off_k);
    } else {
        op.kern_off = 0;
        op.dkern_off = 0;
        double pbuf[4] = {0, 0, 0, 0};
        for (int i = 0; i < op.n_params; ++i) pbuf[i] = gate_params[i];
        CUDA_TRY(cudaMemcpyAsync(base, &op, sizeof(DevOp), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(base + off_par, pbuf, sizeof(pbuf), cudaMemcpyHostToDevice, st));
        build_kernel_tables<<<1, 32, 0, st>>>(reinterpret_cast<DevOp*>(base), 1, nullptr, reinterpret_cast<double*>(base + off_par), 4, 1,
                                              nullptr, reinterpret_cast<cplx*>(base + off_k), d2, reinterpret_cast<cplx*>(base + off_dk), 4 * d2, 1, 0.0);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        K = deriv_param >= 0 ? reinterpret_cast<cplx*>(base + off_dk) + (size_t)deriv_param * d2 : reinterpret_cast<cplx*>(base + off_k);
    }
    time_begin(c, op.dim == 2 ? "gate1q_stream" : "gatekq_stream", st);
    rc = launch_stream_gate(c, op, deriv_param >= 0, d_inout, 0, 1, rows, cols, ld, K, 0, st);
    time_end(c, st);
    return rc;
}

int sqgpu_apply_gate(sqgpu_handle_t c, const sqgpu_gate_desc* gate, const double* gate_params, const double* matrix_pool,
                     int deriv_param, double* inout, int rows, int cols, int stride) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (!gate || !inout || rows <= 0 || cols <= 0 || stride < cols) return fail(SQGPU_ERR_INVALID, "bad arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc;
    if ((rc = c->wMat.ensure((size_t)rows * cols * sizeof(cplx)))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(c->wMat.p, (size_t)cols * sizeof(cplx), inout, (size_t)stride * sizeof(cplx), (size_t)cols * sizeof(cplx), rows, cudaMemcpyHostToDevice, c->stream));
    if ((rc = apply_gate_on_device(c, gate, gate_params, matrix_pool, deriv_param, c->wMat.as<cplx>(), rows, cols, cols, c->stream))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(inout, (size_t)stride * sizeof(cplx), c->wMat.p, (size_t)cols * sizeof(cplx), (size_t)cols * sizeof(cplx), rows, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_apply_gate_dev(sqgpu_handle_t c, const sqgpu_gate_desc* gate, const double* gate_params, const double* matrix_pool,
                         int deriv_param, double* d_inout, int rows, int cols, int stride, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    if (!gate || !d_inout || rows <= 0 || cols <= 0 || stride < cols) return fail(SQGPU_ERR_INVALID, "bad arguments");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return apply_gate_on_device(c, gate, gate_params, matrix_pool, deriv_param, reinterpret_cast<cplx*>(d_inout), rows, cols, stride, (cudaStream_t)stream);
}

// ---- VQE ------------------------------------------------------------------------------------------------------------

int sqgpu_set_hamiltonian_csr(sqgpu_handle_t c, int n_rows, int64_t nnz, const int32_t* indptr, const int32_t* indices, const double* values) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (n_rows <= 0 || nnz < 0 || !indptr || (nnz > 0 && (!indices || !values))) return fail(SQGPU_ERR_INVALID, "bad CSR arguments");
    if (indptr[0] != 0 || indptr[n_rows] != nnz) return fail(SQGPU_ERR_INVALID, "inconsistent CSR indptr");
    if (c->multi) {
        std::lock_guard<std::mutex> lk(c->mtx);
        return multi_set_hamiltonian(c, n_rows, nnz, indptr, indices, values);
    }
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    int rc;
    if ((rc = c->hIndptr.ensure((size_t)(n_rows + 1) * sizeof(int32_t)))) return rc;
    if ((rc = c->hIndices.ensure(std::max<size_t>(1, (size_t)nnz) * sizeof(int32_t)))) return rc;
    if ((rc = c->hValues.ensure(std::max<size_t>(1, (size_t)nnz) * sizeof(cplx)))) return rc;
    wait_idle(c);
    CUDA_TRY(cudaMemcpy(c->hIndptr.p, indptr, (size_t)(n_rows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (nnz) {
        CUDA_TRY(cudaMemcpy(c->hIndices.p, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(c->hValues.p, values, (size_t)nnz * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    c->h_rows = n_rows;
    c->h_nnz = nnz;
    return SQGPU_OK;
}

static int vqe_dev(sqgpu_ctx* c, const double* d_params, int batch, bool with_grad, double* d_energy, double* d_grad, cudaStream_t st);

int sqgpu_vqe_energy_batched_dev(sqgpu_handle_t c, const double* d_params, int batch, double* d_energy, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return vqe_dev(c, d_params, batch, false, d_energy, nullptr, (cudaStream_t)stream);
}

int sqgpu_vqe_energy_grad_batched_dev(sqgpu_handle_t c, const double* d_params, int batch, double* d_energy, double* d_grad, void* stream) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    SQ_NOT_ON_MULTI(c);
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, (cudaStream_t)stream);
    return vqe_dev(c, d_params, batch, true, d_energy, d_grad, (cudaStream_t)stream);
}

static int vqe_host(sqgpu_handle_t c, const double* params, int batch, bool with_grad, double* energy, double* grad) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    if (c->multi) {
        std::lock_guard<std::mutex> lk(c->mtx);
        return multi_vqe(c, params, batch, with_grad, energy, grad);
    }
    if (batch < 0) return fail(SQGPU_ERR_INVALID, "negative batch");
    if (batch == 0) return SQGPU_OK;
    if ((!params && c->n_params > 0) || !energy || (with_grad && !grad && c->n_params > 0)) return fail(SQGPU_ERR_INVALID, "NULL buffer");
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc;
    const size_t np = (size_t)batch * c->n_params;
    if ((rc = c->wParams.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
    if ((rc = c->wCost.ensure((size_t)batch * sizeof(double)))) return rc;
    if (with_grad && (rc = c->wGrad.ensure(std::max<size_t>(1, np) * sizeof(double)))) return rc;
    if (np) CUDA_TRY(cudaMemcpyAsync(c->wParams.p, params, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = vqe_dev(c, c->wParams.as<double>(), batch, with_grad, c->wCost.as<double>(), with_grad ? c->wGrad.as<double>() : nullptr, c->stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(energy, c->wCost.p, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (with_grad && np) CUDA_TRY(cudaMemcpyAsync(grad, c->wGrad.p, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SQGPU_OK;
}

int sqgpu_vqe_energy_batched(sqgpu_handle_t c, const double* params, int batch, double* energy) {
    return vqe_host(c, params, batch, false, energy, nullptr);
}

int sqgpu_vqe_energy_grad_batched(sqgpu_handle_t c, const double* params, int batch, double* energy, double* grad) {
    return vqe_host(c, params, batch, true, energy, grad);
}

// ---- introspection ---------------------------------------------------------------------------------------------------

int sqgpu_launch_count(sqgpu_handle_t c, int64_t* count) {
    if (!c || !count) return fail(SQGPU_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mtx);
    *count = c->multi ? multi_launches(c) : c->launches;
    return SQGPU_OK;
}

// average of the events of one ring (the last RING brackets); resets the ring
static int ring_average(KernelTimer::Ring& r, double* ms, int* launches) {
    *ms = 0;
    *launches = 0;
    if (!r.init || r.n == 0) return SQGPU_OK;
    const int cnt = std::min(r.n, (int)KernelTimer::RING);
    double tot = 0;
    for (int i = 0; i < cnt; ++i) {
        const int slot = (r.n - 1 - i) % KernelTimer::RING;
        CUDA_TRY(cudaEventSynchronize(r.e1[slot]));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, r.e0[slot], r.e1[slot]));
        tot += t;
    }
    *ms = tot / cnt;
    *launches = cnt;
    r.n = 0;
    return SQGPU_OK;
}

int sqgpu_last_kernel_time(sqgpu_handle_t c, char* name, int name_len, double* ms, int* launches) {
    if (!c || !ms || !launches) return fail(SQGPU_ERR_INVALID, "NULL argument");
    if (c->multi) c = multi_first(c);  // introspection of a multi-device handle reports its first device
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    *ms = 0;
    *launches = 0;
    if (name && name_len > 0) name[0] = 0;
    // the ring with the largest accumulated time since the last call is "the dominant kernel"; all rings are reset
    double best_total = -1;
    for (auto& r : c->timer.rings) {
        if (!r.init || r.n == 0) continue;
        double m = 0;
        int n = 0;
        int rc = ring_average(r, &m, &n);
        if (rc) return rc;
        if (m * n > best_total) {
            best_total = m * n;
            *ms = m;
            *launches = n;
            if (name && name_len > 0) {
                strncpy(name, r.name.c_str(), name_len - 1);
                name[name_len - 1] = 0;
            }
        }
    }
    return SQGPU_OK;
}

int sqgpu_kernel_time(sqgpu_handle_t c, const char* name, double* ms, int* launches) {
    if (!c || !name || !ms || !launches) return fail(SQGPU_ERR_INVALID, "NULL argument");
    if (c->multi) c = multi_first(c);  // introspection of a multi-device handle reports its first device
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    *ms = 0;
    *launches = 0;
    for (auto& r : c->timer.rings)
        if (r.init && r.name == name) return ring_average(r, ms, launches);
    return SQGPU_OK;
}

int sqgpu_last_exec_flops(sqgpu_handle_t c, double* tensor_flops, double* scalar_flops) {
    if (!c || !tensor_flops || !scalar_flops) return fail(SQGPU_ERR_INVALID, "NULL argument");
    if (c->multi) c = multi_first(c);  // introspection of a multi-device handle reports its first device
    std::lock_guard<std::mutex> lk(c->mtx);
    *tensor_flops = c->last_flops[0];
    *scalar_flops = c->last_flops[1];
    return SQGPU_OK;
}

int sqgpu_last_launch_shape(sqgpu_handle_t c, int* shape, int n_shape) {
    if (!c || !shape || n_shape < 0) return fail(SQGPU_ERR_INVALID, "NULL argument");
    if (c->multi) c = multi_first(c);  // introspection of a multi-device handle reports its first device
    std::lock_guard<std::mutex> lk(c->mtx);
    for (int i = 0; i < n_shape && i < 6; ++i) shape[i] = c->last_shape[i];
    return SQGPU_OK;
}

int sqgpu_set_option(sqgpu_handle_t c, const char* name, int64_t value) {
    if (!c) return fail(SQGPU_ERR_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(c->mtx);
    if (c->multi) return multi_set_option(c, name, value);
    return option_set(c->opt, name, value);
}

int sqgpu_get_option(sqgpu_handle_t c, const char* name, int64_t* value) {
    if (!c || !name || !value) return fail(SQGPU_ERR_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(c->mtx);
    for (const auto& on : kOptionNames)
        if (!strcmp(on.name, name)) {
            *value = c->opt.*(on.field);
            return SQGPU_OK;
        }
    return fail(SQGPU_ERR_INVALID, "unknown option '%s'", name);
}

int sqgpu_fp64_fma_peak(sqgpu_handle_t c, double* tflops) {
    if (!c || !tflops) return fail(SQGPU_ERR_INVALID, "NULL argument");
    if (c->multi) c = multi_first(c);  // introspection of a multi-device handle reports its first device
    DeviceGuard guard(c->device);
    std::lock_guard<std::mutex> lk(c->mtx);
    CallScope cs(c, c->stream);
    int rc;
    const int blocks = c->sm_count * 8, thr = 256, iters = 4096;
    if ((rc = c->wCost.ensure((size_t)blocks * thr * sizeof(double)))) return rc;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, c->stream));
        fp64_fma_burn<<<blocks, thr, 0, c->stream>>>(c->wCost.as<double>(), iters, 1.0000001);
        CUDA_TRY(cudaEventRecord(e1, c->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        c->launches++;
        const double flops = 2.0 * 16 * (double)iters * blocks * thr;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    // the FP64 tensor-core rate (DMMA m8n8k4, the instruction the executor's block path issues) and, for the record, DFMA and
    // DMMA mixed: all three share one pipe on B200; the reported peak is the larger of the DFMA and DMMA figures
    const bool vb = c->opt.verbose != 0;
    {
        double best_mma = 0, best_mix = 0;
        for (int rep = 0; rep < 4; ++rep) {
            float ms = 0;
            cudaEventRecord(e0, c->stream);
            fp64_dmma_burn<<<blocks, thr, 0, c->stream>>>(c->wCost.as<double>(), iters);
            cudaEventRecord(e1, c->stream);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double fl = 2.0 * 256 * 4 * (double)iters * blocks * (thr / 32);
            if (rep > 0) best_mma = std::max(best_mma, fl / (ms * 1e-3) * 1e-12);
            cudaEventRecord(e0, c->stream);
            fp64_mixed_burn<<<blocks, thr, 0, c->stream>>>(c->wCost.as<double>(), iters, 1.0000001);
            cudaEventRecord(e1, c->stream);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double fl2 = (2.0 * 256 * 2 * (thr / 32) + 2.0 * 16 * thr) * (double)iters * blocks;
            if (rep > 0) best_mix = std::max(best_mix, fl2 / (ms * 1e-3) * 1e-12);
        }
        if (vb)
            fprintf(stderr, "[sqgpu] FP64 peaks: DFMA %.2f TFLOP/s, DMMA m8n8k4 %.2f TFLOP/s, DFMA+DMMA mixed %.2f TFLOP/s\n", best, best_mma, best_mix);
        best = std::max(best, best_mma);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return SQGPU_OK;
}

}  // extern "C"

// VQE device path (defined after the C block so it can use the static helpers above)
#include "vqe_impl.cuh"
// several devices behind one handle
#include "multi.cuh"
// device-resident optimizer inner loops
#include "optim.cuh"

// Profiling builds only (-DSQ_TRACE=events, see exec_fused.cuh): copies the phase trace of the two CTAs on SM 0 to the host and
// re-arms it. Returns the number of long longs written (0 in product builds).
extern "C" long long sqgpu_debug_trace(long long* out, long long max_elems) {
#if SQ_TRACE
    const long long n = (long long)2 * SQ_TRACE * 16 * 2;
    if (!out || max_elems < n) return -n;
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, sq::g_trace, n * sizeof(long long)) != cudaSuccess) return -1;
    const int zero = 0;
    cudaMemcpyToSymbol(sq::g_trace_arrivals, &zero, sizeof(int));
    return n;
#else
    (void)out; (void)max_elems;
    return 0;
#endif
}
