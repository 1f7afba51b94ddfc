// common_GPU.cpp -- see common_GPU.h. Mirrors common_DFE.cpp: dlopen once, bind the C symbols, rethrow failures as
// std::string (the exception type every SQUANDER entry point already throws, Gate.cpp:435-446).
#include "common_GPU.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace sqgpu_bridge {

namespace {

std::string g_lib_path = "libsqgpu.so";
void* g_lib = nullptr;
std::mutex g_lib_mutex;

#define SQGPU_SYMBOLS(X)                                                                                              \
    X(sqgpu_device_count) X(sqgpu_create) X(sqgpu_destroy) X(sqgpu_last_error) X(sqgpu_abi_version) X(sqgpu_upload_matrix) \
    X(sqgpu_set_circuit) X(sqgpu_set_cost) X(sqgpu_cost_batched) X(sqgpu_cost_grad_batched) X(sqgpu_set_hamiltonian_csr)  \
    X(sqgpu_vqe_energy_batched) X(sqgpu_vqe_energy_grad_batched)

#define DECLARE(name) decltype(&name) p_##name = nullptr;
SQGPU_SYMBOLS(DECLARE)
#undef DECLARE

void load_library() {  // == init_dfe_lib's dlopen / dlsym block, common_DFE.cpp:136-176
    std::lock_guard<std::mutex> lk(g_lib_mutex);
    if (g_lib) return;
    void* lib = dlopen(g_lib_path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw std::string("init_gpu_lib: cannot load ") + g_lib_path + ": " + dlerror();
#define LOAD(name)                                              \
    p_##name = reinterpret_cast<decltype(p_##name)>(dlsym(lib, #name)); \
    if (!p_##name) throw std::string("init_gpu_lib: symbol " #name " missing in ") + g_lib_path;
    SQGPU_SYMBOLS(LOAD)
#undef LOAD
    if (p_sqgpu_abi_version() != SQGPU_ABI_VERSION)
        throw std::string("init_gpu_lib: ABI version mismatch between sqgpu.h and ") + g_lib_path;
    g_lib = lib;
}

void check(int rc) {
    if (rc != SQGPU_OK) throw std::string(p_sqgpu_last_error());
}

}  // namespace

void set_library_path(const std::string& path) {
    std::lock_guard<std::mutex> lk(g_lib_mutex);
    g_lib_path = path;
}

int available_gpus() {
    load_library();
    int n = 0;
    if (p_sqgpu_device_count(&n) != SQGPU_OK) return 0;
    return n;
}

std::vector<sqgpu_gate_desc> to_gpu_gates(Gates_block* circuit, std::vector<QGD_Complex16>& pool) {
    std::vector<sqgpu_gate_desc> descs;
    Gates_block* flat = circuit->get_flat_circuit();
    std::vector<Gate*> gates = flat->get_gates();
    descs.reserve(gates.size());
    for (Gate* g : gates) {
        sqgpu_gate_desc d;
        memset(&d, 0, sizeof(d));
        d.type = (int32_t)g->get_type();  // the enum values are shared with sqgpu_gate_type (Gate.h:39-79)
        d.target = g->get_target_qbit();
        d.control = g->get_control_qbit();
        d.target2 = d.control2 = -1;
        d.param_start = g->get_parameter_start_idx();
        d.n_params = g->get_parameter_num();
        const std::vector<int> tq = g->get_target_qbits(), cq = g->get_control_qbits();
        if (g->get_type() == GENERAL_OPERATION) {  // constant local kernel; local index bit j <-> j-th ascending qubit
            std::vector<int> q = g->get_involved_qubits();
            if (q.empty() || q.size() > SQGPU_MAX_GENERAL_QUBITS) { delete flat; throw std::string("to_gpu_gates: GENERAL gate on an unsupported number of qubits"); }
            d.n_qubits = (int32_t)q.size();
            d.target = d.control = -1;
            for (size_t j = 0; j < q.size(); ++j) d.qubits[j] = q[j];
            Matrix k = g->get_matrix();
            d.matrix_off = (int64_t)pool.size();
            for (int r = 0; r < k.rows; ++r) pool.insert(pool.end(), k.get_data() + (size_t)r * k.stride, k.get_data() + (size_t)r * k.stride + k.cols);
        } else {
            if (tq.size() == 2) {  // SWAP, CSWAP, RXX, RYY, RZZ
                d.target = tq[0];
                d.target2 = tq[1];
            }
            if (cq.size() == 2) {  // CCX
                d.control = cq[0];
                d.control2 = cq[1];
            } else if (cq.size() == 1) {
                d.control = cq[0];
            }
        }
        descs.push_back(d);
    }
    delete flat;
    return descs;
}

GPU_Cost_Path::GPU_Cost_Path(int accelerator_num) {
    if (accelerator_num < 1) throw std::string("GPU_Cost_Path: accelerator_num should be at least 1");
    load_library();
    check(p_sqgpu_create(0, &h));
}

GPU_Cost_Path::~GPU_Cost_Path() {
    if (h) p_sqgpu_destroy(h);
}

void GPU_Cost_Path::upload(Matrix& Umtx) {
    check(p_sqgpu_upload_matrix(h, reinterpret_cast<const double*>(Umtx.get_data()), Umtx.rows, Umtx.cols, Umtx.stride));
}

void GPU_Cost_Path::set_circuit(Gates_block* circuit) {
    std::vector<QGD_Complex16> pool;
    std::vector<sqgpu_gate_desc> descs = to_gpu_gates(circuit, pool);
    n_params = circuit->get_parameter_num();
    check(p_sqgpu_set_circuit(h, descs.data(), (int)descs.size(), n_params, circuit->get_qbit_num(),
                              pool.empty() ? nullptr : reinterpret_cast<const double*>(pool.data()), (int64_t)pool.size()));
}

void GPU_Cost_Path::set_hamiltonian(Matrix_sparse& H) {
    check(p_sqgpu_set_hamiltonian_csr(h, H.rows, (int64_t)H.NNZ, H.indptr, H.indices, reinterpret_cast<const double*>(H.data)));
}

void GPU_Cost_Path::set_cost(int variant, int trace_offset, double prev, double c1, double c2) {
    check(p_sqgpu_set_cost(h, variant, trace_offset, prev, c1, c2));
}

double GPU_Cost_Path::cost(Matrix_real& parameters) {
    if ((int)parameters.size() != n_params) throw std::string("Optimization_Interface::optimization_problem: Wrong number of parameters.");
    double f = 0.0;
    check(p_sqgpu_cost_batched(h, parameters.get_data(), 1, &f));
    n_evals += 1;
    return f;
}

void GPU_Cost_Path::cost_grad(Matrix_real& parameters, double* f0, Matrix_real& grad) {
    if ((int)parameters.size() != n_params) throw std::string("Optimization_Interface::optimization_problem_combined: Wrong number of parameters.");
    if ((int)grad.size() != n_params) grad = Matrix_real(1, n_params);
    check(p_sqgpu_cost_grad_batched(h, parameters.get_data(), 1, f0, grad.get_data()));
    n_evals += 1;
}

Matrix_real GPU_Cost_Path::cost_batched(std::vector<Matrix_real>& parameters_vec) {
    const int batch = (int)parameters_vec.size();
    Matrix_real cost_fnc_mtx(batch, 1);
    if (batch == 0) return cost_fnc_mtx;
    std::vector<double> packed((size_t)batch * n_params);
    for (int b = 0; b < batch; ++b) {
        if ((int)parameters_vec[b].size() != n_params) throw std::string("Optimization_Interface::optimization_problem_batched: Wrong number of parameters.");
        memcpy(packed.data() + (size_t)b * n_params, parameters_vec[b].get_data(), sizeof(double) * n_params);
    }
    check(p_sqgpu_cost_batched(h, packed.data(), batch, cost_fnc_mtx.get_data()));
    n_evals += batch;
    return cost_fnc_mtx;
}

double GPU_Cost_Path::vqe_energy(Matrix_real& parameters) {
    double e = 0.0;
    check(p_sqgpu_vqe_energy_batched(h, parameters.get_data(), 1, &e));
    n_evals += 1;
    return e;
}

void GPU_Cost_Path::vqe_energy_grad(Matrix_real& parameters, double* f0, Matrix_real& grad) {
    if ((int)grad.size() != n_params) grad = Matrix_real(1, n_params);
    check(p_sqgpu_vqe_energy_grad_batched(h, parameters.get_data(), 1, f0, grad.get_data()));
    n_evals += 1;
}

}  // namespace sqgpu_bridge

// ---- structural fingerprint (declared in GPU_Cost_Path_Mixin.h) ------------------------------------------------------------
namespace sqgpu_bridge {

namespace {
inline void mix(uint64_t& h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
}
void walk(Gates_block* blk, uint64_t& h) {
    for (Gate* g : blk->get_gates()) {
        mix(h, (uint64_t)g->get_type());
        if (g->get_type() == BLOCK_OPERATION) {
            mix(h, 0xb10cULL);
            walk(static_cast<Gates_block*>(g), h);
            mix(h, 0xe0dULL);
            continue;
        }
        for (int q : g->get_target_qbits()) mix(h, 0x100ULL + (uint64_t)q);
        for (int q : g->get_control_qbits()) mix(h, 0x200ULL + (uint64_t)q);
        mix(h, (uint64_t)g->get_parameter_num());
        if (g->get_type() == GENERAL_OPERATION) mix(h, (uint64_t)(uintptr_t)g);  // constant kernels: identity of the gate object
    }
}
}  // namespace

uint64_t fingerprint(Gates_block* circuit) {
    uint64_t h = 0xcbf29ce484222325ULL;
    mix(h, (uint64_t)circuit->get_qbit_num());
    walk(circuit, h);
    return h;
}

}  // namespace sqgpu_bridge
