// GPU_Cost_Path_Mixin.h -- the drop-in itself: SQUANDER's decomposition / VQE classes with their cost path served by
// libsqgpu.so, WITHOUT touching a line of the reference sources.
//
// The reference routes every optimizer (BFGS, BFGS2, ADAM, GRAD_DESCEND, ...) through three virtual functions:
//     double optimization_problem(Matrix_real&)                                  (Optimization_Interface.h:333, .cpp:634-668)
//     double optimization_problem_non_static(Matrix_real, void*)                 (.h:413, .cpp:1088-1094)
//     void   optimization_problem_combined_non_static(Matrix_real, void*, double*, Matrix_real&)   (.h:444, .cpp:1145-1490)
// (BFGS_Powell gets the static optimization_problem_combined, which forwards to the _non_static virtual, .cpp:1496-1500;
//  ADAM calls the 3-argument virtual optimization_problem_combined, which forwards the same way, .cpp:1510-1514.) The VQE
// class overrides the same three (Variational_Quantum_Eigensolver_Base.h). With_GPU_Cost_Path<Base> overrides them once more
// and forwards to the engine, so `With_GPU_Cost_Path<N_Qubit_Decomposition_adaptive>` IS the reference class -- its
// optimizers, level search, compression, export -- with the GPU under it.
//
// One hook of the four in SURVEY.md §8b is not virtual: optimization_problem_batched (Optimization_Interface.h:385,
// .cpp:939-1033, used by AGENTS / COSINE / the parameter-shift descent). For it the mixin provides
// optimization_problem_batched_GPU and integration/gpu_hooks.patch shows the three-line #ifdef __GPU__ hook a maintainer
// adds at .cpp:944 next to the __DFE__ / __GROQ__ ones.
#pragma once
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "common_GPU.h"

#include "Optimization_Interface.h"
#include "Variational_Quantum_Eigensolver_Base.h"

namespace sqgpu_bridge {

// structural fingerprint of a (nested) gate structure: type, qubits and parameter slots of every gate, depth first. The
// device plan is rebuilt when it changes (add_adaptive_layers, compression, set_custom_gate_structure ...).
uint64_t fingerprint(Gates_block* circuit);

inline bool variant_on_device(cost_function_type v) {
    switch (v) {
        case FROBENIUS_NORM: case FROBENIUS_NORM_CORRECTION1: case FROBENIUS_NORM_CORRECTION2: case HILBERT_SCHMIDT_TEST:
        case HILBERT_SCHMIDT_TEST_CORRECTION1: case HILBERT_SCHMIDT_TEST_CORRECTION2: case SUM_OF_SQUARES: case INFIDELITY:
            return true;
        default:
            return false;  // OSR_ENTANGLEMENT (an SVD per cut) stays on the reference's CPU path
    }
}

template <class Base>
class With_GPU_Cost_Path : public Base {
public:
    using Base::Base;

    // upload_Umtx_to_DFE's counterpart (Optimization_Interface.cpp:1819-1824); also called lazily by the hooks below
    void upload_Umtx_to_GPU() {
        engine().upload(this->Umtx);
        umtx_data = this->Umtx.get_data();
        umtx_rows = this->Umtx.rows;
        umtx_cols = this->Umtx.cols;
    }

    double optimization_problem(Matrix_real& parameters) override {
        if (!variant_on_device(this->cost_fnc)) return Base::optimization_problem(parameters);
        return synced().cost(parameters);
    }

    double optimization_problem_non_static(Matrix_real parameters, void* void_instance) override {
        With_GPU_Cost_Path* instance = reinterpret_cast<With_GPU_Cost_Path*>(void_instance);
        if (!variant_on_device(instance->cost_fnc)) return instance->Base::optimization_problem_non_static(parameters, void_instance);
        instance->increment_num_iters();
        return instance->synced().cost(parameters);
    }

    void optimization_problem_combined_non_static(Matrix_real parameters, void* void_instance, double* f0, Matrix_real& grad) override {
        With_GPU_Cost_Path* instance = reinterpret_cast<With_GPU_Cost_Path*>(void_instance);
        if (!variant_on_device(instance->cost_fnc)) {
            instance->Base::optimization_problem_combined_non_static(parameters, void_instance, f0, grad);
            return;
        }
        instance->synced().cost_grad(parameters, f0, grad);
        instance->increment_num_iters((int)parameters.size() + 1);  // as the DFE branch counts it (.cpp:1310)
    }

    // the body of the non-virtual hook (Optimization_Interface.cpp:944-956, see gpu_hooks.patch)
    Matrix_real optimization_problem_batched_GPU(std::vector<Matrix_real>& parameters_vec) {
        this->increment_num_iters(static_cast<int>(parameters_vec.size()));
        return synced().cost_batched(parameters_vec);
    }

    long long gpu_evaluations() { return gpu ? gpu->evaluations() : 0; }

protected:
#ifdef __GPU__
    // with gpu_hooks.patch applied (-D__GPU__) the reference's own non-virtual optimization_problem_batched forwards to the
    // engine through this hook: AGENTS, COSINE and the parameter-shift descent then run their batches on the GPU unchanged
    bool install_batched_hook() {
        this->gpu_batched_hook = [this](std::vector<Matrix_real>& v) { return this->optimization_problem_batched_GPU(v); };
        return true;
    }
    bool batched_hook_installed = install_batched_hook();
#endif
    GPU_Cost_Path& engine() {
        if (!gpu) gpu.reset(new GPU_Cost_Path(1));
        return *gpu;
    }
    // the engine with matrix, gate structure and cost configuration brought up to date
    GPU_Cost_Path& synced() {
        GPU_Cost_Path& e = engine();
        if (umtx_data != this->Umtx.get_data() || umtx_rows != this->Umtx.rows || umtx_cols != this->Umtx.cols) upload_Umtx_to_GPU();
        const uint64_t fp = fingerprint(this);
        if (!have_fp || fp != circuit_fp) {
            e.set_circuit(this);
            circuit_fp = fp;
            have_fp = true;
        }
        // prev_cost_fnv_val is rewritten by the ADAM engines between evaluations (ADAM.cpp:201): always current
        e.set_cost((int)this->cost_fnc, this->trace_offset, this->prev_cost_fnv_val, this->correction1_scale, this->correction2_scale);
        return e;
    }

    std::unique_ptr<GPU_Cost_Path> gpu;
    const void* umtx_data = nullptr;
    int umtx_rows = 0, umtx_cols = 0;
    uint64_t circuit_fp = 0;
    bool have_fp = false;
};

// VQE: Variational_Quantum_Eigensolver_Base::optimization_problem (...Base.cpp:1088-1121; the reference's __GROQ__ hook sits
// at :1099-1105) and optimization_problem_combined_non_static (:1131-1199) on the windowed state-vector executor.
class VQE_With_GPU_Cost_Path : public Variational_Quantum_Eigensolver_Base {
public:
    VQE_With_GPU_Cost_Path(Matrix_sparse Hamiltonian_in, int qbit_num_in, std::map<std::string, Config_Element>& config_in)
        : Variational_Quantum_Eigensolver_Base(Hamiltonian_in, qbit_num_in, config_in, 0), H(Hamiltonian_in) {}

    // the base keeps its initial state private: this shadows set_initial_state (same signature) to keep the device copy in step
    void set_initial_state(Matrix initial_state_in) {
        Variational_Quantum_Eigensolver_Base::set_initial_state(initial_state_in);
        state = initial_state_in.copy();
        state_dirty = true;
    }

    double optimization_problem(Matrix_real& parameters) override { return synced().vqe_energy(parameters); }

    double optimization_problem_non_static(Matrix_real parameters, void* void_instance) override {
        return reinterpret_cast<VQE_With_GPU_Cost_Path*>(void_instance)->synced().vqe_energy(parameters);
    }

    void optimization_problem_combined_non_static(Matrix_real parameters, void* void_instance, double* f0, Matrix_real& grad) override {
        reinterpret_cast<VQE_With_GPU_Cost_Path*>(void_instance)->synced().vqe_energy_grad(parameters, f0, grad);
    }

protected:
    GPU_Cost_Path& synced() {
        if (!gpu) {
            gpu.reset(new GPU_Cost_Path(1));
            gpu->set_hamiltonian(H);
        }
        if (state.size() == 0) {  // |0...0>, what initialize_zero_state() means (...Base.cpp:1262-1276)
            state = Matrix(1 << qbit_num, 1);
            memset(state.get_data(), 0, sizeof(QGD_Complex16) * state.size());
            state[0].real = 1.0;
            state_dirty = true;
        }
        if (state_dirty) {
            gpu->upload(state);
            state_dirty = false;
        }
        const uint64_t fp = fingerprint(this);
        if (!have_fp || fp != circuit_fp) {
            gpu->set_circuit(this);
            circuit_fp = fp;
            have_fp = true;
        }
        return *gpu;
    }

    Matrix_sparse H;
    Matrix state;
    bool state_dirty = false;
    std::unique_ptr<GPU_Cost_Path> gpu;
    uint64_t circuit_fp = 0;
    bool have_fp = false;
};

}  // namespace sqgpu_bridge
