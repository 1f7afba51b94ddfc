// ref_gpu_harness.cpp -- C entry points that run the reference's OWN optimizers (BFGS_Powell, ADAM) on the reference's own
// decomposition class, with the cost path served either by the reference's CPU code or -- through the drop-in of
// integration/ -- by libsqgpu.so, and that log every cost / gradient evaluation the optimizer makes.
//
// TEST INFRASTRUCTURE ONLY (compiled by `make -C oracle ref_gpu` into oracle/_ref/libsqref_gpu.so together with the reference
// translation units and integration/common_GPU.cpp). tests/test_reference_dropin.py uses it for the teacher-forced trajectory
// check of BASELINE.json's north star: every iterate the reference's optimizer visits while driven by the GPU is re-evaluated
// by the reference's CPU path and must agree within 1e-10.
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../include/sqgpu.h"
#include "../integration/GPU_Cost_Path_Mixin.h"

#include "Adam.h"
#include "Gates_block.h"
#include "N_Qubit_Decomposition_custom.h"

extern "C" void scipy_openblas_set_num_threads(int);

// defined in ref_harness.cpp (same library)
extern "C" void* sqref_circuit_create(int qbit_num, const sqgpu_gate_desc* descs, int n, const double* pool);
extern "C" void sqref_circuit_free(void* c);

namespace {

thread_local std::string g_err;

struct EvalLog {
    std::vector<double> params, cost, grad;  // row-major [n_evals][P], [n_evals], [n_evals][P]
    int n_params = 0;
    void add(Matrix_real& p, double f, Matrix_real& g) {
        n_params = (int)p.size();
        params.insert(params.end(), p.get_data(), p.get_data() + p.size());
        cost.push_back(f);
        grad.insert(grad.end(), g.get_data(), g.get_data() + g.size());
    }
};

// protected members the harness needs, for either flavour
template <class Base>
struct Access : public Base {
    using Base::Base;
    EvalLog log;
    bool logging = false;
    void set_prev_cost(double v) { this->prev_cost_fnv_val = v; }
    void set_scales(double c1, double c2) { this->correction1_scale = c1; this->correction2_scale = c2; }
    void cfg_int(const char* key, long long v) {
        Config_Element e;
        e.set_property(key, v);
        this->config[key] = e;
    }
    void cfg_double(const char* key, double v) {
        Config_Element e;
        e.set_property(key, v);
        this->config[key] = e;
    }
    void optimization_problem_combined_non_static(Matrix_real parameters, void* void_instance, double* f0, Matrix_real& grad) override {
        Base::optimization_problem_combined_non_static(parameters, void_instance, f0, grad);
        Access* self = reinterpret_cast<Access*>(void_instance);
        if (self->logging) self->log.add(parameters, *f0, grad);
    }
};

typedef Access<N_Qubit_Decomposition_custom> CpuDecomp;
typedef Access<sqgpu_bridge::With_GPU_Cost_Path<N_Qubit_Decomposition_custom>> GpuDecomp;

struct Session {
    bool gpu = false;
    std::map<std::string, Config_Element> config;
    CpuDecomp* cpu = nullptr;
    GpuDecomp* dev = nullptr;
    Optimization_Interface* base() { return gpu ? static_cast<Optimization_Interface*>(dev) : static_cast<Optimization_Interface*>(cpu); }
    EvalLog& log() { return gpu ? dev->log : cpu->log; }
};

template <typename F>
int guarded(F f) {
    try {
        f();
        return 0;
    } catch (std::string& e) {
        g_err = e;
    } catch (const char* e) {
        g_err = e;
    } catch (std::exception& e) {
        g_err = e.what();
    } catch (...) {
        g_err = "unknown exception";
    }
    return -1;
}

Matrix wrap_copy(const double* data, int rows, int cols) {
    Matrix m(rows, cols);
    for (int r = 0; r < rows; ++r) memcpy(m.get_data() + (size_t)r * m.stride, data + 2 * (size_t)r * cols, sizeof(QGD_Complex16) * cols);
    return m;
}

}  // namespace

extern "C" {

const char* sqrefgpu_last_error() { return g_err.c_str(); }

int sqrefgpu_set_library_path(const char* path) {
    return guarded([&] { sqgpu_bridge::set_library_path(path); });
}

int sqrefgpu_available_gpus() {
    int n = 0;
    guarded([&] { n = sqgpu_bridge::available_gpus(); });
    return n;
}

// use_gpu = 1: With_GPU_Cost_Path<N_Qubit_Decomposition_custom>; 0: the plain reference class
void* sqrefgpu_session_create(int use_gpu, const double* umtx, int rows, int cols, int qbit_num, const sqgpu_gate_desc* descs,
                              int n_descs, const double* pool) {
    Session* s = nullptr;
    int rc = guarded([&] {
        scipy_openblas_set_num_threads(1);
        s = new Session();
        s->gpu = use_gpu != 0;
        Matrix U = wrap_copy(umtx, rows, cols);
        Gates_block* blk = reinterpret_cast<Gates_block*>(sqref_circuit_create(qbit_num, descs, n_descs, pool));
        if (!blk) throw std::string("ref_gpu_harness: cannot build the gate structure");
        if (s->gpu) {
            s->dev = new GpuDecomp(U, qbit_num, false, s->config, ZEROS, 0);
            s->dev->set_verbose(0);
            s->dev->set_custom_gate_structure(blk);
        } else {
            s->cpu = new CpuDecomp(U, qbit_num, false, s->config, ZEROS, 0);
            s->cpu->set_verbose(0);
            s->cpu->set_custom_gate_structure(blk);
        }
        sqref_circuit_free(blk);
    });
    if (rc) {
        delete s;
        return nullptr;
    }
    return s;
}

void sqrefgpu_session_free(void* h) {
    Session* s = reinterpret_cast<Session*>(h);
    if (!s) return;
    delete s->cpu;
    delete s->dev;
    delete s;
}

int sqrefgpu_param_num(void* h) { return reinterpret_cast<Session*>(h)->base()->get_parameter_num(); }

int sqrefgpu_set_cost(void* h, int variant, int trace_offset, double prev_cost, double c1, double c2) {
    return guarded([&] {
        Session* s = reinterpret_cast<Session*>(h);
        s->base()->set_cost_function_variant((cost_function_type)variant);
        s->base()->set_trace_offset(trace_offset);
        if (s->gpu) { s->dev->set_prev_cost(prev_cost); s->dev->set_scales(c1, c2); }
        else { s->cpu->set_prev_cost(prev_cost); s->cpu->set_scales(c1, c2); }
    });
}

// one evaluation through the class's own virtual dispatch (what an optimizer would call)
int sqrefgpu_cost_grad(void* h, const double* params, int n_params, double* cost, double* grad) {
    return guarded([&] {
        Session* s = reinterpret_cast<Session*>(h);
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        Matrix_real g(1, n_params);
        Optimization_Interface::optimization_problem_combined(p, s->base(), cost, g);
        memcpy(grad, g.get_data(), sizeof(double) * n_params);
    });
}

int sqrefgpu_cost(void* h, const double* params, int n_params, double* cost) {
    return guarded([&] {
        Session* s = reinterpret_cast<Session*>(h);
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        *cost = s->base()->optimization_problem(p);
    });
}

// the non-virtual batched hook: the GPU flavour calls the body gpu_hooks.patch puts behind #ifdef __GPU__
int sqrefgpu_cost_batched(void* h, const double* params, int n_params, int batch, double* cost) {
    return guarded([&] {
        Session* s = reinterpret_cast<Session*>(h);
        std::vector<Matrix_real> vec;
        for (int b = 0; b < batch; ++b) {
            Matrix_real p(1, n_params);
            memcpy(p.get_data(), params + (size_t)b * n_params, sizeof(double) * n_params);
            vec.push_back(p);
        }
        Matrix_real res = s->gpu ? s->dev->optimization_problem_batched_GPU(vec) : s->cpu->optimization_problem_batched(vec);
        for (int b = 0; b < batch; ++b) cost[b] = res[b];
    });
}

// Run the reference's optimizer: alg = 0 ADAM, 1 BFGS (enum optimization_aglorithms, Optimization_Interface.h:50), starting
// from x0, at most max_inner_iterations inner iterations, one outer loop. Every cost+gradient evaluation is logged.
// Returns the number of logged evaluations (< 0 on error); x_out = the optimizer's final parameters, f_out its minimum.
int sqrefgpu_optimize(void* h, int alg, const double* x0, int n_params, long long max_inner_iterations, double eta, double* x_out,
                      double* f_out) {
    int n_evals = -1;
    int rc = guarded([&] {
        Session* s = reinterpret_cast<Session*>(h);
        auto setup = [&](auto* d) {
            d->cfg_int("max_inner_iterations", max_inner_iterations);
            d->cfg_int("max_iteration_loops", 1);
            d->cfg_double("eta", eta);
            d->cfg_double("optimization_tolerance", 1e-30);  // never stop early: both flavours make the same number of steps
            d->log = EvalLog();
            d->logging = true;
        };
        if (s->gpu) setup(s->dev); else setup(s->cpu);
        s->base()->set_optimizer((optimization_aglorithms)alg);
        Matrix_real guess(n_params, 1);
        memcpy(guess.get_data(), x0, sizeof(double) * n_params);
        s->base()->solve_layer_optimization_problem(n_params, guess);
        if (s->gpu) s->dev->logging = false; else s->cpu->logging = false;
        Matrix_real opt = s->base()->get_optimized_parameters();
        memcpy(x_out, opt.get_data(), sizeof(double) * n_params);
        *f_out = s->base()->get_current_minimum();
        n_evals = (int)s->log().cost.size();
    });
    return rc ? -1 : n_evals;
}

// copies of the evaluation log of the last sqrefgpu_optimize: params [n][P], cost [n], grad [n][P]
int sqrefgpu_get_log(void* h, double* params, double* cost, double* grad) {
    return guarded([&] {
        EvalLog& L = reinterpret_cast<Session*>(h)->log();
        memcpy(params, L.params.data(), sizeof(double) * L.params.size());
        memcpy(cost, L.cost.data(), sizeof(double) * L.cost.size());
        memcpy(grad, L.grad.data(), sizeof(double) * L.grad.size());
    });
}

// The shim's flattening on its own (no device needed): the reference Gates_block built from a nested descriptor stream goes
// through sqgpu_bridge::to_gpu_gates and must come back as the flat stream the Python mirror produces. Returns the number
// of gates (< 0 on error); pool_out receives the constant kernels (complex, interleaved), *pool_len their element count.
int sqrefgpu_flatten(int qbit_num, const sqgpu_gate_desc* descs, int n_descs, const double* pool, sqgpu_gate_desc* out, int cap,
                     double* pool_out, long long pool_cap, long long* pool_len) {
    int n = -1;
    int rc = guarded([&] {
        Gates_block* blk = reinterpret_cast<Gates_block*>(sqref_circuit_create(qbit_num, descs, n_descs, pool));
        if (!blk) throw std::string("ref_gpu_harness: cannot build the gate structure");
        std::vector<QGD_Complex16> pl;
        std::vector<sqgpu_gate_desc> flat = sqgpu_bridge::to_gpu_gates(blk, pl);
        const uint64_t fp1 = sqgpu_bridge::fingerprint(blk), fp2 = sqgpu_bridge::fingerprint(blk);
        sqref_circuit_free(blk);
        if (fp1 != fp2) throw std::string("ref_gpu_harness: fingerprint is not deterministic");
        if ((int)flat.size() > cap || (long long)pl.size() > pool_cap) throw std::string("ref_gpu_harness: output buffers too small");
        memcpy(out, flat.data(), sizeof(sqgpu_gate_desc) * flat.size());
        if (!pl.empty()) memcpy(pool_out, pl.data(), sizeof(QGD_Complex16) * pl.size());
        *pool_len = (long long)pl.size();
        n = (int)flat.size();
    });
    return rc ? -1 : n;
}

// ---- the reference's own Adam class (common/Adam.cpp), to pin the host mirror sqo_adam_update and the device kernel ------------
// (run with SQREF_SERIAL=1: the reference advances its bias-correction members inside a TBB parallel_for, Adam.cpp:219-245, which
// is only deterministic when that loop runs on one thread)
void* sqrefgpu_adam_create(double beta1, double beta2, double epsilon, double eta, int n_params) {
    Adam* a = nullptr;
    guarded([&] {
        a = new Adam(beta1, beta2, epsilon, eta);
        a->initialize_moment_and_variance(n_params);
    });
    return a;
}

void sqrefgpu_adam_free(void* a) { delete reinterpret_cast<Adam*>(a); }

int sqrefgpu_adam_update(void* a, double* params, const double* grad, int n_params, double f0) {
    int status = -1;
    guarded([&] {
        Matrix_real p(params, n_params, 1);  // wraps the caller's buffer: updated in place
        Matrix_real g(n_params, 1);
        memcpy(g.get_data(), grad, sizeof(double) * n_params);
        status = reinterpret_cast<Adam*>(a)->update(p, g, f0);
    });
    return status;
}

// ---- the reference's binary gate-list format (Gates_block.cpp:4807-5330), for the round-trip tests of gate_io.py --------------
int sqrefgpu_export_binary(int qbit_num, const sqgpu_gate_desc* descs, int n_descs, const double* params, int n_params, const char* filename) {
    return guarded([&] {
        Gates_block* blk = reinterpret_cast<Gates_block*>(sqref_circuit_create(qbit_num, descs, n_descs, nullptr));
        if (!blk) throw std::string("ref_gpu_harness: cannot build the gate structure");
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        export_gate_list_to_binary(p, blk, std::string(filename), 0);
        sqref_circuit_free(blk);
    });
}

// (the reference's own import_gate_list_from_binary is not exposed: it indexes gate_block_levels[-1] when the top-level block
// completes, Gates_block.cpp:5296-5303, and crashes on files its own exporter wrote -- reference defect 7 in DESIGN.md)

long long sqrefgpu_gpu_evaluations(void* h) {
    Session* s = reinterpret_cast<Session*>(h);
    return s->gpu ? s->dev->gpu_evaluations() : 0;
}

}  // extern "C"
