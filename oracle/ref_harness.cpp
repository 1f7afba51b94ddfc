// ref_harness.cpp -- C entry points around the UNMODIFIED reference classes (compiled from /root/reference by
// oracle/Makefile `make ref` into oracle/_ref/libsqref.so).
//
// TEST INFRASTRUCTURE ONLY. Used (a) to pin the C restatement in sq_oracle.c, (b) to generate the fixtures under
// tests/golden/ (tests/golden/make_golden.py), (c) as bench.py's CPU baseline (`cpu_baseline.kind = "reference"`,
// `--impl reference`). Nothing in the product path links or loads this file.
//
// The harness only *calls* the reference; it takes the same sqgpu_gate_desc stream as the engine
// (include/sqgpu.h) plus SQGPU_BLOCK_BEGIN/END markers so the reference's nested Gates_block layout
// (N_Qubit_Decomposition_adaptive.cpp:1871-1881, 1947-1966) can be rebuilt exactly.

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../include/sqgpu.h"

#include "Gates_block.h"
#include "N_Qubit_Decomposition_custom.h"
#include "N_Qubit_Decomposition_Cost_Function.h"
#include "Variational_Quantum_Eigensolver_Base.h"
#include "matrix_sparse.h"

// The reference runs BLAS single-threaded inside its own TBB tasks (dot(): openblas_set_num_threads(1) / MKL
// equivalents, common/dot.cpp:40-50). The build here leaves BLAS "undefined" (=0), so pin the scipy-bundled OpenBLAS
// to one thread once; parallelism comes from the parallel_for shim exactly as TBB would provide it.
extern "C" void scipy_openblas_set_num_threads(int);

namespace {

thread_local std::string g_err;

struct BlasInit {
    BlasInit() { scipy_openblas_set_num_threads(1); }
} g_blas_init;

struct DecompAccess : public N_Qubit_Decomposition_custom {
    using N_Qubit_Decomposition_custom::N_Qubit_Decomposition_custom;
    void set_prev_cost(double v) { prev_cost_fnv_val = v; }
    void set_scales(double c1, double c2) { correction1_scale = c1; correction2_scale = c2; }
    void set_parallel(int p) {
        Config_Element e;
        e.set_property("parallel", (long long)p);  // read back as long long (Decomposition_Base.cpp:1187-1190)
        config["parallel"] = e;
    }
};

struct RefDecomp {
    DecompAccess* dec;
    std::map<std::string, Config_Element> config;
};

Matrix wrap_copy(const double* data, int rows, int cols, int stride) {
    Matrix m(rows, cols);
    for (int r = 0; r < rows; ++r)
        memcpy(m.get_data() + (size_t)r * m.stride, data + 2 * (size_t)r * stride, sizeof(QGD_Complex16) * cols);
    return m;
}

void add_one(Gates_block* blk, const sqgpu_gate_desc& d, int qbit_num, const double* pool) {
    switch (d.type) {
        case SQGPU_U3: blk->add_u3(d.target); break;
        case SQGPU_RX: blk->add_rx(d.target); break;
        case SQGPU_RY: blk->add_ry(d.target); break;
        case SQGPU_RZ: blk->add_rz(d.target); break;
        case SQGPU_U1: blk->add_u1(d.target); break;
        case SQGPU_U2: blk->add_u2(d.target); break;
        case SQGPU_R: blk->add_r(d.target); break;
        case SQGPU_CRY: blk->add_cry(d.target, d.control); break;
        case SQGPU_CRX: blk->add_crx(d.target, d.control); break;
        case SQGPU_CRZ: blk->add_crz(d.target, d.control); break;
        case SQGPU_CP: blk->add_cp(d.target, d.control); break;
        case SQGPU_CR: blk->add_cr(d.target, d.control); break;
        case SQGPU_CU: blk->add_cu(d.target, d.control); break;
        case SQGPU_CROT: blk->add_crot(d.target, d.control); break;
        case SQGPU_ADAPTIVE: blk->add_adaptive(d.target, d.control); break;
        case SQGPU_CNOT: blk->add_cnot(d.target, d.control); break;
        case SQGPU_CZ: blk->add_cz(d.target, d.control); break;
        case SQGPU_CH: blk->add_ch(d.target, d.control); break;
        case SQGPU_SYC: blk->add_syc(d.target, d.control); break;
        case SQGPU_X: blk->add_x(d.target); break;
        case SQGPU_Y: blk->add_y(d.target); break;
        case SQGPU_Z: blk->add_z(d.target); break;
        case SQGPU_H: blk->add_h(d.target); break;
        case SQGPU_S: blk->add_s(d.target); break;
        case SQGPU_SDG: blk->add_sdg(d.target); break;
        case SQGPU_T: blk->add_t(d.target); break;
        case SQGPU_TDG: blk->add_tdg(d.target); break;
        case SQGPU_SX: blk->add_sx(d.target); break;
        case SQGPU_SXDG: blk->add_sxdg(d.target); break;
        case SQGPU_CCX: blk->add_ccx(d.target, std::vector<int>{d.control, d.control2}); break;
        case SQGPU_SWAP: blk->add_swap(std::vector<int>{d.target, d.target2}); break;
        case SQGPU_CSWAP: blk->add_cswap(std::vector<int>{d.target, d.target2}, std::vector<int>{d.control}); break;
        case SQGPU_RXX: blk->add_rxx(std::vector<int>{d.target, d.target2}); break;
        case SQGPU_RYY: blk->add_ryy(std::vector<int>{d.target, d.target2}); break;
        case SQGPU_RZZ: blk->add_rzz(std::vector<int>{d.target, d.target2}); break;
        case SQGPU_GENERAL: {
            const int dim = 1 << d.n_qubits;
            Matrix km = wrap_copy(pool + 2 * d.matrix_off, dim, dim, dim);
            std::vector<int> tq(d.qubits, d.qubits + d.n_qubits);
            Gate* g = new Gate(qbit_num);
            g->set_matrix(km);
            g->set_target_qbits(tq);
            blk->add_gate(g);
            break;
        }
        default:
            throw std::string("ref_harness: unsupported gate type ") + std::to_string(d.type);
    }
}

// builds gates[*pos ...] into blk until the matching BLOCK_END (or the end of the stream at depth 0)
void build_block(Gates_block* blk, const sqgpu_gate_desc* descs, int n, int* pos, int qbit_num, const double* pool,
                 int depth) {
    while (*pos < n) {
        const sqgpu_gate_desc& d = descs[*pos];
        if (d.type == SQGPU_BLOCK_BEGIN) {
            ++*pos;
            Gates_block* sub = new Gates_block(qbit_num);
            build_block(sub, descs, n, pos, qbit_num, pool, depth + 1);
            blk->add_gate(sub);
        } else if (d.type == SQGPU_BLOCK_END) {
            ++*pos;
            if (depth == 0) throw std::string("ref_harness: unbalanced BLOCK_END");
            return;
        } else {
            add_one(blk, d, qbit_num, pool);
            ++*pos;
        }
    }
    if (depth != 0) throw std::string("ref_harness: missing BLOCK_END");
}

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

}  // namespace

extern "C" {

const char* sqref_last_error() { return g_err.c_str(); }

// ---- Circuit (Gates_block) --------------------------------------------------------------------------------------

void* sqref_circuit_create(int qbit_num, const sqgpu_gate_desc* descs, int n, const double* pool) {
    Gates_block* blk = nullptr;
    int rc = guarded([&] {
        blk = new Gates_block(qbit_num);
        int pos = 0;
        build_block(blk, descs, n, &pos, qbit_num, pool, 0);
    });
    return rc == 0 ? blk : nullptr;
}

void sqref_circuit_free(void* c) { delete reinterpret_cast<Gates_block*>(c); }

int sqref_circuit_param_num(void* c) { return reinterpret_cast<Gates_block*>(c)->get_parameter_num(); }

int sqref_circuit_set_min_fusion(void* c, int mf) {
    return guarded([&] { reinterpret_cast<Gates_block*>(c)->set_min_fusion(mf); });
}

// in-place Gates_block::apply_to on a compact rows x cols matrix
int sqref_circuit_apply(void* c, const double* params, int n_params, double* inout, int rows, int cols, int parallel) {
    return guarded([&] {
        Gates_block* blk = reinterpret_cast<Gates_block*>(c);
        Matrix m = wrap_copy(inout, rows, cols, cols);
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        blk->apply_to(p, m, parallel);
        for (int r = 0; r < rows; ++r)
            memcpy(inout + 2 * (size_t)r * cols, m.get_data() + (size_t)r * m.stride, sizeof(QGD_Complex16) * cols);
    });
}

// Gates_block::apply_derivate_to; out = P compact matrices
int sqref_circuit_apply_derivate(void* c, const double* params, int n_params, const double* in, int rows, int cols,
                                 int parallel, double* out) {
    return guarded([&] {
        Gates_block* blk = reinterpret_cast<Gates_block*>(c);
        Matrix m = wrap_copy(in, rows, cols, cols);
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        std::vector<Matrix> res = blk->apply_derivate_to(p, m, parallel);
        if ((int)res.size() != n_params) throw std::string("ref_harness: derivative count mismatch");
        for (int i = 0; i < n_params; ++i)
            for (int r = 0; r < rows; ++r)
                memcpy(out + 2 * ((size_t)i * rows + r) * cols, res[i].get_data() + (size_t)r * res[i].stride,
                       sizeof(QGD_Complex16) * cols);
    });
}

// ---- decomposition object (cost path) ---------------------------------------------------------------------------

void* sqref_decomp_create(const double* umtx, int rows, int cols, int qbit_num, const sqgpu_gate_desc* descs, int n,
                          const double* pool) {
    RefDecomp* rd = nullptr;
    int rc = guarded([&] {
        rd = new RefDecomp();
        Matrix U = wrap_copy(umtx, rows, cols, cols);
        rd->dec = new DecompAccess(U, qbit_num, false, rd->config, ZEROS, 0);
        rd->dec->set_verbose(0);
        Gates_block* blk = new Gates_block(qbit_num);
        int pos = 0;
        build_block(blk, descs, n, &pos, qbit_num, pool, 0);
        rd->dec->set_custom_gate_structure(blk);
        delete blk;
    });
    if (rc != 0) return nullptr;
    return rd;
}

void sqref_decomp_free(void* h) {
    RefDecomp* rd = reinterpret_cast<RefDecomp*>(h);
    if (!rd) return;
    delete rd->dec;
    delete rd;
}

int sqref_decomp_param_num(void* h) { return reinterpret_cast<RefDecomp*>(h)->dec->get_parameter_num(); }

int sqref_decomp_set_cost(void* h, int variant, int trace_offset, double prev_cost, double c1, double c2) {
    return guarded([&] {
        DecompAccess* d = reinterpret_cast<RefDecomp*>(h)->dec;
        d->set_cost_function_variant((cost_function_type)variant);
        d->set_trace_offset(trace_offset);
        d->set_prev_cost(prev_cost);
        d->set_scales(c1, c2);
    });
}

// parallel: 0 sequential, 1 OpenMP, 2 TBB (Decomposition_Base::get_parallel_configuration reads config["parallel"])
int sqref_decomp_set_parallel(void* h, int parallel) {
    return guarded([&] {
        reinterpret_cast<RefDecomp*>(h)->dec->set_parallel(parallel);
    });
}

int sqref_decomp_cost(void* h, const double* params, int n_params, double* cost) {
    return guarded([&] {
        DecompAccess* d = reinterpret_cast<RefDecomp*>(h)->dec;
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        *cost = d->optimization_problem(p);
    });
}

int sqref_decomp_cost_batched(void* h, const double* params, int n_params, int batch, double* cost) {
    return guarded([&] {
        DecompAccess* d = reinterpret_cast<RefDecomp*>(h)->dec;
        std::vector<Matrix_real> vec;
        for (int b = 0; b < batch; ++b) {
            Matrix_real p(1, n_params);
            memcpy(p.get_data(), params + (size_t)b * n_params, sizeof(double) * n_params);
            vec.push_back(p);
        }
        Matrix_real res = d->optimization_problem_batched(vec);
        for (int b = 0; b < batch; ++b) cost[b] = res[b];
    });
}

int sqref_decomp_cost_grad(void* h, const double* params, int n_params, double* cost, double* grad) {
    return guarded([&] {
        DecompAccess* d = reinterpret_cast<RefDecomp*>(h)->dec;
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        Matrix_real g(1, n_params);
        d->optimization_problem_combined(p, d, cost, g);
        memcpy(grad, g.get_data(), sizeof(double) * n_params);
    });
}

// ---- VQE (state-vector) object ---------------------------------------------------------------------------------------

struct RefVQE {
    Variational_Quantum_Eigensolver_Base* vqe;
    std::map<std::string, Config_Element> config;
    std::vector<QGD_Complex16> hdata;
    std::vector<int> hind, hptr;
};

// ansatz: 0 HEA, 1 HEA_ZYZ (generate_circuit(layers, inner_blocks)); descs != NULL: custom gate structure instead
void* sqref_vqe_create(int qbit_num, int n_rows, int nnz, const int* indptr, const int* indices, const double* values,
                       int ansatz, int layers, int inner_blocks, const sqgpu_gate_desc* descs, int n_descs) {
    RefVQE* rv = nullptr;
    int rc = guarded([&] {
        rv = new RefVQE();
        rv->hdata.resize(nnz);
        memcpy(rv->hdata.data(), values, sizeof(QGD_Complex16) * nnz);
        rv->hind.assign(indices, indices + nnz);
        rv->hptr.assign(indptr, indptr + n_rows + 1);
        Matrix_sparse H(rv->hdata.data(), n_rows, n_rows, nnz, rv->hind.data(), rv->hptr.data());
        rv->vqe = new Variational_Quantum_Eigensolver_Base(H, qbit_num, rv->config, 0);
        rv->vqe->set_verbose(0);
        // |0...0> given explicitly: the reference's own initialize_zero_state() offsets a QGD_Complex16* by 2 elements
        // while counting doubles (…Base.cpp:1269-1274), leaving amplitude 1 uninitialised and overrunning the buffer.
        Matrix psi0(1 << qbit_num, 1);
        memset(psi0.get_data(), 0, sizeof(QGD_Complex16) * psi0.size());
        psi0[0].real = 1.0;
        rv->vqe->set_initial_state(psi0);
        if (descs) {
            Gates_block* blk = new Gates_block(qbit_num);
            int pos = 0;
            build_block(blk, descs, n_descs, &pos, qbit_num, nullptr, 0);
            rv->vqe->set_custom_gate_structure(blk);
            delete blk;
        } else {
            rv->vqe->set_ansatz(ansatz == 1 ? HEA_ZYZ : HEA);
            rv->vqe->generate_circuit(layers, inner_blocks);
        }
    });
    if (rc != 0) return nullptr;
    return rv;
}

void sqref_vqe_free(void* h) {
    RefVQE* rv = reinterpret_cast<RefVQE*>(h);
    if (!rv) return;
    delete rv->vqe;
    delete rv;
}

int sqref_vqe_param_num(void* h) { return reinterpret_cast<RefVQE*>(h)->vqe->get_parameter_num(); }

int sqref_vqe_energy(void* h, const double* params, int n_params, double* energy) {
    return guarded([&] {
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        *energy = reinterpret_cast<RefVQE*>(h)->vqe->optimization_problem(p);
    });
}

int sqref_vqe_energy_grad(void* h, const double* params, int n_params, double* energy, double* grad) {
    return guarded([&] {
        Variational_Quantum_Eigensolver_Base* v = reinterpret_cast<RefVQE*>(h)->vqe;
        Matrix_real p(1, n_params);
        memcpy(p.get_data(), params, sizeof(double) * n_params);
        Matrix_real g(1, n_params);
        v->optimization_problem_combined_non_static(p, v, energy, g);
        memcpy(grad, g.get_data(), sizeof(double) * n_params);
    });
}

// ---- standalone cost functions on a given matrix (N_Qubit_Decomposition_Cost_Function.cpp) -----------------------

// out[0..5]: Re/Im main trace (with offset), Re/Im one-bit-flip sum, Re/Im two-bit-flip sum, as the reference
// routines return them (get_trace*, get_cost_function_with_correction2 give the real parts with offset).
int sqref_traces(const double* mtx, int rows, int cols, int qbit_num, int trace_offset, double* out) {
    return guarded([&] {
        Matrix m = wrap_copy(mtx, rows, cols, cols);
        Matrix_real fr = get_cost_function_with_correction2(m, qbit_num, trace_offset);
        out[0] = fr[0];
        out[1] = fr[1];
        out[2] = fr[2];
        if (trace_offset == 0 && rows >= cols) {
            Matrix tr = get_trace_with_correction2(m, qbit_num);
            for (int i = 0; i < 3; ++i) {
                out[3 + 2 * i] = tr[i].real;
                out[4 + 2 * i] = tr[i].imag;
            }
        }
    });
}

}  // extern "C"
