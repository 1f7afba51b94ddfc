/*
 * sq_oracle.h -- CPU restatement (plain C11) of the SQUANDER decomposition hot path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker. The product (sequential-quantum-gate-decomposer_b200/) never links, loads or calls it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py checks every function here against the reference's own
 * translation units compiled into oracle/_ref/libsqref.so (in this container), and tests/test_oracle_golden.py
 * checks it against fixtures generated from that library (tests/golden/, generator committed beside them).
 *
 * All matrices are row-major interleaved complex128; see include/sqgpu.h for the descriptor and conventions.
 */
#ifndef SQ_ORACLE_H
#define SQ_ORACLE_H

#include <stdint.h>
#include "../include/sqgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Forward kernel of one gate from its own parameter slice. Writes dim x dim complex (dim = 2, or 4 for
 * RXX/RYY/RZZ/SWAP-like two-target gates) into `kernel`; returns dim, or <0 if the type has no dense kernel here. */
int sqo_gate_kernel(int type, const double* gate_params, double* kernel);

/* Derivative kernel with respect to the gate's pidx-th parameter (same shape as the forward kernel). */
int sqo_gate_derivative_kernel(int type, const double* gate_params, int pidx, double* kernel);

/* number of parameters a gate type takes, -1 for unknown types */
int sqo_gate_param_count(int type);

/* [a;b] <- K [a;b] on every row pair of bit `target`, rows whose `control` bits are set; deriv != 0 zero-fills the
 * inactive rows (kernels/apply_kernel_to_input.cpp:33-115). control / control2 = -1 when absent. */
void sqo_apply_kernel_to_input(const double* k2x2, double* input, int rows, int cols, int stride, int deriv,
                               int target, int control, int control2);

/* dense 2^k x 2^k kernel on ascending qubits (kernels/apply_large_kernel_to_input.cpp:123-213) with optional control */
void sqo_apply_large_kernel_to_input(const double* kernel, double* input, int rows, int cols, int stride,
                                     const int* qubits, int k, int control, int deriv);

/* one gate of a descriptor list (params = the circuit's whole parameter vector); deriv_param < 0: the gate itself */
int sqo_apply_gate(const sqgpu_gate_desc* g, const double* params, const double* pool, int deriv_param, double* input,
                   int rows, int cols, int stride);

/* Gates_block::apply_to (Gates_block.cpp:605-710) on a flattened descriptor list */
int sqo_apply_circuit(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool,
                      double* input, int rows, int cols, int stride);

/* Gates_block::apply_derivate_to (Gates_block.cpp:1011-1150), prefix-only route: out = n_params compact matrices */
int sqo_apply_derivate(const sqgpu_gate_desc* gates, int n_gates, int n_params, const double* params,
                       const double* pool, const double* input, int rows, int cols, int stride, double* out);

/* trace terms of N_Qubit_Decomposition_Cost_Function.cpp:73-162,191-404,482-664:
 * out[2*t + {0,1}] = {Re, Im} of sum_j M[(j+off) ^ mask, j] over t = 0 (mask 0), 1 (all one-bit masks),
 * 2 (all two-bit masks). */
void sqo_traces(const double* mtx, int rows, int cols, int stride, int qbit_num, int trace_offset, double* out6);

/* cost from trace terms (Optimization_Interface::calculate_cost_function, Optimization_Interface.cpp:677-735) */
double sqo_cost_from_traces(int variant, const double* tr6, int cols, double prev_cost, double c1, double c2);

/* gradient component from the trace terms of the circuit (tr6) and of one derivative matrix (dtr6)
 * (Optimization_Interface.cpp:1397-1458) */
double sqo_grad_from_traces(int variant, const double* tr6, const double* dtr6, int cols, double prev_cost, double c1,
                            double c2);

/* SUM_OF_SQUARES cost (N_Qubit_Decomposition_Cost_Function.cpp:443-458) */
double sqo_cost_sum_of_squares(const double* mtx, int rows, int cols, int stride);

/* Optimization_Interface::optimization_problem (Optimization_Interface.cpp:634-668) */
int sqo_cost(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool, const double* umtx,
             int rows, int cols, int stride, int qbit_num, int variant, int trace_offset, double prev_cost, double c1,
             double c2, double* cost);

/* Optimization_Interface::optimization_problem_combined_non_static (Optimization_Interface.cpp:1145-1490) */
int sqo_cost_grad(const sqgpu_gate_desc* gates, int n_gates, int n_params, const double* params, const double* pool,
                  const double* umtx, int rows, int cols, int stride, int qbit_num, int variant, int trace_offset,
                  double prev_cost, double c1, double c2, double* cost, double* grad);

/* CSR SpMV y = H x  (common/common.cpp:403-436) and Re<x|y> */
void sqo_csr_matvec(int n_rows, const int32_t* indptr, const int32_t* indices, const double* values, const double* x,
                    double* y);

/* VQE energy Re<psi|H|psi>, psi = C(params) state0  (Variational_Quantum_Eigensolver_Base.cpp:584-624,1088-1121) */
int sqo_vqe_energy(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool,
                   const double* state0, int n_rows, const int32_t* indptr, const int32_t* indices,
                   const double* values, double* energy);

/* VQE energy + gradient, grad_i = 2 Re <d_i psi|H|psi> (…Base.cpp:1131-1199) */
int sqo_vqe_energy_grad(const sqgpu_gate_desc* gates, int n_gates, int n_params, const double* params,
                        const double* pool, const double* state0, int n_rows, const int32_t* indptr,
                        const int32_t* indices, const double* values, double* energy, double* grad);

/* energy and the gradient entries of the parameters in sample[0..n_sample) only (same derivative route) */
int sqo_vqe_energy_grad_sampled(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool,
                                const double* state0, int n_rows, const int32_t* indptr, const int32_t* indices,
                                const double* values, const int32_t* sample, int n_sample, double* energy, double* grad);

/* Adam::update (common/Adam.cpp:120-262), sequential semantics; host mirror of the device-resident ADAM loop */
typedef struct sqo_adam_state {
    double beta1_t, beta2_t, f0_mean, decreasing_test, f0_prev;
    int f0_idx, decreasing_idx, iter_t;
    double f0_vec[100];
    int decreasing_vec[20];
} sqo_adam_state;
void sqo_adam_reset(sqo_adam_state* s);
int sqo_adam_update(sqo_adam_state* s, double* params, const double* grad, double* mom, double* var, int n, double f0,
                    double eta, double beta1, double beta2, double epsilon);

#ifdef __cplusplus
}
#endif
#endif
