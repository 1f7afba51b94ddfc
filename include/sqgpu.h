/*
 * sqgpu.h -- C-ABI of the B200 (sm_100a) decomposition hot-path engine.
 *
 * This is the drop-in boundary for ONE path of SQUANDER (rakytap/sequential-quantum-gate-decomposer):
 * apply a parametrised gate structure to a 2^n x C complex128 matrix (C = 1: state vector), reduce the
 * trace cost and its parameter gradient, for a batch of parameter vectors.
 *
 * The boundary mirrors the accelerator plug-in the reference already has for Maxeler DFE boards
 * (dlopen'd C symbols, squander/src-cpp/common/common_DFE.cpp:47-53,160-166) -- same life cycle
 * (init / upload matrix / evaluate gate sets / release), but fp64 end to end and with the gate structure sent
 * once instead of once per call. Each entry point names the reference interface it replaces.
 *
 * Conventions
 *  - complex128 buffers are interleaved {re, im} doubles == QGD_Complex16 (common/include/QGDTypes.h:38-43),
 *    row-major, element (r, c) at data[2*(r*stride + c)]  (common/include/matrix_base.hpp:38-59).
 *  - every function returns SQGPU_OK (0) or a negative sqgpu_status; sqgpu_last_error() gives the text
 *    (the reference throws std::string, Gate.cpp:435-446; the host shim rethrows it -- see INTEGRATION.md).
 *  - host pointers are borrowed for the duration of the call only; the library owns all device memory.
 *  - a handle is bound to one CUDA device; calls on one handle are serialised by an internal mutex
 *    (the DFE bridge serialises with a rw-mutex, common_DFE.cpp:33,61-68), and the work they enqueue is ordered on
 *    the device too: every call makes its stream wait for the previous call's work, whatever stream that used, so
 *    the handle's workspaces are never shared by two calls in flight; uploads block until the handle is idle.
 *  - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *    SQGPU_ERR_NO_DEVICE.
 */
#ifndef SQGPU_H_INCLUDED
#define SQGPU_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQGPU_ABI_VERSION 2

typedef enum sqgpu_status {
    SQGPU_OK = 0,
    SQGPU_ERR_NO_DEVICE = -1,   /* no CUDA device / driver, or device index out of range            */
    SQGPU_ERR_INVALID = -2,     /* bad argument (sizes, qubit indices, unsupported gate type, ...)   */
    SQGPU_ERR_STATE = -3,       /* call order: matrix / circuit / hamiltonian not set yet            */
    SQGPU_ERR_CUDA = -4,        /* a CUDA runtime call failed; text in sqgpu_last_error()            */
    SQGPU_ERR_UNSUPPORTED = -5, /* valid in the reference, not implemented on the device path (yet)  */
    SQGPU_ERR_NOMEM = -6        /* device or pinned-host allocation failed                           */
} sqgpu_status;

/* Gate type codes: numerically identical to the reference enum gate_type
 * (squander/src-cpp/gates/include/Gate.h:39-79) so descriptors can be filled straight from Gate::get_type(). */
typedef enum sqgpu_gate_type {
    SQGPU_GENERAL = 1,
    SQGPU_CZ = 4,
    SQGPU_CNOT = 5,
    SQGPU_CH = 6,
    SQGPU_U3 = 7,
    SQGPU_RY = 8,
    SQGPU_RX = 9,
    SQGPU_RZ = 10,
    SQGPU_X = 12,
    SQGPU_SX = 13,
    SQGPU_CRY = 14,
    SQGPU_SYC = 15,
    SQGPU_BLOCK = 16, /* never sent to the engine: blocks are flattened (Gates_block::get_flat_circuit, Gates_block.cpp:3827-3856) */
    SQGPU_ADAPTIVE = 18,
    SQGPU_Y = 23,
    SQGPU_Z = 24,
    SQGPU_H = 25,
    SQGPU_CROT = 27,
    SQGPU_R = 28,
    SQGPU_T = 29,
    SQGPU_TDG = 30,
    SQGPU_U1 = 31,
    SQGPU_U2 = 32,
    SQGPU_CR = 33,
    SQGPU_S = 34,
    SQGPU_SDG = 35,
    SQGPU_CU = 36,
    SQGPU_CP = 38,
    SQGPU_CRX = 39,
    SQGPU_CRZ = 40,
    SQGPU_CCX = 41,
    SQGPU_SWAP = 42,
    SQGPU_CSWAP = 43,
    SQGPU_RXX = 44,
    SQGPU_RYY = 45,
    SQGPU_RZZ = 46,
    SQGPU_SXDG = 47,
    /* descriptor-stream markers understood only by the oracle harness (oracle/ref_harness.cpp) to rebuild the
     * reference's nested Gates_block structure; the engine rejects them. */
    SQGPU_BLOCK_BEGIN = 1001,
    SQGPU_BLOCK_END = 1002
} sqgpu_gate_type;

/* Cost-function variants: numerically identical to the reference enum cost_function_type
 * (squander/src-cpp/decomposition/include/Optimization_Interface.h:43-45). Every variant listed here is implemented on the
 * device path; the reference's OSR_ENTANGLEMENT (an SVD per cut) and its VQE / GQML tags are not cost variants of this path. */
typedef enum sqgpu_cost_variant {
    SQGPU_FROBENIUS_NORM = 0,
    SQGPU_FROBENIUS_NORM_CORRECTION1 = 1,
    SQGPU_FROBENIUS_NORM_CORRECTION2 = 2,
    SQGPU_HILBERT_SCHMIDT_TEST = 3,
    SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1 = 4,
    SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2 = 5,
    SQGPU_SUM_OF_SQUARES = 6,
    SQGPU_INFIDELITY = 9
} sqgpu_cost_variant;

#define SQGPU_MAX_GENERAL_QUBITS 5

/* One gate of the flattened circuit, in application order (gate 0 is applied first:
 * Gates_block::apply_to_inner, Gates_block.cpp:683-708). Replaces DFEgate_kernel_type
 * (common/include/common_DFE.h:62-70), which carried fixed-point angles; here the angles stay in the fp64
 * parameter vector and the descriptor only says where to read them. */
typedef struct sqgpu_gate_desc {
    int32_t type;        /* sqgpu_gate_type                                                                      */
    int32_t target;      /* target qubit (Gate::target_qbit); first target for SWAP/CSWAP/RXX/RYY/RZZ; -1 GENERAL */
    int32_t control;     /* control qubit (Gate::control_qbit) or -1                                              */
    int32_t target2;     /* second target (SWAP, CSWAP, RXX, RYY, RZZ) or -1                                       */
    int32_t control2;    /* second control (CCX) or -1                                                            */
    int32_t param_start; /* index of the gate's first parameter (Gate::parameter_start_idx)                       */
    int32_t n_params;    /* Gate::parameter_num; must match the type                                              */
    int32_t n_qubits;    /* GENERAL only: k = number of involved qubits, 1..SQGPU_MAX_GENERAL_QUBITS              */
    int32_t qubits[8];   /* GENERAL only: the k involved qubits, ascending; local-index bit j <-> qubits[j]
                            (apply_large_kernel_to_input.cpp:160-169)                                             */
    int64_t matrix_off;  /* GENERAL only: offset (complex elements) of the row-major 2^k x 2^k kernel in the pool */
} sqgpu_gate_desc;

typedef struct sqgpu_ctx* sqgpu_handle_t;

/* ---- life cycle ------------------------------------------------------------------------------------------- */

/* replaces get_accelerator_avail_num (common_DFE.cpp:47). *count = visible CUDA devices with cc >= 10.0. */
int sqgpu_device_count(int* count);

/* replaces initialize_DFE(accelerator_num) (common_DFE.cpp:52,178-184): bind a context to CUDA device `device`. */
int sqgpu_create(int device, sqgpu_handle_t* out);

/* replaces releive_DFE (common_DFE.cpp:49,110-113). */
int sqgpu_destroy(sqgpu_handle_t h);

/* replaces initialize_DFE(accelerator_num) for accelerator_num = G > 1 and the reference's split of the batched cost path over
 * accelerators / MPI ranks (Optimization_Interface.cpp:806-832, 962-1004): ONE handle over n_devices GPUs of this process
 * (devices = NULL: 0 .. n_devices-1). The handle takes the same calls as a single-device one -- sqgpu_upload_matrix,
 * sqgpu_set_circuit, sqgpu_set_cost, sqgpu_set_option, sqgpu_cost_batched, sqgpu_cost_grad_batched, sqgpu_set_hamiltonian_csr,
 * sqgpu_vqe_energy[_grad]_batched, sqgpu_destroy -- and shards the work itself:
 *   SQGPU_SHARD_BATCH    parameter vectors over the devices, results written straight into the caller's arrays;
 *   SQGPU_SHARD_COLUMNS  columns of the matrix over the devices, one ncclAllReduce of the raw trace terms per evaluation on
 *                        the devices' compute streams (two for the Hilbert-Schmidt correction variants);
 *   SQGPU_SHARD_AUTO     columns for matrices with >= 2048 columns (n >= 11), batch otherwise (and for state vectors).
 * The device-pointer (_dev), apply and raw-trace entry points are per device: they fail with SQGPU_ERR_UNSUPPORTED here. */
typedef enum sqgpu_shard_mode { SQGPU_SHARD_AUTO = 0, SQGPU_SHARD_BATCH = 1, SQGPU_SHARD_COLUMNS = 2 } sqgpu_shard_mode;
int sqgpu_create_multi(int n_devices, const int* devices, int mode, sqgpu_handle_t* out);
/* number of devices behind a handle (1 for sqgpu_create) and the sharding mode in force after the last upload */
int sqgpu_multi_info(sqgpu_handle_t h, int* n_devices, int* mode);

/* text of the last failure on the calling thread ("" if none). Never NULL. */
const char* sqgpu_last_error(void);

int sqgpu_abi_version(void);

/* ---- inputs ------------------------------------------------------------------------------------------------ */

/* replaces load2LMEM(QGD_Complex16*, rows, cols) (common_DFE.cpp:50,129-132) and the per-evaluation
 * Umtx.copy_to(matrix_new) (Optimization_Interface.cpp:661-663): upload the 2^n x cols matrix (cols = 1: state
 * vector) once; it stays resident until the next upload. */
int sqgpu_upload_matrix(sqgpu_handle_t h, const double* data, int rows, int cols, int stride);

/* replaces the per-call DFE descriptor flattening (Gates_block::convert_to_DFE_gates, Gates_block.cpp:4202-4278):
 * send the flattened gate structure once. `matrix_pool` (may be NULL) holds the constant kernels of GENERAL gates,
 * pool_len complex elements. */
int sqgpu_set_circuit(sqgpu_handle_t h, const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num,
                      const double* matrix_pool, int64_t pool_len);

/* What the host planner makes of a gate structure, WITHOUT a device (no handle): the same lowering, block fusion
 * (the device analogue of Gates_block's fusion rule, Gates_block.cpp:632-681) and window scheduling that
 * sqgpu_set_circuit runs. stats[0..SQGPU_PLAN_STATS): ops of the <=2-qubit plan, ops of the <=3-qubit plan, window
 * segments, ops of the largest segment, window width, complex elements per parameter set of the kernel / derivative-kernel
 * / W tables of the 3-qubit plan, raw dense 3-5 qubit ops, gates inside fused blocks. Validation errors are the ones
 * sqgpu_set_circuit reports. */
#define SQGPU_PLAN_STATS 10
int sqgpu_plan_stats(const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool,
                     int64_t pool_len, int64_t* stats, int n_stats);
/* the same with planner options, "name=value,name=value" (names of sqgpu_set_option), NULL = defaults */
int sqgpu_plan_stats_opt(const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool,
                         int64_t pool_len, const char* options, int64_t* stats, int n_stats);

/* The op list the planner produces, for inspection (docs, tests): ops[i*8 .. i*8+8) = {dim, q0..q4 (-1: unused), n_params,
 * n_members} of op i; which = 2 / 3: the <=2- / <=3-qubit block plans, 0: the window plan in segment order, 11 / 12 / 13: the
 * cluster plans for 2 / 4 / 8 CTAs (qubits are row-bit positions there; a RESPLIT op has dim 0 and {local row bit, cluster-rank
 * bit}; *n_ops = 0: the circuit has no cluster plan). No device needed. */
int sqgpu_plan_ops(const sqgpu_gate_desc* gates, int n_gates, int n_params, int qbit_num, const double* matrix_pool,
                   int64_t pool_len, const char* options, int which, int32_t* ops, int cap, int* n_ops);

/* Per-handle switches; the defaults are the product configuration. They are what the reference keeps in its `config` map
 * (Decomposition_Base::config, e.g. "use_float", "parallel"; Decomposition_Base.cpp:1115-1190) for this path. Planner
 * options take effect with the next sqgpu_set_circuit. Names:
 *   no_fuse (0/1)          every gate stays its own op (parity tests cover the raw-op paths of the executor with it)
 *   max_fuse_qubits (2/3)  widest fused block of the shared-memory executor
 *   fuse_consecutive (0/1) fuse runs of consecutive gates only (no commuting reorder)
 *   force_stream (0/1)     cost / gradient through the one-op-per-launch streaming executor
 *   vqe_stream (0/1)       VQE through the streaming kernels instead of the windowed executor
 *   window (1..30)         window width of the state-vector / tall-matrix segment planner (default 11)
 *   tall_window (0/1)      matrices whose column does not fit shared memory go through the windowed executor (default 1)
 *   async_tiles (0/1)      windowed executor, -DSQ_WIN_BULK=1 builds only: double-buffered tiles in the forward segments
 *   cluster (0/1/2)        thread-block-cluster executor (2 / 4 / 8 CTAs share a column over distributed shared memory): 0 off,
 *                          1 (default) where a column fits one CTA only once per SM (n = 12 gradient) or no window plan exists,
 *                          2 also instead of the windowed executor (gradient n >= 13, cost n >= 14)
 *   const_fuse_qubits (0, 4, 5)  constant sub-circuits (gates without parameters) are multiplied out on the host into dense
 *                          kernels of up to this many qubits where that needs fewer flops than 3-qubit blocks (default 4; 0: off)
 *   split_tables (0/1/2)   derivative kernel tables of fused blocks built by one warp per block member in a second kernel instead of
 *                          one warp per block: 0 off, 1 (default) for batches <= 8 -- it shortens what a single evaluation of a
 *                          BFGS line search waits for --, 2 always; the tables are bit-identical either way
 *   split, split_force, threads, ctas_per_sm   CTA-shape experiments of the launch planner
 *   verbose (0/1) */
int sqgpu_set_option(sqgpu_handle_t h, const char* name, int64_t value);
int sqgpu_get_option(sqgpu_handle_t h, const char* name, int64_t* value);

/* replaces Optimization_Interface::set_cost_function_variant / set_trace_offset and the members
 * prev_cost_fnv_val, correction1_scale, correction2_scale (Optimization_Interface.h:83-89, .cpp:74-76,1785-1800). */
int sqgpu_set_cost(sqgpu_handle_t h, int variant, int trace_offset, double prev_cost_fnv_val,
                   double correction1_scale, double correction2_scale);

/* ---- the hot path ------------------------------------------------------------------------------------------ */

/* replaces Optimization_Interface::optimization_problem_batched (Optimization_Interface.cpp:939-1033; the DFE
 * version calcqgdKernelDFE, common_DFE.cpp:51,61-68): cost[b] = f(params[b*P .. b*P+P)), b < batch. */
int sqgpu_cost_batched(sqgpu_handle_t h, const double* params, int batch, double* cost);

/* replaces Optimization_Interface::optimization_problem_combined_non_static (Optimization_Interface.cpp:1145-1490),
 * batched (the reference has no batched cost+gradient entry; batch = 1 is the reference call):
 * cost[b] and grad[b*P + i] = d cost / d params[b*P + i]. */
int sqgpu_cost_grad_batched(sqgpu_handle_t h, const double* params, int batch, double* cost, double* grad);

/* Raw trace terms before the non-linear cost formulas, for column-sharded multi-GPU runs (the caller sums them over
 * ranks, then calls sqgpu_cost_from_traces). Mirrors the {trace, correction1, correction2} triple calcqgdKernelDFE
 * returns per gate set (Optimization_Interface.cpp:806-832). Layout: traces[b][k][t][2], k = 0 the circuit itself,
 * k = 1..P the P derivatives (only if with_grad), t < 3 = {main diagonal, one-bit-flip, two-bit-flip sums}, {re, im}.
 * The normalisation of the cost formulas (the summed column count) is applied by sqgpu_cost_from_traces. */
int sqgpu_traces_batched(sqgpu_handle_t h, const double* params, int batch, int with_grad, double* traces);

/* cost (and gradient if grad != NULL) from (possibly rank-summed) trace terms; cols_total = summed column count. */
int sqgpu_cost_from_traces(sqgpu_handle_t h, const double* traces, int batch, int with_grad, int cols_total,
                           double* cost, double* grad);

/* replaces Gates_block::apply_to(parameters, input) (Gates_block.cpp:605-710) for the Circuit.apply_to binding
 * (qgd_Circuit_Wrapper.cpp:861-995): transform `inout` (rows x cols, any cols >= 1) in place. Does not touch the
 * resident matrix. */
int sqgpu_apply(sqgpu_handle_t h, const double* params, double* inout, int rows, int cols, int stride);

/* replaces Gates_block::apply_derivate_to (Gates_block.cpp:1011-1150): out holds P matrices of rows x cols
 * (compact, stride = cols), out[i] = d(C(params) * in)/d params[i], with the reference's conventions (zero rows
 * where a controlled gate is inactive, kernels/apply_kernel_to_input.cpp:93-97). */
int sqgpu_apply_derivative(sqgpu_handle_t h, const double* params, const double* in, int rows, int cols, int stride,
                           double* out);

/* single gate on a matrix / state vector: Gate::apply_to(parameters, input) (Gate.cpp:500-526 -> apply_kernel_to,
 * Gate.cpp:1477-1768). `deriv_param` < 0: the gate itself; otherwise its derivative with respect to its
 * deriv_param-th parameter (Gate::apply_derivative_to_precomputed, Gate.cpp:644-706). */
int sqgpu_apply_gate(sqgpu_handle_t h, const sqgpu_gate_desc* gate, const double* gate_params,
                     const double* matrix_pool, int deriv_param, double* inout, int rows, int cols, int stride);

/* ---- VQE (state-vector) path ------------------------------------------------------------------------------- */

/* replaces the Matrix_sparse Hamiltonian member of Variational_Quantum_Eigensolver_Base
 * (common/include/matrix_sparse.h:38; ctor variational_quantum_eigensolver/...Base.cpp): CSR, complex128 values,
 * int32 indices, n_rows = 2^n. */
int sqgpu_set_hamiltonian_csr(sqgpu_handle_t h, int n_rows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                              const double* values);

/* replaces Variational_Quantum_Eigensolver_Base::optimization_problem (…Base.cpp:1088-1121) batched over parameter
 * sets: energy[b] = Re <psi(params_b)| H |psi(params_b)>, psi = C(params_b) * (resident state vector). */
int sqgpu_vqe_energy_batched(sqgpu_handle_t h, const double* params, int batch, double* energy);

/* replaces Variational_Quantum_Eigensolver_Base::optimization_problem_combined_non_static (…Base.cpp:1131-1199):
 * grad[b*P+i] = 2 Re <d_i psi| H |psi>. */
int sqgpu_vqe_energy_grad_batched(sqgpu_handle_t h, const double* params, int batch, double* energy, double* grad);

/* ---- device-resident variants (inputs/outputs are device pointers on the handle's device; enqueue on `stream`,
 *      a cudaStream_t passed as void*; no host synchronisation) -- what bench.py's `value` times ------------ */

int sqgpu_cost_batched_dev(sqgpu_handle_t h, const double* d_params, int batch, double* d_cost, void* stream);
int sqgpu_cost_grad_batched_dev(sqgpu_handle_t h, const double* d_params, int batch, double* d_cost, double* d_grad,
                                void* stream);
int sqgpu_traces_batched_dev(sqgpu_handle_t h, const double* d_params, int batch, int with_grad, double* d_traces,
                             void* stream);
int sqgpu_cost_from_traces_dev(sqgpu_handle_t h, const double* d_traces, int batch, int with_grad, int cols_total,
                               double* d_cost, double* d_grad, void* stream);
/* Column sharding driven from OUTSIDE the library (one process per GPU, torch.distributed / MPI for the exchange; the
 * reference's rectangular-Umtx + trace_offset semantics, N_Qubit_Decomposition_Cost_Function.cpp:147-153): the resident matrix
 * of this handle is U[:, col_begin : col_begin + cols) of a matrix with cols_total columns. The shard's row offset then enters
 * the trace terms of EVERY cost variant (the user's trace_offset keeps entering the Frobenius family only, as in the
 * reference). col_begin = 0, cols_total = 0 switches sharding off. */
int sqgpu_set_shard(sqgpu_handle_t h, int col_begin, int cols_total);
/* gradient traces of a column shard for the Hilbert-Schmidt correction variants (4, 5), whose gradient functional takes its
 * weights from the traces of the circuit itself: d_global_traces0 [batch][3][2] = sqgpu_traces_batched_dev(with_grad = 0)
 * summed over all shards. For every other variant sqgpu_traces_batched_dev(with_grad = 1) is enough. */
int sqgpu_grad_traces_with_global_dev(sqgpu_handle_t h, const double* d_params, int batch, const double* d_global_traces0,
                                      double* d_traces, void* stream);
int sqgpu_apply_gate_dev(sqgpu_handle_t h, const sqgpu_gate_desc* gate, const double* gate_params,
                         const double* matrix_pool, int deriv_param, double* d_inout, int rows, int cols, int stride,
                         void* stream);
int sqgpu_vqe_energy_batched_dev(sqgpu_handle_t h, const double* d_params, int batch, double* d_energy, void* stream);
int sqgpu_vqe_energy_grad_batched_dev(sqgpu_handle_t h, const double* d_params, int batch, double* d_energy, double* d_grad,
                                      void* stream);

/* ---- device-resident optimizer inner loops ------------------------------------------------------------ */

/* replaces the per-iteration host round trip of solve_layer_optimization_problem_ADAM (optimization_engines/ADAM.cpp:199-330:
 * optimization_problem_combined on the accelerator, Adam::update on the host): `batch` independent ADAM trajectories whose
 * parameters, moments and scalar optimizer state stay on the device. The update is Adam::update (common/Adam.cpp:120-262) in
 * its sequential semantics -- the bias-correction products advance once per parameter, as the reference's loop does -- without
 * FMA contraction, so a trajectory equals the host-driven loop (evaluate, read back, update on the host) bit for bit.
 * Defaults of the reference: eta 1e-3, beta1 0.68, beta2 0.8, epsilon 1e-4 (Adam::Adam, Adam.cpp:31-36).
 * The randomisation / restart policy of ADAM.cpp stays with the caller: sqgpu_adam_steps returns the cost of every step and
 * sqgpu_adam_get the local-minimum flag Adam::update returns, so the host applies it between calls. */
int sqgpu_adam_init(sqgpu_handle_t h, const double* theta0 /* [batch][P] */, int batch, double eta, double beta1, double beta2,
                    double epsilon);
/* n_steps iterations of { f, g = cost+gradient(theta); best = min(best, f); theta = adam_update(theta, g, f) } enqueued on the
 * device without a host synchronisation (one CUDA graph replay per step after the first). cost_history (NULL or
 * [n_steps][batch]): f of every step, i.e. the cost BEFORE that step's update. */
int sqgpu_adam_steps(sqgpu_handle_t h, int n_steps, double* cost_history);
/* current parameters, lowest cost seen and the parameters that gave it (ADAM.cpp:219-222), Adam::update's status flag
 * (1: converged to a local minimum); any pointer may be NULL */
int sqgpu_adam_get(sqgpu_handle_t h, double* theta, double* best_cost, double* best_theta, int* status);

/* replaces the shift batches of the parameter-shift engines -- COSINE evaluates optimization_problem_batched on batch_size copies
 * of theta with ONE parameter each moved by pi/2, then by pi (optimization_engines/COSINE.cpp:255-291; GRAD_DESCEND_PARAMETER_
 * SHIFT_RULE.cpp likewise): shifted[s][b][p] = cost(params_b + shifts[s] e_p) for EVERY parameter p, cost[b] = cost(params_b),
 * from one adjoint sweep per parameter set (about three forward passes) instead of one forward pass per shifted parameter. The
 * trace functional is linear in each gate's kernel, so the sweep of sqgpu_cost_grad_batched run on K(theta_p + shift) - K(theta_p)
 * in place of dK/dtheta_p returns the exact change of the trace; the sweep itself does not depend on the shift, so further shifts
 * cost a table build and a reduction only. Cost variants 0, 1, 2, 3, 9 (functions of one linear trace functional); the others
 * return SQGPU_ERR_UNSUPPORTED. shifts: n_shifts >= 1 host values != 0. Single-device handles. */
int sqgpu_cost_shifted_batched(sqgpu_handle_t h, const double* params, int batch, const double* shifts, int n_shifts, double* cost,
                               double* shifted);
/* device buffers d_params / d_cost / d_shifted, `shifts` stays a host array */
int sqgpu_cost_shifted_batched_dev(sqgpu_handle_t h, const double* d_params, int batch, const double* shifts, int n_shifts, double* d_cost,
                                   double* d_shifted, void* stream);

/* replaces the one-evaluation-per-trial-point line search of BFGS_Powell (common/BFGS_Powell.cpp:70-200) by ONE batch: the k
 * points x + alphas[j] * dir are formed on the device, cost[j] = f(x + alphas[j] dir) and, if dphi != NULL, the directional
 * derivatives dphi[j] = grad f(x + alphas[j] dir) . dir come back (2 P + k doubles up, k or 2 k doubles down). */
int sqgpu_line_search_batched(sqgpu_handle_t h, const double* x, const double* dir, const double* alphas, int k, double* cost,
                              double* dphi);

/* ---- introspection for bench.py / tests -------------------------------------------------------------------- */

/* number of kernels this library has launched on the handle since creation (bench.py's gpu_launches). */
int sqgpu_launch_count(sqgpu_handle_t h, int64_t* count);

/* name and average device time (ms, CUDA events on the launching stream) of the kernel that took the most device time since
 * the last call -- bench.py's roofline numerator; resets every per-kernel event ring. name_len includes the terminating NUL. */
int sqgpu_last_kernel_time(sqgpu_handle_t h, char* name, int name_len, double* ms, int* launches);
/* the same for one kernel by name ("fused_exec<GRAD>", "fused_exec<WINDOW_FWD>", "fused_exec<WINDOW_BWD>", "gate1q_stream", ...);
 * resets that ring only. launches = 0: the kernel has not run since the last reset. */
int sqgpu_kernel_time(sqgpu_handle_t h, const char* name, double* ms, int* launches);

/* launch geometry of the fused executor in the last cost / gradient evaluation: shape[0..6) = log2(tile columns), threads per
 * CTA, column chunks (grid.x), tiles per CTA, dynamic shared memory bytes, cluster size. shape[0] = -1: it has not run (the
 * evaluation went down the streaming path). The parity tests use it to prove they exercise the instantiation bench.py times. */
int sqgpu_last_launch_shape(sqgpu_handle_t h, int* shape, int n_shape);

/* FP64 flops the fused executor issued in its last cost / gradient launch (whole batch), split into tensor-pipe (DMMA m8n8k4,
 * 512 flops each) and scalar (DFMA) work: counted from the launch's own op list, i.e. what the kernel EXECUTES, not the
 * per-gate algorithmic figure. bench.py's roofline.frac uses this; profiles/ holds the ncu sm__ops_path_tensor_src_fp64
 * capture that validates the count. */
int sqgpu_last_exec_flops(sqgpu_handle_t h, double* tensor_flops, double* scalar_flops);

/* measured FP64 throughput of the handle's device in TFLOP/s: the larger of a DFMA and a DMMA (mma.sync m8n8k4.f64, the
 * instruction the executor's block path issues) burn kernel -- both run on the same pipe. Roofline denominator of the
 * shared-memory executor, which is bound by the FP64 tensor pipe and not by HBM (MEASURED_PEAKS.json has no FP64 figure). */
int sqgpu_fp64_fma_peak(sqgpu_handle_t h, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SQGPU_H_INCLUDED */
