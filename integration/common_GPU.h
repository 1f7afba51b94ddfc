// common_GPU.h -- host-side C++ bridge between SQUANDER's classes and the C-ABI of libsqgpu.so (include/sqgpu.h).
//
// This is the file a SQUANDER maintainer adds next to squander/src-cpp/common/common_DFE.{h,cpp}: the same role (load the
// accelerator library once, bind its C symbols, turn failures into the std::string exceptions the Python wrappers already
// catch -- qgd_N_Qubit_Decompositions_Wrapper.cpp:1471-1483), for the B200 engine instead of the Maxeler DFE. It contains
// no numerics: every cost / gradient value comes out of libsqgpu.so.
//
// Reference interfaces it stands in for, entry by entry:
//   init_dfe_lib / load2LMEM / releive_DFE          (common/common_DFE.cpp:110-184)      -> GPU_Cost_Path::GPU_Cost_Path, upload, ~GPU_Cost_Path
//   Gates_block::convert_to_DFE_gates*              (gates/Gates_block.cpp:4202-4278)    -> to_gpu_gates
//   calcqgdKernelDFE + the formulas around it       (Optimization_Interface.cpp:806-832, 1277-1370) -> cost / cost_grad / cost_batched
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "sqgpu.h"

#include "Gates_block.h"
#include "matrix.h"
#include "matrix_real.h"
#include "matrix_sparse.h"

namespace sqgpu_bridge {

// Where to find libsqgpu.so. Default: "libsqgpu.so" through the dynamic loader's search path, as common_DFE.cpp does for
// libqgdDFE.so (common_DFE.h:44-46). Call before the first GPU_Cost_Path is created.
void set_library_path(const std::string& path);

// number of usable devices (get_accelerator_avail_num, common_DFE.cpp:47)
int available_gpus();

// Gates_block -> flat descriptor stream + constant-kernel pool: one walk of get_flat_circuit() (Gates_block.cpp:3827-3856)
std::vector<sqgpu_gate_desc> to_gpu_gates(Gates_block* circuit, std::vector<QGD_Complex16>& pool);

// One engine handle bound to one decomposition / VQE object (the DFE bridge binds the board to one owner through
// `id` / `initialize_id`, common_DFE.cpp:58, 178-184).
class GPU_Cost_Path {
public:
    explicit GPU_Cost_Path(int accelerator_num);  // accelerator_num >= 1: number of GPUs requested (this class drives device 0;
                                                  // more devices go through sqgpu_create_multi)
    ~GPU_Cost_Path();
    GPU_Cost_Path(const GPU_Cost_Path&) = delete;
    GPU_Cost_Path& operator=(const GPU_Cost_Path&) = delete;

    // upload_Umtx_to_DFE (Optimization_Interface.cpp:1819-1824): the matrix (or, cols = 1, the initial state) stays resident
    void upload(Matrix& Umtx);
    // the gate structure, once per structure (not once per evaluation as the DFE descriptors were)
    void set_circuit(Gates_block* circuit);
    void set_hamiltonian(Matrix_sparse& H);
    // Optimization_Interface members that enter the cost formulas (Optimization_Interface.h:83-89)
    void set_cost(int variant, int trace_offset, double prev_cost_fnv_val, double correction1_scale, double correction2_scale);

    double cost(Matrix_real& parameters);                                               // optimization_problem
    void cost_grad(Matrix_real& parameters, double* f0, Matrix_real& grad);             // optimization_problem_combined
    Matrix_real cost_batched(std::vector<Matrix_real>& parameters_vec);                 // optimization_problem_batched
    double vqe_energy(Matrix_real& parameters);                                         // VQE optimization_problem
    void vqe_energy_grad(Matrix_real& parameters, double* f0, Matrix_real& grad);       // VQE optimization_problem_combined

    long long evaluations() const { return n_evals; }

private:
    sqgpu_handle_t h = nullptr;
    int n_params = 0;
    long long n_evals = 0;
};

}  // namespace sqgpu_bridge
