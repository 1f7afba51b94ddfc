// sq_types.cuh -- device-side program representation and complex helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sq {

typedef double2 cplx;  // {x = re, y = im} == QGD_Complex16 (common/include/QGDTypes.h:38-43)

__host__ __device__ __forceinline__ cplx cmake(double re, double im) { return make_double2(re, im); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// acc + a*b
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx acc) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
    return acc;
}
// acc + conj(a)*b
__device__ __forceinline__ cplx cfmac(cplx a, cplx b, cplx acc) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
    return acc;
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx czero() { return make_double2(0.0, 0.0); }

// One operation of the flattened circuit as the kernels see it. 1-qubit gates (optionally controlled by a bit mask)
// have dim == 2; everything else (two-target gates, GENERAL blocks) is a dense dim x dim kernel on ascending qubits.
struct DevOp {
    int32_t type;         // sqgpu_gate_type
    int32_t dim;          // 2, 4, 8, 16 or 32
    int32_t target;       // dim == 2: target qubit
    uint32_t ctrl_mask;   // all of these row-index bits must be set for the gate to act
    int32_t nq;           // dim > 2: number of qubits
    int32_t q[5];         // dim > 2: ascending qubits, local bit j <-> q[j]
    int32_t param_start;  // first parameter
    int32_t n_params;
    int32_t kern_off;     // offset (complex) into the per-parameter-set kernel table; -1: constant kernel in the pool
    int32_t dkern_off;    // offset (complex) into the per-parameter-set derivative-kernel table (n_params kernels)
    int32_t w_off;        // offset (complex) into the per-parameter-set W accumulator (dim*dim), -1 if no parameters
    int32_t pad;
    int64_t pool_off;     // constant kernel offset (complex) in the pool
};

// insert a zero bit at position t of idx (all higher bits move up)
__host__ __device__ __forceinline__ int insert_zero(int idx, int t) {
    return ((idx >> t) << (t + 1)) | (idx & ((1 << t) - 1));
}

}  // namespace sq
