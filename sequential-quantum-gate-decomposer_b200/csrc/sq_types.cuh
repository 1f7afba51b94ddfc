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
__host__ __device__ __forceinline__ cplx czero() { return make_double2(0.0, 0.0); }

// One operation of the device program. The host planner (sqgpu.cu: plan_blocks) fuses runs of consecutive gates that act
// inside one or two qubits into a single dense 2x2 / 4x4 "block" op (type == SQ_OP_BLOCK) whose matrix -- and the
// derivative of that matrix with respect to every parameter of its member gates -- is built per parameter set by
// build_kernel_tables. Gates that cannot be fused (controls outside the pair, 3+ qubit dense kernels) stay "raw":
// a 1-qubit gate with a control bit mask (dim == 2) or a dense dim x dim kernel on ascending qubits.
enum { SQ_OP_BLOCK = 2000,
       // cluster executor only: exchange local row bit `target` with cluster-rank bit `nq` (no arithmetic), see exec_fused.cuh
       SQ_OP_RESPLIT = 2001 };

struct DevOp {
    int32_t type;         // sqgpu_gate_type of a raw op, or SQ_OP_BLOCK
    int32_t dim;          // 2, 4, 8, 16 or 32
    int32_t target;       // dim == 2: target qubit
    uint32_t ctrl_mask;   // raw ops: all of these row-index bits must be set for the gate to act (0 for blocks)
    int32_t nq;           // dim > 2: number of qubits
    int32_t q[5];         // dim > 2: ascending qubits, local bit j <-> q[j]
    int32_t param_start;  // raw ops: first parameter
    int32_t n_params;     // parameters of the op (blocks: of all members)
    int32_t kern_off;     // offset (complex) into the per-parameter-set kernel table; -1: constant kernel in the pool
    int32_t dkern_off;    // offset (complex) into the per-parameter-set derivative table (n_params kernels of dim*dim)
    int32_t w_off;        // offset (complex) into the per-parameter-set W accumulator (dim*dim), -1 if no parameters
    int32_t member_off;   // blocks: first member in the member table
    int32_t n_members;
    int32_t nfix;         // number of fixed row-index bits when enumerating the op's groups (targets + controls)
    int32_t fix[6];       // those bit positions, ascending; unused entries = 30 (insert_zero(.,30) is the identity)
    int32_t dtab;         // raw dense 3-/4-qubit ops with a constant kernel: 1 + index of the op's DMMA fragment table, else 0
    int64_t pool_off;     // constant kernel offset (complex) in the pool
};

// A gate inside a fused block, in block-local coordinates.
struct DevMember {
    int32_t type;         // sqgpu_gate_type
    int32_t dim;          // 2: 1-qubit kernel (optionally controlled by another block qubit); 4: two-target kernel
    int32_t tl;           // local bit of the target (dim == 2) / of the lower target (dim == 4)
    int32_t cl;           // dim == 2: local control bit or -1
    int32_t param_start;  // first parameter in the circuit's parameter vector
    int32_t n_params;
    int32_t slot0;        // index of the member's first derivative kernel inside the block's derivative table
    int32_t tl2;          // dim == 4: local bit of the higher target
    int64_t pool_off;     // GENERAL members: constant kernel in the pool
};

static const int SQ_MAX_MEMBERS = 24;  // gates per fused block

// insert a zero bit at position t of idx (all higher bits move up)
__host__ __device__ __forceinline__ int insert_zero(int idx, int t) {
    return ((idx >> t) << (t + 1)) | (idx & ((1 << t) - 1));
}

// software PDEP: the bits of v, lowest first, placed on the set bits of mask, lowest first
__host__ __device__ __forceinline__ unsigned deposit_bits(unsigned v, unsigned mask) {
    unsigned out = 0;
    while (mask && v) {
        const unsigned low = mask & (0u - mask);
        if (v & 1u) out |= low;
        v >>= 1;
        mask ^= low;
    }
    return out;
}

}  // namespace sq
