// exec_fused.cuh -- the fused circuit executor: one CTA keeps a tile of CT columns of the 2^n x C matrix in shared
// memory, runs the WHOLE device program on it, and emits only trace partials (cost), trace + W partials (adjoint
// gradient) or the transformed tile (apply / materialised derivative).
//
// It replaces, for one column tile, the reference's per-gate passes over the full matrix:
//   Gates_block::apply_to_inner forward loop        (gates/Gates_block.cpp:683-708)
//   apply_kernel_to_input row-pair update           (gates/kernels/apply_kernel_to_input.cpp:52-112)
//   apply_nqbit_kernel_to_matrix_input_impl         (gates/kernels/apply_large_kernel_to_input.cpp:123-213)
//   get_cost_function / get_trace* diagonals        (decomposition/N_Qubit_Decomposition_Cost_Function.cpp:73-664)
// and, for the gradient, the P materialised derivative matrices of Gates_block::apply_derivate_to
// (gates/Gates_block.cpp:1011-1150) by an adjoint sweep: with a_k the column after k ops and beta_k the row functional
// e_r^T G_{N-1}...G_{k+1}, the derivative of the trace term wrt a parameter of op k is sum_groups beta_k^T dK a_k.
// The executor accumulates W'_k[r][c] = sum_{groups, columns} beta_k[r] p_k[c] with p_k = K a_k the column AFTER the op
// (so the accumulation does not wait for the un-applied column); since a_k = K^dagger p_k, the wanted
// W_k[r][c] = sum beta_k[r] a_k[c] equals (W'_k conj(K))[r][c] and reduce_partials contracts W'_k with dK K^dagger
// (for fused blocks dK is the product-rule derivative of the block matrix, gate_kernels.cuh).
// Inactive (control = 0) pairs contribute nothing, which is exactly the reference's "zero rows in the derivative"
// convention (apply_kernel_to_input.cpp:93-97).
//
// The hot ops are the planner's fused blocks: dense 8x8 (three qubits) or 4x4 (two qubits) complex kernels without
// controls, run on the FP64 tensor cores (mma.sync m8n8k4.f64 on the real embedding of the kernel, block_dmma_forward /
// block_dmma_backward); a block costs one shared-memory round trip and one barrier instead of one per gate. Single-qubit
// blocks, controlled leftovers and derivative ops take scalar paths in the same kernel; raw GENERAL 3-5 qubit kernels the
// dense DMMA paths.
//
// Data layout in shared memory: element (row i, tile column c) at elem(i, c), 16 B each, a GF(2)-linear bijection that
// XORs images of the row bits into the three bank-group bits (see elem() below), so that the 128-bit fragment accesses of
// the DMMA paths are conflict-free and every address splits into B0(batch) ^ slot(lane).
//
// Window mode (state vectors too long for one tile: VQE; column chunks of matrices with n >= 13): the rows of the tile are
// the configurations of an arbitrary subset of `n_win` qubits (ExecArgs::wmask), the columns those of the other bits;
// MODE_APPLY / MODE_BWD run one SEGMENT of the window plan (sqgpu.cu: build_window_plan) per launch. Tiles arrive by
// per-element asynchronous copies (the whole tile in flight), CTAs are long-lived and take tiles round robin, the last op of
// a segment stores straight to HBM.
//
// Cluster mode (CLU; n = 12 gradients by default): the CTAs of a thread-block cluster share a column tile, RESPLIT ops
// exchange the split qubits through distributed shared memory (sqgpu.cu: build_cluster_plan).
//
// Complex arithmetic of the block paths: three real products per complex product, the second and third accumulated on
// register copies of the first (SQ_BLOCK_3M / SQ_DENSE_3M below): 6 / 18 DMMA per batch forward / backward.
#pragma once
#include <cooperative_groups.h>
#include "sq_types.cuh"
#include "../../include/sqgpu.h"

namespace sq {

// MODE_COST: forward sweep + trace terms.  MODE_GRAD: + adjoint (backward) sweep.  MODE_APPLY: forward sweep, tile stored.
// MODE_BWD: one segment of the adjoint sweep over a state vector that lives in HBM -- the column tile `a` (out) and the row
// functional `beta` are loaded, swept backward through the segment's ops and stored again (windowed VQE executor).
enum { MODE_COST = 0, MODE_GRAD = 1, MODE_APPLY = 2, MODE_BWD = 3 };

struct OpTab;
struct DenseTab;
struct DenseTab5;

struct ExecArgs {
    const cplx* in;          // input matrix (row-major, leading dimension ld_in)
    cplx* out;               // MODE_APPLY: output (may alias in)
    long long in_ystride;    // element stride of `in` per blockIdx.y (0: shared input)
    long long out_ystride;   // element stride of `out` per blockIdx.y
    int ld_in, ld_out;
    int rows, cols, n;       // rows = 2^n
    int ct, log_ct;          // tile width (columns), power of two
    int tiles, tiles_per_cta;
    const DevOp* ops;
    int n_ops;
    const cplx* ktab;        // [ysets][kern_total]
    int kern_total;
    const cplx* dktab;       // [ysets][dkern_total]
    int dkern_total;
    const cplx* pool;
    const struct OpTab* optabs;  // [ysets][optab_stride] lookup tables of the DMMA block path (build_optabs)
    int optab_stride;        // tables per parameter set (0: n_ops)
    unsigned wmask;          // window mode (state vectors): bit mask of the `n_win` qubits that form the tile's rows; element
                             // (row r, column j) of y lives at in[y * ystride + deposit(r, wmask) | deposit(j, ~wmask)]; 0: matrix
    cplx* beta;              // MODE_BWD: the row functional, same layout and stride as out
    int sum_sq;              // SUM_OF_SQUARES cost: trace slot 0 carries sum |M_ij - delta_ij|^2, beta_N = 2 conj(M - I)
    const struct DenseTab* dense_tabs;  // fragment tables of the raw dense 3-/4-qubit ops (build_dense_tabs), NULL: none
    const struct DenseTab5* dense_tabs5;  // ... of the 5-qubit ops
    int k_shared;            // 1: every blockIdx.y uses kernel-table set 0 (materialised derivative: one parameter set)
    const int* deriv_op;     // MODE_APPLY: per blockIdx.y the op whose derivative kernel is applied (NULL: none)
    const int* deriv_slot;   //             and which of its derivative kernels
    int trace_offset;
    int n_trace_types;       // 1: main diagonal only, 2: + one-bit-flip sums, 3: + two-bit-flip sums
    double* tr_part;         // [y][chunks][6]
    cplx* w_part;            // [y][chunks][w_total]
    int w_total;
    int w_in_smem;           // accumulate W over the CTA's tiles in shared memory; else in the CTA's own (zero-initialised)
                             // slice of w_part with fire-and-forget reductions: one thread owns each address, so the
                             // summation order stays fixed
    const cplx* omega;       // [y][3] weights of the three trace types in the functional whose gradient is taken
    int w_direct;            // (not w_in_smem) 1: one slice of w_part per (parameter set, CTA, warp), written by that warp alone
    int dense_stage;         // complex elements of kernel staging for the generic dense path (0: none)
    int wmax;                // max dim*dim over parametric ops (complex), >= 4
    int dbuf;                // window forward segments: two tile buffers, the next tile streams in while this one is computed
    int dns;                 // 1: the program has raw dense ops (or a materialised block derivative): launch the DNS = true kernels
    int rho;                 // cluster executor: log2(CTAs per cluster); `rows` is then the rows ONE CTA holds (2^(n - rho))
    signed char fin_pos[32]; // cluster executor: where logical qubit q sits after the forward sweep: local row bit p (p < 32) or
                             // cluster-rank bit p - 32
};

static const int FUSED_THREADS = 512;

// ---- shared-memory layout ----------------------------------------------------------------------------------------
// Element (row r, tile column c) lives at complex index elem(r, c), a GF(2)-linear bijection of the (row, column) bits: the
// three address bits that select the 16 B bank group (address mod 8) are XORed with images of the row bits. Row bit k has
// the image (1, 2, 3)[k % 3] in the two bank bits above tile-column bit 0 (for CT = 1: a 3-bit image), so rows that differ
// in two block qubits of different residue, and columns that differ in bit 0, fall into 8 different bank groups: the
// quarter-warp of a 128-bit fragment access of the DMMA block path is conflict-free. Linearity makes every address of that
// path  B0(batch) ^ slot(lane).
__host__ __device__ constexpr unsigned bits_mod(int m, unsigned sel, int from) {
    unsigned r = 0;
    for (int b = from; b < 31; ++b)
        if ((sel >> ((b - from) % m)) & 1u) r |= 1u << b;
    return r;
}
template <int LOG_CT>
__device__ __forceinline__ int elem(int r, int c) {
    constexpr unsigned MA = bits_mod(3, 0b101u, 0);  // row bits k with k % 3 in {0, 2}: image bit 0
    constexpr unsigned MB = bits_mod(3, 0b110u, 0);  // row bits k with k % 3 in {1, 2}: image bit 1
    const int pa = __popc((unsigned)r & MA) & 1, pb = __popc((unsigned)r & MB) & 1;
    if (LOG_CT >= 3) return ((r << LOG_CT) | c) ^ ((pa | (pb << 1)) << 1);
    if (LOG_CT == 2) return ((((r & ~1) | pa) << 2) | c) ^ (pb << 1);
    if (LOG_CT == 1) return ((((r & ~3) | pa | (pb << 1))) << 1) | c;
    // CT = 1: bank bits are row bits 0..2; row bit k >= 3 has image (7, 3, 5, 6)[(k - 3) % 4]
    constexpr unsigned M0 = bits_mod(4, 0b0111u, 3), M1 = bits_mod(4, 0b1011u, 3), M2 = bits_mod(4, 0b1101u, 3);
    return r ^ ((__popc((unsigned)r & M0) & 1) | ((__popc((unsigned)r & M1) & 1) << 1) | ((__popc((unsigned)r & M2) & 1) << 2));
}

// reduce 8 per-lane doubles over the warp; lanes with (lane & 3) == 0 end up holding the total of value
// index ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1) in v[0]   (9 double shuffles instead of 40)
__device__ __forceinline__ void warp_reduce8(double* v, int lane) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 16;
        const double send = up ? v[i] : v[i + 4];
        const double keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 8;
        const double send = up ? v[i] : v[i + 2];
        const double keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    {
        const bool up = lane & 4;
        const double send = up ? v[0] : v[1];
        const double keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 4);
    }
    v[0] += __shfl_xor_sync(full, v[0], 2);
    v[0] += __shfl_xor_sync(full, v[0], 1);
}

// per-warp W partial (complex w[n]) -> swarp slot; n = 4 or 16
template <int N>
__device__ __forceinline__ void warp_store_w(const cplx* w, double* slot, int lane, bool direct = false) {
#pragma unroll
    for (int part = 0; part < N / 4; ++part) {
        double v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[2 * e] = w[part * 4 + e].x;
            v[2 * e + 1] = w[part * 4 + e].y;
        }
        warp_reduce8(v, lane);
        if ((lane & 3) == 0) {
            double* dst = slot + part * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            if (direct) atomicAdd(dst, v[0]);  // the warp's own slice of w_part (global): fire-and-forget reduction
            else *dst = v[0];
        }
    }
}

// ---- dense k-qubit blocks on the FP64 tensor cores ----------------------------------------------------------------
// A dense 2^k x 2^k complex kernel acting on a group of 2^k amplitudes is a real (2^(k+1)) x (2^(k+1)) matrix acting on
// the interleaved {re, im} vector:  Kreal[2r+a][2c+b] = { K.re if a == b;  -K.im if (a,b) = (0,1);  +K.im if (1,0) }.
// One warp transforms 8 (group, column) items per step with mma.sync.m8n8k4.f64 (DMMA): D[8 x 8 items] +=
// Kreal[8 x 4] * X[4 x 8 items], RT = 2^(k+1)/8 row tiles x KS = 2^(k+1)/4 k-steps. Replaces the reference's gather /
// dense matvec / scatter loop for 4-5 qubit kernels (apply_large_kernel_to_input.cpp:160-199,
// apply_large_kernel_to_input_AVX.cpp:90-136).
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

static const int DMMA_PAD = 4;  // doubles of row padding of the staged real kernel (conflict-free 8 x 4 fragment loads)

// forward application of a raw dense op (no controls) on the tile; requires nitems % 8 == 0. skr: staged real kernel
// [DIMR][DIMR + DMMA_PAD]; spat: row pattern of local index l. All warps of the CTA take part.
template <int LOG_CT, int KQ>
__device__ __forceinline__ void dense_dmma_forward(cplx* sa, const double* skr, const int* spat, const DevOp& op, int rows,
                                                   int tid, int nthr) {
    constexpr int CT = 1 << LOG_CT;
    constexpr int DIMR = 2 << KQ, RT = DIMR / 8, KS = DIMR / 4, LD = DIMR + DMMA_PAD;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    double* sad = reinterpret_cast<double*>(sa);
    const int nitems = (rows >> KQ) << LOG_CT;
    const int m = lane >> 2, kk = lane & 3;
    int q[KQ];
#pragma unroll
    for (int j = 0; j < KQ; ++j) q[j] = op.q[j];
    auto item_base = [&](int item, int& c) {
        c = item & (CT - 1);
        int base = item >> LOG_CT;
#pragma unroll
        for (int j = 0; j < KQ; ++j) base = insert_zero(base, q[j]);
        return base;
    };
    // A fragments: in registers when they fit (k <= 4: 32 doubles), else re-read from shared memory per DMMA
    constexpr bool A_IN_REGS = (RT * KS <= 32);
    double areg[A_IN_REGS ? RT * KS : 1];
    if (A_IN_REGS) {
#pragma unroll
        for (int rt = 0; rt < RT; ++rt)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) areg[rt * KS + ks] = skr[(8 * rt + m) * LD + 4 * ks + kk];
    }
    for (int b0 = warp * 8; b0 < nitems; b0 += nwarps * 8) {
        // B operand: lane holds real component 4*ks + kk of item b0 + m
        double bfrag[KS];
        {
            int c;
            const int base = item_base(b0 + m, c);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int comp = 4 * ks + kk;
                const int row = base | spat[comp >> 1];
                bfrag[ks] = sad[(elem<LOG_CT>(row, c)) * 2 + (comp & 1)];
            }
        }
        double d[RT][2];
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            d[rt][0] = 0.0;
            d[rt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double a = A_IN_REGS ? areg[rt * KS + ks] : skr[(8 * rt + m) * LD + 4 * ks + kk];
                dmma_m8n8k4(d[rt][0], d[rt][1], a, bfrag[ks]);
            }
        }
        __syncwarp();  // every lane has read its inputs of this 8-item batch before anyone overwrites them
        // D: lane holds output component 8*rt + m of items b0 + 2*kk and b0 + 2*kk + 1
        int c0, c1;
        const int base0 = item_base(b0 + 2 * kk, c0), base1 = item_base(b0 + 2 * kk + 1, c1);
#pragma unroll
        for (int rt = 0; rt < RT; ++rt) {
            const int comp = 8 * rt + m;
            const int pat = spat[comp >> 1];
            sad[(elem<LOG_CT>(base0 | pat, c0)) * 2 + (comp & 1)] = d[rt][0];
            sad[(elem<LOG_CT>(base1 | pat, c1)) * 2 + (comp & 1)] = d[rt][1];
        }
        __syncwarp();
    }
}

// ---- fused 2-/3-qubit blocks on the FP64 tensor cores -------------------------------------------------------------
// D[8 items x 8 out comps] += X[8 items x 4 in comps] * KrealT[4 in comps x 8 out comps] with mma.sync.m8n8k4.f64: the DATA
// is the A operand (lane l supplies item l >> 2, component slot l & 3) and the kernel the B operand. Lane (i, j) loads the
// complex amplitudes dep(j, u) of item i (one 128-bit load per u; .x feeds k-step 2u, .y k-step 2u + 1) and receives in its
// D fragment exactly the same amplitudes of the result (out comps 2j, 2j + 1 of n-tile u): loads and stores of a lane hit
// the same addresses, 16 B wide, and no lane touches another lane's elements.
template <int LOG_CT, int KQ>
struct BlockGeom {
    static constexpr int CT = 1 << LOG_CT;
    static constexpr int LOGG = 3 - LOG_CT;  // log2(groups per 8-item batch)
    int q[3];   // block qubits in kernel order: amplitude bit j <-> q[j] (unused = 30)
    int qs[3];  // the same, ascending (where zero bits are inserted)
    int Pq[3];  // address image (complex units) of row bit q[j]
    int F[3];   // address image of the j-th lowest row bit that is NOT a block qubit (group-offset bits inside a batch)

    __device__ __forceinline__ void init(int q0, int q1, int q2) {
        q[0] = q0; q[1] = q1; q[2] = q2;
        {
            int s0 = q0, s1 = KQ > 1 ? q1 : 30, s2 = KQ > 2 ? q2 : 30;
            if (s0 > s1) { const int t = s0; s0 = s1; s1 = t; }
            if (s1 > s2) { const int t = s1; s1 = s2; s2 = t; }
            if (s0 > s1) { const int t = s0; s0 = s1; s1 = t; }
            qs[0] = s0; qs[1] = s1; qs[2] = s2;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) Pq[j] = (j < KQ) ? elem<LOG_CT>(1 << q[j], 0) : 0;
        int f = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            while (f == q0 || f == q1 || f == q2) ++f;
            F[j] = elem<LOG_CT>(1 << f, 0);
            ++f;
        }
    }
    // B0 (complex units) of the batch starting at item b0 (b0 % 8 == 0). The block qubits are in KERNEL order (amplitude bit
    // j <-> q[j]); the cluster planner's row-bit relabelling can leave them unsorted, the zero bits go in ascending order.
    __device__ __forceinline__ int batch_base(int b0) const {
        int base = b0 >> LOG_CT;
#pragma unroll
        for (int j = 0; j < KQ; ++j) base = insert_zero(base, qs[j]);
        return elem<LOG_CT>(base, 0);
    }
    // lane-constant part of the address (complex units) of local amplitude `amp` of batch item `item` (0..7)
    __device__ __forceinline__ int slot(int item, int amp) const {
        const int go = item >> LOG_CT;
        int a = item & (CT - 1);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j < LOGG && ((go >> j) & 1)) a ^= F[j];
#pragma unroll
        for (int j = 0; j < KQ; ++j)
            if ((amp >> j) & 1) a ^= Pq[j];
        return a;
    }
};

// Per-(parameter set, op) lookup table of the DMMA block path, written by build_optabs and prefetched into shared memory
// with cp.async while the previous op runs (double-buffered).
static const int B0TAB = 64;  // batch bases precomputed per op (tiles with more batches fall back to the index arithmetic)
// Where the DMMA loops get the base address of a batch from (compile-time switch; measured on B200, C3 / C5 energy,
// profiles/r2_variants.jsonl):
//   0: index arithmetic at the top of every iteration (1 245 evals/s / 2 458);  1: the per-op table b0[] in shared memory
//   (1 243 / 2 534, the default);  2: index arithmetic for the NEXT batch issued before the current batch's DMMAs (1 228 / 2 429)
#ifndef SQ_B0_MODE
#define SQ_B0_MODE 1
#endif
// Complex arithmetic of the 3-qubit blocks on the tensor cores (compile-time switch):
//   0: the real embedding -- an 8 x 8 complex block is a 16 x 16 real matrix, 8 DMMA per 8 items forward, 24 backward;
//   1 (default): three real products per complex product ("3M": with K = C + iD and x = u + iv,
//      k1 = C (u + v),  Re = k1 - (C + D) v,  Im = k1 + (D - C) u), each an 8 x 8 real matrix-vector product = 2 DMMA per 8
//      items: 6 DMMA forward, 18 backward (K^dagger p, K^T beta and the W' outer products each drop from 8 to 6). The second
//      and third product accumulate on top of register copies of k1 (the DMMA's addend; the table holds C, D - C and -(C + D)),
//      so the only FP64 additions left are the operand sums u + v: 2 / 8 DADD per batch instead of the 8 / 20 of the first 3M
//      form (t1 = C u, t2 = D v, t3 = (C + D)(u + v), Re = t1 - t2, Im = t3 - t1 - t2), whose DADDs held 30 % of the warp
//      samples of the gradient kernel and ~20 % of its FP64-pipe cycles (DADD and DMMA share one pipe,
//      profiles/r2_dmma_dadd_mix.txt). 6 / 12 kernel-fragment registers instead of 16 / 32. The rounding differs from the
//      4-product form by O(eps) per product (normwise; the parity tests hold the 1e-12 / 1e-10 bars for both settings).
#ifndef SQ_BLOCK_3M
#define SQ_BLOCK_3M 1
#endif
// How the op tables travel to shared memory: 0 (default): per-thread cp.async (LDGSTS) groups; 1: bulk-async copies issued by
// one thread (cp.async.bulk, the TMA engine: UBLKCP in SASS) with mbarrier completion. Measured: the bulk ring is 2.2 % SLOWER
// on C3 (1 217 vs 1 245 evals/s) and 1.5 % on C5 -- a 4.9 KB table per ~3 us op is too small for the TMA engine to beat 1.2
// LDGSTS per thread, and the mbarrier try_wait (60-90 cycles) lands on every warp's critical path once per op. Kept as a
// switch with its measurement; not the product configuration.
#ifndef SQ_TAB_BULK
#define SQ_TAB_BULK 0
#endif
// Window modes only: 1: the op tables use the bulk-async / mbarrier ring there, which frees the cp.async groups for a
// double-buffered tile pipeline in the forward segments (two single-column tile buffers, tile t + 1 requested before tile t is
// computed: option async_tiles). Measured on C5 (profiles/r2_variants_win3.jsonl): forward sweep 20.2 ms per 64 sets against
// 17.6 ms for the default (the bulk ring alone: 18.9 ms). The tile round trips it hides were not the limiter: 73-79 % of
// the warp samples sit inside the DMMA loops, which run the FP64 pipe at ~80 %, and waits for global loads are 6 %
// (profiles/r2_ncu_vqe_window_src.md); the narrower single-column tiles conflict more in shared memory.
// 0 (default): LDGSTS ring, one tile buffer, whole tile requested at once.
#ifndef SQ_WIN_BULK
#define SQ_WIN_BULK 0
#endif
// 1: the per-lane operands of the next DMMA block op (BlkPre) are loaded BEFORE the end-of-op barrier of the current one, which
// needs the next op's table one barrier earlier (tables are waited for two ops ahead instead of one); 0 (default): after the
// barrier, at the top of the op. Measured (profiles/r2_variants_exec.jsonl): the early load keeps ~50 more registers alive
// across the barrier (spills at the 128-register cap) and loses 11 % on C3.
#ifndef SQ_PRELOAD
#define SQ_PRELOAD 0
#endif
// How the batches of a block op are dealt to the warps: 1 (default): contiguous runs (a run of exactly four comes with ONE
// 128-bit load of its four bases); 0: round robin (round-1 behaviour, one table read per batch).
#ifndef SQ_BATCH_CONTIG
#define SQ_BATCH_CONTIG 1
#endif
// Experiment: the CTA that arrives second on an SM starts SQ_STAGGER nanoseconds late, so that the two resident CTAs do not
// run their per-op phases (tensor loop / barrier + operand prologue) in lock-step.
#ifndef SQ_STAGGER
#define SQ_STAGGER 0
#endif
// Where a warp's W' partial of an op goes: 0: into the warp's shared-memory slot; after the end-of-op barrier the CTA folds the
// slots and issues one reduction per element into ITS slice of w_part (round-1 scheme: half of the warps spend ~600 cycles per
// op in that fold while the others are already in the next op -- the phase trace shows them finishing last every time);
// 1: every warp owns a slice of w_part and adds its partial straight from the accumulator registers with fire-and-forget
// reductions (RED.ADD.F64): no shared-memory slots, no fold, nothing of W' behind the barrier; each address still has exactly
// one writer (bit-reproducible). Measured (profiles/r2_variants_exec.jsonl): 8x the reduction traffic and an L2 working set
// of 196 MB instead of 25 MB cost more than the fold: C3 1 211 against 1 339 evals/s, batch-1 latency 2.7 against 1.7 ms.
// 0 (default): the shared-memory slots, with the fold spread over ALL warps (see the fold below).
#ifndef SQ_W_DIRECT
#define SQ_W_DIRECT 0
#endif
// When the warps' W' partials of an op are folded: 0 (default): right after the op's end-of-op barrier, by the first nd / 32
// warps (which then enter the next op ~600 clocks late -- three dependent DADDs queue behind the other warps' DMMAs -- and make
// the others wait at the next barrier: 14 % of the gradient kernel's warp samples sit on the instruction after the barriers);
// 1: one op later and BEFORE that op's barrier, by the warps that reach it first (a ticket from a shared-memory counter picks
// the 32-element quarter a warp folds), i.e. in the slack of the fast warps. Same values, same summation order (bit-identical
// results) -- and measured SLOWER: C3 1 391 against 1 498 evals/s, C5 backward sweep 48.4 against 44.6 ms
// (profiles/README_r2.md): the ticket (atomic + shuffle) sits on every warp's path once per op and the fast warps' slack is
// smaller than the fold. Kept as a switch with its measurement.
#ifndef SQ_FOLD_EARLY
#define SQ_FOLD_EARLY 0
#endif

// Raw dense 3-/4-qubit kernels (GENERAL blocks): 1 (default): the same three-product form as the fused blocks -- 24 instead of
// 32 DMMA per 8 items for a 16 x 16 kernel, 6 instead of 8 for an 8 x 8 one, and 24 / 6 kernel-fragment registers per lane
// instead of 32 / 8 doubles; 0: the real embedding. 5-qubit kernels keep the real embedding (their fragments stream from L1).
#ifndef SQ_DENSE_3M
#define SQ_DENSE_3M 1
#endif
// The same for 5-qubit kernels (32 x 32): 96 instead of 128 DMMA per batch, but 16 more live operand registers -- measured: the
// kernel spills (820 B at the 128-register cap), 64 5-qubit blocks at n = 12 run at 67.4 instead of 93.1 evals/s, and the extra
// code costs the C3 gradient kernel 1.3 % (1 479 against 1 498 evals/s). 0 (default): 5-qubit kernels keep the real embedding.
// Window segments: 1 (default): the last op of a segment writes its results straight to the state in HBM (no store phase, no
// shared-memory round trip); 0: separate store loop. Measured on C5: forward 15.0 against 15.5 ms, backward 41.5 against 43.9 ms.
#ifndef SQ_FUSE_STORE
#define SQ_FUSE_STORE 1
#endif
#ifndef SQ_DENSE5_3M
#define SQ_DENSE5_3M 0
#endif

struct OpTab {
    double frag[3][8][32];  // [mode: K, K^dagger, K^T][t * KS + s][lane]: kernel (B operand) fragments
    int slot[4][32];        // [u]: load/store slot of amplitude dep(j, u); [2 + h]: W' operand slot of item half h
    int widx[32];           // 3-qubit blocks: positions of the lane's two W' outputs, idx0 | idx1 << 8
    int b0[B0TAB];          // B0 (complex units) of the batch of items [8 i, 8 i + 8): the op's qubits inserted as zeros, then elem()
};

// shared-memory copy of the part of an OpTab one sweep direction needs: the forward sweep takes frag[0] (K), the backward
// sweep frag[1..2] (K^dagger, K^T); both take the slots (4.7 KB instead of 6.8 KB per buffer)
struct OpTabS {
    double frag[2][8][32];
    int slot[4][32];
    int widx[32];
    int b0[B0TAB];
};
static_assert(sizeof(OpTab) % 16 == 0 && sizeof(OpTabS) % 16 == 0, "bulk copies move multiples of 16 bytes");

// local amplitude index from the lane's amplitude slot j (2 bits) and the k-step pair u: block-qubit position ju carries u
__device__ __forceinline__ int dep3(int j, int u, int ju) {
    const int ja = (ju == 0) ? 1 : 0, jb = (ju == 2) ? 1 : 2;
    return ((j & 1) << ja) | (((j >> 1) & 1) << jb) | (u << ju);
}
// W' row/column enumeration: bit 0 of rho sits on block-qubit position pa, the others ascend
__device__ __forceinline__ int permw3(int rho, int pa) {
    const int pb = (pa == 0) ? 1 : 0, pc = (pa == 2) ? 1 : 2;
    return ((rho & 1) << pa) | (((rho >> 1) & 1) << pb) | (((rho >> 2) & 1) << pc);
}

__device__ __forceinline__ double kreal_entry(const cplx* __restrict__ K, int dim, int mode, int amp_out, int a, int amp_in, int b) {
    const cplx e = (mode == 0) ? K[amp_out * dim + amp_in] : K[amp_in * dim + amp_out];
    const double im = (mode == 1) ? -e.y : e.y;
    return (a == b) ? e.x : (a ? im : -im);
}

template <int LOG_CT>
__device__ __forceinline__ void fill_optab(OpTab* T, const DevOp& op, const cplx* __restrict__ K) {
    __shared__ int s_choice[2];
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (op.dim == 8) {
        BlockGeom<LOG_CT, 3> G;
        G.init(op.q[0], op.q[1], op.q[2]);
        if (tid == 0) {
            // pick the amplitude-to-lane assignments whose quarter-warp (lanes 0..7) hits the most bank groups
            int best_ju = 0, best_pa = 0, best_m = -1, best_w = -1;
            for (int cand = 0; cand < 3; ++cand) {
                unsigned seen_m = 0, seen_w = 0;
                for (int l = 0; l < 8; ++l) {
                    seen_m |= 1u << (G.slot(l >> 2, dep3(l & 3, 0, cand)) & 7);
                    seen_w |= 1u << (G.slot(l & 3, permw3(l >> 2, cand)) & 7);
                }
                if (__popc(seen_m) > best_m) { best_m = __popc(seen_m); best_ju = cand; }
                if (__popc(seen_w) > best_w) { best_w = __popc(seen_w); best_pa = cand; }
            }
            s_choice[0] = best_ju;
            s_choice[1] = best_pa;
        }
        __syncthreads();
        const int ju = s_choice[0], pa = s_choice[1];
#if SQ_BLOCK_3M
        // frag[mode][m * 2 + s][lane]: B-operand fragment (lane = (k = lane & 3, n = lane >> 2)) of k-step s of the REAL 8 x 8
        // matrix m in {0: C, 1: D - C, 2: -(C + D)} of the mode's kernel C + iD (mode 0: K, 1: K^dagger, 2: K^T):
        // entry [out amplitude o(n)][in amplitude c(4 s + k)] with o(n) = dep3(n >> 1, n & 1), c(4 s + k) = dep3(k, s) -- the
        // lane's two inputs c(j), c(4 + j) are its two outputs o(2 j), o(2 j + 1): loads and stores hit the same elements
        for (int e = tid; e < 3 * 6 * 32; e += nthr) {
            const int mode = e / (6 * 32), ms = (e / 32) % 6, lane = e & 31, m = ms >> 1, st = ms & 1;
            const int n = lane >> 2, k = lane & 3;
            const int ao = dep3(n >> 1, n & 1, ju), ai = dep3(k, st, ju);
            const cplx kv = (mode == 0) ? K[ao * 8 + ai] : K[ai * 8 + ao];
            const double cre = kv.x, dim_ = (mode == 1) ? -kv.y : kv.y;
            T->frag[mode][ms][lane] = (m == 0) ? cre : (m == 1 ? dim_ - cre : -(cre + dim_));
        }
#else
        for (int e = tid; e < 3 * 8 * 32; e += nthr) {
            const int mode = e >> 8, ts = (e >> 5) & 7, lane = e & 31, t = ts >> 2, s = ts & 3;
            const int n = lane >> 2, k = lane & 3;
            T->frag[mode][ts][lane] = kreal_entry(K, 8, mode, dep3(n >> 1, t, ju), n & 1, dep3(k, s >> 1, ju), s & 1);
        }
#endif
        for (int e = tid; e < 4 * 32; e += nthr) {
            const int sidx = e >> 5, lane = e & 31;
            T->slot[sidx][lane] = (sidx < 2) ? G.slot(lane >> 2, dep3(lane & 3, sidx, ju))
                                             : G.slot((lane & 3) + 4 * (sidx - 2), permw3(lane >> 2, pa));
        }
        for (int lane = tid; lane < 32; lane += nthr) {
            const int r = permw3(lane >> 2, pa), c0 = permw3(2 * (lane & 3), pa), c1 = permw3(2 * (lane & 3) + 1, pa);
            T->widx[lane] = (r * 8 + c0) | ((r * 8 + c1) << 8);
        }
        for (int i = tid; i < B0TAB; i += nthr) T->b0[i] = G.batch_base(8 * i);
    } else {
        BlockGeom<LOG_CT, 2> G;
        G.init(op.q[0], op.q[1], 30);
        for (int e = tid; e < 3 * 2 * 32; e += nthr) {
            const int mode = e >> 6, s = (e >> 5) & 1, lane = e & 31;
            const int n = lane >> 2, k = lane & 3;
            T->frag[mode][s][lane] = kreal_entry(K, 4, mode, n >> 1, n & 1, k, s);
        }
        for (int e = tid; e < 4 * 32; e += nthr) {
            const int sidx = e >> 5, lane = e & 31, m = lane >> 2, kk = lane & 3;
            // W' operands of 2-qubit blocks are single doubles: component m of item kk + 4h (double units)
            T->slot[sidx][lane] = (sidx < 2) ? G.slot(m, kk) : 2 * G.slot(kk + 4 * (sidx - 2), m >> 1) + (m & 1);
        }
        for (int i = tid; i < B0TAB; i += nthr) T->b0[i] = G.batch_base(8 * i);
    }
}

// One CTA per (parameter set, op); ops the DMMA block path cannot take are skipped (their tables are never read).
__global__ void build_optabs(const DevOp* __restrict__ ops, int n_ops, const cplx* __restrict__ ktab, int kern_total,
                             const cplx* __restrict__ pool, int log_ct, OpTab* __restrict__ tabs) {
    const int b = blockIdx.x / n_ops, k = blockIdx.x - b * n_ops;
    const DevOp op = ops[k];
    if (op.ctrl_mask != 0 || (op.dim != 4 && op.dim != 8)) return;
    const cplx* K = op.kern_off >= 0 ? ktab + (size_t)b * kern_total + op.kern_off : pool + op.pool_off;
    OpTab* T = tabs + (size_t)b * n_ops + k;
    switch (log_ct) {
        case 0: fill_optab<0>(T, op, K); break;
        case 1: fill_optab<1>(T, op, K); break;
        case 2: fill_optab<2>(T, op, K); break;
        default: fill_optab<3>(T, op, K); break;
    }
}

// Per-lane operands of one DMMA block op, read from the op's table in shared memory: kernel fragments, load/store slots, the
// W' operand slots and output positions and -- when the warp's share of the tile is exactly four batches -- the four batch
// bases. The executor loads them for the NEXT op while it waits at the end-of-op barrier of the current one (blk_preload
// below, called between the last store of the loop and the barrier): warps finish their loops at different times, so these
// ~20 shared-memory loads and their latency overlap with the other warps' tensor work instead of following the barrier,
// where all warps of the CTA would sit in them at the same time with the tensor pipe idle.
struct BlkPre {
    double kd[2][4];  // forward: fragments of K; backward: of K^dagger   (3M: [0][0..3] = C0 C1 D0 D1, [1][0..1] = S0 S1)
    double kt[2][4];  // backward: fragments of K^T
    int sl[2], slw[2], widx;
    int b0v[4];
    bool ok, have_b0;
};

template <int KQ, bool BWD>
__device__ __forceinline__ void blk_preload(BlkPre& R, const OpTabS* T, int nitems, int lane, int warp, int nwarps) {
    constexpr int NT = (KQ == 3) ? 2 : 1, KS = 2 * NT;
    constexpr bool M3 = (KQ == 3) && (SQ_BLOCK_3M != 0);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        R.sl[t] = T->slot[t][lane];
#pragma unroll
        for (int s = 0; s < KS; ++s)
            if (!M3 || t * KS + s < 6) {
                R.kd[t][s] = T->frag[0][t * KS + s][lane];
                if (BWD) R.kt[t][s] = T->frag[1][t * KS + s][lane];
            }
    }
    if (BWD) {
        R.slw[0] = T->slot[2][lane];
        R.slw[1] = T->slot[3][lane];
        if (KQ == 3) R.widx = T->widx[lane];
    }
    // batches are dealt to the warps in contiguous runs; a run of exactly four comes with one 128-bit load of its bases
    R.have_b0 = (SQ_B0_MODE == 1) && (nitems >> 3) == 4 * nwarps && (nitems >> 3) <= B0TAB;
    if (R.have_b0) {
        if (SQ_BATCH_CONTIG) {
            const int4 v = *reinterpret_cast<const int4*>(&T->b0[4 * warp]);
            R.b0v[0] = v.x; R.b0v[1] = v.y; R.b0v[2] = v.z; R.b0v[3] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) R.b0v[i] = T->b0[warp + i * nwarps];
        }
    }
    R.ok = true;
}

// Called at the end of the executor's non-block branches: a preloaded BlkPre is never consumed there (the preload happens only
// when the next op is a block), but the compiler cannot know that and would keep ~50 operand registers alive (spilled)
// through those register-hungry paths. Overwriting them with constants ends their live ranges.
__device__ __forceinline__ void blk_kill(BlkPre& R) {
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int s = 0; s < 4; ++s) R.kd[t][s] = R.kt[t][s] = 0.0;
    R.sl[0] = R.sl[1] = R.slw[0] = R.slw[1] = R.widx = 0;
    R.b0v[0] = R.b0v[1] = R.b0v[2] = R.b0v[3] = 0;
    R.ok = R.have_b0 = false;
}

// runs body(B0) for every batch of this warp: contiguous runs of ceil(nb / nwarps) batches per warp
template <int LOG_CT, int KQ, typename BODY>
__device__ __forceinline__ void blk_for_batches(const BlkPre& R, const OpTabS* T, int q0, int q1, int q2, int nitems, int warp,
                                                int nwarps, BODY&& body) {
    if (R.have_b0) {
        // not unrolled: four copies of the body would hoist four batches' worth of addresses (register pressure)
#pragma unroll 1
        for (int i = 0; i < 4; ++i) body(i < 2 ? (i == 0 ? R.b0v[0] : R.b0v[1]) : (i == 2 ? R.b0v[2] : R.b0v[3]));
        return;
    }
    const int nb = nitems >> 3, nbw = (nb + nwarps - 1) / nwarps;
    const int i0 = SQ_BATCH_CONTIG ? warp * nbw : warp, i1 = SQ_BATCH_CONTIG ? min(nb, i0 + nbw) : nb, di = SQ_BATCH_CONTIG ? 1 : nwarps;
    const bool tab_b0 = (SQ_B0_MODE == 1) && nb <= B0TAB;
    BlockGeom<LOG_CT, KQ> G;
    if (!tab_b0) G.init(q0, q1, q2);
    for (int i = i0; i < i1; i += di) body(tab_b0 ? T->b0[i] : G.batch_base(8 * i));
}

// forward: x <- K x for every group of the tile. One warp handles 8 (group, column) items per step.
// GST: the results are written to the state in HBM (gdst[sglob[slot]], window segments' last op) instead of the tile.
template <int LOG_CT, int KQ, bool GST>
__device__ __forceinline__ void block_dmma_forward(cplx* sa, const OpTabS* T, BlkPre& R, int q0, int q1, int q2, int rows, int tid, int nthr,
                                                   cplx* __restrict__ gdst, const int* sglob) {
    constexpr int NT = (KQ == 3) ? 2 : 1;
    constexpr bool M3 = (KQ == 3) && (SQ_BLOCK_3M != 0);
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int nitems = (rows >> KQ) << LOG_CT;
    if (!R.ok) blk_preload<KQ, false>(R, T, nitems, lane, warp, nwarps);
    blk_for_batches<LOG_CT, KQ>(R, T, q0, q1, q2, nitems, warp, nwarps, [&](int B0) {
        cplx x[NT], d[NT];
#pragma unroll
        for (int u = 0; u < NT; ++u) {
            x[u] = sa[B0 ^ R.sl[u]];
            d[u] = czero();
        }
        if (M3) {
            // k1 = C (u + v) over the two k-steps, then Re = k1 - (C + D) v and Im = k1 + (D - C) u accumulate ON TOP of copies of
            // k1 (the DMMA's addend): no additions after the products. The D fragment pair holds outputs o(2j), o(2j+1), which
            // are the amplitudes the lane loaded as x[0], x[1].
            double k1[2] = {0.0, 0.0};
            dmma_m8n8k4(k1[0], k1[1], x[0].x + x[0].y, R.kd[0][0]);
            dmma_m8n8k4(k1[0], k1[1], x[NT - 1].x + x[NT - 1].y, R.kd[0][1]);
            double re[2] = {k1[0], k1[1]}, im[2] = {k1[0], k1[1]};
            dmma_m8n8k4(re[0], re[1], x[0].y, R.kd[NT - 1][0]);
            dmma_m8n8k4(im[0], im[1], x[0].x, R.kd[0][2]);
            dmma_m8n8k4(re[0], re[1], x[NT - 1].y, R.kd[NT - 1][1]);
            dmma_m8n8k4(im[0], im[1], x[NT - 1].x, R.kd[0][3]);
            d[0] = cmake(re[0], im[0]);
            d[NT - 1] = cmake(re[1], im[1]);
        } else {
#pragma unroll
            for (int u = 0; u < NT; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    dmma_m8n8k4(d[t].x, d[t].y, x[u].x, R.kd[t][2 * u]);
                    dmma_m8n8k4(d[t].x, d[t].y, x[u].y, R.kd[t][2 * u + 1]);
                }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (GST) gdst[(size_t)(unsigned)sglob[B0 ^ R.sl[t]]] = d[t];
            else sa[B0 ^ R.sl[t]] = d[t];
        }
    });
    R.ok = false;
}

// backward step of the adjoint sweep: W' += beta p^T (outer product over items), a <- K^dagger p, beta <- K^T beta, all on
// DMMA; the warp's W' (DIM x DIM complex) is written to wslot after the loop.
template <int LOG_CT, int KQ, bool GST>
__device__ __forceinline__ void block_dmma_backward(cplx* sa, cplx* sb, const OpTabS* T, BlkPre& R, int q0, int q1, int q2, int rows,
                                                    bool has_w, cplx* wslot, bool wdirect, int tid, int nthr,
                                                    cplx* __restrict__ ga, cplx* __restrict__ gb, const int* sglob) {
    constexpr int DIM = 1 << KQ, NT = (KQ == 3) ? 2 : 1;
    constexpr bool M3 = (KQ == 3) && (SQ_BLOCK_3M != 0);
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int nitems = (rows >> KQ) << LOG_CT;
    if (!R.ok) blk_preload<KQ, true>(R, T, nitems, lane, warp, nwarps);
    double pacc[NT][NT][2];
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b) pacc[a][b][0] = pacc[a][b][1] = 0.0;
    blk_for_batches<LOG_CT, KQ>(R, T, q0, q1, q2, nitems, warp, nwarps, [&](int B0) {
        cplx p[NT], be[NT], da[NT], db[NT];
#pragma unroll
        for (int u = 0; u < NT; ++u) {
            p[u] = sa[B0 ^ R.sl[u]];
            be[u] = sb[B0 ^ R.sl[u]];
            da[u] = czero();
            db[u] = czero();
        }
        if (has_w) {
            if (M3) {
                // three real outer products per half: T1 += br pr^T, T2 += bi pi^T, T3 += (br + bi)(pr + pi)^T
                // (pacc[0][0] = T1, pacc[NT-1][NT-1] = T2, pacc[0][NT-1] = T3)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const cplx bw = sb[B0 ^ R.slw[h]], pw = sa[B0 ^ R.slw[h]];
                    dmma_m8n8k4(pacc[0][0][0], pacc[0][0][1], bw.x, pw.x);
                    dmma_m8n8k4(pacc[NT - 1][NT - 1][0], pacc[NT - 1][NT - 1][1], bw.y, pw.y);
                    dmma_m8n8k4(pacc[0][NT - 1][0], pacc[0][NT - 1][1], bw.x + bw.y, pw.x + pw.y);
                }
            } else if (KQ == 3) {
                // lane (rho, it): amplitude permw(rho) of item it + 4h of beta (A operand, rows = rho, tile = re/im) and of p
                // (B operand, columns = rho)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const cplx bw = sb[B0 ^ R.slw[h]], pw = sa[B0 ^ R.slw[h]];
                    dmma_m8n8k4(pacc[0][0][0], pacc[0][0][1], bw.x, pw.x);
                    dmma_m8n8k4(pacc[0][NT - 1][0], pacc[0][NT - 1][1], bw.x, pw.y);
                    dmma_m8n8k4(pacc[NT - 1][0][0], pacc[NT - 1][0][1], bw.y, pw.x);
                    dmma_m8n8k4(pacc[NT - 1][NT - 1][0], pacc[NT - 1][NT - 1][1], bw.y, pw.y);
                }
            } else {
                const double* sad = reinterpret_cast<const double*>(sa);
                const double* sbd = reinterpret_cast<const double*>(sb);
#pragma unroll
                for (int h = 0; h < 2; ++h) dmma_m8n8k4(pacc[0][0][0], pacc[0][0][1], sbd[(2 * B0) ^ R.slw[h]], sad[(2 * B0) ^ R.slw[h]]);
            }
        }
        if (M3) {
            // as in the forward step: k1 = C (u + v), Re = k1 - (C + D) v, Im = k1 + (D - C) u, for a <- K^dagger p (fragments kd)
            // and beta <- K^T beta (fragments kt), the two chains interleaved
            double ka[2] = {0.0, 0.0}, kb[2] = {0.0, 0.0};
#pragma unroll
            for (int st = 0; st < 2; ++st) {
                const cplx pp = p[st == 0 ? 0 : NT - 1], bb = be[st == 0 ? 0 : NT - 1];
                dmma_m8n8k4(ka[0], ka[1], pp.x + pp.y, R.kd[0][st]);
                dmma_m8n8k4(kb[0], kb[1], bb.x + bb.y, R.kt[0][st]);
            }
            double are[2] = {ka[0], ka[1]}, aim[2] = {ka[0], ka[1]}, bre[2] = {kb[0], kb[1]}, bim[2] = {kb[0], kb[1]};
#pragma unroll
            for (int st = 0; st < 2; ++st) {
                const cplx pp = p[st == 0 ? 0 : NT - 1], bb = be[st == 0 ? 0 : NT - 1];
                dmma_m8n8k4(are[0], are[1], pp.y, R.kd[NT - 1][st]);
                dmma_m8n8k4(bre[0], bre[1], bb.y, R.kt[NT - 1][st]);
                dmma_m8n8k4(aim[0], aim[1], pp.x, R.kd[0][2 + st]);
                dmma_m8n8k4(bim[0], bim[1], bb.x, R.kt[0][2 + st]);
            }
            da[0] = cmake(are[0], aim[0]);
            da[NT - 1] = cmake(are[1], aim[1]);
            db[0] = cmake(bre[0], bim[0]);
            db[NT - 1] = cmake(bre[1], bim[1]);
        } else {
#pragma unroll
            for (int u = 0; u < NT; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    dmma_m8n8k4(da[t].x, da[t].y, p[u].x, R.kd[t][2 * u]);
                    dmma_m8n8k4(db[t].x, db[t].y, be[u].x, R.kt[t][2 * u]);
                    dmma_m8n8k4(da[t].x, da[t].y, p[u].y, R.kd[t][2 * u + 1]);
                    dmma_m8n8k4(db[t].x, db[t].y, be[u].y, R.kt[t][2 * u + 1]);
                }
        }
        if (has_w) __syncwarp();  // the W' operands (other lanes' elements) are read before anybody overwrites them
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (GST) {
                const size_t off = (size_t)(unsigned)sglob[B0 ^ R.sl[t]];
                ga[off] = da[t];
                gb[off] = db[t];
            } else {
                sa[B0 ^ R.sl[t]] = da[t];
                sb[B0 ^ R.sl[t]] = db[t];
            }
        }
    });
    if (has_w) {
        if (KQ == 3) {
            // lane (rho, jj) holds P[(rho, a)][(2jj + e, b)] in pacc[a][b][e]:
            // W'[rho][gamma] = (P[re][re] - P[im][im]) + i (P[re][im] + P[im][re])
            const int wi = R.widx;
            cplx w0, w1;
            if (M3) {  // W' = (T1 - T2) + i (T3 - T1 - T2)
                w0 = cmake(pacc[0][0][0] - pacc[NT - 1][NT - 1][0], pacc[0][NT - 1][0] - pacc[0][0][0] - pacc[NT - 1][NT - 1][0]);
                w1 = cmake(pacc[0][0][1] - pacc[NT - 1][NT - 1][1], pacc[0][NT - 1][1] - pacc[0][0][1] - pacc[NT - 1][NT - 1][1]);
            } else {
                w0 = cmake(pacc[0][0][0] - pacc[NT - 1][NT - 1][0], pacc[0][NT - 1][0] + pacc[NT - 1][0][0]);
                w1 = cmake(pacc[0][0][1] - pacc[NT - 1][NT - 1][1], pacc[0][NT - 1][1] + pacc[NT - 1][0][1]);
            }
            if (wdirect) {
                double* wd = reinterpret_cast<double*>(wslot);
                atomicAdd(wd + 2 * (wi & 255), w0.x);
                atomicAdd(wd + 2 * (wi & 255) + 1, w0.y);
                atomicAdd(wd + 2 * ((wi >> 8) & 255), w1.x);
                atomicAdd(wd + 2 * ((wi >> 8) & 255) + 1, w1.y);
            } else {
                wslot[wi & 255] = w0;
                wslot[(wi >> 8) & 255] = w1;
            }
        } else {
            // lane (m, kk) holds P[m][2kk], P[m][2kk+1]; rows 2r (even m) and 2r+1 (odd m) combine to
            // W'[r][c] = (P[2r][2c] - P[2r+1][2c+1]) + i (P[2r][2c+1] + P[2r+1][2c]),  r = m/2, c = kk
            const int m = lane >> 2, kk = lane & 3;
            const double o0 = __shfl_xor_sync(0xffffffffu, pacc[0][0][0], 4);
            const double o1 = __shfl_xor_sync(0xffffffffu, pacc[0][0][1], 4);
            if ((m & 1) == 0) {
                if (wdirect) {
                    double* wd = reinterpret_cast<double*>(wslot + (m >> 1) * DIM + kk);
                    atomicAdd(wd, pacc[0][0][0] - o1);
                    atomicAdd(wd + 1, pacc[0][0][1] + o0);
                } else {
                    wslot[(m >> 1) * DIM + kk] = cmake(pacc[0][0][0] - o1, pacc[0][0][1] + o0);
                }
            }
        }
    }
    R.ok = false;
}

// Raw dense 3-/4-qubit kernels (GENERAL blocks; no controls) in the same formulation as the fused blocks: data = A operand,
// 128-bit per-lane loads / stores of the same elements, kernel fragments (NT x KS doubles per lane) in registers. Their
// kernels are constants of the circuit, so the lane-ordered fragment table is built once per (circuit, tile width) by
// build_dense_tabs and read straight from global memory / L2 at the start of the op (32 coalesced 256 B loads per warp);
// the two block-qubit positions that index the lane's amplitude slot are picked for the fewest bank conflicts.
struct DenseTab {
    double frag[1024];  // [t * KS + ks][lane], NT * KS <= 32
    int sl[4][32];      // [u][lane]: load/store slot (complex units) of amplitude dep(lane & 3, u) of item lane >> 2
};
// 5-qubit kernels: 8 n-tiles x 16 k-steps = 128 fragments per lane do not fit registers; they are read per DMMA from this
// table through L1 (32 KB, shared by every warp of the SM, coalesced 256 B per warp load)
struct DenseTab5 {
    double frag[4096];
    int sl[8][32];
};

template <int LOG_CT, int KQ, typename TAB>
__device__ __forceinline__ void dense_tab_fill(TAB* T, int* schoice, const cplx* __restrict__ K, const DevOp& op, int tid, int nthr, int mode = 0) {
    constexpr int CT = 1 << LOG_CT, LOGG = 3 - LOG_CT, DIM = 1 << KQ, NT = DIM / 4, KS = 2 * NT;
    int Pq[KQ], F[3];
    unsigned qmask = 0;
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
        Pq[j] = elem<LOG_CT>(1 << op.q[j], 0);
        qmask |= 1u << op.q[j];
    }
    {
        int f = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            while ((qmask >> f) & 1u) ++f;
            F[j] = elem<LOG_CT>(1 << f, 0);
            ++f;
        }
    }
    // amplitude index of lane slot j (2 bits, on positions ja < jb) and k-step group u (the other positions, ascending)
    auto dep = [&](int j, int u, int ja, int jb) {
        int amp = ((j & 1) << ja) | (((j >> 1) & 1) << jb), ub = 0;
#pragma unroll
        for (int pos = 0; pos < KQ; ++pos)
            if (pos != ja && pos != jb) {
                amp |= ((u >> ub) & 1) << pos;
                ++ub;
            }
        return amp;
    };
    auto slot = [&](int item, int amp) {
        const int go = item >> LOG_CT;
        int a = item & (CT - 1);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j < LOGG && ((go >> j) & 1)) a ^= F[j];
#pragma unroll
        for (int j = 0; j < KQ; ++j)
            if ((amp >> j) & 1) a ^= Pq[j];
        return a;
    };
    if (tid == 0) {
        int best = -1, bja = 0, bjb = 1;
        for (int ja = 0; ja < KQ; ++ja)
            for (int jb = ja + 1; jb < KQ; ++jb) {
                unsigned seen = 0;
                for (int l = 0; l < 8; ++l) seen |= 1u << (slot(l >> 2, dep(l & 3, 0, ja, jb)) & 7);
                if (__popc(seen) > best) {
                    best = __popc(seen);
                    bja = ja;
                    bjb = jb;
                }
            }
        schoice[0] = bja;
        schoice[1] = bjb;
    }
    __syncthreads();
    const int ja = schoice[0], jb = schoice[1];
    if ((KQ <= 4 && SQ_DENSE_3M) || (KQ == 5 && SQ_DENSE5_3M)) {
        // three-product form: frag[(m * NTL + nt) * NT + ks][lane], m in {0: C, 1: D - C, 2: -(C + D)} of K = C + iD, n-tile nt
        // (8 output amplitudes), k-step ks (4 input amplitudes). Lane (k = lane & 3, n = lane >> 2) holds the entry
        // [out amplitude dep(n >> 1, (n & 1) | 2 nt)][in amplitude dep(k, ks)]: the lane's inputs dep(j, u), u < NT, are its outputs
        constexpr int NTL = DIM / 8;
        for (int e = tid; e < 3 * NTL * NT * 32; e += nthr) {
            const int l = e & 31, idx = e >> 5, m = idx / (NTL * NT), rem = idx - m * NTL * NT, nt = rem / NT, ks = rem - nt * NT;
            const int n = l >> 2, k = l & 3;
            const int ao = dep(n >> 1, (n & 1) | (nt << 1), ja, jb), ai = dep(k, ks, ja, jb);
            const cplx kv = (mode == 0) ? K[ao * DIM + ai] : K[ai * DIM + ao];  // mode 1: K^dagger, 2: K^T
            const double cre = kv.x, dim_ = (mode == 1) ? -kv.y : kv.y;
            T->frag[e] = (m == 0) ? cre : (m == 1 ? dim_ - cre : -(cre + dim_));
        }
    } else {
        for (int e = tid; e < NT * KS * 32; e += nthr) {
            const int ts = e >> 5, l = e & 31, t = ts / KS, ks = ts - t * KS, n = l >> 2, k = l & 3;
            T->frag[e] = kreal_entry(K, DIM, mode, dep(n >> 1, t, ja, jb), n & 1, dep(k, ks >> 1, ja, jb), ks & 1);
        }
    }
    for (int e = tid; e < NT * 32; e += nthr) T->sl[e >> 5][e & 31] = slot((e & 31) >> 2, dep(e & 3, e >> 5, ja, jb));
}

// dtab > 0: 1 + index into the 3-/4-qubit tables; dtab < 0: -1 - index into the 5-qubit tables
template <int LOG_CT>
__global__ void build_dense_tabs(const DevOp* __restrict__ ops, int n_ops, const cplx* __restrict__ pool, DenseTab* __restrict__ tabs,
                                 DenseTab5* __restrict__ tabs5) {
    __shared__ int schoice[2];
    const DevOp op = ops[blockIdx.x];
    if (op.dtab == 0) return;
    if (op.dtab < 0) {
        dense_tab_fill<LOG_CT, 5>(tabs5 + (-1 - op.dtab), schoice, pool + op.pool_off, op, threadIdx.x, blockDim.x);
        return;
    }
    // three tables per op: K (forward sweep), K^dagger and K^T (adjoint sweep: a <- K^dagger a, beta <- K^T beta)
    for (int mode = 0; mode < 3; ++mode) {
        DenseTab* T = tabs + 3 * (op.dtab - 1) + mode;
        if (op.nq == 3) dense_tab_fill<LOG_CT, 3>(T, schoice, pool + op.pool_off, op, threadIdx.x, blockDim.x, mode);
        else dense_tab_fill<LOG_CT, 4>(T, schoice, pool + op.pool_off, op, threadIdx.x, blockDim.x, mode);
        __syncthreads();
    }
}

// 5-qubit kernels: 128 DMMA per 8-item batch in the real embedding (96 in the three-product form, SQ_DENSE5_3M), kernel
// fragments streamed from the (L1-resident) table
template <int LOG_CT>
__device__ __forceinline__ void dense_dmma_forward5(cplx* sa, const DenseTab5* __restrict__ T, const DevOp& op, int rows, int tid, int nthr) {
    constexpr int KQ = 5, NT = 8;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    if constexpr (SQ_DENSE5_3M != 0) {
        // per n-tile (8 output amplitudes): k1 = C (u + v) over the 8 k-steps, then Re = k1 - (C + D) v and Im = k1 + (D - C) u on
        // copies of k1; the lane's outputs of n-tile nt are the amplitudes it loaded as x[2 nt], x[2 nt + 1]
        constexpr int NTL = 4;
        int sl3[NT], q3[KQ];
#pragma unroll
        for (int t = 0; t < NT; ++t) sl3[t] = T->sl[t][lane];
#pragma unroll
        for (int j = 0; j < KQ; ++j) q3[j] = op.q[j];
        const double* __restrict__ fr = T->frag + lane;
        const int nitems3 = (rows >> KQ) << LOG_CT;
        for (int b0 = warp * 8; b0 < nitems3; b0 += nwarps * 8) {
            int base = b0 >> LOG_CT;
#pragma unroll
            for (int j = 0; j < KQ; ++j) base = insert_zero(base, q3[j]);
            const int B0 = elem<LOG_CT>(base, 0);
            cplx x[NT];
            double sm[NT];
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                x[u] = sa[B0 ^ sl3[u]];
                sm[u] = x[u].x + x[u].y;
            }
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                double k1[2] = {0.0, 0.0};
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) dmma_m8n8k4(k1[0], k1[1], sm[ks], __ldg(fr + ((0 * NTL + nt) * NT + ks) * 32));
                double re[2] = {k1[0], k1[1]}, im[2] = {k1[0], k1[1]};
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) {
                    dmma_m8n8k4(re[0], re[1], x[ks].y, __ldg(fr + ((2 * NTL + nt) * NT + ks) * 32));
                    dmma_m8n8k4(im[0], im[1], x[ks].x, __ldg(fr + ((1 * NTL + nt) * NT + ks) * 32));
                }
                // a lane's stores hit elements it loaded itself (all of them are in x[] by now): no hazard with other lanes
                sa[B0 ^ sl3[2 * nt]] = cmake(re[0], im[0]);
                sa[B0 ^ sl3[2 * nt + 1]] = cmake(re[1], im[1]);
            }
        }
    } else {
    constexpr int KS = 16, TH = 4;  // real embedding: n-tiles are processed in two halves of TH (8 accumulator registers live)
    int sl[NT], q[KQ];
#pragma unroll
    for (int t = 0; t < NT; ++t) sl[t] = T->sl[t][lane];
#pragma unroll
    for (int j = 0; j < KQ; ++j) q[j] = op.q[j];
    const int nitems = (rows >> KQ) << LOG_CT;
    for (int b0 = warp * 8; b0 < nitems; b0 += nwarps * 8) {
        int base = b0 >> LOG_CT;
#pragma unroll
        for (int j = 0; j < KQ; ++j) base = insert_zero(base, q[j]);
        const int B0 = elem<LOG_CT>(base, 0);
        cplx x[NT];
#pragma unroll
        for (int u = 0; u < NT; ++u) x[u] = sa[B0 ^ sl[u]];
#pragma unroll 1
        for (int th = 0; th < NT / TH; ++th) {
            const double* __restrict__ fr = T->frag + (size_t)th * TH * KS * 32 + lane;
            cplx d[TH];
#pragma unroll
            for (int t = 0; t < TH; ++t) d[t] = czero();
#pragma unroll
            for (int u = 0; u < NT; ++u)
#pragma unroll
                for (int t = 0; t < TH; ++t) {
                    dmma_m8n8k4(d[t].x, d[t].y, x[u].x, __ldg(fr + (t * KS + 2 * u) * 32));
                    dmma_m8n8k4(d[t].x, d[t].y, x[u].y, __ldg(fr + (t * KS + 2 * u + 1) * 32));
                }
            // a lane's stores hit the elements it loaded itself (all of them are in x[] by now): no hazard with other lanes
#pragma unroll
            for (int t = 0; t < TH; ++t) sa[B0 ^ __ldg(&T->sl[th * TH + t][lane])] = d[t];  // (slot re-read: no dynamic register index)
        }
    }
}
}


template <int LOG_CT, int KQ>
__device__ __forceinline__ void dense_dmma_forward2(cplx* sa, const DenseTab* __restrict__ T, const DevOp& op, int rows, int tid, int nthr) {
    constexpr int DIM = 1 << KQ, NT = DIM / 4;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    if constexpr (SQ_DENSE_3M != 0) {
        // k1 = C (u + v), Re = k1 - (C + D) v, Im = k1 + (D - C) u per n-tile (see block_dmma_forward): 3 * NTL * NT DMMA per batch
        constexpr int NTL = DIM / 8;
        double f[3][NTL][NT];
        int sl[NT];
#pragma unroll
        for (int u = 0; u < NT; ++u) sl[u] = T->sl[u][lane];
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) f[m][nt][ks] = T->frag[((m * NTL + nt) * NT + ks) * 32 + lane];
        int q[KQ];
#pragma unroll
        for (int j = 0; j < KQ; ++j) q[j] = op.q[j];
        const int nitems = (rows >> KQ) << LOG_CT;
        for (int b0 = warp * 8; b0 < nitems; b0 += nwarps * 8) {
            int base = b0 >> LOG_CT;
#pragma unroll
            for (int j = 0; j < KQ; ++j) base = insert_zero(base, q[j]);
            const int B0 = elem<LOG_CT>(base, 0);
            cplx x[NT], d[NT];
            double sm[NT];
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                x[u] = sa[B0 ^ sl[u]];
                sm[u] = x[u].x + x[u].y;
            }
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                double k1[2] = {0.0, 0.0};
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) dmma_m8n8k4(k1[0], k1[1], sm[ks], f[0][nt][ks]);
                double re[2] = {k1[0], k1[1]}, im[2] = {k1[0], k1[1]};
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) {
                    dmma_m8n8k4(re[0], re[1], x[ks].y, f[2][nt][ks]);
                    dmma_m8n8k4(im[0], im[1], x[ks].x, f[1][nt][ks]);
                }
                d[2 * nt] = cmake(re[0], im[0]);
                d[2 * nt + 1] = cmake(re[1], im[1]);
            }
#pragma unroll
            for (int t = 0; t < NT; ++t) sa[B0 ^ sl[t]] = d[t];
        }
    } else {
        constexpr int KS = 2 * NT;
        double kf[NT][KS];
        int sl[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            sl[t] = T->sl[t][lane];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) kf[t][ks] = T->frag[(t * KS + ks) * 32 + lane];
        }
        int q[KQ];
#pragma unroll
        for (int j = 0; j < KQ; ++j) q[j] = op.q[j];
        const int nitems = (rows >> KQ) << LOG_CT;
        for (int b0 = warp * 8; b0 < nitems; b0 += nwarps * 8) {
            int base = b0 >> LOG_CT;
#pragma unroll
            for (int j = 0; j < KQ; ++j) base = insert_zero(base, q[j]);
            const int B0 = elem<LOG_CT>(base, 0);
            cplx x[NT], d[NT];
#pragma unroll
            for (int u = 0; u < NT; ++u) {
                x[u] = sa[B0 ^ sl[u]];
                d[u] = czero();
            }
#pragma unroll
            for (int u = 0; u < NT; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    dmma_m8n8k4(d[t].x, d[t].y, x[u].x, kf[t][2 * u]);
                    dmma_m8n8k4(d[t].x, d[t].y, x[u].y, kf[t][2 * u + 1]);
                }
#pragma unroll
            for (int t = 0; t < NT; ++t) sa[B0 ^ sl[t]] = d[t];
        }
    }
}

// compact per-op record staged in shared memory (uniform reads, no global latency on the critical path)
struct SOp {
    int32_t dim;       // 2 / 4 / 8 for block ops
    int32_t q0, q1, q2;  // block ops: qubits ascending (unused entries 30)
    int32_t kern_off;  // kernel-table offset, -1: pool
    int32_t w_off;
    int32_t kind;      // 0: scalar 2x2 block, 2: DMMA 4x4 / 8x8 block, 1: generic path (controls, raw dense, derivative op)
    int32_t pool_lo, pool_hi;
};

static const int TAB_RING = 4;   // shared-memory ring of op tables: the table of op k + TAB_RING - 1 is requested when op k starts

#if SQ_STAGGER
__device__ int g_sm_arrivals[256];
#endif
// Profiling build (-DSQ_TRACE=events): the first two CTAs that land on SM 0 record, per op and warp, the SM clock at the start
// and at the end of the op's work loop (profiles/trace_phases.py reads them through sqgpu_debug_trace and computes how much
// the two CTAs' tensor phases overlap).
#ifndef SQ_TRACE
#define SQ_TRACE 0
#endif
#if SQ_TRACE
__device__ long long g_trace[2][SQ_TRACE][16][2];
__device__ int g_trace_arrivals;
#define SQ_TRACE_MARK(which)                                                                          \
    if (trace_slot >= 0 && trace_ev < SQ_TRACE && lane == 0 && warp < 16) g_trace[trace_slot][trace_ev][warp][which] = clock64();
#define SQ_TRACE_NEXT() ++trace_ev;
#else
#define SQ_TRACE_MARK(which)
#define SQ_TRACE_NEXT()
#endif

// CLU: cluster executor (cost / gradient of matrices whose column -- or column + row functional -- does not fit ONE CTA's
// shared memory, n = 12...15): the 2^rho CTAs of a thread-block cluster share a column tile, CTA `rank` holding the rows whose
// `rho` SPLIT qubits spell its rank. Ops never touch a split qubit: the host planner (build_cluster_plan) inserts RESPLIT ops
// that exchange a split qubit with a local one through distributed shared memory (each CTA swaps half of its tile with ONE
// partner) before an op needs it, and rewrites every op's qubits to the row-bit positions they have at that point. Everything
// else -- block path, tables, W' slices, trace partials (one chunk per CTA) -- is the single-CTA executor.
// DNS: the kernel carries the paths of RAW dense ops (GENERAL blocks, multiplied-out constant sub-circuits, controlled
// two-target gates, materialised block derivatives): dense DMMA forward / adjoint steps and the generic scalar path. Circuits
// without such ops -- every fused decomposition structure -- run the DNS = false instantiation: the mere presence of those
// branches costs the block path registers and schedule (measured: C5 backward sweep 45.7 ms with the dense adjoint step
// compiled in, 42.8 ms without; C3 1 496 -> 1 527 evals/s and 128 -> 112 registers without any dense path).
template <int MODE, int LOG_CT, bool CLU = false, bool DNS = true>
__global__ void __launch_bounds__(FUSED_THREADS, 1) fused_exec(const ExecArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int CT = 1 << LOG_CT;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int rows = A.rows;
    const int y = blockIdx.y;
    const int kset = A.k_shared ? 0 : y;
    const cplx* __restrict__ ktab = A.ktab + (size_t)kset * A.kern_total;
    const cplx* __restrict__ dktab = A.dktab ? A.dktab + (size_t)kset * A.dkern_total : nullptr;

    cplx* sa = reinterpret_cast<cplx*>(smem_raw);
    constexpr bool HAS_B = (MODE == MODE_GRAD || MODE == MODE_BWD);
    constexpr bool WIN = (MODE == MODE_APPLY || MODE == MODE_BWD);  // window mode is compiled only where it is used
    const bool dbuf = (MODE == MODE_APPLY) && A.dbuf && A.wmask;        // two tile buffers (window forward segments)
    cplx* const sa0 = sa;
    cplx* sb = sa + (size_t)rows * CT * (dbuf ? 2 : 1);                  // the row functional beta (HAS_B only)
    cplx* sk = HAS_B ? sb + (size_t)rows * CT : sb;                       // raw dense kernel staging
    OpTabS* stab = reinterpret_cast<OpTabS*>(sk + A.dense_stage);         // [TAB_RING] DMMA block lookup tables (one sweep direction);
                                                                          // the 2 x 2 kernel of a single-qubit block travels in the same ring
    unsigned long long* tbar = reinterpret_cast<unsigned long long*>(stab + TAB_RING);  // [TAB_RING] mbarriers of the table ring
    const bool wdirect = HAS_B && SQ_W_DIRECT && A.w_direct;              // W' partials go from the registers to the warp's own global slice
    cplx* swarp = reinterpret_cast<cplx*>(tbar + TAB_RING);               // [2][nwarps][wmax] (not with wdirect)
    cplx* swacc = swarp + ((HAS_B && !wdirect) ? 2 * nwarps * A.wmax : 0);  // [w_total] if w_in_smem
    double* sred = reinterpret_cast<double*>(swacc + ((HAS_B && A.w_in_smem) ? A.w_total : 0));  // [nwarps][6]
    double* stsum = sred + nwarps * 6;                                   // [6] running trace sums of this CTA over its tiles
    SOp* sops = reinterpret_cast<SOp*>(stsum + 6);                       // [n_ops]
    int* sglob = reinterpret_cast<int*>(sops + A.n_ops);                 // window mode: [rows * CT] element offset in the state of the
                                                                         // amplitude stored in shared-memory slot e (elem() order)

    const int chunk = blockIdx.x;
    const int nchunks = gridDim.x;
    // cluster executor: rank of this CTA inside its cluster, cluster index (tiles are dealt to clusters)
    unsigned crank = 0;
    int tile_owner = chunk;
    __shared__ int s_fold_cnt[2];  // SQ_FOLD_EARLY: arrival tickets of the current / next op
    __shared__ signed char sfin[CLU ? 32 : 1];  // cluster executor: A.fin_pos (indexed dynamically: kept out of the parameter space)
    if constexpr (CLU) {
        crank = cooperative_groups::this_cluster().block_rank();
        tile_owner = chunk >> A.rho;
        if (tid == 0) {
#pragma unroll
            for (int qb = 0; qb < 32; ++qb) sfin[qb] = A.fin_pos[qb];
        }
    }
#if SQ_TRACE
    int trace_slot = -1, trace_ev = 0;
    {
        __shared__ int s_trace_slot;
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_trace_slot = -1;
            if (smid == 0 && (MODE == MODE_GRAD || MODE == MODE_BWD)) {
                const int a = atomicAdd(&g_trace_arrivals, 1);
                if (a < 2) s_trace_slot = a;
            }
        }
        __syncthreads();
        trace_slot = s_trace_slot;
    }
#endif
#if SQ_STAGGER
    if (MODE == MODE_GRAD || MODE == MODE_BWD || MODE == MODE_COST) {
        __shared__ int s_arrival;
        if (tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_arrival = atomicAdd(&g_sm_arrivals[smid & 255], 1);
        }
        __syncthreads();
        if (s_arrival & 1) __nanosleep(SQ_STAGGER);
    }
#endif
    const int deriv_op = (MODE == MODE_APPLY && A.deriv_op) ? A.deriv_op[y] : -1;
    const int deriv_slot = (MODE == MODE_APPLY && A.deriv_slot) ? A.deriv_slot[y] : 0;

    // ---- stage the op table ------------------------------------------------------------------------------------------
    for (int k = tid; k < A.n_ops; k += nthr) {
        const DevOp op = A.ops[k];
        SOp s;
        s.dim = op.dim;
        s.kern_off = op.kern_off;
        s.w_off = op.w_off;
        s.pool_lo = (int32_t)(op.pool_off & 0xffffffffLL);
        s.pool_hi = (int32_t)(op.pool_off >> 32);
        s.kind = 1;
        if (op.ctrl_mask == 0 && k != deriv_op) {
            if (op.dim == 2) s.kind = 0;
            else if ((op.dim == 4 || op.dim == 8) && ((((rows >> op.nq) << LOG_CT) & 7) == 0) && nthr >= 32) s.kind = 2;
        }
        s.q0 = op.dim == 2 ? op.target : op.q[0];
        s.q1 = op.dim == 2 ? 30 : op.q[1];
        s.q2 = op.dim == 8 ? op.q[2] : 30;
        if (CLU && op.type == SQ_OP_RESPLIT) {  // cluster executor: q0 = local row bit, q1 = cluster-rank bit
            s.kind = 3;
            s.q0 = op.target;
            s.q1 = op.nq;
        }
        sops[k] = s;
    }
    if (HAS_B && A.w_in_smem) {
        for (int e = tid; e < A.w_total; e += nthr) swacc[e] = czero();
    }
    if (WIN && A.wmask) {
        // offset of (row r, tile column c) = deposit(r, wmask) | deposit(c, ~wmask); deposit(r) = deposit(r & 63) | deposit(r & ~63):
        // two small tables (built with the software PDEP loop by 64 + rows / 64 threads, in the not yet loaded tile area),
        // then one scatter into slot order -- the tile loads, the stores and the fused store of a segment's last op all index
        // by shared-memory slot
        int* slo = reinterpret_cast<int*>(sa);
        int* shi = slo + 64;
        int* scol = shi + max(rows >> 6, 1);
        for (int t = tid; t < 64 + max(rows >> 6, 1); t += nthr)
            slo[t] = (int)deposit_bits(t < 64 ? (unsigned)t : (unsigned)(t - 64) << 6, A.wmask);
        if (tid < CT) scol[tid] = (int)deposit_bits((unsigned)tid, ~A.wmask);
        __syncthreads();
        for (int e = tid; e < rows * CT; e += nthr) {
            const int r = e >> LOG_CT, cc = e & (CT - 1);
            sglob[elem<LOG_CT>(r, cc)] = slo[r & 63] | shi[r >> 6] | scol[cc];
        }
        __syncthreads();
    }
    if (tid < 6) stsum[tid] = 0.0;
    if (tid < 2) s_fold_cnt[tid] = 0;
    __syncthreads();

    // the 2 x 2 kernel of single-qubit block k (kind 0): 64 bytes that take the place of the table in the op's ring slot
    auto kernel_of = [&](int k) -> const cplx* {
        const SOp s = sops[k];
        return s.kern_off >= 0 ? ktab + s.kern_off : A.pool + (((long long)s.pool_hi << 32) | (unsigned)s.pool_lo);
    };

    // DMMA block table of op k: built per (parameter set, op) by build_optabs. ONE thread requests it with two bulk-async
    // copies (cp.async.bulk, the TMA engine: no registers, no per-thread address arithmetic, no instruction issue in the
    // other 255 threads) into ring slot k % TAB_RING while earlier ops compute; completion is signalled on the slot's
    // mbarrier (expect_tx / complete_tx), which every thread waits on (tab_acquire) right before it reads the table. Tables
    // are requested TAB_RING - 1 ops ahead; a slot is rewritten only after the end-of-op barrier of its previous user.
    const OpTab* __restrict__ gtabs = A.optabs + (size_t)kset * (A.optab_stride ? A.optab_stride : A.n_ops);
    // The ring comes in two flavours. TAB_BULK (window modes, or -DSQ_TAB_BULK=1): ONE thread requests a table with two
    // bulk-async copies (cp.async.bulk, the TMA engine; UBLKCP in SASS) and every thread waits on the slot's mbarrier right
    // before the op -- this leaves the per-thread cp.async groups to the TILE pipeline of the window modes, whose prefetch of the
    // next tile must stay in flight across several ops. Otherwise (cost / gradient over a resident matrix): per-thread cp.async
    // (LDGSTS) copies, one commit group per op, 2 % faster on C3 (profiles/r2_variants.jsonl) because no mbarrier try_wait sits
    // on every warp's path once per op.
    constexpr bool TAB_BULK = (SQ_TAB_BULK != 0) || ((SQ_WIN_BULK != 0) && WIN);
    if (TAB_BULK && tid == 0) {
        for (int i = 0; i < TAB_RING; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(tbar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (TAB_BULK) __syncthreads();  // nobody polls a barrier before it is initialised
    unsigned tab_parity = 0;  // bit s: phase parity the next wait on ring slot s expects (identical in every thread)
    auto tab_prefetch = [&](int k, bool bwd) {
        constexpr int FRAG = 8 * 32 * 8, TAIL = (int)(sizeof(OpTabS) - 2 * FRAG);  // bytes of one fragment set / of slots + widx + b0
        if (TAB_BULK) {
            if (tid != 0 || k < 0 || k >= A.n_ops || sops[k].kind == 1 || (CLU && sops[k].kind == 3)) return;
            const char* src = reinterpret_cast<const char*>(gtabs + k);
            const int slot = k & (TAB_RING - 1);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stab + slot);
            const unsigned bar = (unsigned)__cvta_generic_to_shared(tbar + slot);
            if (sops[k].kind == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(64) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                             "l"(kernel_of(k)), "r"(64), "r"(bar)
                             : "memory");
                return;
            }
            const int nfrag = bwd ? 2 * FRAG : FRAG, src_off = bwd ? FRAG : 0;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nfrag + TAIL) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src + src_off), "r"(nfrag), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 2 * FRAG),
                         "l"(src + 3 * FRAG), "r"(TAIL), "r"(bar)
                         : "memory");
            return;
        }
        // per-thread cp.async (LDGSTS) copies, one commit group per op in sweep order (empty for ops without a table): when an
        // op ends, "at most TAB_RING - 2 groups pending" means the NEXT op's table has landed; the end-of-op barrier publishes
        // it. With SQ_PRELOAD the operands of the next op are read BEFORE that barrier, so tables are waited for one op
        // earlier ("at most TAB_RING - 3 pending") and every table is published by the barrier one op before its first reader.
        if (!(k < 0 || k >= A.n_ops || sops[k].kind == 1 || (CLU && sops[k].kind == 3))) {
            const char* src = reinterpret_cast<const char*>(gtabs + k);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(stab + (k & (TAB_RING - 1)));
            if (sops[k].kind == 0) {
                if (tid < 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * tid), "l"(kernel_of(k) + tid));
            } else {
                const int nfrag = bwd ? 2 * FRAG : FRAG, src_off = bwd ? FRAG : 0;
                for (int e = tid; e < (nfrag + TAIL) / 16; e += nthr) {
                    const int b = e * 16;
                    const int d_off = (b < nfrag) ? b : 2 * FRAG + (b - nfrag);
                    const int s_off = (b < nfrag) ? src_off + b : 3 * FRAG + (b - nfrag);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + d_off), "l"(src + s_off));
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto tab_acquire = [&](int k) {  // bulk ring: the table of op k (if it has one) has landed in its ring slot
        if (k < 0 || k >= A.n_ops || sops[k].kind == 1 || (CLU && sops[k].kind == 3)) return;
        const int slot = k & (TAB_RING - 1);
        const unsigned bar = (unsigned)__cvta_generic_to_shared(tbar + slot);
        const unsigned parity = (tab_parity >> slot) & 1u;
        unsigned done;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        } while (!done);
        tab_parity ^= 1u << slot;
    };
    // bulk ring: the wait for the NEXT op's table sits in front of the end-of-op barrier, where its latency overlaps the wait
    // for the slowest warp (the table was requested TAB_RING - 1 ops ago and has long landed)
    auto tab_ready = [&](int next_k) {
        if (TAB_BULK) tab_acquire(next_k);
    };
    auto tab_end_of_op = [&]() {
        if (!TAB_BULK) asm volatile("cp.async.wait_group %0;" ::"n"(TAB_RING - 2 - (SQ_PRELOAD ? 1 : 0)) : "memory");
    };
    auto tab_prime = [&](int first, int step, bool bwd) {  // requests for the first TAB_RING - 1 ops of a sweep
#pragma unroll
        for (int j = 0; j < TAB_RING - 1; ++j) tab_prefetch(first + j * step, bwd);
        if (TAB_BULK) tab_acquire(first);
        else tab_end_of_op();
    };

    // Window mode, tile pipeline: every element of a tile travels as one 16 B asynchronous copy (LDGSTS) straight into its
    // swizzled shared-memory slot -- no staging registers, so the WHOLE tile is in flight at once (the state comes from HBM /
    // L2: a register-staged gather is latency-bound) -- as one cp.async group per tile (the op tables use the mbarrier ring in
    // these modes, so the groups belong to the tiles alone). Forward segments with two tile buffers (A.dbuf) request tile
    // t + 1 before they start on tile t; the last op of a segment writes its results straight to the state in HBM
    // (block_dmma_forward / block_dmma_backward with GST), so a tile has no separate store phase either.
    auto tile_request = [&](int t, cplx* bufa) {
        const size_t colbase = (size_t)y * A.in_ystride + deposit_bits((unsigned)(t * CT), ~A.wmask);
        const cplx* __restrict__ src = (MODE == MODE_BWD ? A.out : A.in) + colbase;
        const cplx* __restrict__ srcb = (MODE == MODE_BWD) ? A.beta + colbase : nullptr;
        for (int e = tid; e < rows * CT; e += nthr) {
            const size_t off = (size_t)(unsigned)sglob[e];
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(bufa + e)), "l"(src + off));
            if (MODE == MODE_BWD)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sb + e)), "l"(srcb + off));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto tile_of = [&](int ti) { return (WIN && A.wmask) ? ti * nchunks + chunk : tile_owner * A.tiles_per_cta + ti; };
    // logical row -> (owning rank, local row) under the layout the forward sweep ends in (cluster executor)
    auto locate = [&](int lrow_logical, int& local) -> bool {
        if constexpr (CLU) {
            unsigned rk = 0;
            int loc = 0;
            for (int qb = 0; qb < A.n; ++qb)
                if ((lrow_logical >> qb) & 1) {
                    const int ps = sfin[qb];
                    if (ps >= 32) rk |= 1u << (ps - 32);
                    else loc |= 1 << ps;
                }
            local = loc;
            return rk == crank;
        } else {
            local = lrow_logical;
            return true;
        }
    };
    // cluster executor, RESPLIT: exchange local row bit j with rank bit i. The elements of this CTA whose bit j differs from
    // its rank bit i trade places with the partner's elements whose bit j equals it (same pair index = row with bit j removed);
    // each CTA of the pair performs half of the swaps (pair parity), reading and writing the partner's tile through
    // distributed shared memory.
    auto resplit = [&](int j, int i, bool with_b) {
        if constexpr (CLU) {
            auto cluster = cooperative_groups::this_cluster();
            cluster.sync();  // every CTA of the cluster has finished the previous op
            const unsigned partner = crank ^ (1u << i);
            const int my_bit = (crank >> i) & 1;
            cplx* pa = cluster.map_shared_rank(sa, partner);
            cplx* pb = cluster.map_shared_rank(sb, partner);
            const int npairs = (rows >> 1) << LOG_CT;
            for (int t = tid; 2 * t + my_bit < npairs; t += nthr) {
                const int pid = 2 * t + my_bit;
                const int cc = pid & (CT - 1);
                const int x0 = insert_zero(pid >> LOG_CT, j);
                const int em = elem<LOG_CT>(x0 | ((my_bit ^ 1) << j), cc), ep = elem<LOG_CT>(x0 | (my_bit << j), cc);
                const cplx vm = sa[em], vp = pa[ep];
                sa[em] = vp;
                pa[ep] = vm;
                if (HAS_B && with_b) {
                    const cplx bm = sb[em], bp = pb[ep];
                    sb[em] = bp;
                    pb[ep] = bm;
                }
            }
            cluster.sync();  // the exchange is complete before anyone touches the tiles again
        }
    };
    if (WIN && dbuf && A.tiles_per_cta > 0 && tile_of(0) < A.tiles) tile_request(tile_of(0), sa0);

    for (int ti = 0; ti < A.tiles_per_cta; ++ti) {
        // window mode deals the tiles round robin: the CTAs of a wave then work on neighbouring tiles at the same time, and the
        // 32 B / 64 B pieces of one DRAM burst that belong to different tiles meet in L2 instead of being fetched twice
        const int tile = tile_of(ti);
        if (tile >= A.tiles) break;
        const int j0 = tile * CT;
        const int valid = min(CT, A.cols - j0);

        // ---- load the tile ------------------------------------------------------------------------------------
        {
            const int first_op = (MODE == MODE_BWD) ? A.n_ops - 1 : 0;
            if (WIN && A.wmask) {
                if (dbuf) {
                    sa = sa0 + (size_t)(ti & 1) * rows * CT;
                    const bool more = ti + 1 < A.tiles_per_cta && tile_of(ti + 1) < A.tiles;
                    if (more) tile_request(tile_of(ti + 1), sa0 + (size_t)((ti + 1) & 1) * rows * CT);
                    if (A.n_ops > 0) tab_prime(first_op, 1, false);
                    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                } else {
                    tile_request(tile, sa);
                    if (A.n_ops > 0) tab_prime(first_op, MODE == MODE_BWD ? -1 : 1, MODE == MODE_BWD);
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
            } else {
                // cluster executor: the sweep starts with the top rho qubits as split qubits: CTA `crank` loads rows [crank * rows, ...)
                const cplx* __restrict__ src = A.in + (size_t)y * A.in_ystride + j0 + (CLU ? (size_t)crank * rows * A.ld_in : 0);
                for (int e = tid; e < rows * CT; e += nthr) {
                    const int i = e >> LOG_CT, c = e & (CT - 1);
                    cplx v = czero();
                    if (c < valid) v = src[(size_t)i * A.ld_in + c];
                    sa[elem<LOG_CT>(i, c)] = v;
                }
                if (A.n_ops > 0) tab_prime(first_op, MODE == MODE_BWD ? -1 : 1, MODE == MODE_BWD);
            }
        }
        __syncthreads();
        // window mode: the state in HBM that the last op of the segment writes to (fused store)
        cplx* __restrict__ gst_a = nullptr;
        cplx* __restrict__ gst_b = nullptr;
        if (WIN && A.wmask) {
            const size_t colbase = (size_t)y * A.out_ystride + deposit_bits((unsigned)j0, ~A.wmask);
            gst_a = A.out + colbase;
            if (MODE == MODE_BWD) gst_b = A.beta + colbase;
        }
        bool stored = false;

        // ---- forward sweep: op 0 first (Gates_block.cpp:683) ---------------------------------------------------
        BlkPre R;  // operands of the next DMMA block op, loaded ahead of the barrier (SQ_PRELOAD)
        R.ok = false;
        for (int k = 0; k < (MODE == MODE_BWD ? 0 : A.n_ops); ++k) {
            const SOp s = sops[k];
            const bool have_next = k + 1 < A.n_ops;
            tab_prefetch(k + TAB_RING - 1, false);
            const cplx* __restrict__ km = reinterpret_cast<const cplx*>(stab + (k & (TAB_RING - 1)));
            SQ_TRACE_MARK(0)
            if (CLU && s.kind == 3) {
                resplit(s.q0, s.q1, false);
            } else if (s.kind == 2) {
                bool fused = false;
                if constexpr (MODE == MODE_APPLY) {
                    if (SQ_FUSE_STORE && gst_a && k == A.n_ops - 1) {  // last op of a window segment: results go straight to the state in HBM
                        if (s.dim == 8) block_dmma_forward<LOG_CT, 3, true>(sa, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, s.q2, rows, tid, nthr, gst_a, sglob);
                        else block_dmma_forward<LOG_CT, 2, true>(sa, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, 30, rows, tid, nthr, gst_a, sglob);
                        fused = stored = true;
                    }
                }
                if (!fused) {
                    if (s.dim == 8) block_dmma_forward<LOG_CT, 3, false>(sa, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, s.q2, rows, tid, nthr, nullptr, nullptr);
                    else block_dmma_forward<LOG_CT, 2, false>(sa, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, 30, rows, tid, nthr, nullptr, nullptr);
                }
            } else if (s.kind == 0) {
                const cplx k00 = km[0], k01 = km[1], k10 = km[2], k11 = km[3];
                const int tbit = 1 << s.q0;
                const int nitems = (rows >> 1) << LOG_CT;
                for (int item = tid; item < nitems; item += nthr) {
                    const int c = item & (CT - 1);
                    const int i0 = insert_zero(item >> LOG_CT, s.q0);
                    const int e0 = elem<LOG_CT>(i0, c), e1 = elem<LOG_CT>(i0 | tbit, c);
                    const cplx a0 = sa[e0], a1 = sa[e1];
                    sa[e0] = cfma(k01, a1, cmul(k00, a0));
                    sa[e1] = cfma(k11, a1, cmul(k10, a0));
                }
            } else {
                // ---- generic path: controlled gates, raw dense kernels, derivative kernels ------------------------
                const DevOp& op = A.ops[k];
                const bool deriv = (MODE == MODE_APPLY) && (k == deriv_op);
                const int dim = op.dim;
                const cplx* __restrict__ K = deriv ? dktab + op.dkern_off + deriv_slot * dim * dim
                                                    : (op.kern_off >= 0 ? ktab + op.kern_off : A.pool + op.pool_off);
                if (dim == 2) {
                    const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
                    const int tbit = 1 << op.target;
                    const unsigned cm = op.ctrl_mask;
                    if (!deriv) {
                        const int f0 = op.fix[0], f1 = op.fix[1], f2 = op.fix[2];
                        const int nitems = (rows >> op.nfix) << LOG_CT;
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            const int i0 = insert_zero(insert_zero(insert_zero(item >> LOG_CT, f0), f1), f2) | cm;
                            const int e0 = elem<LOG_CT>(i0, c), e1 = elem<LOG_CT>(i0 | tbit, c);
                            const cplx a0 = sa[e0], a1 = sa[e1];
                            sa[e0] = cfma(k01, a1, cmul(k00, a0));
                            sa[e1] = cfma(k11, a1, cmul(k10, a0));
                        }
                    } else {  // derivative kernel: inactive pairs are zero-filled (apply_kernel_to_input.cpp:93-97)
                        const int nitems = (rows >> 1) << LOG_CT;
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            const int i0 = insert_zero(item >> LOG_CT, op.target);
                            const int e0 = elem<LOG_CT>(i0, c), e1 = elem<LOG_CT>(i0 | tbit, c);
                            if ((i0 & cm) == cm) {
                                const cplx a0 = sa[e0], a1 = sa[e1];
                                sa[e0] = cfma(k01, a1, cmul(k00, a0));
                                sa[e1] = cfma(k11, a1, cmul(k10, a0));
                            } else {
                                sa[e0] = czero();
                                sa[e1] = czero();
                            }
                        }
                    }
                } else if (DNS) {
                    // dense dim x dim kernel on ascending qubits (apply_large_kernel_to_input.cpp:160-199)
                    const int nq = op.nq;
                    const int nitems = (rows >> nq) << LOG_CT;
                    const bool use_dmma = !deriv && op.ctrl_mask == 0 && nq >= 3 && (nitems & 7) == 0;
                    if (use_dmma && nq <= 4) {
                        const DenseTab* T = (A.dense_tabs && op.dtab > 0) ? A.dense_tabs + 3 * (op.dtab - 1) : nullptr;
                        if (!T) {  // no precomputed table (parametric dense op): build it in the staging area
                            DenseTab* Ts = reinterpret_cast<DenseTab*>(sk);
                            int* schoice = reinterpret_cast<int*>(Ts + 1);
                            if (nq == 3) dense_tab_fill<LOG_CT, 3>(Ts, schoice, K, op, tid, nthr);
                            else dense_tab_fill<LOG_CT, 4>(Ts, schoice, K, op, tid, nthr);
                            __syncthreads();
                            T = Ts;
                        }
                        if (nq == 3) dense_dmma_forward2<LOG_CT, 3>(sa, T, op, rows, tid, nthr);
                        else dense_dmma_forward2<LOG_CT, 4>(sa, T, op, rows, tid, nthr);
                    } else if (use_dmma && A.dense_tabs5 && op.dtab < 0) {
                        dense_dmma_forward5<LOG_CT>(sa, A.dense_tabs5 + (-1 - op.dtab), op, rows, tid, nthr);
                    } else if (use_dmma) {
                        // 5 qubits: stage the real embedding of K (padded rows) and the local-index -> row-bit pattern
                        const int dimr = 2 * dim, ld = dimr + DMMA_PAD;
                        double* skr = reinterpret_cast<double*>(sk);
                        int* spat = reinterpret_cast<int*>(skr + dimr * ld);
                        for (int e = tid; e < dimr * dimr; e += nthr) {
                            const int r = e / dimr, cidx = e - r * dimr;
                            const cplx kv = K[(r >> 1) * dim + (cidx >> 1)];
                            const int a = r & 1, b = cidx & 1;
                            skr[r * ld + cidx] = (a == b) ? kv.x : (a ? kv.y : -kv.y);
                        }
                        for (int l = tid; l < dim; l += nthr) {
                            int r = 0;
                            for (int j = 0; j < nq; ++j) r |= ((l >> j) & 1) << op.q[j];
                            spat[l] = r;
                        }
                        __syncthreads();
                        dense_dmma_forward<LOG_CT, 5>(sa, skr, spat, op, rows, tid, nthr);
                    } else {
                        for (int e = tid; e < dim * dim; e += nthr) sk[e] = K[e];
                        __syncthreads();
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            int base = item >> LOG_CT;
                            for (int j = 0; j < nq; ++j) base = insert_zero(base, op.q[j]);
                            const bool active = (base & op.ctrl_mask) == op.ctrl_mask;
                            if (!active && !deriv) continue;
                            cplx v[32];
                            for (int l = 0; l < dim; ++l) {
                                int r = base;
                                for (int j = 0; j < nq; ++j) r |= ((l >> j) & 1) << op.q[j];
                                v[l] = active ? sa[elem<LOG_CT>(r, c)] : czero();
                            }
                            for (int ro = 0; ro < dim; ++ro) {
                                cplx acc = czero();
                                if (active)
                                    for (int l = 0; l < dim; ++l) acc = cfma(sk[ro * dim + l], v[l], acc);
                                int r = base;
                                for (int j = 0; j < nq; ++j) r |= ((ro >> j) & 1) << op.q[j];
                                sa[elem<LOG_CT>(r, c)] = acc;
                            }
                        }
                    }
                }
            }
            SQ_TRACE_MARK(1)
            SQ_TRACE_NEXT()
            if (SQ_PRELOAD && s.kind != 2) blk_kill(R);
            tab_ready(k + 1);
            if (SQ_PRELOAD && have_next) {
                const SOp sn = sops[k + 1];
                if (sn.kind == 2 && sn.dim == 8) blk_preload<3, false>(R, stab + ((k + 1) & (TAB_RING - 1)), (rows >> 3) << LOG_CT, lane, warp, nwarps);
            }
            tab_end_of_op();
            __syncthreads();
        }

        if (MODE == MODE_APPLY) {
            if (WIN && A.wmask) {
                if (stored) continue;  // the segment's last op wrote the tile (its end-of-op barrier frees the buffer)
                for (int e = tid; e < rows * CT; e += nthr) gst_a[(size_t)(unsigned)sglob[e]] = sa[e];
            } else {
                cplx* __restrict__ dst = A.out + (size_t)y * A.out_ystride + j0;
                for (int e = tid; e < rows * CT; e += nthr) {
                    const int i = e >> LOG_CT, c = e & (CT - 1);
                    if (c < valid) dst[(size_t)i * A.ld_out + c] = sa[elem<LOG_CT>(i, c)];
                }
            }
            __syncthreads();
            continue;
        }

        // ---- trace terms: sum_j M[(j+off) ^ mask, j] (N_Qubit_Decomposition_Cost_Function.cpp:137-160,191-404) ----
        if (MODE != MODE_BWD) {
            double t[6] = {0, 0, 0, 0, 0, 0};
            const int off = A.trace_offset;
            if (A.sum_sq) {
                // get_cost_function_sum_of_squares (N_Qubit_Decomposition_Cost_Function.cpp:443-457): every element of the tile
                for (int e = tid; e < rows * CT; e += nthr) {
                    const int i = e >> LOG_CT, c = e & (CT - 1);
                    if (c < valid) {
                        const cplx v = sa[elem<LOG_CT>(i, c)];
                        const double dr = v.x - ((i == j0 + c + off) ? 1.0 : 0.0);
                        t[0] += dr * dr + v.y * v.y;
                    }
                }
            } else if (tid < valid) {
                int lr;
                if (locate(j0 + tid + off, lr)) {
                    const cplx v = sa[elem<LOG_CT>(lr, tid)];
                    t[0] = v.x;
                    t[1] = v.y;
                }
            }
            if (A.n_trace_types > 1) {
                for (int e = tid; e < A.n * CT; e += nthr) {
                    const int c = e & (CT - 1), qb = e >> LOG_CT;
                    int lr;
                    if (c < valid && locate((j0 + c + off) ^ (1 << qb), lr)) {
                        const cplx v = sa[elem<LOG_CT>(lr, c)];
                        t[2] += v.x;
                        t[3] += v.y;
                    }
                }
            }
            if (A.n_trace_types > 2) {
                int e = 0;
                for (int q1 = 0; q1 < A.n - 1; ++q1)
                    for (int q2 = q1 + 1; q2 < A.n; ++q2)
                        for (int c = 0; c < valid; ++c, ++e)
                            if (e % nthr == tid) {
                                int lr;
                                if (locate((j0 + c + off) ^ ((1 << q1) | (1 << q2)), lr)) {
                                    const cplx v = sa[elem<LOG_CT>(lr, c)];
                                    t[4] += v.x;
                                    t[5] += v.y;
                                }
                            }
            }
            const int nt = 2 * A.n_trace_types;
            for (int i = 0; i < nt; ++i)
                for (int sft = 16; sft > 0; sft >>= 1) t[i] += __shfl_xor_sync(0xffffffffu, t[i], sft);
            if (lane == 0)
                for (int i = 0; i < nt; ++i) sred[warp * 6 + i] = t[i];
            __syncthreads();
            if (tid == 0) {
                for (int i = 0; i < nt; ++i) {
                    double sum = 0;
                    for (int w = 0; w < nwarps; ++w) sum += sred[w * 6 + i];
                    stsum[i] += sum;
                }
            }
        }

        if (HAS_B) {
          if (MODE == MODE_GRAD) {
            // ---- beta_N = sum_t omega_t * sum_{masks of type t} e_{(j+off)^mask} per column ---------------------
            for (int e = tid; e < rows * CT; e += nthr) sb[e] = czero();
            if (A.n_ops > 0) tab_prime(A.n_ops - 1, -1, true);
            __syncthreads();
            if (A.sum_sq) {
                // the functional of the gradient is Re sum conj(Upartial_ij) dM_ij with Upartial = 2 (M - I)
                // (get_deriv_sum_of_squares :459-475, real_trace_conj_dot): beta_N = conj(Upartial) per column
                const int off = A.trace_offset;
                for (int e = tid; e < rows * CT; e += nthr) {
                    const int i = e >> LOG_CT, c = e & (CT - 1);
                    if (c < valid) {
                        const cplx v = sa[elem<LOG_CT>(i, c)];
                        sb[elem<LOG_CT>(i, c)] = cmake(2.0 * (v.x - ((i == j0 + c + off) ? 1.0 : 0.0)), -2.0 * v.y);
                    }
                }
            } else {
                const int off = A.trace_offset;
                const cplx w0 = A.omega[(size_t)y * 3 + 0];
                {
                    int lr;
                    if (tid < valid && locate(j0 + tid + off, lr)) sb[elem<LOG_CT>(lr, tid)] = w0;
                }
                if (A.n_trace_types > 1) {
                    const cplx w1 = A.omega[(size_t)y * 3 + 1];
                    for (int e = tid; e < A.n * CT; e += nthr) {
                        const int c = e & (CT - 1), qb = e >> LOG_CT;
                        int lr;
                        if (c < valid && locate((j0 + c + off) ^ (1 << qb), lr)) sb[elem<LOG_CT>(lr, c)] = w1;
                    }
                }
                if (A.n_trace_types > 2) {
                    const cplx w2 = A.omega[(size_t)y * 3 + 2];
                    int e = 0;
                    for (int q1 = 0; q1 < A.n - 1; ++q1)
                        for (int q2 = q1 + 1; q2 < A.n; ++q2)
                            for (int c = 0; c < valid; ++c, ++e)
                                if (e % nthr == tid) {
                                    int lr;
                                    if (locate((j0 + c + off) ^ ((1 << q1) | (1 << q2)), lr)) sb[elem<LOG_CT>(lr, c)] = w2;
                                }
                }
            }
            __syncthreads();
          }

            // ---- backward sweep ---------------------------------------------------------------------------------
            int buf = 0;
            // fold of one op's W' partials: element e of the op's dim x dim complex block (as doubles) summed over the warps'
            // slots in a fixed order (a pairwise tree: three dependent DADDs instead of seven) and ONE fire-and-forget reduction
            // into the CTA's slice: each address has a single writer, so the sums are bit-reproducible
            auto fold_elem = [&](int e, int fbuf, int w_off) {
                const double* base = reinterpret_cast<const double*>(swarp + (size_t)fbuf * nwarps * A.wmax) + e;
                const size_t wstride = (size_t)A.wmax * 2;
                double sum;
                if (nwarps == 8) {
                    const double a0 = base[0], a1 = base[wstride], a2 = base[2 * wstride], a3 = base[3 * wstride];
                    const double a4 = base[4 * wstride], a5 = base[5 * wstride], a6 = base[6 * wstride], a7 = base[7 * wstride];
                    sum = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
                } else {
                    sum = 0;
                    for (int w = 0; w < nwarps; ++w) sum += base[w * wstride];
                }
                if (A.w_in_smem) reinterpret_cast<double*>(swacc + w_off)[e] += sum;
                else atomicAdd(reinterpret_cast<double*>(A.w_part + ((size_t)y * nchunks + chunk) * A.w_total + w_off) + e, sum);
            };
            int pend_nd = 0, pend_buf = 0, pend_off = 0, step = 0;  // SQ_FOLD_EARLY: the previous parametric op's fold, still to do
            for (int k = A.n_ops - 1; k >= 0; --k) {
                const SOp s = sops[k];
                const bool have_next = k > 0;
                tab_prefetch(k - (TAB_RING - 1), true);
                const bool has_w = s.w_off >= 0;
                cplx* wslot_c = wdirect ? A.w_part + (((size_t)y * nchunks + chunk) * nwarps + warp) * A.w_total + max(s.w_off, 0)
                                        : swarp + (size_t)(buf * nwarps + warp) * A.wmax;
                double* wslot = reinterpret_cast<double*>(wslot_c);
                const cplx* __restrict__ km = reinterpret_cast<const cplx*>(stab + (k & (TAB_RING - 1)));
                int wdim = s.dim;
                SQ_TRACE_MARK(0)
                if (CLU && s.kind == 3) {
                    resplit(s.q0, s.q1, true);
                } else if (s.kind == 2) {
                    bool fused = false;
                    if constexpr (MODE == MODE_BWD) {
                        if (SQ_FUSE_STORE && gst_a && k == 0) {  // last op of the segment's backward sweep: a and beta go straight to HBM
                            if (s.dim == 8) block_dmma_backward<LOG_CT, 3, true>(sa, sb, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, s.q2, rows, has_w, wslot_c, wdirect, tid, nthr, gst_a, gst_b, sglob);
                            else block_dmma_backward<LOG_CT, 2, true>(sa, sb, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, 30, rows, has_w, wslot_c, wdirect, tid, nthr, gst_a, gst_b, sglob);
                            fused = stored = true;
                        }
                    }
                    if (!fused) {
                        if (s.dim == 8) block_dmma_backward<LOG_CT, 3, false>(sa, sb, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, s.q2, rows, has_w, wslot_c, wdirect, tid, nthr, nullptr, nullptr, nullptr);
                        else block_dmma_backward<LOG_CT, 2, false>(sa, sb, stab + (k & (TAB_RING - 1)), R, s.q0, s.q1, 30, rows, has_w, wslot_c, wdirect, tid, nthr, nullptr, nullptr, nullptr);
                    }
                } else if (s.kind == 0) {
                    const cplx k00 = km[0], k01 = km[1], k10 = km[2], k11 = km[3];
                    const int tbit = 1 << s.q0;
                    cplx W[4] = {czero(), czero(), czero(), czero()};
                    const int nitems = (rows >> 1) << LOG_CT;
                    for (int item = tid; item < nitems; item += nthr) {
                        const int c = item & (CT - 1);
                        const int i0 = insert_zero(item >> LOG_CT, s.q0);
                        const int e0 = elem<LOG_CT>(i0, c), e1 = elem<LOG_CT>(i0 | tbit, c);
                        const cplx p0 = sa[e0], p1 = sa[e1], b0 = sb[e0], b1 = sb[e1];
                        sa[e0] = cfmac(k10, p1, cfmac(k00, p0, czero()));
                        sa[e1] = cfmac(k11, p1, cfmac(k01, p0, czero()));
                        if (has_w) {
                            W[0] = cfma(b0, p0, W[0]);
                            W[1] = cfma(b0, p1, W[1]);
                            W[2] = cfma(b1, p0, W[2]);
                            W[3] = cfma(b1, p1, W[3]);
                        }
                        sb[e0] = cfma(k10, b1, cmul(k00, b0));
                        sb[e1] = cfma(k11, b1, cmul(k01, b0));
                    }
                    if (has_w) warp_store_w<4>(W, wslot, lane, wdirect);
                } else {
                    const DevOp& op = A.ops[k];
                    const cplx* __restrict__ K = op.kern_off >= 0 ? ktab + op.kern_off : A.pool + op.pool_off;
                    wdim = op.dim;
                    if (op.dim == 2) {
                        const cplx k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3];
                        const int tbit = 1 << op.target;
                        const unsigned cm = op.ctrl_mask;
                        const int f0 = op.fix[0], f1 = op.fix[1], f2 = op.fix[2];
                        cplx W[4] = {czero(), czero(), czero(), czero()};
                        const int nitems = (rows >> op.nfix) << LOG_CT;
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            const int i0 = insert_zero(insert_zero(insert_zero(item >> LOG_CT, f0), f1), f2) | cm;
                            const int e0 = elem<LOG_CT>(i0, c), e1 = elem<LOG_CT>(i0 | tbit, c);
                            const cplx p0 = sa[e0], p1 = sa[e1], b0 = sb[e0], b1 = sb[e1];
                            sa[e0] = cfmac(k10, p1, cfmac(k00, p0, czero()));
                            sa[e1] = cfmac(k11, p1, cfmac(k01, p0, czero()));
                            if (has_w) {
                                W[0] = cfma(b0, p0, W[0]);
                                W[1] = cfma(b0, p1, W[1]);
                                W[2] = cfma(b1, p0, W[2]);
                                W[3] = cfma(b1, p1, W[3]);
                            }
                            sb[e0] = cfma(k10, b1, cmul(k00, b0));
                            sb[e1] = cfma(k11, b1, cmul(k01, b0));
                        }
                        if (has_w) warp_store_w<4>(W, wslot, lane, wdirect);
                    } else if (DNS && A.dense_tabs && op.dtab > 0 && !has_w && op.ctrl_mask == 0 && ((((rows >> op.nq) << LOG_CT) & 7) == 0)) {
                        // constant dense 3-/4-qubit kernel (GENERAL blocks, multiplied-out constant sub-circuits): the adjoint
                        // step is two forward-style products on the tensor cores with the op's K^dagger and K^T tables
                        const DenseTab* T = A.dense_tabs + 3 * (op.dtab - 1);
                        if (op.nq == 3) {
                            dense_dmma_forward2<LOG_CT, 3>(sa, T + 1, op, rows, tid, nthr);
                            dense_dmma_forward2<LOG_CT, 3>(sb, T + 2, op, rows, tid, nthr);
                        } else {
                            dense_dmma_forward2<LOG_CT, 4>(sa, T + 1, op, rows, tid, nthr);
                            dense_dmma_forward2<LOG_CT, 4>(sb, T + 2, op, rows, tid, nthr);
                        }
                    } else if (DNS) {
                        // raw dense op (controlled two-target gates, GENERAL blocks, or a block too small for the tensor
                        // path): thread per (group, column), local arrays
                        const int dim = op.dim, nq = op.nq;
                        for (int e = tid; e < dim * dim; e += nthr) sk[e] = K[e];
                        __syncthreads();
                        const int nitems = (rows >> nq) << LOG_CT;
                        cplx wl[64];  // parametric dense ops have dim <= 8
                        if (has_w)
                            for (int e = 0; e < dim * dim; ++e) wl[e] = czero();
                        for (int item = tid; item < nitems; item += nthr) {
                            const int c = item & (CT - 1);
                            int base = item >> LOG_CT;
                            for (int j = 0; j < nq; ++j) base = insert_zero(base, op.q[j]);
                            if ((base & op.ctrl_mask) != op.ctrl_mask) continue;
                            cplx pv[32], bv[32];
                            int addr[32];
                            for (int l = 0; l < dim; ++l) {
                                int r = base;
                                for (int j = 0; j < nq; ++j) r |= ((l >> j) & 1) << op.q[j];
                                addr[l] = elem<LOG_CT>(r, c);
                                pv[l] = sa[addr[l]];
                                bv[l] = sb[addr[l]];
                            }
                            for (int ro = 0; ro < dim; ++ro) {
                                cplx acc = czero(), bacc = czero();
                                for (int l = 0; l < dim; ++l) {
                                    acc = cfmac(sk[l * dim + ro], pv[l], acc);   // (K^dagger p)[ro]
                                    bacc = cfma(sk[l * dim + ro], bv[l], bacc);  // (K^T beta)[ro]
                                }
                                sa[addr[ro]] = acc;
                                sb[addr[ro]] = bacc;
                                if (has_w)
                                    for (int r2 = 0; r2 < dim; ++r2) wl[r2 * dim + ro] = cfma(bv[r2], pv[ro], wl[r2 * dim + ro]);
                            }
                        }
                        if (has_w) {
                            for (int part = 0; part < dim * dim / 4; ++part) {
                                double v[8];
                                for (int e = 0; e < 4; ++e) {
                                    v[2 * e] = wl[part * 4 + e].x;
                                    v[2 * e + 1] = wl[part * 4 + e].y;
                                }
                                warp_reduce8(v, lane);
                                if ((lane & 3) == 0) {
                                    double* dst = wslot + part * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                                    if (wdirect) atomicAdd(dst, v[0]);
                                    else *dst = v[0];
                                }
                            }
                        }
                    }
                }
                SQ_TRACE_MARK(1)
                SQ_TRACE_NEXT()
                if (SQ_PRELOAD && s.kind != 2) blk_kill(R);
                tab_ready(k - 1);
                if (SQ_PRELOAD && have_next) {
                    const SOp sn = sops[k - 1];
                    if (sn.kind == 2 && sn.dim == 8) blk_preload<3, true>(R, stab + ((k - 1) & (TAB_RING - 1)), (rows >> 3) << LOG_CT, lane, warp, nwarps);
                }
                if (SQ_FOLD_EARLY && pend_nd > 0) {
                    // the PREVIOUS parametric op's fold, by the warps that get here first: ticket t folds elements [32 t, 32 t + 32)
                    int ticket = 0;
                    if (lane == 0) ticket = atomicAdd(&s_fold_cnt[step & 1], 1);
                    ticket = __shfl_sync(0xffffffffu, ticket, 0);
                    for (int e0 = ticket * 32; e0 < pend_nd; e0 += nwarps * 32)
                        if (e0 + lane < pend_nd) fold_elem(e0 + lane, pend_buf, pend_off);
                    pend_nd = 0;
                }
                tab_end_of_op();
                __syncthreads();
                if (SQ_FOLD_EARLY && tid == 0) s_fold_cnt[step & 1] = 0;  // next used two barriers from now
                ++step;
                if (has_w && !wdirect) {
                    const int nd = 2 * wdim * wdim;  // doubles
                    if (SQ_FOLD_EARLY) {
                        pend_nd = nd;
                        pend_buf = buf;
                        pend_off = s.w_off;
                    } else {
                        // only the first nd / 32 warps fold; the others are already in the next op (spreading the fold over all
                        // warps was measured slower: 1 282 against 1 339 evals/s on C3, profiles/r2_variants_exec.jsonl)
                        for (int e = tid; e < nd; e += nthr) fold_elem(e, buf, s.w_off);
                    }
                    buf ^= 1;
                }
            }
            if (SQ_FOLD_EARLY && pend_nd > 0)  // the last parametric op of the sweep (its slots were published by the barrier above)
                for (int e = tid; e < pend_nd; e += nthr) fold_elem(e, pend_buf, pend_off);
            __syncthreads();
            if (MODE == MODE_BWD && !stored) {
                for (int e = tid; e < rows * CT; e += nthr) {
                    const size_t off = (size_t)(unsigned)sglob[e];
                    gst_a[off] = sa[e];
                    gst_b[off] = sb[e];
                }
                __syncthreads();
            }
        }
    }

    if constexpr (CLU) cooperative_groups::this_cluster().sync();  // no CTA leaves while a partner may still touch its tile
    if (MODE == MODE_COST || MODE == MODE_GRAD) {
        if (tid == 0) {
            double* dst = A.tr_part + ((size_t)y * nchunks + chunk) * 6;
            for (int i = 0; i < 6; ++i) dst[i] = stsum[i];
        }
        if (MODE == MODE_GRAD && A.w_in_smem) {
            __syncthreads();
            cplx* dst = A.w_part + ((size_t)y * nchunks + chunk) * A.w_total;
            for (int e = tid; e < A.w_total; e += nthr) dst[e] = swacc[e];
        }
    }
}

}  // namespace sq
