"""Multi-GPU sharding of the cost path: one process per GPU, torch.distributed for the plumbing.

Two independent axes, both already present in the reference:

  * ``mode="batch"``   -- parameter vectors are independent (Optimization_Interface.cpp:1009-1025); every rank evaluates a
                          contiguous slice of the batch on the full matrix and ONE all-gather returns all costs/gradients
                          (the reference's MPI_Allgather, Optimization_Interface.cpp:962-1004).
  * ``mode="columns"`` -- left multiplication never mixes columns (kernels/apply_kernel_to_input.cpp:69-89); rank r holds
                          U[:, r*w:(r+1)*w] and evaluates with trace_offset = r*w -- exactly the reference's rectangular
                          Umtx + trace_offset semantics (N_Qubit_Decomposition_Cost_Function.cpp:147-153). ONE all-reduce
                          (sum) of the raw trace terms [B x (1+P) x 3 x {Re,Im}] happens BEFORE the non-linear cost
                          formulas (the Hilbert-Schmidt variants need the full complex trace,
                          Optimization_Interface.cpp:1414-1419), like the DFE path's gather of trace triples (:806-832).

The engine is injected (``engine_factory``) so the host logic can be exercised on CPU with the gloo backend and a
test double; the product factory is ``Engine`` (CUDA, NCCL).
"""
import numpy as np

from . import abi


def column_shard(cols, rank, world):
    """[begin, end) of rank's column block; blocks differ by at most one column"""
    base, rem = divmod(cols, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def batch_shard(batch, rank, world):
    base, rem = divmod(batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def trace_pass_variant(variant):
    """cost variant the shard engines run for the TRACE pass: the Frobenius family keeps its own variant (it fixes how
    many trace types are needed and the real weights of the functional); the trace-modulus variants (3, 9) need the
    shard's row offset, which only the Frobenius family honours, with the same single trace type."""
    if variant in (abi.FROBENIUS_NORM, abi.FROBENIUS_NORM_CORRECTION1, abi.FROBENIUS_NORM_CORRECTION2):
        return variant
    if variant in (abi.HILBERT_SCHMIDT_TEST, abi.INFIDELITY):
        return abi.FROBENIUS_NORM
    raise Exception("cost variant %d is not supported with column sharding" % variant)


class ShardedCost:
    """Cost / cost+gradient over ``world`` ranks. All ranks call the same methods with the same arguments and get the same
    results back (lock-step, as the reference's MPI build does with MPI_Bcast / MPI_Allgather)."""

    def __init__(self, Umtx, circuit, variant=abi.FROBENIUS_NORM, mode="batch", prev_cost=1.0, c1=1 / 1.7, c2=0.5,
                 engine_factory=None, device=None, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if mode not in ("batch", "columns"):
            raise Exception("mode should be 'batch' or 'columns'")
        self.mode = mode
        self.variant = int(variant)
        self.cfg = (float(prev_cost), float(c1), float(c2))
        U = np.ascontiguousarray(Umtx, dtype=np.complex128)
        self.rows, self.cols = U.shape
        self.n_params = circuit.get_Parameter_Num()
        if engine_factory is None:
            from .engine import Engine

            engine_factory = Engine
        self.engine = engine_factory(self.rank if device is None else device)
        if mode == "columns":
            if self.cols < self.world:
                raise Exception("fewer columns than ranks")
            b, e = column_shard(self.cols, self.rank, self.world)
            self.col_begin = b
            self.engine.upload_matrix(np.ascontiguousarray(U[:, b:e]))
            self.engine.set_circuit(circuit)
            self._trace_variant = trace_pass_variant(self.variant)
        else:
            self.engine.upload_matrix(U)
            self.engine.set_circuit(circuit)
            self.engine.set_cost(self.variant, 0, *self.cfg)

    # ---- collectives on host tensors (gloo) or device tensors (nccl) ------------------------------------------------
    def _tensor(self, a):
        import torch

        t = torch.from_numpy(np.ascontiguousarray(a))
        if self.dist.is_initialized() and self.dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        return t

    def _all_reduce_sum(self, a):
        if self.world == 1:
            return a
        t = self._tensor(a)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def _all_gather_rows(self, a, counts):
        """gather row blocks of unequal height (counts[r] rows on rank r) into one array on every rank"""
        if self.world == 1:
            return a
        import torch

        width = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
        mx = max(counts)
        pad = np.zeros((mx, width), dtype=a.dtype)
        pad[: a.shape[0]] = a.reshape(a.shape[0], width)
        t = self._tensor(pad)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        parts = [o.cpu().numpy()[: counts[r]] for r, o in enumerate(out)]
        return np.concatenate(parts, axis=0).reshape((sum(counts),) + a.shape[1:])

    # ---- the sharded hot path ------------------------------------------------------------------------------------------
    def cost_grad(self, params, with_grad=True):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p.reshape(1, -1)
        B = p.shape[0]
        if self.mode == "batch":
            counts = [batch_shard(B, r, self.world)[1] - batch_shard(B, r, self.world)[0] for r in range(self.world)]
            b, e = batch_shard(B, self.rank, self.world)
            if with_grad:
                if e > b:
                    c, g = self.engine.cost_grad_batched(p[b:e])
                else:
                    c, g = np.zeros(0), np.zeros((0, self.n_params))
                packed = np.concatenate([c.reshape(-1, 1), g], axis=1)
                allp = self._all_gather_rows(packed, counts)
                return allp[:, 0].copy(), allp[:, 1:].copy()
            c = self.engine.cost_batched(p[b:e]) if e > b else np.zeros(0)
            return self._all_gather_rows(c.reshape(-1, 1), counts)[:, 0].copy()
        # columns: shard traces -> one all-reduce -> cost formulas on the summed traces
        self.engine.set_cost(self._trace_variant, self.col_begin, *self.cfg)
        tr = self.engine.traces_batched(p, with_grad)
        tr = self._all_reduce_sum(tr)
        self.engine.set_cost(self.variant, 0, *self.cfg)
        return self.engine.cost_from_traces(tr, with_grad, self.cols)

    def cost(self, params):
        return self.cost_grad(params, with_grad=False)

    def close(self):
        if hasattr(self.engine, "close"):
            self.engine.close()


class ShardedVQE:
    """VQE energy / energy+gradient over ``world`` ranks. A 2^n state vector is never split: the parameter
    sets are sharded (rank r evaluates a contiguous slice on its own copy of the initial state and Hamiltonian) and one
    all-gather returns all energies and gradients, as for ``ShardedCost(mode="batch")``."""

    def __init__(self, state0, circuit, indptr, indices, data, engine_factory=None, device=None, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_params = circuit.get_Parameter_Num()
        if engine_factory is None:
            from .engine import Engine

            engine_factory = Engine
        self.engine = engine_factory(self.rank if device is None else device)
        self.engine.upload_matrix(np.ascontiguousarray(state0, dtype=np.complex128).reshape(-1))
        self.engine.set_circuit(circuit)
        self.engine.set_hamiltonian_csr(indptr, indices, data)

    _tensor = ShardedCost._tensor
    _all_gather_rows = ShardedCost._all_gather_rows

    def energy_grad(self, params, with_grad=True):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p.reshape(1, -1)
        B = p.shape[0]
        counts = [batch_shard(B, r, self.world)[1] - batch_shard(B, r, self.world)[0] for r in range(self.world)]
        b, e = batch_shard(B, self.rank, self.world)
        if with_grad:
            if e > b:
                en, g = self.engine.vqe_energy_grad_batched(p[b:e])
            else:
                en, g = np.zeros(0), np.zeros((0, self.n_params))
            packed = np.concatenate([np.asarray(en).reshape(-1, 1), np.asarray(g).reshape(e - b, self.n_params)], axis=1)
            allp = self._all_gather_rows(packed, counts)
            return allp[:, 0].copy(), allp[:, 1:].copy()
        en = self.engine.vqe_energy_batched(p[b:e]) if e > b else np.zeros(0)
        return self._all_gather_rows(np.asarray(en).reshape(-1, 1), counts)[:, 0].copy()

    def energy(self, params):
        return self.energy_grad(params, with_grad=False)

    def close(self):
        if hasattr(self.engine, "close"):
            self.engine.close()
