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


HS_CORRECTION = (abi.HILBERT_SCHMIDT_TEST_CORRECTION1, abi.HILBERT_SCHMIDT_TEST_CORRECTION2)


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
            # the shard's row offset enters the trace terms of every cost variant (sqgpu_set_shard); the cost formulas run on
            # the summed traces with the full column count
            self.engine.set_shard(b, self.cols)
            self.engine.set_cost(self.variant, 0, *self.cfg)
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
            if self._on_device() and self.world > 1:
                return self._batch_on_device(p, b, e, counts, with_grad)
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
        # columns: shard traces -> one all-reduce -> cost formulas on the summed traces. The Hilbert-Schmidt correction
        # variants take the weights of their gradient functional from the summed traces of the circuit itself: one more,
        # small all-reduce in front of the gradient pass.
        hs_corr = with_grad and self.variant in HS_CORRECTION
        if self._on_device():
            return self._columns_on_device(p, with_grad, hs_corr)
        if hs_corr:
            tr0 = self._all_reduce_sum(self.engine.traces_batched(p, False))
            tr = self.engine.grad_traces_with_global(p, tr0)
        else:
            tr = self.engine.traces_batched(p, with_grad)
        tr = self._all_reduce_sum(tr)
        return self.engine.cost_from_traces(tr, with_grad, self.cols)

    def _on_device(self):
        """NCCL backend + the CUDA engine: parameters, traces and results stay in device tensors, the all-reduce runs on the
        compute stream between the executor and the cost formulas (no host round trip before the final read-back)"""
        return (self.dist.is_initialized() and self.dist.get_backend(self.group) == "nccl"
                and hasattr(self.engine, "traces_batched_dev"))

    def _columns_on_device(self, p, with_grad, hs_corr):
        import torch

        B, P = p.shape
        n_k = 1 + (P if with_grad else 0)
        st = torch.cuda.current_stream().cuda_stream
        d_p = torch.from_numpy(p).cuda()
        tr = torch.empty(B * n_k * 6, dtype=torch.float64, device="cuda")
        if hs_corr:
            tr0 = torch.empty(B * 6, dtype=torch.float64, device="cuda")
            self.engine.traces_batched_dev(d_p.data_ptr(), B, False, tr0.data_ptr(), st)
            if self.world > 1:
                self.dist.all_reduce(tr0, op=self.dist.ReduceOp.SUM, group=self.group)
            self.engine.grad_traces_with_global_dev(d_p.data_ptr(), B, tr0.data_ptr(), tr.data_ptr(), st)
        else:
            self.engine.traces_batched_dev(d_p.data_ptr(), B, with_grad, tr.data_ptr(), st)
        if self.world > 1:
            self.dist.all_reduce(tr, op=self.dist.ReduceOp.SUM, group=self.group)
        cost = torch.empty(B, dtype=torch.float64, device="cuda")
        grad = torch.empty(B * P, dtype=torch.float64, device="cuda") if with_grad else None
        self.engine.cost_from_traces_dev(tr.data_ptr(), B, with_grad, self.cols, cost.data_ptr(), grad.data_ptr() if with_grad else 0, st)
        if with_grad:
            return cost.cpu().numpy(), grad.cpu().numpy().reshape(B, P)
        return cost.cpu().numpy()

    def _batch_on_device(self, p, b, e, counts, with_grad):
        """batch sharding with device tensors end to end: one all-gather of [max slice x (1 + P)] per evaluation"""
        import torch

        B, P = p.shape
        width = 1 + (P if with_grad else 0)
        mx = max(counts)
        st = torch.cuda.current_stream().cuda_stream
        packed = torch.zeros(mx * width, dtype=torch.float64, device="cuda")  # [cost(mx) | grad(mx * P)], zero padded
        nb = e - b
        if nb:
            d_p = torch.from_numpy(np.ascontiguousarray(p[b:e])).cuda()
            if with_grad:
                self.engine.cost_grad_batched_dev(d_p.data_ptr(), nb, packed.data_ptr(), packed.data_ptr() + 8 * mx, st)
            else:
                self.engine.cost_batched_dev(d_p.data_ptr(), nb, packed.data_ptr(), st)
        gathered = torch.empty(self.world * mx * width, dtype=torch.float64, device="cuda")
        self.dist.all_gather_into_tensor(gathered, packed, group=self.group)
        g = gathered.cpu().numpy().reshape(self.world, mx * width)
        cost = np.concatenate([g[r, : counts[r]] for r in range(self.world)])
        if not with_grad:
            return cost
        grad = np.concatenate([g[r, mx: mx + counts[r] * P].reshape(counts[r], P) for r in range(self.world)], axis=0)
        return cost, grad

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
