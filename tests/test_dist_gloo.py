"""N > 1 host logic on CPU: world_size = 2 over the gloo backend (127.0.0.1 rendezvous). The CUDA engine is replaced by
a test double built on the C oracle (tests may use the oracle; the product never does), so what is exercised here is
the sharding plan, the single collective per evaluation and the lock-step results of dist.ShardedCost."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers as H

abi = H.abi


class OracleEngine:
    """engine double with the Engine methods dist.ShardedCost uses, evaluated by oracle/sq_oracle.c"""

    def __init__(self, device=0):
        import pyoracle

        self.port = pyoracle.Port()
        self.variant, self.off, self.cfg = 0, 0, (1.0, 1 / 1.7, 0.5)
        self.shard = 0

    def set_shard(self, col_begin, cols_total):
        self.shard = col_begin

    def _off(self):
        # sqgpu_set_shard: the shard's row offset enters every variant; the user's trace_offset the Frobenius family only
        return (self.off if self.variant <= 2 else 0) + self.shard

    def upload_matrix(self, U):
        self.U = np.ascontiguousarray(U)

    def set_circuit(self, circuit):
        self.descs, self.pool = circuit.descriptors()
        self.P = circuit.get_Parameter_Num()
        self.n = circuit.qbit_num

    def set_cost(self, variant, trace_offset, prev, c1, c2):
        self.variant, self.off, self.cfg = variant, trace_offset, (prev, c1, c2)

    def _omega(self, tr0=None):
        """complex weights of the three trace types in the functional whose gradient is taken (reduce.cuh: make_omega)"""
        prev, c1, c2 = self.cfg
        sp = np.sqrt(prev)
        if self.variant in (4, 5):
            T = tr0[:, 0] - 1j * tr0[:, 1]  # conj of the summed traces of the circuit itself
            return (T[0], sp * c1 * T[1], sp * c2 * T[2] if self.variant == 5 else 0)
        return {0: (1, 0, 0), 1: (1, sp * c1, 0), 2: (1, sp * c1, sp * c2), 3: (1, 0, 0), 9: (1, 0, 0), 6: (1, 0, 0)}[self.variant]

    def traces_batched(self, params, with_grad, tr0=None):
        out = np.zeros((len(params), 1 + (self.P if with_grad else 0), 3, 2))
        for b, p in enumerate(params):
            m = self.port.apply_circuit(self.descs, p, self.U, self.pool)
            out[b, 0] = self.port.traces(m, self.n, self._off()).reshape(3, 2)
            if with_grad:
                w = self._omega(None if tr0 is None else tr0[b, 0])
                d = self.port.apply_derivate(self.descs, self.P, p, self.U, self.pool)
                for k in range(self.P):
                    t = self.port.traces(d[k], self.n, self._off()).reshape(3, 2)
                    tc = t[:, 0] + 1j * t[:, 1]
                    dl = w[0] * tc[0] + w[1] * tc[1] + w[2] * tc[2]
                    out[b, 1 + k, 0] = (dl.real, dl.imag)
        return out

    def grad_traces_with_global(self, params, tr0):
        return self.traces_batched(params, True, tr0)

    def cost_from_traces(self, tr, with_grad, cols_total):
        prev, c1, c2 = self.cfg
        cost = np.array([self.port.cost_from_traces(self.variant, t[0].reshape(-1), cols_total, prev, c1, c2) for t in tr])
        if not with_grad:
            return cost
        n = float(cols_total)
        grad = np.zeros((len(tr), self.P))
        for b, t in enumerate(tr):
            T, dl = t[0, 0], t[1:, 0]
            if self.variant <= 2:
                grad[b] = (1.0 - dl[:, 0] / n) - 1.0
            elif self.variant == 3:
                grad[b] = -2.0 / n / n * (T[0] * dl[:, 0] + T[1] * dl[:, 1])
            elif self.variant in (4, 5):
                grad[b] = -2.0 / n / n * dl[:, 0]
            else:
                grad[b] = -2.0 / n / (n + 1) * (T[0] * dl[:, 0] + T[1] * dl[:, 1])
        return cost, grad

    def cost_batched(self, params):
        prev, c1, c2 = self.cfg
        return np.array([self.port.cost(self.descs, p, self.U, self.n, self.variant, self.off, prev, c1, c2, self.pool) for p in params])


    def cost_grad_batched(self, params):
        prev, c1, c2 = self.cfg
        res = [self.port.cost_grad(self.descs, self.P, p, self.U, self.n, self.variant, self.off, prev, c1, c2, self.pool) for p in params]
        return np.array([r[0] for r in res]), np.array([r[1] for r in res])


class OracleVQEEngine:
    """engine double with the Engine methods dist.ShardedVQE uses"""

    def __init__(self, device=0):
        import pyoracle

        self.port = pyoracle.Port()

    def upload_matrix(self, psi0):
        self.psi0 = np.ascontiguousarray(psi0)

    def set_circuit(self, circuit):
        self.descs, self.pool = circuit.descriptors()
        self.P = circuit.get_Parameter_Num()

    def set_hamiltonian_csr(self, indptr, indices, data):
        self.H = (indptr, indices, data)

    def vqe_energy_grad_batched(self, params):
        res = [self.port.vqe_energy_grad(self.descs, self.P, p, self.psi0, *self.H, pool=self.pool) for p in params]
        return np.array([r[0] for r in res]), np.array([r[1] for r in res])

    def vqe_energy_batched(self, params):
        return self.vqe_energy_grad_batched(params)[0]


def _vqe_problem():
    n = 4
    circ = H.hea_zyz_circuit(n, 1)
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[0] = 1.0
    return n, circ, psi0, H.heisenberg_csr(n), H.random_params(circ.get_Parameter_Num(), seed=9, batch=5)


def _worker(rank, world, port_no, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sq = H.sq
        n = 4
        circ = H.adaptive_circuit(n, 1)
        P = circ.get_Parameter_Num()
        U = H.random_unitary(1 << n).conj().T.copy()
        params = H.random_params(P, batch=5)
        res = {}
        for mode in ("batch", "columns"):
            for variant in (0, 2, 3, 4, 5, 9):
                sc = sq.dist.ShardedCost(U, circ, variant=variant, mode=mode, prev_cost=0.37, engine_factory=OracleEngine)
                c, g = sc.cost_grad(params)
                res[(mode, variant)] = (c, g, sc.cost(params))
        _, vcirc, psi0, (ip, ix, dv), vparams = _vqe_problem()
        sv = sq.dist.ShardedVQE(psi0, vcirc, ip, ix, dv, engine_factory=OracleVQEEngine)
        en, gr = sv.energy_grad(vparams)
        res["vqe"] = (en, gr, sv.energy(vparams))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_plans():
    d = H.sq.dist
    for total, world in ((1024, 8), (23, 4), (5, 2), (256, 3)):
        blocks = [d.column_shard(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == total
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert d.HS_CORRECTION == (abi.HILBERT_SCHMIDT_TEST_CORRECTION1, abi.HILBERT_SCHMIDT_TEST_CORRECTION2)


def test_world2_gloo_matches_single_process(port):
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 4
    circ = H.adaptive_circuit(n, 1)
    d, pool = circ.descriptors()
    P = circ.get_Parameter_Num()
    U = H.random_unitary(1 << n).conj().T.copy()
    params = H.random_params(P, batch=5)
    _, vcirc, psi0, (ip, ix, dv), vparams = _vqe_problem()
    vd, vpool = vcirc.descriptors()
    en0, gr0, ee0 = results[0].pop("vqe")
    en1, gr1, ee1 = results[1].pop("vqe")
    assert (en0 == en1).all() and (gr0 == gr1).all() and (ee0 == ee1).all()
    for b in range(len(vparams)):
        e_ref, g_ref = port.vqe_energy_grad(vd, vcirc.get_Parameter_Num(), vparams[b], psi0, ip, ix, dv, pool=vpool)
        assert abs(en0[b] - e_ref) < 1e-12 and abs(ee0[b] - e_ref) < 1e-12 and np.abs(gr0[b] - g_ref).max() < 1e-12
    for key, (c0, g0, cc0) in results[0].items():
        mode, variant = key
        c1, g1, cc1 = results[1][key]
        assert (c0 == c1).all() and (g0 == g1).all() and (cc0 == cc1).all()  # lock-step: identical on every rank
        for b in range(5):
            f_ref, g_ref = port.cost_grad(d, P, params[b], U, n, variant, 0, 0.37)
            assert abs(c0[b] - f_ref) < 1e-12 and abs(cc0[b] - f_ref) < 1e-12
            assert np.abs(g0[b] - g_ref).max() < 1e-12
