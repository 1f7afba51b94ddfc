"""Secondary measurements for the BASELINE.json configs other than the headline one (bench.py times C3).
usage: python profiles/bench_configs.py [c2] [c4] [c5] [c3cost]   -> one JSON line per config"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import squander_b200 as sq


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def c3cost():
    n, L, B = 10, 4, 256
    c = H.adaptive_circuit(n, L)
    e = sq.Engine(0)
    e.upload_matrix(np.ascontiguousarray(H.random_unitary(1 << n).conj().T))
    e.set_circuit(c)
    e.set_cost(0)
    p = H.random_params(c.get_Parameter_Num(), batch=B)
    t = timeit(lambda: e.cost_batched(p))
    return {"config": "C3 cost only (Optimization_Problem_Batch)", "n": n, "batch": B, "evals_per_s": B / t, "kernel": e.last_kernel_time()}


def c2():
    import golden_cases as G
    g = G.load("C2_19CNOT")
    e = sq.Engine(0)
    e.upload_matrix(g.U)
    e.set_circuit_raw(g.descs, g.pool, g.P, g.n)
    e.set_cost(3)
    B = 1
    p = np.repeat(g.params[:1], B, axis=0)
    t = timeit(lambda: e.cost_grad_batched(p), reps=20, warm=3)
    p256 = np.repeat(g.params[:1], 256, axis=0)
    t256 = timeit(lambda: e.cost_grad_batched(p256), reps=5, warm=2)
    return {"config": "C2 19CNOT.qasm cost+grad, HS-test cost", "n": g.n, "P": g.P, "evals_per_s_batch1": B / t, "latency_ms_batch1": t * 1e3,
            "evals_per_s_batch256": 256 / t256}


def c4():
    n, M, B = 12, 64, 64
    rng = np.random.default_rng(7)
    c = sq.Circuit(n)
    for m in range(M):
        qs = sorted(int(q) for q in rng.choice(n, 4, replace=False))
        c.add_GENERAL(H.random_unitary(16, seed=1000 + m), qs)
        if m % 8 == 7:
            for q in range(n):
                c.add_U3(q)
    e = sq.Engine(0)
    U = np.ascontiguousarray(H.random_unitary(1 << n).conj().T)
    e.upload_matrix(U)
    e.set_circuit(c)
    e.set_cost(0)
    p = H.random_params(c.get_Parameter_Num(), batch=B)
    t = timeit(lambda: e.cost_batched(p), reps=2)
    flops = B * (M * 8.0 * 16 * 16 * (1 << n) / 16 * (1 << n) + 8 * n * 28.0 * (1 << n) / 2 * (1 << n))
    return {"config": "C4 n=12, 64 GENERAL 4-qubit blocks + U3 layers, cost only", "batch": B, "evals_per_s": B / t,
            "TFLOP/s_algorithmic": flops / t / 1e12, "kernel": e.last_kernel_time()}


def c4q5():
    """C4 with 5-qubit blocks: 32 x 32 kernels, 256 flop per amplitude"""
    n, M, B = 12, 64, 32
    rng = np.random.default_rng(9)
    c = sq.Circuit(n)
    for m in range(M):
        qs = sorted(int(q) for q in rng.choice(n, 5, replace=False))
        c.add_GENERAL(H.random_unitary(32, seed=2000 + m), qs)
    e = sq.Engine(0)
    e.upload_matrix(np.ascontiguousarray(H.random_unitary(1 << n).conj().T))
    e.set_circuit(c)
    e.set_cost(0)
    p = H.random_params(max(c.get_Parameter_Num(), 1), batch=B)[:, : c.get_Parameter_Num()]
    t = timeit(lambda: e.cost_batched(p), reps=2)
    flops = B * M * 8.0 * 32 * (1 << n) * (1 << n)
    return {"config": "C4 variant: n=12, 64 GENERAL 5-qubit blocks, cost only", "batch": B, "evals_per_s": B / t,
            "TFLOP/s_algorithmic": flops / t / 1e12, "kernel": e.last_kernel_time()}


def c5(n=20, layers=10, B=64):
    indptr, indices, data = H.heisenberg_csr_fast(n)
    c = H.hea_zyz_circuit(n, layers)
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[0] = 1
    e = sq.Engine(0)
    e.upload_matrix(psi0)
    e.set_circuit(c)
    e.set_hamiltonian_csr(indptr, indices, data)
    p = H.random_params(c.get_Parameter_Num(), batch=B)
    t = timeit(lambda: e.vqe_energy_batched(p), reps=2)
    tg = timeit(lambda: e.vqe_energy_grad_batched(p[:16]), reps=1)
    return {"config": "C5 n=%d Heisenberg VQE, HEA_ZYZ %d layers" % (n, layers), "gates": len(c.descriptors()[0]), "P": c.get_Parameter_Num(),
            "nnz": int(data.size), "energy_evals_per_s": B / t, "energy_grad_evals_per_s": 16 / tg}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3cost", "c4", "c5"]
    for w in which:
        print(json.dumps({"c2": c2, "c4": c4, "c4q5": c4q5, "c5": c5, "c3cost": c3cost}[w]()), flush=True)
