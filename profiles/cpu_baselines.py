"""CPU baselines of the BASELINE.json configurations other than the headline one: the reference's OWN
code (oracle/_ref/libsqref.so = the reference's translation units compiled in place, OpenMP-backed TBB shim, scipy's OpenBLAS)
timed on the host cores of the box, next to the device figures of the same inputs when a GPU is visible.
    python profiles/cpu_baselines.py [c1] [c2] [c4] [c5]      -> one JSON line per configuration
Bounded samples: C4 is timed on a circuit truncated to its first 8 of 64 blocks with 256 of 4096 columns (cost-only is linear in
both), C5 on single evaluations. Threads: OMP_NUM_THREADS = all cores, set before the library loads."""
import json
import os
import sys
import time

NCORES = len(os.sched_getaffinity(0))
os.environ["OMP_NUM_THREADS"] = str(NCORES)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import pyoracle
import squander_b200 as sq

ref = pyoracle.Ref()
try:
    import torch

    GPU = torch.cuda.is_available()
except Exception:
    GPU = False


def timeit(fn, budget_s=4.0, min_reps=2):
    fn()
    t0 = time.perf_counter()
    reps = 0
    while reps < min_reps or time.perf_counter() - t0 < budget_s:
        fn()
        reps += 1
        if reps >= 2000:
            break
    return (time.perf_counter() - t0) / reps, reps


def decomp_case(name, U, circ, variant, with_grad=True):
    d, pool = circ.descriptors()
    n = circ.qbit_num
    P = circ.get_Parameter_Num()
    x = H.random_params(P, seed=3)
    out = {"config": name, "n": n, "P": P, "cores": NCORES, "cost_variant": variant}
    for parallel, label in ((0, "1_core"), (2, "all_cores")):
        dec = ref.decomp(U, n, d, pool)
        dec.set_cost(variant, 0)
        dec.set_parallel(parallel)
        s, reps = timeit((lambda: dec.cost_grad(x)) if with_grad else (lambda: dec.cost(x)))
        out["reference_%s_evals_per_s" % label] = round(1.0 / s, 3)
        out["reference_%s_reps" % label] = reps
    if GPU:
        e = sq.Engine(0)
        e.upload_matrix(U)
        e.set_circuit(circ)
        e.set_cost(variant, 0)
        X = np.repeat(x.reshape(1, -1), 256, axis=0)
        f1 = (lambda: e.cost_grad_batched(x.reshape(1, -1))) if with_grad else (lambda: e.cost_batched(x.reshape(1, -1)))
        fb = (lambda: e.cost_grad_batched(X)) if with_grad else (lambda: e.cost_batched(X))
        s1, _ = timeit(f1, 1.0)
        sb, _ = timeit(fb, 1.0)
        out["gpu_evals_per_s_batch1_host_call"] = round(1.0 / s1, 1)
        out["gpu_evals_per_s_batch256_host_call"] = round(256.0 / sb, 1)
        e.close()
    print(json.dumps(out), flush=True)


def c1():
    import golden_cases as G

    decomp_case("C1: n=4 data/Umtx.mat, adaptive L=3, cost+grad", G.load("C1_L3").U, H.adaptive_circuit(4, 3), 0)


def c2():
    import golden_cases as G

    g = G.load("C2_19CNOT")
    c = sq.Circuit(5)
    for r in g.descs:
        t = int(r["type"])
        if t == sq.abi.U3: c.add_U3(int(r["target"]))
        elif t == sq.abi.RX: c.add_RX(int(r["target"]))
        elif t == sq.abi.RY: c.add_RY(int(r["target"]))
        elif t == sq.abi.RZ: c.add_RZ(int(r["target"]))
        elif t == sq.abi.CZ: c.add_CZ(int(r["target"]), int(r["control"]))
        elif t == sq.abi.CNOT: c.add_CNOT(int(r["target"]), int(r["control"]))
    decomp_case("C2: n=5 19CNOT.qasm, Hilbert-Schmidt-test cost, cost+grad", g.U, c, 3)


def c4():
    n, M = 12, 8
    rng = np.random.default_rng(7)
    c = sq.Circuit(n)
    for m in range(M):
        c.add_GENERAL(H.random_unitary(16, seed=1000 + m), sorted(int(q) for q in rng.choice(n, 4, replace=False)))
        if m % 8 == 7:
            for q in range(n):
                c.add_U3(q)
    cols = 256
    rs = np.random.default_rng(1)
    U = np.ascontiguousarray((rs.normal(size=(1 << n, cols)) + 1j * rs.normal(size=(1 << n, cols))) / np.sqrt(2 << n))
    d, pool = c.descriptors()
    P = c.get_Parameter_Num()
    x = H.random_params(P, seed=3)
    out = {"config": "C4: n=12 GENERAL 4-qubit blocks, cost only", "sample": "first 8 of 64 blocks (+ one U3 layer), 256 of 4096 columns",
           "cores": NCORES, "P": P}
    for parallel, label in ((0, "1_core"), (2, "all_cores")):
        # Gates_block::apply_to on a copy of U = optimization_problem without the (negligible) trace; the reference's
        # decomposition class clones its gates and the clone of a multi-target GENERAL gate loses its target list
        rc = ref.circuit(n, d, pool)
        s, reps = timeit(lambda: rc.apply(x, U, parallel), 4.0)
        # cost-only is linear in the gate count and in the columns: 8x the blocks, 16x the columns
        out["reference_%s_sample_s" % label] = round(s, 5)
        out["reference_%s_evals_per_s_full_config" % label] = round(1.0 / (s * 8 * 16), 4)
    out["extrapolation"] = "time x 8 (blocks) x 16 (columns): every gate is one pass over the matrix, columns are independent"
    print(json.dumps(out), flush=True)


def c5():
    n, layers = 20, 10
    ip, ix, dat = H.heisenberg_csr_fast(n)
    out = {"config": "C5: n=20 Heisenberg VQE, HEA_ZYZ 10 layers", "cores": NCORES, "nnz": int(len(dat))}
    v = ref.vqe(n, ip, ix, dat, ansatz="HEA_ZYZ", layers=layers, inner_blocks=1)
    P = v.n_params
    out["P"] = P
    x = H.random_params(P, seed=5)
    s, reps = timeit(lambda: v.energy(x), 8.0)
    out["reference_energy_evals_per_s"] = round(1.0 / s, 3)
    out["reference_energy_reps"] = reps
    out["reference_energy_grad_evals_per_s_estimate"] = round(1.0 / (s * (1 + P)), 5)
    out["estimate_note"] = "the reference's VQE gradient applies the circuit once per parameter (…Base.cpp:1131-1199): (1 + P) energy-like passes"
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2", "c4", "c5"]
    for w in which:
        globals()[w]()
