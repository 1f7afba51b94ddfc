"""N3: constant sub-circuits multiplied out into dense 16 x 16 kernels against leaving them to the 3-qubit block planner.
n = 10, `stretches` x (U3 layer + `gates` random constant gates on 4 qubits), cost and cost+gradient, batch 64.
usage: python profiles/bench_const_fuse.py [stretches] [gates]   -> one JSON line"""
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

stretches = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gates = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n, B = 10, 64
c = H.const_heavy_circuit(n, stretches, gates, seed=3)
P = c.get_Parameter_Num()
U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)
ps = H.random_params(P, seed=1, batch=B)
out = {"workload": "n=10, %d x (U3 layer + %d constant gates on 4 qubits), P=%d, batch %d" % (stretches, gates, P, B)}
ref = None
for name, opts in (("dense_kernels", {}), ("blocks_only", {"const_fuse_qubits": 0})):
    e = sq.Engine(0, options=opts)
    e.upload_matrix(U)
    e.set_circuit(c)
    e.set_cost(0, 0)
    f = e.cost_batched(ps)
    t0 = time.perf_counter()
    for _ in range(3):
        f = e.cost_batched(ps)
    tc = (time.perf_counter() - t0) / 3
    fg, g = e.cost_grad_batched(ps)
    t0 = time.perf_counter()
    for _ in range(3):
        fg, g = e.cost_grad_batched(ps)
    tg = (time.perf_counter() - t0) / 3
    out[name] = {"ops_plan3": len(sq.abi.plan_ops(c, which=3, **opts)), "cost_evals_per_s": B / tc, "cost_grad_evals_per_s": B / tg}
    if ref is None:
        ref = (f, g)
    else:
        out["max_abs_diff_cost"] = float(np.abs(f - ref[0]).max())
        out["max_abs_diff_grad"] = float(np.abs(g - ref[1]).max())
    e.close()
print(json.dumps(out))
