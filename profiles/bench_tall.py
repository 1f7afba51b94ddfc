"""Matrices too tall for one shared-memory column (gradient n >= 13, cost n >= 14): the windowed executor on column chunks
against the one-op-per-launch streaming fallback (option tall_window = 0).
usage: python profiles/bench_tall.py [n] [levels] [cols] [batch]   -> one JSON line"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cols = int(sys.argv[3]) if len(sys.argv) > 3 else 512
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 4
c = H.adaptive_circuit(n, levels)
P = c.get_Parameter_Num()
rng = np.random.default_rng(1)
U = np.ascontiguousarray((rng.standard_normal((1 << n, cols)) + 1j * rng.standard_normal((1 << n, cols))) / np.sqrt(1 << n))
theta = H.random_params(P, seed=3, batch=batch)
out = {"workload": "n=%d adaptive L=%d (P=%d), %d columns, batch %d, cost+grad" % (n, levels, P, cols, batch)}
res = {}
for name, opts in (("windowed", {}), ("streaming", {"tall_window": 0})):
    e = sq.Engine(0, options=opts)
    e.set_circuit(c)
    e.upload_matrix(U)
    e.set_cost(0, 0)
    f, g = e.cost_grad_batched(theta)  # warm-up
    reps = 3 if name == "windowed" else 1
    t0 = time.perf_counter()
    for _ in range(reps):
        f, g = e.cost_grad_batched(theta)
    dt = (time.perf_counter() - t0) / reps
    res[name] = (f, g)
    out[name] = {"s_per_call": dt, "evals_per_s": batch / dt, "kernel": e.last_kernel_time()[0], "launches": e.launch_count(),
                 "plan": e.plan_stats() if hasattr(e, "plan_stats") else None}
    e.close()
out["speedup"] = out["streaming"]["s_per_call"] / out["windowed"]["s_per_call"]
out["max_rel_diff_grad"] = float(np.abs(res["windowed"][1] - res["streaming"][1]).max() / max(1.0, np.abs(res["streaming"][1]).max()))
print(json.dumps(out))
