"""Matrices too tall for one CTA's shared memory (gradient n >= 13, cost n >= 14): the cluster executor (2 / 4 CTAs of a
thread-block cluster share a column), the windowed executor on column chunks (option cluster = 0) and the one-op-per-launch
streaming fallback (tall_window = 0); n <= 12: the single-CTA executor against clusters of two.
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

sys_args = [a for a in sys.argv if not a.startswith("--")]
n = int(sys_args[1]) if len(sys_args) > 1 else 13
levels = int(sys_args[2]) if len(sys_args) > 2 else 1
cols = int(sys_args[3]) if len(sys_args) > 3 else 512
batch = int(sys_args[4]) if len(sys_args) > 4 else 4
c = H.adaptive_circuit(n, levels)
P = c.get_Parameter_Num()
rng = np.random.default_rng(1)
U = np.ascontiguousarray((rng.standard_normal((1 << n, cols)) + 1j * rng.standard_normal((1 << n, cols))) / np.sqrt(1 << n))
theta = H.random_params(P, seed=3, batch=batch)
out = {"workload": "n=%d adaptive L=%d (P=%d), %d columns, batch %d, cost+grad" % (n, levels, P, cols, batch)}
res = {}
variants = (("cluster", {"cluster": 2}), ("windowed", {"cluster": 0}), ("streaming", {"tall_window": 0, "cluster": 0}))
if "--no-stream" in sys.argv:
    variants = variants[:2]
if n <= 12:
    variants = (("single_cta", {"cluster": 0}), ("cluster", {}))
for name, opts in variants:
    e = sq.Engine(0, options=opts)
    e.set_circuit(c)
    e.upload_matrix(U)
    e.set_cost(0, 0)
    f, g = e.cost_grad_batched(theta)  # warm-up
    reps = 1 if name == "streaming" else 3
    t0 = time.perf_counter()
    for _ in range(reps):
        f, g = e.cost_grad_batched(theta)
    dt = (time.perf_counter() - t0) / reps
    res[name] = (f, g)
    shape = e.last_launch_shape()
    out[name] = {"s_per_call": dt, "evals_per_s": batch / dt, "kernel": e.last_kernel_time()[0], "launches": e.launch_count(),
                 "cluster": shape["cluster"], "threads": shape["threads"], "smem": shape["smem"]}
    e.close()
names = [v[0] for v in variants]
out["speedup_first_vs_last"] = out[names[-1]]["s_per_call"] / out[names[0]]["s_per_call"]
out["max_rel_diff_grad"] = max(float(np.abs(res[a][1] - res[names[-1]][1]).max() / max(1.0, np.abs(res[names[-1]][1]).max())) for a in names[:-1])
print(json.dumps(out))
