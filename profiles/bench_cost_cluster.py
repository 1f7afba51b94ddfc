import sys, time, json
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import numpy as np, helpers as H, squander_b200 as sq
out = {}
for n, cols, B in ((13, 1024, 8), (12, 2048, 8)):
    c = H.adaptive_circuit(n, 2)
    P = c.get_Parameter_Num()
    rng = np.random.default_rng(1)
    U = np.ascontiguousarray((rng.standard_normal((1 << n, cols)) + 1j * rng.standard_normal((1 << n, cols))) / np.sqrt(1 << n))
    th = H.random_params(P, seed=3, batch=B)
    for name, opts in (("single", {"cluster": 0}), ("default", {})):
        e = sq.Engine(0, options=opts); e.set_circuit(c); e.upload_matrix(U); e.set_cost(0, 0)
        f = e.cost_batched(th)
        t0 = time.perf_counter()
        for _ in range(3): f = e.cost_batched(th)
        dt = (time.perf_counter() - t0) / 3
        out["n%d_cost_%s" % (n, name)] = {"evals_per_s": round(B / dt, 1), "shape": e.last_launch_shape(), "f0": float(f[0])}
        e.close()
print(json.dumps(out))
