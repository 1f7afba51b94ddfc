"""N1: steps per second of a single ADAM trajectory, device-resident loop (sqgpu_adam_steps) vs the host-driven loop (cost+gradient
through the host C-ABI call, Adam::update on the host), and the batched line search vs k separate evaluations.
usage: python profiles/bench_optim.py [--shift]  -> one JSON line per configuration (--shift: the COSINE / AGENTS engines)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import pyoracle
import squander_b200 as sq

port = pyoracle.Port()


def run(name, U, circ, steps):
    P = circ.get_Parameter_Num()
    e = sq.Engine(0)
    e.upload_matrix(U)
    e.set_circuit(circ)
    e.set_cost(0, 0)
    x0 = H.random_params(P, seed=3)
    out = {"config": name, "P": P, "steps": steps}
    for label, opts in (("device_graph", {}), ("device_launches", {"no_graph": 1})):
        for k, v in {"no_graph": 0, **opts}.items():
            e.set_option(k, v)
        e.adam_init(x0.reshape(1, -1))
        e.adam_steps(5)
        e.adam_init(x0.reshape(1, -1))
        t0 = time.perf_counter()
        hist = e.adam_steps(steps)
        out[label + "_steps_per_s"] = round(steps / (time.perf_counter() - t0), 1)
    x = x0.copy()
    opt = port.adam(P)
    for _ in range(3):
        e.cost_grad_batched(x.reshape(1, -1))
    t0 = time.perf_counter()
    hh = []
    for _ in range(steps):
        f, g = e.cost_grad_batched(x.reshape(1, -1))
        hh.append(f[0])
        opt.update(x, g[0], f[0])
    out["host_driven_steps_per_s"] = round(steps / (time.perf_counter() - t0), 1)
    out["trajectories_identical"] = bool((np.array(hh) == hist[:, 0]).all())
    out["speedup_graph"] = round(out["device_graph_steps_per_s"] / out["host_driven_steps_per_s"], 2)
    # line search: 16 trial points
    d = np.random.default_rng(1).standard_normal(P)
    alphas = np.linspace(0, 1, 16)
    e.line_search_batched(x0, d, alphas)
    t0 = time.perf_counter()
    for _ in range(10):
        e.line_search_batched(x0, d, alphas)
    tb = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for _ in range(3):
        for a in alphas:
            e.cost_grad_batched((x0 + a * d).reshape(1, -1))
    ts = (time.perf_counter() - t0) / 3
    out["line_search_16_points_ms_batched"] = round(tb * 1e3, 3)
    out["line_search_16_points_ms_one_by_one"] = round(ts * 1e3, 3)
    e.close()
    print(json.dumps(out), flush=True)


def run_shift_engines(name, U, circ, iters):
    """COSINE / AGENTS iterations per second over the device's batched cost: the arrangement of optimize.cosine / optimize.agents
    (two, resp. one, batched call per iteration) against the reference's call pattern for the same iteration (COSINE.cpp:255-523:
    two batched calls of batch_size sets, then ~12 dependent single evaluations of the golden-section line search)."""
    P = circ.get_Parameter_Num()
    e = sq.Engine(0)
    e.upload_matrix(U)
    e.set_circuit(circ)
    e.set_cost(0, 0)
    x0 = H.random_params(P, seed=3)
    out = {"config": name, "P": P, "iterations": iters}
    bs = min(64, P)
    sq.optimize.cosine(e.cost_batched, x0, np.random.default_rng(1), batch_size=bs, max_iter=2, tol=0)
    t0 = time.perf_counter()
    _, f, it, ne = sq.optimize.cosine(e.cost_batched, x0, np.random.default_rng(1), batch_size=bs, max_iter=iters, tol=0, check_for_convergence=False)
    dt = time.perf_counter() - t0
    out.update(cosine_iters_per_s=round(it / dt, 2), cosine_cost_evals_per_s=round(ne / dt, 1), cosine_final_cost=f)
    # the shift batch from two adjoint sweeps (sqgpu_cost_shifted_batched) instead of 2 x batch_size forward passes
    sq.optimize.cosine(e.cost_batched, x0, np.random.default_rng(1), batch_size=bs, max_iter=2, tol=0, cost_shifted=e.cost_shifted_batched)
    t0 = time.perf_counter()
    _, f2, it, ne = sq.optimize.cosine(e.cost_batched, x0, np.random.default_rng(1), batch_size=bs, max_iter=iters, tol=0, check_for_convergence=False,
                                       cost_shifted=e.cost_shifted_batched)
    dt = time.perf_counter() - t0
    out.update(cosine_sweep_iters_per_s=round(it / dt, 2), cosine_sweep_final_cost=f2, cosine_sweep_same_result=bool(abs(f2 - f) < 1e-9 * max(1, abs(f))))
    # ... and with ALL parameters updated per iteration (batch_size = P costs the same two sweeps)
    t0 = time.perf_counter()
    _, f3, it, ne = sq.optimize.cosine(e.cost_batched, x0, np.random.default_rng(1), batch_size=P, max_iter=iters, tol=0, check_for_convergence=False,
                                       cost_shifted=e.cost_shifted_batched)
    dt = time.perf_counter() - t0
    out.update(cosine_sweep_all_params_iters_per_s=round(it / dt, 2), cosine_sweep_all_params_final_cost=f3)
    # the reference's call pattern for one iteration on the same engine
    X = np.repeat(x0.reshape(1, -1), bs, axis=0)
    t0 = time.perf_counter()
    for _ in range(max(2, iters // 4)):
        e.cost_batched(X)
        e.cost_batched(X)
        for _ in range(12):
            e.cost_batched(x0.reshape(1, -1))
    out["cosine_reference_call_pattern_iters_per_s"] = round(max(2, iters // 4) / (time.perf_counter() - t0), 2)
    t0 = time.perf_counter()
    _, f, it, ne = sq.optimize.agents(e.cost_batched, x0, np.random.default_rng(1), agent_num=64, max_iter=iters, tol=0, agent_lifetime=max(10, iters // 4))
    dt = time.perf_counter() - t0
    out.update(agents_iters_per_s=round(it / dt, 2), agents_cost_evals_per_s=round(ne / dt, 1), agents_final_cost=f)
    e.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    import golden_cases as G

    if "--shift" in sys.argv:
        run_shift_engines("C1: n=4 Umtx.mat, adaptive L=3", G.load("C1_L3").U, H.adaptive_circuit(4, 3), 400)
        run_shift_engines("n=8 adaptive L=2", np.ascontiguousarray(H.random_unitary(256, seed=123).conj().T), H.adaptive_circuit(8, 2), 200)
        run_shift_engines("C3: n=10 adaptive L=4", np.ascontiguousarray(H.random_unitary(1024, seed=123).conj().T), H.adaptive_circuit(10, 4), 40)
        sys.exit(0)

    run("C1: n=4 Umtx.mat, adaptive L=3", G.load("C1_L3").U, H.adaptive_circuit(4, 3), 2000)
    g = G.load("C2_19CNOT")
    c2 = sq.Circuit(5)  # rebuild the 19-CNOT structure from its stored descriptors
    e = None
    names = {v: k for k, v in vars(sq.abi).items() if isinstance(v, int) and k.isupper() and k not in ("OK",)}
    for r in g.descs:
        t = int(r["type"])
        if t == sq.abi.U3: c2.add_U3(int(r["target"]))
        elif t == sq.abi.RX: c2.add_RX(int(r["target"]))
        elif t == sq.abi.RY: c2.add_RY(int(r["target"]))
        elif t == sq.abi.RZ: c2.add_RZ(int(r["target"]))
        elif t == sq.abi.CZ: c2.add_CZ(int(r["target"]), int(r["control"]))
        elif t == sq.abi.CNOT: c2.add_CNOT(int(r["target"]), int(r["control"]))
    run("C2: n=5 19CNOT.qasm", g.U, c2, 2000)
    run("n=8 adaptive L=2", np.ascontiguousarray(H.random_unitary(256, seed=123).conj().T), H.adaptive_circuit(8, 2), 500)
    run("C3: n=10 adaptive L=4", np.ascontiguousarray(H.random_unitary(1024, seed=123).conj().T), H.adaptive_circuit(10, 4), 100)
