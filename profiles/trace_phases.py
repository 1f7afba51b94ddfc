"""Phase trace of the fused executor (profiling build: OUT=var_trace.so build.sh -DSQ_TRACE=4096 [-DSQ_PRELOAD=0]):
the first two CTAs on SM 0 record, per op and warp, the SM clock at the start and end of the op's work loop. Prints how the
CTAs' loop phases overlap: fraction of time 0 / 1 / 2 CTAs have at least one warp inside a loop, and per-op statistics.
    SQGPU_LIB=.../var_trace.so python profiles/trace_phases.py [c3|c5]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import squander_b200 as sq

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
EV = int(os.environ.get("SQ_TRACE_EV", "4096"))
lib = sq.abi.load_library()
lib.sqgpu_debug_trace.restype = C.c_longlong
lib.sqgpu_debug_trace.argtypes = [C.c_void_p, C.c_longlong]
buf = np.zeros((2, EV, 16, 2), dtype=np.int64)

if which == "c3":
    n, L, B = 10, 4, 256
    c = H.adaptive_circuit(n, L)
    e = sq.Engine(0)
    e.upload_matrix(np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T))
    e.set_circuit(c)
    e.set_cost(0, 0)
    p = H.random_params(c.get_Parameter_Num(), batch=B)
    e.cost_grad_batched(p)
    nwarps = 8
else:
    n, layers, B = 20, 10, 64
    ip, ix, dat = H.heisenberg_csr_fast(n)
    c = H.hea_zyz_circuit(n, layers)
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[0] = 1
    e = sq.Engine(0)
    e.upload_matrix(psi0)
    e.set_circuit(c)
    e.set_hamiltonian_csr(ip, ix, dat)
    p = H.random_params(c.get_Parameter_Num(), seed=5, batch=B)
    e.vqe_energy_grad_batched(p)
    nwarps = 8
got = lib.sqgpu_debug_trace(buf.ctypes.data, buf.size)
assert got == buf.size, got
out = {"workload": which}
t = buf[:, :, :nwarps, :].astype(np.float64)
valid = (t[:, :, 0, 0] > 0) & (t[:, :, 0, 1] > 0)
nev = int(min(valid[0].sum(), valid[1].sum()))
out["events_per_cta"] = nev
if nev > 100:
    lo, hi = 50, nev - 10  # skip the start-up
    t = t[:, lo:hi]
    t0 = t[:, :, :, 0].min()
    t -= t0
    start, end = t[:, :, :, 0], t[:, :, :, 1]
    # CTA-level: loop phase of an op = [first warp starts, last warp ends]; warp-level durations
    cta_s, cta_e = start.min(axis=2), end.max(axis=2)
    out["op_period_clks"] = float(np.diff(cta_s, axis=1).mean())
    out["warp_loop_clks_mean"] = float((end - start).mean())
    out["cta_loop_span_clks_mean"] = float((cta_e - cta_s).mean())
    out["warp_end_skew_clks_mean"] = float((end.max(axis=2) - end.min(axis=2)).mean())
    out["gap_between_ops_clks_mean"] = float((cta_s[:, 1:] - cta_e[:, :-1]).mean())
    # timeline: number of CTAs with any warp in a loop, sampled every 16 clks over the common span
    T0, T1 = max(cta_s[0, 0], cta_s[1, 0]), min(cta_e[0, -1], cta_e[1, -1])
    grid = np.arange(T0, T1, 16.0)
    inloop = np.zeros((2, grid.size), dtype=np.int32)
    wcount = np.zeros(grid.size, dtype=np.int32)
    for ci in range(2):
        for w in range(nwarps):
            s_idx = np.searchsorted(grid, start[ci, :, w])
            e_idx = np.searchsorted(grid, end[ci, :, w])
            d = np.zeros(grid.size + 1, dtype=np.int32)
            np.add.at(d, s_idx, 1)
            np.add.at(d, e_idx, -1)
            cur = np.cumsum(d[:-1])
            wcount += cur
            inloop[ci] |= (cur > 0)
    both = inloop.sum(axis=0)
    out["frac_time_0_ctas_in_loop"] = float((both == 0).mean())
    out["frac_time_1_cta_in_loop"] = float((both == 1).mean())
    out["frac_time_2_ctas_in_loop"] = float((both == 2).mean())
    out["warps_in_loop_hist"] = [float((wcount == k).mean()) for k in range(2 * nwarps + 1)]
    # per scheduler (SM sub-partition s holds warps w with w % 4 == s of both CTAs): warps inside a loop at a time
    sm_hist = np.zeros(5)
    for sp in range(4):
        cnt = np.zeros(grid.size, dtype=np.int32)
        for ci in range(2):
            for w in range(sp, nwarps, 4):
                d = np.zeros(grid.size + 1, dtype=np.int32)
                np.add.at(d, np.searchsorted(grid, start[ci, :, w]), 1)
                np.add.at(d, np.searchsorted(grid, end[ci, :, w]), -1)
                cnt += np.cumsum(d[:-1])
        for k in range(5):
            sm_hist[k] += (cnt == k).mean() / 4
    out["warps_in_loop_per_scheduler_hist"] = sm_hist.tolist()
    dur = end - start
    out["warp_loop_clks_by_warp_cta0"] = dur[0].mean(axis=0).tolist()
    out["warp_loop_clks_by_warp_cta1"] = dur[1].mean(axis=0).tolist()
    out["warp_loop_clks_p10_p50_p90"] = np.percentile(dur, [10, 50, 90]).tolist()
    # order in which the warps of a CTA finish an op: how often each warp is the last one
    last = end.argmax(axis=2)
    out["last_warp_hist_cta0"] = np.bincount(last[0], minlength=nwarps).tolist()
    first = end.argmin(axis=2)
    out["first_warp_hist_cta0"] = np.bincount(first[0], minlength=nwarps).tolist()
    # start skew: how far apart the warps of a CTA leave the barrier / prologue
    out["warp_start_skew_clks_mean"] = float((start.max(axis=2) - start.min(axis=2)).mean())
    # phase offset between the CTAs: start of CTA 1's op relative to CTA 0's nearest op start, in units of the period
    per = out["op_period_clks"]
    idx = np.searchsorted(cta_s[0], cta_s[1]) - 1
    ok = (idx >= 0) & (idx < cta_s.shape[1])
    ph = ((cta_s[1][ok] - cta_s[0][idx[ok]]) / per) % 1.0
    out["phase_offset_hist10"] = np.histogram(ph, bins=10, range=(0, 1))[0].tolist()
print(json.dumps(out))
