#!/usr/bin/env python
"""bench.py -- cost+gradient throughput of the decomposition hot path (BASELINE.json metric).

Workload (config.workload = "C3"): BASELINE.json configs[2] -- 10-qubit random unitary (QR of a seeded Gaussian, the
recipe of the reference's tests/gates/test_circuit.py:93-102), adaptive gate structure with L = 4 levels (550 gates,
P = 1290 parameters), a batch of 256 parameter vectors (default_rng(42).random * 2 pi), Frobenius trace cost
(variant 0) and its full parameter gradient. One "step" = cost+gradient for the whole batch.

  python bench.py --gpus N --steps K --warmup W            our engine (libsqgpu.so through the C-ABI)
  python bench.py --impl reference --gpus N ...             the reference's OWN CPU code (oracle/_ref/libsqref.so) on
                                                            the host cores: the SAME structure, one parameter vector per step

N > 1 (launched by torch.distributed.run, one rank per GPU): batch entries are independent
(Optimization_Interface.cpp:1009-1025). Default `--scaling strong`: the configuration's 256 parameter vectors are split over
the ranks (256 / N each) and one NCCL all-gather returns all costs / gradients to every rank -- the device analogue of the
reference's MPI_Allgather (:962-1004). `--scaling weak` keeps 256 vectors per GPU.

Besides the headline line's keys the JSON carries, under "secondary" (skipped with --no-secondary):
  C4_columns  BASELINE configs[3]: n = 12, 64 fused 4-qubit blocks, batch 64, cost only, the COLUMNS of U sharded over the
              ranks (rank r holds U[:, r w:(r+1) w] with trace_offset = r w) and ONE all-reduce of the raw trace terms
              [64 x 3 x 2] per evaluation inside the timed region -- the north star's ">= 6x from 1 to 8 GPUs at 12 qubits".
  C5_vqe      BASELINE configs[4]: n = 20 Heisenberg VQE energy + gradient, 1024 parameter sets split over the ranks.
  latency     one cost+gradient evaluation (batch 1): what a BFGS line search sees; the shift batch of a COSINE iteration from one
              adjoint sweep against the explicit batch of shifted parameter sets.

One JSON line on stdout (rank 0).
"""
import os
import sys

# The reference arm runs the reference's TBB kernels through an OpenMP-backed shim and zgemm through the scipy-bundled
# OpenBLAS. torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which silently serialised that arm at N > 1 in
# round 1: the thread counts are pinned here, BEFORE numpy / libgomp / OpenBLAS are loaded, and the count really used is
# read back from the runtimes and reported.
if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"] or "--impl=reference" in sys.argv:
    _cores = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    os.environ["OMP_NUM_THREADS"] = _cores
    os.environ["OPENBLAS_NUM_THREADS"] = _cores
    os.environ["OMP_DYNAMIC"] = "FALSE"

import argparse
import ctypes
import importlib
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "cost+grad evals/s (10-qubit unitary decomposition)"
UNIT = "evals/s"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fused_exec<GRAD> launch of the default workload (n=10, L=4, batch 256):
# an OFFLINE ncu capture, not measured in this run -- see TRAFFIC_SOURCE. null for every other workload.
TRAFFIC_DEFAULT_WORKLOAD = 762308352 + 537867264
TRAFFIC_SOURCE = {"kind": "offline_capture", "file": "profiles/r2_ncu_fused_grad_b256.csv", "commit": "4edd65a",
                  "note": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch; refreshed per round when the kernel changes"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--qubits", type=int, default=10)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="budget of the cpu_baseline leg of our arm")
    ap.add_argument("--ref-seconds", type=float, default=150.0, help="budget of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--cpu-leg", action="store_true", help=argparse.SUPPRESS)  # internal: cpu_baseline in a clean subprocess
    return ap.parse_args()


def workload(args):
    import helpers as H

    n, L = args.qubits, args.levels
    circ = H.adaptive_circuit(n, L)
    U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)  # examples/decomposition/example.py:62
    return circ, U


def flops_per_eval(descs, rows, cols):
    """ALGORITHMIC real flops of one cost+gradient evaluation by the adjoint sweep, per gate (SURVEY.md §8d "14 flop/amplitude"):
    forward 28 flop per active row pair and column; backward 28 (un-apply) + 28 (row functional) + 32 (W accumulation,
    parametric gates only). A side figure: the roofline fraction uses the flops the kernel really issues."""
    fwd = 0.0
    tot = 0.0
    for d in descs:
        ctrl = (1 if d["control"] >= 0 else 0) + (1 if d["control2"] >= 0 else 0)
        if d["type"] == 1:  # GENERAL
            k = int(d["n_qubits"])
            f = 8.0 * (1 << k) * (1 << k) * (rows >> k) * cols
            fwd += f
            tot += 3 * f
            continue
        pairs = (rows // 2) >> ctrl
        f = 28.0 * pairs * cols
        fwd += f
        tot += f * 3 + (32.0 * pairs * cols if d["n_params"] > 0 else 0.0)
    return fwd, tot


def stream_bytes_per_eval(descs, rows, cols):
    """bytes the reference's per-gate streaming algorithm moves for ONE forward pass"""
    b = 16.0 * rows * cols
    for d in descs:
        ctrl = (1 if d["control"] >= 0 else 0) + (1 if d["control2"] >= 0 else 0)
        b += 32.0 * (rows >> ctrl) * cols
    return b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that load anything under oracle/)
# ---------------------------------------------------------------------------------------------------------------------

def host_threads():
    """threads the OpenMP runtime and the scipy OpenBLAS of this process will really use (read back, not assumed)"""
    # openblas_threads = 1 is the reference's own design: it runs BLAS single-threaded INSIDE its TBB tasks (the reference calls
    # openblas_set_num_threads(1), N_Qubit_Decomposition_adaptive.cpp:293-299; oracle/ref_harness.cpp does the same), so the
    # parallelism of this arm is omp_max_threads, the width of the OpenMP-backed TBB shim
    out = {"omp_max_threads": None, "openblas_threads": None, "affinity_cores": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count(),
           "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS")}
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_get_max_threads.restype = ctypes.c_int
        out["omp_max_threads"] = int(gomp.omp_get_max_threads())
    except Exception:
        pass
    try:
        import glob
        import scipy

        lib = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))[0]
        ob = ctypes.CDLL(lib)
        fn = getattr(ob, "scipy_openblas_get_num_threads64_", None) or getattr(ob, "scipy_openblas_get_num_threads", None)
        if fn is not None:
            fn.restype = ctypes.c_int
            out["openblas_threads"] = int(fn())
    except Exception:
        pass
    return out


def cpu_reference(args, budget_s, steps):
    """Times the reference's own optimization_problem_combined (cost + full gradient, parallel = 2) on the host cores on the
    SAME configuration as our arm: the n-qubit matrix and the full adaptive structure (550 gates, P = 1290 at the default),
    ONE parameter vector per step (our arm's step holds `batch` of them; the metric is evaluations per second either way).

    One evaluation costs P dense 2^n zgemm's plus the prefix / suffix products (Gates_block.cpp:358-428): tens of seconds.
    So: (1) two truncated structures (the first m sub-blocks + the final U3 layer) are timed as warm-up and give the fitted
    law t(P) = a + b P; (2) if the law predicts that one full evaluation fits the budget, the full structure is timed for as many
    steps as fit (at least one) and `value` = 1 / mean(step) is a direct measurement -- "extrapolated": false; otherwise the
    value is the fitted law at P_full and says so."""
    import helpers as H
    import pyoracle
    import squander_b200 as sq

    t_start = time.perf_counter()
    n, L = args.qubits, args.levels
    U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)
    pairs = [(t, c) for t in range(n) for c in range(t + 1, n)] * L
    ref_ok = pyoracle.Ref.available() or os.path.isdir("/root/reference")

    def build(m):
        c = sq.Circuit(n)
        for t, cq in pairs[:m]:
            layer = sq.Circuit(n)
            layer.add_U3(t)
            layer.add_U3(cq)
            layer.add_adaptive(t, cq)
            c.add_Circuit(layer)
        fin = sq.Circuit(n)
        for q in range(n):
            fin.add_U3(q)
        c.add_Circuit(fin)
        return c

    if ref_ok:
        ref = pyoracle.Ref()
        kind = "reference"

        def run(circ, p):
            d, pool = circ.descriptors(nested=True)
            dec = ref.decomp(U, n, d, pool)
            dec.set_parallel(2)
            dec.set_cost(args.variant)
            t0 = time.perf_counter()
            dec.cost_grad(p)
            return time.perf_counter() - t0
    else:
        port = pyoracle.Port()
        kind = "port"

        def run(circ, p):
            d, pool = circ.descriptors()
            t0 = time.perf_counter()
            port.cost_grad(d, circ.get_Parameter_Num(), p, U, n, args.variant)
            return time.perf_counter() - t0

    thr = host_threads()
    cores = (thr["omp_max_threads"] or 1) if kind == "reference" else 1
    rng = np.random.default_rng(42)
    full = build(len(pairs))
    P_full = full.get_Parameter_Num()
    # (1) two truncation levels: warm-up + the fitted law (the larger the levels, the closer the law: the fixed part of an
    # evaluation is seconds)
    levels = []
    for m in ((8, 32) if budget_s >= 90 else (4, 24)):
        c = build(min(m, len(pairs)))
        p = rng.random(c.get_Parameter_Num()) * 2 * np.pi
        levels.append((c.get_Parameter_Num(), run(c, p)))
    (P1, t1), (P2, t2) = levels
    b = max((t2 - t1) / max(P2 - P1, 1), 1e-9)
    a = max(t1 - b * P1, 0.0)
    t_pred = a + b * P_full
    remaining = budget_s - (time.perf_counter() - t_start)
    times = []
    if t_pred * 1.15 <= remaining:
        # (2) the full structure, as many steps as fit
        p_full = np.random.default_rng(42).random(P_full) * 2 * np.pi  # row 0 of our arm's batch
        n_fit = int(max(1, min(steps, remaining // (t_pred * 1.15))))
        for _ in range(n_fit):
            times.append(run(full, p_full))
            if budget_s - (time.perf_counter() - t_start) < 1.15 * times[-1]:
                break
        t_step = float(np.mean(times))
        extrapolated = False
        sample = ("full structure: %d-qubit matrix, %d gates, all %d parameters, 1 parameter vector per step, %d step(s) timed; "
                  "reference optimization_problem_combined (parallel=2, OpenMP-backed TBB shim, scipy OpenBLAS zgemm)"
                  % (n, 3 * len(pairs) + n, P_full, len(times)))
    else:
        t_step = t_pred
        extrapolated = True
        sample = ("the full structure (predicted %.0f s per evaluation) does not fit the %.0f s budget: value = fitted law t(P) = %.3g + %.3g P "
                  "at P = %d from truncated structures with P = %d and %d" % (t_pred, budget_s, a, b, P_full, P1, P2))
    return {"value": 1.0 / t_step, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
            "sample_seconds_per_step": t_step, "steps_timed": max(len(times), 1), "extrapolated": extrapolated,
            "threads": thr, "fit": {"law": "t(P) = a + b P", "a_s": a, "b_s_per_param": b, "levels": [{"P": P1, "s": t1}, {"P": P2, "s": t2}],
                                    "predicted_full_s": t_pred}}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_reference(args, budget_s=args.cpu_seconds if args.cpu_leg else args.ref_seconds, steps=args.steps)
    if args.cpu_leg:
        print(json.dumps(cb), flush=True)
        return
    n_gates = 3 * (args.qubits * (args.qubits - 1) // 2) * args.levels + args.qubits
    P = 7 * (args.qubits * (args.qubits - 1) // 2) * args.levels + 3 * args.qubits
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cb["sample_seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: n=%d random unitary, adaptive L=%d (%d gates, P=%d), cost+grad, Frobenius trace cost"
                   % (args.qubits, args.levels, n_gates, P), "qubits": args.qubits, "levels": args.levels, "batch": 1,
                   "cost_variant": args.variant, "note": "one parameter vector per step on the full structure; rank 0 only, all host threads"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args):
    """our arm's cpu_baseline leg: the reference arm's measurement in a CLEAN process (its own OpenMP / OpenBLAS thread
    settings, no torch runtime in the way), bounded by --cpu-seconds"""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-leg", "--qubits", str(args.qubits), "--levels", str(args.levels),
           "--variant", str(args.variant), "--steps", "1", "--cpu-seconds", str(args.cpu_seconds)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=args.cpu_seconds * 3 + 120, env=env)
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return {"error": (r.stderr or r.stdout)[-500:]}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------

def single_gate_microbench(sq, eng, torch, peak_gbs):
    """G1: HBM GB/s of the single-gate streaming kernels: U3 (three targets), CRY and CNOT on the reference's own shape
    (2^12 x 256, tests/gates/test_float32_performance.py:23-24), on 2^10 x 2^10, 2^12 x 2^12 (256 MiB > L2), 2^20 x 1 and 2^24 x 1.
    Shapes below the 126 MB L2 are flushed out of it between repetitions (timed per launch by the library's CUDA events)."""
    import helpers as H
    abi = sq.abi
    out = []
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for rows, cols in ((1 << 12, 256), (1 << 10, 1 << 10), (1 << 12, 1 << 12), (1 << 20, 1), (1 << 24, 1)):
        n = int(np.log2(rows))
        buf = torch.zeros(rows * cols * 2, dtype=torch.float64, device="cuda")
        buf[0::2] = 1.0 / np.sqrt(rows)
        in_l2 = rows * cols * 16 < (120 << 20)
        cases = [("U3", n - 1, -1), ("U3", n // 2, -1), ("U3", 0, -1), ("CRY", 1, n - 1), ("CNOT", 0, n // 2)]
        for name, t, c in cases:
            circ = sq.Circuit(n)
            H.add_named(circ, name, [t, c if c >= 0 else (t + 1) % n, 0])
            d, pool = circ.descriptors()
            gp = np.array([0.3, 0.7, 1.1][: int(d[0]["n_params"])], dtype=np.float64)
            dd = np.ascontiguousarray(d[:1])
            cargs = (eng._h, dd.ctypes.data_as(ctypes.POINTER(abi.GateDesc)), abi.as_dp(gp) if gp.size else None, None, -1,
                     buf.data_ptr(), rows, cols, cols, stream)
            for _ in range(3):
                abi.check(eng.lib, eng.lib.sqgpu_apply_gate_dev(*cargs))
            torch.cuda.synchronize()
            eng.last_kernel_time()  # reset the library's per-kernel CUDA-event rings
            reps = 10
            for _ in range(reps):
                if in_l2:
                    flush.zero_()
                abi.check(eng.lib, eng.lib.sqgpu_apply_gate_dev(*cargs))
            torch.cuda.synchronize()
            _, ms, nl = eng.last_kernel_time()  # events bracket the streaming kernel only, on the launching stream
            touched = rows * cols / (2 if c >= 0 else 1)
            gbs = 32.0 * touched / (ms * 1e-3) / 1e9
            out.append({"gate": name, "target": t, "control": c, "rows": rows, "cols": cols, "GB/s": round(gbs, 1),
                        "frac_of_measured_hbm": round(gbs / peak_gbs, 3), "us": round(ms * 1e3, 2), "l2_flushed": bool(in_l2)})
        del buf
    return out


def gpu_random_unitary(torch, dim, seed):
    """the QR recipe of helpers.random_unitary with the factorisation on the GPU (bench SETUP only; n = 12 takes minutes on one
    host thread under torchrun)"""
    rng = np.random.default_rng(seed)
    a = torch.from_numpy(rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))).cuda()
    q, r = torch.linalg.qr(a)
    dr = torch.diagonal(r)
    return (q * (dr / dr.abs())).contiguous()


def max_over_ranks(torch, dist, world, x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def secondary_c4(sq, torch, dist, rank, local_rank, world, steps=3, warmup=2):
    """C4: n = 12, 64 Haar 16 x 16 blocks on random 4-qubit subsets + a U3 layer after every 8th, cost only,
    batch 64, columns of U sharded over the ranks, one all-reduce of the raw traces per evaluation, strong scaling."""
    import helpers as H

    n, M, B = 12, 64, 64
    rng = np.random.default_rng(7)
    c = sq.Circuit(n)
    for m in range(M):
        qs = sorted(int(q) for q in rng.choice(n, 4, replace=False))
        c.add_GENERAL(H.random_unitary(16, seed=1000 + m), qs)
        if m % 8 == 7:
            for q in range(n):
                c.add_U3(q)
    P = c.get_Parameter_Num()
    cols = 1 << n
    w = cols // world
    Uq = gpu_random_unitary(torch, cols, 123)  # Q; the target passed to the engine is Q^dagger (example.py:62)
    Ush = Uq.conj().T[:, rank * w:(rank + 1) * w].contiguous()
    eng = sq.Engine(local_rank)
    eng.upload_matrix(Ush.cpu().numpy())
    del Uq, Ush
    eng.set_circuit(c)
    eng.set_cost(0, rank * w)
    params = torch.from_numpy(H.random_params(P, batch=B)).cuda()
    traces = torch.zeros(B * 6, dtype=torch.float64, device="cuda")
    cost = torch.zeros(B, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        eng.traces_batched_dev(params.data_ptr(), B, False, traces.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(traces, op=dist.ReduceOp.SUM)
        eng.cost_from_traces_dev(traces.data_ptr(), B, False, cols, cost.data_ptr(), 0, stream.cuda_stream)

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        e0.record(stream)
        step()
        e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(torch, dist, world, float(np.mean([a.elapsed_time(b) for a, b in evs])))
    tf, sf = eng.last_exec_flops()
    kms, kn = eng.kernel_time("fused_exec<COST>")
    out = {"workload": "C4: n=12, 64 GENERAL 4-qubit blocks + U3 layers (P=%d), batch %d, cost only; columns sharded %d x %d, one all-reduce of "
                       "[%d x 3 x 2] trace terms per evaluation inside the timed region" % (P, B, world, w, B),
           "scaling": "strong", "n_gpus": world, "evals_per_s": B / (ms * 1e-3), "ms_per_step": ms, "cost0": float(cost[0].item()),
           "kernel_ms_rank0": kms, "executed_tensor_tflops_rank0": (tf / (kms * 1e-3) / 1e12) if kms > 0 else None,
           "input_bytes_resident": int(w) * cols * 16, "l2": "inputs larger than L2 at N <= 2 (256 / 128 MiB per GPU); smaller shards are re-read from L2"}
    eng.close()
    return out


def secondary_c5(sq, torch, dist, rank, local_rank, world, steps=2, warmup=1):
    """C5: n = 20 Heisenberg VQE (3-regular graph seed 31415, HEA_ZYZ 10 layers, P = 1140), energy + gradient for 1024 parameter
    sets split over the ranks (128 per GPU at N = 8), one all-gather of energies and gradients."""
    import helpers as H

    n, layers, B = 20, 10, 1024
    ip, ix, dat = H.heisenberg_csr_fast(n)
    c = H.hea_zyz_circuit(n, layers)
    P = c.get_Parameter_Num()
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[0] = 1.0
    eng = sq.Engine(local_rank)
    eng.upload_matrix(psi0)
    eng.set_circuit(c)
    eng.set_hamiltonian_csr(ip, ix, dat)
    Bl = B // world
    params = torch.from_numpy(H.random_params(P, seed=5, batch=B)[rank * Bl:(rank + 1) * Bl].copy()).cuda()
    out_l = torch.zeros(Bl * (1 + P), dtype=torch.float64, device="cuda")
    out_all = torch.zeros(B * (1 + P), dtype=torch.float64, device="cuda") if world > 1 else None
    stream = torch.cuda.current_stream()

    def step():
        eng.vqe_energy_grad_batched_dev(params.data_ptr(), Bl, out_l.data_ptr(), out_l.data_ptr() + 8 * Bl, stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(out_all, out_l)

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.last_kernel_time()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        e0.record(stream)
        step()
        e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = max_over_ranks(torch, dist, world, float(np.mean([a.elapsed_time(b) for a, b in evs])))
    fwd_ms, fwd_n = eng.kernel_time("fused_exec<WINDOW_FWD>")
    bwd_ms, bwd_n = eng.kernel_time("fused_exec<WINDOW_BWD>")
    # algorithmic HBM bytes of the windowed executor: one read + one write of every state (16 B per amplitude) per segment,
    # two states (psi, lambda) in the backward sweep
    st = sq.abi.plan_stats(c)
    seg = st["segments"]
    state_bytes = 16.0 * (1 << n)
    peaks, _ = measured_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    res = {"workload": "C5: n=20 Heisenberg VQE, HEA_ZYZ 10 layers (P=%d), energy+gradient, %d parameter sets over %d GPU(s)" % (P, B, world),
           "scaling": "strong", "n_gpus": world, "evals_per_s": B / (ms * 1e-3), "ms_per_step": ms, "energy0": float(out_l[0].item()),
           "window_segments": seg, "window_fwd_ms": fwd_ms, "window_bwd_ms": bwd_ms}
    if fwd_ms > 0 and fwd_n > 0:
        # the timer brackets all segments of one slice of parameter sets; slices per step = fwd_n / steps
        sets_per_bracket = Bl * steps / fwd_n
        res["window_fwd_hbm_GB/s"] = round(2 * state_bytes * seg * sets_per_bracket / (fwd_ms * 1e-3) / 1e9, 1)
        res["window_fwd_hbm_frac"] = round(res["window_fwd_hbm_GB/s"] / hbm, 3)
    if bwd_ms > 0 and bwd_n > 0:
        sets_per_bracket = Bl * steps / bwd_n
        res["window_bwd_hbm_GB/s"] = round(4 * state_bytes * seg * sets_per_bracket / (bwd_ms * 1e-3) / 1e9, 1)
        res["window_bwd_hbm_frac"] = round(res["window_bwd_hbm_GB/s"] / hbm, 3)
    eng.close()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sqgpu engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sq = importlib.import_module("sequential-quantum-gate-decomposer_b200")
    circ, U = workload(args)
    descs, _ = circ.descriptors()
    n, P = args.qubits, circ.get_Parameter_Num()
    strong = args.scaling == "strong"
    Bg = args.batch if strong else args.batch * world          # global batch
    if Bg % world:
        raise SystemExit("bench.py: the global batch %d does not divide over %d ranks" % (Bg, world))
    B = Bg // world                                             # per rank
    all_params = np.random.default_rng(42).random((Bg, P)) * 2 * np.pi
    params = np.ascontiguousarray(all_params[rank * B:(rank + 1) * B])
    eng = sq.Engine(local_rank)
    eng.upload_matrix(U)
    eng.set_circuit(circ)
    eng.set_cost(args.variant, 0)

    stream = torch.cuda.current_stream()
    d_params = torch.from_numpy(params).cuda()
    d_out = torch.empty(B * (1 + P), dtype=torch.float64, device="cuda")  # [cost(B) | grad(B*P)]
    d_cost = d_out[:B]
    d_grad = d_out[B:]
    d_all = torch.empty(world * B * (1 + P), dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        eng.cost_grad_batched_dev(d_params.data_ptr(), B, d_cost.data_ptr(), d_grad.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    eng.last_kernel_time()  # reset the per-kernel event rings
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in evs:
        flush.zero_()  # L2 flush between timed iterations, outside the timed pair
        e0.record(stream)
        step()
        e1.record(stream)
    barrier()
    launches = eng.launch_count() - launches0
    kname, kms, klaunches = eng.last_kernel_time()
    tensor_flops, scalar_flops = eng.last_exec_flops()
    shape = eng.last_launch_shape()
    step_ms = max_over_ranks(torch, dist, world, float(np.mean([a.elapsed_time(b) for a, b in evs])))
    value = Bg / (step_ms * 1e-3)
    # A short timed region (8 GPUs: ~0.1 s) ends before nvidia-smi (200 ms period) has printed a sample: the same step keeps
    # running, untimed and uncounted, until ~1 s of this load has been sampled. The count derives from the max-over-ranks step
    # time, so every rank runs the same number of steps (the step holds a collective).
    extra_steps = 0
    if step_ms * args.steps < 900.0:
        extra_steps = int(min(500, np.ceil(1000.0 / max(step_ms, 1e-3))))
        for _ in range(extra_steps):
            step()
        torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    if extra_steps:
        clocks["note"] = "timed region %.0f ms: %d untimed extra steps of the same load so that nvidia-smi (200 ms period) samples it" % (step_ms * args.steps, extra_steps)

    # ---- e2e: the same step through the host-buffer C-ABI call (pinned numpy in, numpy out) -------------------
    h_params = torch.from_numpy(params).pin_memory().numpy()
    h_cost = torch.empty(B, dtype=torch.float64).pin_memory().numpy()
    h_grad = torch.empty((B, P), dtype=torch.float64).pin_memory().numpy()
    abi = sq.abi

    def e2e_step():
        abi.check(eng.lib, eng.lib.sqgpu_cost_grad_batched(eng._h, abi.as_dp(h_params), B, abi.as_dp(h_cost), abi.as_dp(h_grad)))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(torch, dist, world, (time.perf_counter() - t0) / args.steps)
    assert np.allclose(h_cost, d_cost.cpu().numpy(), rtol=0, atol=1e-12)

    # ---- latency: ONE cost+gradient evaluation (what a BFGS line search or an ADAM step sees) ---------------------
    lat = {}
    if rank == 0:
        one = h_params[:1].copy()
        c1 = np.zeros(1)
        g1 = np.zeros((1, P))
        for _ in range(3):
            abi.check(eng.lib, eng.lib.sqgpu_cost_grad_batched(eng._h, abi.as_dp(one), 1, abi.as_dp(c1), abi.as_dp(g1)))
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            abi.check(eng.lib, eng.lib.sqgpu_cost_grad_batched(eng._h, abi.as_dp(one), 1, abi.as_dp(c1), abi.as_dp(g1)))
        lat["cost_grad_batch1_ms_host_call"] = (time.perf_counter() - t0) / reps * 1e3
        eng.last_kernel_time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(reps):
            eng.cost_grad_batched_dev(d_params.data_ptr(), 1, d_cost.data_ptr(), d_grad.data_ptr(), stream.cuda_stream)
        ev1.record(stream)
        torch.cuda.synchronize()
        lat["cost_grad_batch1_ms_device"] = ev0.elapsed_time(ev1) / reps
        lat["kernel"] = eng.last_kernel_time()[0]
        lat["launch_shape_batch1"] = eng.last_launch_shape()
        t0 = time.perf_counter()
        for _ in range(reps):
            abi.check(eng.lib, eng.lib.sqgpu_cost_batched(eng._h, abi.as_dp(one), 1, abi.as_dp(c1)))
        lat["cost_batch1_ms_host_call"] = (time.perf_counter() - t0) / reps * 1e3
        lat["evals_per_s_batch1"] = 1e3 / lat["cost_grad_batch1_ms_host_call"]
        # the shift batch of a COSINE iteration (COSINE.cpp:255-291: 64 parameters x shifts pi/2 and pi): ONE adjoint sweep that
        # returns the shifted costs of all P parameters (sqgpu_cost_shifted_batched) against the explicit batch of 128 sets
        if args.variant in (0, 1, 2, 3, 9):
            try:
                sh = np.array([np.pi / 2, np.pi])
                fs = np.zeros((2, 1, P))
                for _ in range(2):
                    abi.check(eng.lib, eng.lib.sqgpu_cost_shifted_batched(eng._h, abi.as_dp(one), 1, abi.as_dp(sh), 2, abi.as_dp(c1), abi.as_dp(fs)))
                t0 = time.perf_counter()
                for _ in range(reps):
                    abi.check(eng.lib, eng.lib.sqgpu_cost_shifted_batched(eng._h, abi.as_dp(one), 1, abi.as_dp(sh), 2, abi.as_dp(c1), abi.as_dp(fs)))
                lat["shift_batch_one_sweep_ms_host_call"] = (time.perf_counter() - t0) / reps * 1e3
                idx = np.random.default_rng(1).choice(P, min(64, P), replace=False)
                X = np.repeat(one, 2 * idx.size, axis=0)
                X[np.arange(idx.size), idx] += np.pi / 2
                X[idx.size + np.arange(idx.size), idx] += np.pi
                cx = np.zeros(2 * idx.size)
                abi.check(eng.lib, eng.lib.sqgpu_cost_batched(eng._h, abi.as_dp(X), 2 * idx.size, abi.as_dp(cx)))
                t0 = time.perf_counter()
                for _ in range(3):
                    abi.check(eng.lib, eng.lib.sqgpu_cost_batched(eng._h, abi.as_dp(X), 2 * idx.size, abi.as_dp(cx)))
                lat["shift_batch_explicit_%d_sets_ms_host_call" % (2 * idx.size)] = (time.perf_counter() - t0) / 3 * 1e3
                lat["shift_batch_max_abs_difference"] = float(max(np.abs(fs[0, 0, idx] - cx[:idx.size]).max(), np.abs(fs[1, 0, idx] - cx[idx.size:]).max()))
            except Exception as ex:  # the headline number must not depend on it
                lat["shift_batch_error"] = repr(ex)

    secondary = {}
    if not args.no_secondary:
        if world > 1 and rank == 0:
            secondary["note"] = "each entry is measured in this same launch, after the headline loop"
        # the other scaling mode of the headline workload, for the record
        if world > 1:
            alt_B = args.batch if strong else args.batch // world
            if alt_B >= 1:
                ap_ = torch.from_numpy(np.ascontiguousarray((np.random.default_rng(43).random((alt_B, P)) * 2 * np.pi))).cuda()
                a_out = torch.empty(alt_B * (1 + P), dtype=torch.float64, device="cuda")
                a_all = torch.empty(world * alt_B * (1 + P), dtype=torch.float64, device="cuda")

                def alt_step():
                    eng.cost_grad_batched_dev(ap_.data_ptr(), alt_B, a_out.data_ptr(), a_out.data_ptr() + 8 * alt_B, stream.cuda_stream)
                    dist.all_gather_into_tensor(a_all, a_out)

                alt_step()
                barrier()
                aev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
                for e0, e1 in aev:
                    flush.zero_()
                    e0.record(stream)
                    alt_step()
                    e1.record(stream)
                barrier()
                ams = max_over_ranks(torch, dist, world, float(np.mean([a.elapsed_time(b) for a, b in aev])))
                secondary["C3_" + ("weak" if strong else "strong")] = {"batch_per_gpu": alt_B, "global_batch": alt_B * world, "ms_per_step": ams,
                                                                       "evals_per_s": alt_B * world / (ams * 1e-3)}
                del ap_, a_out, a_all
        try:
            secondary["C4_columns"] = secondary_c4(sq, torch, dist, rank, local_rank, world)
        except Exception as ex:  # the headline number must not depend on the secondary workloads
            secondary["C4_columns"] = {"error": repr(ex)}
            if world > 1:
                raise
        try:
            secondary["C5_vqe"] = secondary_c5(sq, torch, dist, rank, local_rank, world)
        except Exception as ex:
            secondary["C5_vqe"] = {"error": repr(ex)}
            if world > 1:
                raise

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    fp64_peak = eng.fp64_fma_peak()
    fwd_flops, alg_flops = flops_per_eval(descs, 1 << n, 1 << n)
    k_s = kms * 1e-3 if kms > 0 else step_ms * 1e-3
    exec_flops = tensor_flops + scalar_flops
    achieved_tf = exec_flops / k_s / 1e12
    sb = stream_bytes_per_eval(descs, 1 << n, 1 << n)
    default_wl = (n, args.levels, B, args.variant) == (10, 4, 256, 0)
    roofline = {
        "kernel": kname, "bound": "tensor", "achieved": round(achieved_tf, 3), "peak": round(fp64_peak, 3), "unit": "TFLOP/s",
        "frac": round(achieved_tf / fp64_peak, 4) if fp64_peak > 0 else None,
        "traffic": TRAFFIC_DEFAULT_WORKLOAD if default_wl else None, "traffic_source": TRAFFIC_SOURCE if default_wl else None,
        "definition": "achieved = FP64 flops the kernel ISSUES per launch (DMMA m8n8k4 = 512 flops each + scalar DFMA paths, counted by the "
                      "library from the launch's own op list, sqgpu_last_exec_flops; validated against ncu sm__ops_path_tensor_src_fp64 in "
                      "profiles/) / the kernel's CUDA-event time on the launching stream; frac = achieved / peak",
        "peak_source": "FP64 DMMA m8n8k4 / DFMA burn kernels run in this process (sqgpu_fp64_fma_peak, the larger of the two: one pipe); "
                       "MEASURED_PEAKS.json holds only HBM and bf16 figures; record with clocks in profiles/r2_fp64_peak.json",
        "kernel_ms": round(kms, 4), "kernel_launches_timed": klaunches, "launch_shape": shape,
        "executed_tensor_flops_per_launch": tensor_flops, "executed_scalar_flops_per_launch": scalar_flops,
        "algorithmic": {"flops_per_launch": alg_flops * B, "TFLOP/s": round(alg_flops * B / k_s / 1e12, 3),
                        "frac": round(alg_flops * B / k_s / 1e12 / fp64_peak, 4) if fp64_peak > 0 else None,
                        "note": "per-gate adjoint count (SURVEY 8d); larger than what the fused blocks execute -- a side figure, not the roofline fraction"},
        "hbm_equivalent": {"bytes_per_eval_streaming": 4 * sb, "achieved_GB/s": round(4 * sb * B / k_s / 1e9, 1),
                           "peak_GB/s": peaks.get("hbm_gbs"), "peak_source": peak_src,
                           "frac": round(4 * sb * B / k_s / 1e9 / peaks.get("hbm_gbs", 6650.0), 2),
                           "note": "bandwidth the reference's per-gate streaming algorithm would need for the same evals/s"},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C3: n=%d random unitary, adaptive L=%d (%d gates, P=%d), global batch %d (%d per GPU), cost+grad, Frobenius trace cost"
                   % (n, args.levels, len(descs), P, Bg, B), "qubits": n, "levels": args.levels, "batch_per_gpu": B,
                   "global_batch": Bg, "cost_variant": args.variant, "parallelism": "batch-sharded x%d, one all-gather per step" % world,
                   "l2": "256 MB flush write between timed iterations"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": Bg / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes),
                "d2h_bytes_per_step": int(h_cost.nbytes + h_grad.nbytes), "ms_per_step": e2e_s * 1e3},
        "roofline": roofline, "latency": lat,
    }
    if secondary:
        line["secondary"] = secondary
    if not args.no_microbench and world == 1:
        try:
            line["single_gate_hbm"] = single_gate_microbench(sq, eng, torch, peaks.get("hbm_gbs", 6650.0))
        except Exception as ex:  # the headline number must not depend on the microbenchmark
            line["single_gate_hbm"] = {"error": str(ex)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_subprocess(args)
        except Exception as ex:
            line["cpu_baseline"] = {"error": str(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
