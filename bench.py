#!/usr/bin/env python
"""bench.py -- cost+gradient throughput of the decomposition hot path (BASELINE.json metric).

Workload (config.workload = "C3"): BASELINE.json configs[2] -- 10-qubit random unitary (QR of a seeded Gaussian, the
recipe of the reference's tests/gates/test_circuit.py:93-102), adaptive gate structure with L = 4 levels (550 gates,
P = 1290 parameters), a batch of 256 parameter vectors (default_rng(42).random * 2 pi), Frobenius trace cost
(variant 0) and its full parameter gradient. One "step" = cost+gradient for the whole batch.

  python bench.py --gpus N --steps K --warmup W            our engine (libsqgpu.so through the C-ABI)
  python bench.py --impl reference --gpus N ...             the reference's OWN CPU code (oracle/_ref/libsqref.so) on
                                                            the host cores, bounded sample (see cpu_sample())

N > 1 (launched by torch.distributed.run): batch entries are independent (Optimization_Interface.cpp:1009-1025), so
every rank evaluates its own 256 parameter vectors on the full matrix ("weak" scaling) and one NCCL all-gather returns
all costs/gradients to every rank -- the device analogue of the reference's MPI_Allgather (:962-1004).

One JSON line on stdout (rank 0).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "cost+grad evals/s (10-qubit unitary decomposition)"
UNIT = "evals/s"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE fused_exec<GRAD> launch of the default workload (n=10, L=4, batch 256),
# from the `ncu --set full` capture summarised in profiles/r1_ncu_fused_grad_v10_b256.csv (789.6 MB read + 587.3 MB written).
# Algorithmic HBM bytes of the same launch: U once (16.8 MB) + block tables (187 MB) + W partials written once (3.2 GB would
# be the naive figure; they are reduced in L2) -- the kernel is tensor-pipe bound, HBM runs at 0.04 % of peak.
TRAFFIC_DEFAULT_WORKLOAD = 1376906752
# tensor-pipe flops the same launch EXECUTES (sm__ops_path_tensor_src_fp64.sum): the planner's fused blocks need fewer flops
# than the per-gate algorithmic count that `achieved` is defined on, so both fractions are reported
EXECUTED_TENSOR_FLOPS_DEFAULT_WORKLOAD = 5634997092352
TRAFFIC_NOTE = "ncu --set full capture of this launch (profiles/r1_ncu_fused_grad_v10_b256.csv); null for non-default workloads"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=10)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-microbench", action="store_true")
    return ap.parse_args()


def workload(args, rank=0):
    import helpers as H

    n, L = args.qubits, args.levels
    circ = H.adaptive_circuit(n, L)
    U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)  # examples/decomposition/example.py:62
    P = circ.get_Parameter_Num()
    rng = np.random.default_rng(42 + rank)
    params = rng.random((args.batch, P)) * 2 * np.pi
    return circ, U, params


def flops_per_eval(descs, rows, cols):
    """algorithmic real flops of one cost+gradient evaluation by the adjoint sweep:
    forward 28 flop per active row pair and column (4 complex mul + 2 complex add, SURVEY.md §8d "14 flop/amplitude");
    backward 28 (un-apply) + 28 (row functional) + 32 (W accumulation, parametric gates only)."""
    fwd = 0.0
    tot = 0.0
    for d in descs:
        ctrl = (1 if d["control"] >= 0 else 0) + (1 if d["control2"] >= 0 else 0)
        if d["type"] == 1:  # GENERAL
            k = int(d["n_qubits"])
            per_group = 8.0 * (1 << k) * (1 << k)
            groups = rows >> k
            f = per_group * groups * cols
            fwd += f
            tot += 3 * f
            continue
        pairs = (rows // 2) >> ctrl
        f = 28.0 * pairs * cols
        fwd += f
        tot += f * 3 + (32.0 * pairs * cols if d["n_params"] > 0 else 0.0)
    return fwd, tot


def stream_bytes_per_eval(descs, rows, cols):
    """bytes the reference's per-gate streaming algorithm moves for ONE forward pass:
    32 B per touched amplitude per gate + one read of U."""
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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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

def cpu_sample(args, budget_s, steps=1, warmup=0):
    """Times the reference's own optimization_problem_combined on the host cores on a BOUNDED sample.

    One evaluation of the full structure costs P dense 2^n zgemm's (Gates_block.cpp:358-428) -- minutes on a
    workstation -- so the sample is ONE parameter vector on the same matrix with the gate structure truncated to its
    first m sub-blocks (+ the final U3 layer), m chosen so that a step fits the budget; the rate is scaled by
    P_sample / P_full (the reference's gradient cost is linear in the number of parameters: one suffix zgemm and one
    derivative-kernel pass per parameter, one prefix + one suffix product per gate)."""
    import helpers as H
    import pyoracle
    import squander_b200 as sq

    n, L = args.qubits, args.levels
    full = H.adaptive_circuit(n, L)
    P_full = full.get_Parameter_Num()
    U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)
    ref_ok = pyoracle.Ref.available() or os.path.isdir("/root/reference")
    cores = os.cpu_count() or 1
    pairs = [(t, c) for t in range(n) for c in range(t + 1, n)] * L

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
        cores = 1

        def run(circ, p):
            d, pool = circ.descriptors()
            t0 = time.perf_counter()
            port.cost_grad(d, circ.get_Parameter_Num(), p, U, n, args.variant)
            return time.perf_counter() - t0

    rng = np.random.default_rng(42)
    # calibrate on 2 sub-blocks, then pick m for the per-step budget
    c0 = build(2)
    p0 = rng.random(c0.get_Parameter_Num()) * 2 * np.pi
    run(c0, p0)
    t_cal = run(c0, p0)
    per_param = t_cal / c0.get_Parameter_Num()
    per_step_budget = budget_s / max(1, steps + warmup)
    m = int(max(2, min(len(pairs), (per_step_budget / per_param - 3 * n) / 7)))
    circ = build(m)
    P_s = circ.get_Parameter_Num()
    p = rng.random(P_s) * 2 * np.pi
    for _ in range(warmup):
        run(circ, p)
    times = [run(circ, p) for _ in range(max(1, steps))]
    t_step = float(np.mean(times))
    evals_per_s_full = (1.0 / t_step) * (P_s / P_full)
    sample = ("1 parameter vector, %d-qubit matrix, gate structure truncated to the first %d of %d sub-blocks + final U3 layer "
              "(%d of %d parameters), reference optimization_problem_combined (parallel=2, OpenMP-backed TBB shim); "
              "rate scaled by %d/%d to the full structure" % (n, m, len(pairs), P_s, P_full, P_s, P_full))
    return {"value": evals_per_s_full, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
            "sample_seconds_per_step": t_step, "steps": len(times)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_sample(args, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cb["sample_seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: n=%d adaptive L=%d cost+grad, Frobenius trace cost" % (args.qubits, args.levels),
                   "qubits": args.qubits, "levels": args.levels, "batch": 1},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------

def single_gate_microbench(sq, eng, torch, peak_gbs):
    """G1: HBM GB/s of the single-gate streaming kernels on a 2^12 x 2^12 matrix (256 MiB > L2) and a 2^24 state vector."""
    import helpers as H
    abi = sq.abi
    out = []
    stream = torch.cuda.current_stream().cuda_stream
    for rows, cols in ((1 << 12, 1 << 12), (1 << 24, 1)):
        n = int(np.log2(rows))
        buf = torch.zeros(rows * cols * 2, dtype=torch.float64, device="cuda")
        buf[0::2] = 1.0 / np.sqrt(rows)
        cases = [("U3", n - 1, -1), ("U3", n // 2, -1), ("U3", 0, -1), ("CRY", 1, n - 1), ("CNOT", 0, n // 2)]
        for name, t, c in cases:
            circ = sq.Circuit(n)
            H.add_named(circ, name if name != "CRY" else "CRY", [t, c if c >= 0 else (t + 1) % n, 0])
            d, pool = circ.descriptors()
            gp = np.array([0.3, 0.7, 1.1][: int(d[0]["n_params"])], dtype=np.float64)
            dd = np.ascontiguousarray(d[:1])
            import ctypes as C
            args = (eng._h, dd.ctypes.data_as(C.POINTER(abi.GateDesc)), abi.as_dp(gp) if gp.size else None, None, -1,
                    buf.data_ptr(), rows, cols, cols, stream)
            for _ in range(3):
                abi.check(eng.lib, eng.lib.sqgpu_apply_gate_dev(*args))
            torch.cuda.synchronize()
            eng.last_kernel_time()  # reset the library's per-kernel CUDA-event ring
            reps = 10
            for _ in range(reps):
                abi.check(eng.lib, eng.lib.sqgpu_apply_gate_dev(*args))
            torch.cuda.synchronize()
            _, ms, nl = eng.last_kernel_time()  # events bracket the streaming kernel only, on the launching stream
            touched = rows * cols / (2 if c >= 0 else 1)
            gbs = 32.0 * touched / (ms * 1e-3) / 1e9
            out.append({"gate": name, "target": t, "control": c, "rows": rows, "cols": cols, "GB/s": round(gbs, 1),
                        "frac_of_measured_hbm": round(gbs / peak_gbs, 3)})
        del buf
    return out


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
    circ, U, params = workload(args, rank)
    descs, _ = circ.descriptors()
    n, P, B = args.qubits, circ.get_Parameter_Num(), args.batch
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
    eng.last_kernel_time()  # reset the per-kernel event ring
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
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    kname, kms, klaunches = eng.last_kernel_time()
    step_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    t = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = B * world / (step_ms * 1e-3)

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
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert np.allclose(h_cost, d_cost.cpu().numpy(), rtol=0, atol=1e-12)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    fp64_peak = eng.fp64_fma_peak()
    fwd_flops, tot_flops = flops_per_eval(descs, 1 << n, 1 << n)
    k_s = kms * 1e-3 if kms > 0 else step_ms * 1e-3
    achieved_tf = tot_flops * B / k_s / 1e12
    sb = stream_bytes_per_eval(descs, 1 << n, 1 << n)
    default_wl = (n, args.levels, B, args.variant) == (10, 4, 256, 0)
    roofline = {
        "kernel": kname, "bound": "tensor", "achieved": round(achieved_tf, 3), "peak": round(fp64_peak, 3), "unit": "TFLOP/s",
        "frac": round(achieved_tf / fp64_peak, 4) if fp64_peak > 0 else None, "traffic": TRAFFIC_DEFAULT_WORKLOAD if default_wl else None,
        "peak_source": "FP64 tensor-core (DMMA m8n8k4) / DFMA burn kernels run in this process (sqgpu_fp64_fma_peak, the larger of "
                       "the two: they share one pipe); MEASURED_PEAKS.json holds only HBM and bf16 figures, not usable for an f64 path",
        "traffic_note": TRAFFIC_NOTE,
        "kernel_ms": round(kms, 4), "kernel_launches_timed": klaunches,
        "algorithmic_flops_per_launch": tot_flops * B,
        "executed_tensor_flops_per_launch": EXECUTED_TENSOR_FLOPS_DEFAULT_WORKLOAD if default_wl else None,
        "frac_executed": round(EXECUTED_TENSOR_FLOPS_DEFAULT_WORKLOAD / k_s / 1e12 / fp64_peak, 4) if (default_wl and fp64_peak > 0) else None,
        "note": "the executor keeps column tiles in shared memory and runs the fused blocks on the FP64 tensor cores, so that pipe "
                "bounds it, not HBM; achieved/frac use the per-gate ALGORITHMIC flop count (SURVEY 8d), frac_executed the flops the "
                "tensor pipe really executes for it (ncu; block fusion needs ~21 % fewer); hbm_equivalent is the "
                "bandwidth the reference's per-gate streaming algorithm would need for the same evals/s",
        "hbm_equivalent": {"bytes_per_eval_streaming": 4 * sb, "achieved_GB/s": round(4 * sb * B / k_s / 1e9, 1),
                           "peak_GB/s": peaks.get("hbm_gbs"), "peak_source": peak_src,
                           "frac": round(4 * sb * B / k_s / 1e9 / peaks.get("hbm_gbs", 6650.0), 2)},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C3: n=%d random unitary, adaptive L=%d (%d gates, P=%d), batch %d per GPU, cost+grad, Frobenius trace cost"
                   % (n, args.levels, len(descs), P, B), "qubits": n, "levels": args.levels, "batch_per_gpu": B,
                   "global_batch": B * world, "cost_variant": args.variant, "parallelism": "batch-sharded x%d" % world,
                   "l2": "256 MB flush write between timed iterations"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": B * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes),
                "d2h_bytes_per_step": int(h_cost.nbytes + h_grad.nbytes), "ms_per_step": e2e_s * 1e3},
        "roofline": roofline,
    }
    if not args.no_microbench and world == 1:
        try:
            line["single_gate_hbm"] = single_gate_microbench(sq, eng, torch, peaks.get("hbm_gbs", 6650.0))
        except Exception as ex:  # the headline number must not depend on the microbenchmark
            line["single_gate_hbm"] = {"error": str(ex)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_sample(args, budget_s=args.cpu_seconds)
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
