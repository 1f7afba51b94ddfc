"""Record of the FP64 roofline denominator: the library's DFMA / DMMA m8n8k4 burn kernels
(sqgpu_fp64_fma_peak) with the SM clock and throttle reasons sampled while they run.
usage: python profiles/fp64_peak.py > profiles/r2_fp64_peak.json"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import squander_b200 as sq

samples = []
stop = False


def sampler():
    q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.active,power.draw,temperature.gpu"
    while not stop:
        try:
            out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                samples.append([s.strip() for s in out.split(",")])
        except Exception:
            pass
        time.sleep(0.05)


e = sq.Engine(0, options={"verbose": 1})
e.fp64_fma_peak()  # warm-up (context, clocks)
t = threading.Thread(target=sampler)
t.start()
runs = [e.fp64_fma_peak() for _ in range(5)]
stop = True
t.join()
sm = sorted(float(s[0]) for s in samples if s and s[0].replace(".", "").isdigit())
print(json.dumps({
    "what": "sqgpu_fp64_fma_peak: max(DFMA burn, DMMA m8n8k4 burn) TFLOP/s, 5 runs; details of each burn on stderr (option verbose)",
    "tflops_runs": runs, "tflops_median": sorted(runs)[len(runs) // 2],
    "theory": "148 SMs x 64 DFMA/clk x 2 flop x 1.965 GHz = 37.2 TFLOP/s",
    "clocks": {"sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
               "sm_max_mhz": float(samples[0][1]) if samples else None,
               "throttle_reasons": sorted(set(s[2] for s in samples)), "power_w_max": max((float(s[3]) for s in samples), default=None),
               "samples": len(samples)},
}))
