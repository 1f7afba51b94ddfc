"""Single-evaluation latency (what a BFGS line search waits for) with the derivative tables built by one warp per fused block
(split_tables = 0) and by one warp per block member (split_tables = 1, the default for batches <= 8; 2 = always):
    python profiles/bench_latency.py            -> one JSON line per configuration
C3 structure (n = 10, 4 levels, P = 1290), C1-like (n = 4, 3 levels) and the batch-256 throughput with the split forced on."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch
import helpers as H
import squander_b200 as sq

st = torch.cuda.current_stream()


def timed(fn, reps, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n, L in ((10, 4), (4, 3), (8, 2)):
    c = H.adaptive_circuit(n, L)
    P = c.get_Parameter_Num()
    U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)
    p = torch.from_numpy(np.random.default_rng(42).random((256, P)) * 2 * np.pi).cuda()
    o = torch.zeros(256 * (1 + P), dtype=torch.float64, device="cuda")
    line = {"n": n, "levels": L, "P": P}
    ref = None
    for mode in (0, 1, 2):
        e = sq.Engine(0, options={"split_tables": mode})
        e.upload_matrix(U)
        e.set_circuit(c)
        e.set_cost(0, 0)
        for B, reps in ((1, 50), (4, 20), (256, 3)):
            ms = timed(lambda: e.cost_grad_batched_dev(p.data_ptr(), B, o.data_ptr(), o.data_ptr() + 8 * 256, st.cuda_stream), reps)
            line["split%d_batch%d_ms" % (mode, B)] = round(ms, 4)
        res = o[: 256 * (1 + P)].clone()
        if ref is None:
            ref = res
        line["split%d_bit_identical" % mode] = bool(torch.equal(ref, res))
        e.close()
    print(json.dumps(line), flush=True)
