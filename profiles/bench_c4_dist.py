"""C4 scaling run: n = 12, 64 random 4-qubit GENERAL blocks + U3 layers, cost only, batch 64, the columns of
U sharded over the ranks (rank r holds U[:, r*w:(r+1)*w], trace_offset = r*w, one all-reduce of the raw traces).
STRONG scaling: the total work is fixed. Launch (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/bench_c4_dist.py

Timing: K steps bracketed by barrier + torch.cuda.synchronize(), max over ranks; each step goes through the public host
API (ShardedCost.cost: numpy parameters in, numpy costs out), so H2D/D2H copies and the NCCL all-reduce are inside."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import helpers as H
import squander_b200 as sq

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n, M, B = 12, 64, 64
steps, warmup = 3, 2
rng = np.random.default_rng(7)
c = sq.Circuit(n)
for m in range(M):
    qs = sorted(int(q) for q in rng.choice(n, 4, replace=False))
    c.add_GENERAL(H.random_unitary(16, seed=1000 + m), qs)
    if m % 8 == 7:
        for q in range(n):
            c.add_U3(q)
U = np.ascontiguousarray(H.random_unitary(1 << n).conj().T)
params = H.random_params(c.get_Parameter_Num(), batch=B)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if world > 1:
    sc = sq.dist.ShardedCost(U, c, variant=0, mode="columns", device=local)
    run = lambda: sc.cost(params)
else:
    eng = sq.Engine(local)
    eng.upload_matrix(U)
    eng.set_circuit(c)
    eng.set_cost(0)
    run = lambda: eng.cost_batched(params)

for _ in range(warmup):
    out = run()
barrier()
t0 = time.perf_counter()
for _ in range(steps):
    out = run()
barrier()
dt = (time.perf_counter() - t0) / steps
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "C4 n=12, 64 GENERAL 4-qubit blocks + U3 layers, cost only, columns sharded", "n_gpus": world,
                      "batch": B, "scaling": "strong", "evals_per_s": B / float(t.item()), "ms_per_step": float(t.item()) * 1e3,
                      "cost0": float(np.asarray(out)[0])}), flush=True)
if world > 1:
    dist.destroy_process_group()
