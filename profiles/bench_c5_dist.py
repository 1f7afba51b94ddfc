"""C5 scaling run: n = 20 Heisenberg VQE, HEA_ZYZ 10 layers, energy + gradient, 128 parameter sets per GPU
(1024 over 8 GPUs), parameter sets sharded over the ranks (dist.ShardedVQE; a 2^20 state is never split). WEAK scaling.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 profiles/bench_c5_dist.py

Timing: K steps through the public host API (numpy parameters in, numpy energies/gradients out, NCCL all-gather inside),
bracketed by barrier + torch.cuda.synchronize(), max over ranks."""
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

n, layers, per_gpu = 20, 10, 128
steps, warmup = 2, 1
indptr, indices, data = H.heisenberg_csr_fast(n)
c = H.hea_zyz_circuit(n, layers)
psi0 = np.zeros(1 << n, dtype=np.complex128)
psi0[0] = 1
params = H.random_params(c.get_Parameter_Num(), batch=per_gpu * world)
sv = sq.dist.ShardedVQE(psi0, c, indptr, indices, data, device=local)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(warmup):
    en, gr = sv.energy_grad(params)
barrier()
t0 = time.perf_counter()
for _ in range(steps):
    en, gr = sv.energy_grad(params)
barrier()
dt = (time.perf_counter() - t0) / steps
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "C5 n=20 Heisenberg VQE, HEA_ZYZ 10 layers, energy+gradient, parameter sets sharded", "n_gpus": world,
                      "sets_per_gpu": per_gpu, "scaling": "weak", "energy_grad_evals_per_s": per_gpu * world / float(t.item()),
                      "ms_per_step": float(t.item()) * 1e3, "energy0": float(en[0])}), flush=True)
if world > 1:
    dist.destroy_process_group()
