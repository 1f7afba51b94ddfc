"""Multi-GPU check, launched by torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_dist_gpu.py

Every rank computes cost+gradient through dist.ShardedCost in both sharding modes and compares with the CPU ORACLE
(oracle/sq_oracle.c, pinned to the reference) at the stated 1e-10, and with a single-GPU evaluation of the full problem on
its own device (bit-for-bit in batch mode, 1e-12 in column mode where the summation order of the trace differs).
tests/test_gpu_parity.py::test_multi_gpu_sharding_matches_oracle spawns this script when two devices are visible."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import helpers as H
import pyoracle
import squander_b200 as sq

port = pyoracle.Port()

rank = int(os.environ["RANK"])
local = int(os.environ.get("LOCAL_RANK", rank))
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 6
circ = H.adaptive_circuit(n, 2)
P = circ.get_Parameter_Num()
descs, pool = circ.descriptors()
U = H.random_unitary(1 << n).conj().T.copy()
params = H.random_params(P, batch=11)


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max()))


ref = sq.Engine(local)
ref.upload_matrix(U)
ref.set_circuit(circ)
ok = True
for variant in (0, 2, 3, 4, 5, 6, 9):
    ref.set_cost(variant, 0, 0.37)
    c_ref, g_ref = ref.cost_grad_batched(params)
    oracle = [port.cost_grad(descs, P, params[b], U, n, variant, 0, 0.37) for b in range(len(params))]
    c_orc = np.array([o[0] for o in oracle])
    g_orc = np.array([o[1] for o in oracle])
    for mode in ("batch", "columns"):
        sc = sq.dist.ShardedCost(U, circ, variant=variant, mode=mode, prev_cost=0.37, device=local)
        c, g = sc.cost_grad(params)
        c2 = sc.cost(params)
        tol = 0.0 if mode == "batch" else 1e-12
        err = max(np.abs(c - c_ref).max(), np.abs(g - g_ref).max(), np.abs(c2 - c_ref).max())
        err_orc = max(rel_err(c, c_orc), rel_err(g, g_orc))
        good = err <= tol and err_orc <= 1e-10
        ok = ok and good
        if rank == 0:
            print("variant %d mode %-7s world %d vs single GPU %.3e, vs oracle (rel) %.3e %s" % (variant, mode, world, err, err_orc, "ok" if good else "FAIL"), flush=True)
        sc.close()
# VQE: parameter sets sharded over the ranks (the partial-sum order may depend on the slice size: 1e-12)
nv = 10
vc = H.hea_zyz_circuit(nv, 2)
psi0 = np.zeros(1 << nv, dtype=np.complex128)
psi0[0] = 1.0
ip, ix, dv = H.heisenberg_csr(nv)
vp = H.random_params(vc.get_Parameter_Num(), seed=3, batch=13)
single = sq.Engine(local)
single.upload_matrix(psi0)
single.set_circuit(vc)
single.set_hamiltonian_csr(ip, ix, dv)
e_ref, g_ref = single.vqe_energy_grad_batched(vp)
sv = sq.dist.ShardedVQE(psi0, vc, ip, ix, dv, device=local)
e_sh, g_sh = sv.energy_grad(vp)
err = max(np.abs(e_sh - e_ref).max(), np.abs(g_sh - g_ref).max(), np.abs(sv.energy(vp) - e_ref).max())
vd = vc.descriptors()[0]
vo = [port.vqe_energy_grad(vd, vc.get_Parameter_Num(), vp[b], psi0, ip, ix, dv) for b in (0, len(vp) - 1)]
err_orc = max(rel_err(e_sh[[0, -1]], [o[0] for o in vo]), rel_err(g_sh[[0, -1]], [o[1] for o in vo]))
ok = ok and err <= 1e-12 and err_orc <= 1e-10
if rank == 0:
    print("VQE batch sharding world %d vs single GPU %.3e, vs oracle (rel) %.3e %s" % (world, err, err_orc, "ok" if (err <= 1e-12 and err_orc <= 1e-10) else "FAIL"), flush=True)
sv.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
if rank == 0:
    print("DIST_GPU_OK" if t.item() == 1 else "DIST_GPU_FAIL", flush=True)
sys.exit(0 if t.item() == 1 else 1)
