"""C5 (n = 20 Heisenberg VQE, HEA_ZYZ 10 layers) through the windowed executor: device time of the forward and backward
segment sweeps per slice of parameter sets, and energy / energy+gradient evaluations per second.
usage: python profiles/bench_vqe_window.py [batch] [reps] [name=value ...engine options]   -> one JSON line"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import squander_b200 as sq

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opts = dict((a.split("=")[0], int(a.split("=")[1])) for a in sys.argv[3:])
n, layers = 20, 10
indptr, indices, data = H.heisenberg_csr_fast(n)
c = H.hea_zyz_circuit(n, layers)
psi0 = np.zeros(1 << n, dtype=np.complex128)
psi0[0] = 1
e = sq.Engine(0, options=opts)
e.upload_matrix(psi0)
e.set_circuit(c)
e.set_hamiltonian_csr(indptr, indices, data)
p = H.random_params(c.get_Parameter_Num(), batch=batch)
out = {"workload": "C5 n=20 HEA_ZYZ x10 (P=%d), %d parameter sets" % (c.get_Parameter_Num(), batch), "options": opts}
E = e.vqe_energy_batched(p)
t0 = time.perf_counter()
for _ in range(reps):
    E = e.vqe_energy_batched(p)
out["energy_evals_per_s"] = batch * reps / (time.perf_counter() - t0)
Eg, g = e.vqe_energy_grad_batched(p)
e.kernel_time("fused_exec<WINDOW_FWD>")
e.kernel_time("fused_exec<WINDOW_BWD>")
t0 = time.perf_counter()
for _ in range(reps):
    Eg, g = e.vqe_energy_grad_batched(p)
out["energy_grad_evals_per_s"] = batch * reps / (time.perf_counter() - t0)
out["fwd_ms_per_slice"] = e.kernel_time("fused_exec<WINDOW_FWD>")
out["bwd_ms_per_slice"] = e.kernel_time("fused_exec<WINDOW_BWD>")
out["energy0"] = float(Eg[0])
out["grad_checksum"] = float(np.abs(g).sum())
print(json.dumps(out))
