"""ncu driver for the VQE path: C5 (n = 20 Heisenberg, HEA_ZYZ 10 layers), energy + gradient for a few parameter sets.
usage: python profiles/prof_vqe.py [batch] [grad|energy]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import squander_b200 as sq

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "grad"
n, layers = 20, 10
indptr, indices, data = H.heisenberg_csr_fast(n)
c = H.hea_zyz_circuit(n, layers)
psi0 = np.zeros(1 << n, dtype=np.complex128)
psi0[0] = 1
e = sq.Engine(0)
e.upload_matrix(psi0)
e.set_circuit(c)
e.set_hamiltonian_csr(indptr, indices, data)
p = H.random_params(c.get_Parameter_Num(), batch=batch)
for _ in range(2):
    out = e.vqe_energy_grad_batched(p) if mode == "grad" else e.vqe_energy_batched(p)
print(mode, batch, e.last_kernel_time())
