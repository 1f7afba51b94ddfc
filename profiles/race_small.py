"""Small workloads for compute-sanitizer (racecheck / memcheck): the fused executor in gradient mode (bulk-async table ring,
mbarriers, W' reduction) and the windowed state-vector executor (forward + backward segments).
usage: compute-sanitizer --tool racecheck python profiles/race_small.py"""
import sys; sys.path[:0] = [".", "tests", "oracle"]
import numpy as np, helpers as H, squander_b200 as sq
n, L = 7, 1
c = H.adaptive_circuit(n, L); P = c.get_Parameter_Num()
U = np.ascontiguousarray(H.random_unitary(1 << n).conj().T)[:, :16].copy(); p = H.random_params(P, seed=1, batch=2)
e = sq.Engine(0); e.upload_matrix(U); e.set_circuit(c); e.set_cost(0, 0)
f1, g1 = e.cost_grad_batched(p)
print("grad", f1, e.last_kernel_time())
e.close()
nv = 8
vc = H.hea_zyz_circuit(nv, 2)
ip, ix, dv = H.heisenberg_csr_fast(nv)
psi0 = np.zeros(1 << nv, dtype=np.complex128); psi0[0] = 1
e = sq.Engine(0, options={"window": 5}); e.upload_matrix(psi0); e.set_circuit(vc); e.set_hamiltonian_csr(ip, ix, dv)
en, g = e.vqe_energy_grad_batched(H.random_params(vc.get_Parameter_Num(), seed=3, batch=3))
print("vqe", en, e.last_kernel_time())
