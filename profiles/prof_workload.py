"""Small driver for ncu captures: one cost+grad (or cost) batch of the C3 workload (n=10, L=4) through the C-ABI.
usage: python profiles/prof_workload.py [batch] [grad|cost] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import squander_b200 as sq

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "grad"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n, L = 10, 4
c = H.adaptive_circuit(n, L)
U = np.ascontiguousarray(H.random_unitary(1 << n).conj().T)
eng = sq.Engine(0)
eng.upload_matrix(U)
eng.set_circuit(c)
eng.set_cost(0, 0)
p = H.random_params(c.get_Parameter_Num(), batch=batch)
for _ in range(reps):
    out = eng.cost_grad_batched(p) if mode == "grad" else eng.cost_batched(p)
print(mode, batch, eng.last_kernel_time(), np.asarray(out[0] if mode == "grad" else out)[:2])
