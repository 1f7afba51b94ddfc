"""ncu driver for the raw dense 4-qubit DMMA path: C4 (n = 12, 64 GENERAL 4-qubit blocks + U3 layers), cost only, batch 8."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
import squander_b200 as sq

n, M, B = 12, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 8
rng = np.random.default_rng(7)
c = sq.Circuit(n)
for m in range(M):
    qs = sorted(int(q) for q in rng.choice(n, 4, replace=False))
    c.add_GENERAL(H.random_unitary(16, seed=1000 + m), qs)
    if m % 8 == 7:
        for q in range(n):
            c.add_U3(q)
e = sq.Engine(0)
e.upload_matrix(np.ascontiguousarray(H.random_unitary(1 << n).conj().T))
e.set_circuit(c)
e.set_cost(0)
p = H.random_params(c.get_Parameter_Num(), batch=B)
for _ in range(2):
    out = e.cost_batched(p)
print(B, e.last_kernel_time(), out[:2])
