"""Generates tests/golden/golden_r2.npz from the REFERENCE'S OWN CODE (oracle/_ref/libsqref.so): the benchmarked
configurations, which round 1 only pinned through identities and finite differences.

    python tests/golden/make_golden_r2.py        (in the container that has /root/reference; ~2 minutes)

  C3_n10_cols8   BASELINE configs[2] as bench.py builds it -- n = 10, adaptive L = 4 (550 gates, P = 1290), Haar U (seed 123)
                 conjugate-transposed -- restricted to the 8 columns [80, 88) with trace_offset = 80 (the rectangular-Umtx +
                 trace-offset semantics of the reference, tests/decomposition/test_optmization_problem_combined.py:156-170),
                 cost and all 1290 gradient entries for variants 0 and 3, parameters default_rng(42).random(P) * 2 pi.
                 Inputs are regenerated from the seeds by the test (helpers), only outputs are stored.
  C5_n10_vqe     the C5 recipe (Heisenberg on a random 3-regular graph, seed 31415, HEA_ZYZ, |0..0>) at n = 10, 3 layers: energy
                 and gradient from Variational_Quantum_Eigensolver_Base itself. Its gradient materialises dense 2^n x 2^n
                 suffix products from n = 7 on (Gates_block.cpp:358-428, should_use_suffix), which does not fit host memory
                 at n >= 14; the n = 20 check of the GPU tests therefore uses the C port, which this case and C5_n16 pin.
  C5_n16_vqe     the same recipe at n = 16, 10 layers: energy only (forward pass) from the reference class.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import helpers as H
import pyoracle

ref = pyoracle.Ref()
out = {}

# ---- C3 at the benchmarked size, column slice ------------------------------------------------------------------------
n, L, c0, ncols = 10, 4, 80, 8
circ = H.adaptive_circuit(n, L)
P = circ.get_Parameter_Num()
U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)
Us = np.ascontiguousarray(U[:, c0:c0 + ncols])
params = np.random.default_rng(42).random(P) * 2 * np.pi  # row 0 of bench.py's batch
d_nested, pool = circ.descriptors(nested=True)
dec = ref.decomp(Us, n, d_nested, pool)
dec.set_parallel(2)
costs, grads = [], []
for v in (0, 3):
    dec.set_cost(v, c0, 1.0, 1 / 1.7, 0.5)
    t0 = time.time()
    f, g = dec.cost_grad(params)
    print("C3_n10_cols8 variant", v, "cost", f, "|grad|max", np.abs(g).max(), "%.1f s" % (time.time() - t0))
    costs.append(f)
    grads.append(g)
out["C3_n10_cols8/meta"] = np.array([n, L, c0, ncols, P], dtype=np.int64)
out["C3_n10_cols8/variants"] = np.array([0, 3], dtype=np.int64)
out["C3_n10_cols8/cost"] = np.array(costs)
out["C3_n10_cols8/grad"] = np.array(grads)

# ---- C5 recipe at n = 10 (energy + gradient) and n = 16 (energy) --------------------------------------------------------
for name, n5, layers, with_grad in (("C5_n10_vqe", 10, 3, True), ("C5_n16_vqe", 16, 10, False)):
    indptr, indices, data = H.heisenberg_csr(n5)
    vq = ref.vqe(n5, indptr, indices, data, ansatz="HEA_ZYZ", layers=layers, inner_blocks=1)
    P5 = vq.n_params
    assert P5 == H.hea_zyz_circuit(n5, layers).get_Parameter_Num()
    p5 = np.random.default_rng(11).random(P5) * 2 * np.pi
    t0 = time.time()
    if with_grad:
        e5, g5 = vq.energy_grad(p5)
        out[name + "/grad"] = g5
    else:
        e5 = vq.energy(p5)
    print(name, "energy", e5, "%.1f s" % (time.time() - t0))
    out[name + "/meta"] = np.array([n5, layers, P5], dtype=np.int64)
    out[name + "/energy"] = np.array([e5])
    del vq

np.savez_compressed(os.path.join(HERE, "golden_r2.npz"), **out)
print("wrote golden_r2.npz", os.path.getsize(os.path.join(HERE, "golden_r2.npz")), "bytes")
