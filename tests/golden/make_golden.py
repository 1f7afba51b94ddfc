"""Generates tests/golden/golden_r1.npz from the REFERENCE'S OWN CODE (oracle/_ref/libsqref.so, built from
/root/reference by oracle/Makefile). Run in the container that has /root/reference:

    python tests/golden/make_golden.py

The reference ships no stored output vectors for this path, so these fixtures are outputs of the
reference itself on its own fixed inputs (data/Umtx.mat, data/19CNOT.qasm) and on the seeded recipes of its tests.
Every case stores its inputs too, so nothing under /root/reference is needed to replay it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import helpers as H
import pyoracle
import squander_b200 as sq
from scipy.io import loadmat

REF_DATA = "/root/reference/data"
ref = pyoracle.Ref()
out = {}


def add_case(name, circ, U, params_list, variants, trace_offset=0, prev=1.0, n=None, with_matrices=False):
    n = n or circ.qbit_num
    d_nested, pool = circ.descriptors(nested=True)
    d_flat, pool_f = circ.descriptors()
    out[name + "/descs"] = d_flat.view(np.uint8)
    out[name + "/pool"] = pool_f
    out[name + "/U"] = U
    out[name + "/params"] = np.array(params_list)
    out[name + "/meta"] = np.array([n, circ.get_Parameter_Num(), trace_offset], dtype=np.int64)
    out[name + "/prev"] = np.array([prev])
    out[name + "/variants"] = np.array(variants, dtype=np.int64)
    dec = ref.decomp(U, n, d_nested, pool)
    dec.set_parallel(0)
    costs = np.zeros((len(variants), len(params_list)))
    grads = np.zeros((len(variants), len(params_list), circ.get_Parameter_Num()))
    for vi, v in enumerate(variants):
        dec.set_cost(v, trace_offset, prev, 1 / 1.7, 0.5)
        for pi, p in enumerate(params_list):
            f, g = dec.cost_grad(p)
            assert abs(dec.cost(p) - f) < 1e-13
            costs[vi, pi] = f
            grads[vi, pi] = g
    out[name + "/cost"] = costs
    out[name + "/grad"] = grads
    if with_matrices:
        rc = ref.circuit(n, d_nested, pool)
        out[name + "/applied"] = rc.apply(params_list[0], U)
        dm = rc.apply_derivate(params_list[0], U)
        idx = np.linspace(0, circ.get_Parameter_Num() - 1, 5).astype(int)
        out[name + "/deriv_idx"] = idx
        out[name + "/deriv"] = dm[idx]
    print(name, "P =", circ.get_Parameter_Num(), "cost[0] =", costs[:, 0])


# C1: data/Umtx.mat, adaptive L = 1..5 (BASELINE configs[0]); the examples pass Umtx.conj().T (example.py:62)
Umtx = np.ascontiguousarray(loadmat(os.path.join(REF_DATA, "Umtx.mat"))["Umtx"].astype(np.complex128))
for L in (1, 3, 5):
    c = H.adaptive_circuit(4, L)
    ps = [H.random_params(c.get_Parameter_Num(), seed=42), H.random_params(c.get_Parameter_Num(), seed=43) * 0.1]
    add_case("C1_L%d" % L, c, np.ascontiguousarray(Umtx.conj().T), ps, [0, 3, 9], with_matrices=(L == 1))

# C2: data/19CNOT.qasm re-optimisation (BASELINE configs[1]); target = Pauli exponent at alpha = 1.8236...
# (tests/decomposition/test_parametric_circuit.py:52-115, 247), cost variant 3 (:182)
circ2, p_qasm = sq.qasm.load(os.path.join(REF_DATA, "19CNOT.qasm"))
tc, tp = H.pauli_exponent_circuit(1.823631161607293)
target = ref.circuit(5, tc.descriptors(nested=True)[0]).apply(tp, np.eye(32, dtype=np.complex128))
rng = np.random.default_rng(7)
ps2 = [p_qasm, p_qasm + 0.05 * rng.standard_normal(p_qasm.size), rng.random(p_qasm.size) * 2 * np.pi]
add_case("C2_19CNOT", circ2, np.ascontiguousarray(target.conj().T), ps2, [3, 0, 4], prev=0.25, with_matrices=True)
out["C2_19CNOT/qasm_params"] = p_qasm

# C3-small: the C3 structure at n = 6 (same generator, seeds of SURVEY.md §8d), all cost variants
c3 = H.adaptive_circuit(6, 2)
U3 = np.ascontiguousarray(H.random_unitary(64, seed=123).conj().T)
add_case("C3_n6", c3, U3, [H.random_params(c3.get_Parameter_Num(), seed=42)], [0, 1, 2, 3, 4, 5, 6, 9], prev=0.37)

# trace offset / rectangular U (tests/decomposition/test_optmization_problem_combined.py:123-184 at n = 6)
c4 = H.adaptive_circuit(6, 1)
p4 = H.random_params(c4.get_Parameter_Num(), seed=3)
full = ref.circuit(6, c4.descriptors(nested=True)[0]).apply(p4, np.eye(64, dtype=np.complex128))
Urect = np.ascontiguousarray(full[17:40, :].conj().T)
add_case("OFFSET_n6", c4, Urect, [p4, H.random_params(c4.get_Parameter_Num(), seed=4)], [0, 1, 2], trace_offset=17)

# every gate family in one circuit: cost/gradient through the decomposition object (no GENERAL gates there: the
# reference's Gate::clone drops the target qubits of a GENERAL gate, so set_custom_gate_structure cannot carry them)
c5 = H.random_circuit(5, 60, seed=11)
add_case("MIXED_n5", c5, H.random_unitary(32, seed=5), [H.random_params(c5.get_Parameter_Num(), seed=12)], [0, 3, 6], with_matrices=True)
assert {int(t) for t in c5.descriptors()[0]["type"]} >= {sq.abi.CROT, sq.abi.SYC}  # the two-branch and Sycamore kernels are in

# GENERAL 2/3-qubit blocks mixed with every gate family through Circuit.apply_to, matrix and state-vector input
c6 = H.random_circuit(5, 60, seed=13, general_k=(2, 3))
d6n, pool6 = c6.descriptors(nested=True)
d6, pool6f = c6.descriptors()
rc6 = ref.circuit(5, d6n, pool6)
p6 = H.random_params(c6.get_Parameter_Num(), seed=14)
psi = H.random_state(32)
U6 = H.random_unitary(32, seed=6)[:, :9].copy()
out["GENERAL_n5/descs"] = d6.view(np.uint8)
out["GENERAL_n5/pool"] = pool6f
out["GENERAL_n5/params"] = p6
out["GENERAL_n5/meta"] = np.array([5, c6.get_Parameter_Num(), 0], dtype=np.int64)
out["GENERAL_n5/state_in"] = psi
out["GENERAL_n5/state_out"] = rc6.apply(p6, psi)
out["GENERAL_n5/U"] = U6
out["GENERAL_n5/applied"] = rc6.apply(p6, U6)

np.savez_compressed(os.path.join(HERE, "golden_r1.npz"), **out)
print("wrote", os.path.join(HERE, "golden_r1.npz"), os.path.getsize(os.path.join(HERE, "golden_r1.npz")), "bytes")
