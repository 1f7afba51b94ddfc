"""Loader for tests/golden/golden_r1.npz (outputs of the reference's own code, see tests/golden/make_golden.py)."""
import os

import numpy as np

import helpers as H

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_r1.npz")
COST_CASES = ["C1_L1", "C1_L3", "C1_L5", "C2_19CNOT", "C3_n6", "OFFSET_n6", "MIXED_n5"]
MATRIX_CASES = ["C1_L1", "C2_19CNOT", "MIXED_n5"]


class Case:
    def __init__(self, z, name):
        g = lambda k: z[name + "/" + k]
        self.name = name
        self.descs = np.frombuffer(g("descs").tobytes(), dtype=H.abi.GATE_DESC_DTYPE).copy()
        self.pool = g("pool")
        self.n, self.P, self.trace_offset = (int(x) for x in g("meta"))
        self.params = g("params")
        for k in ("U", "prev", "variants", "cost", "grad", "applied", "deriv_idx", "deriv", "state_in", "state_out"):
            key = name + "/" + k
            setattr(self, k, z[key] if key in z.files else None)


def load(name):
    return Case(np.load(PATH), name)


PATH_R2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_r2.npz")


def c3_n10_slice():
    """BASELINE configs[2] exactly as bench.py builds it, restricted to 8 columns with trace_offset (inputs are regenerated from
    the seeds; outputs come from the reference's own code, tests/golden/make_golden_r2.py).
    Returns (circuit, U_slice, params, trace_offset, variants, cost[variant], grad[variant][P])."""
    z = np.load(PATH_R2)
    n, L, c0, ncols, P = (int(x) for x in z["C3_n10_cols8/meta"])
    circ = H.adaptive_circuit(n, L)
    assert circ.get_Parameter_Num() == P
    U = np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)
    Us = np.ascontiguousarray(U[:, c0:c0 + ncols])
    params = np.random.default_rng(42).random(P) * 2 * np.pi
    return circ, Us, params, c0, [int(v) for v in z["C3_n10_cols8/variants"]], z["C3_n10_cols8/cost"], z["C3_n10_cols8/grad"]


def c5_vqe(name):
    """(n, circuit, params, csr triple, energy, grad or None) of a C5-recipe case of golden_r2.npz"""
    z = np.load(PATH_R2)
    n, layers, P = (int(x) for x in z[name + "/meta"])
    circ = H.hea_zyz_circuit(n, layers)
    assert circ.get_Parameter_Num() == P
    params = np.random.default_rng(11).random(P) * 2 * np.pi
    grad = z[name + "/grad"] if (name + "/grad") in z.files else None
    return n, circ, params, H.heisenberg_csr_fast(n), float(z[name + "/energy"][0]), grad
