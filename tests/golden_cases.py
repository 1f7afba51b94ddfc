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
