"""Host-side logic above the C-ABI that needs no GPU: the CRY -> CZ / CNOT finalisation of the adaptive decomposition
(N_Qubit_Decomposition_adaptive::replace_trivial_CRY_gates, decomposition/N_Qubit_Decomposition_adaptive.cpp:1398-1590),
checked against the oracle's matrices."""
import numpy as np
import pytest

import helpers as H
import pyoracle

sq = H.sq


@pytest.fixture(scope="module")
def port():
    return pyoracle.Port()


def circuit_matrix(port, c, params):
    d, pool = c.descriptors()
    eye = np.eye(1 << c.qbit_num, dtype=np.complex128)
    return port.apply_circuit(d, np.asarray(params, dtype=np.float64), eye, pool)


def test_replace_trivial_cry_gates_keeps_the_unitary(port):
    """half turns become RX CZ RX RZ + a global phase, identities are dropped, everything else becomes RY CNOT RY CNOT: the
    rewritten circuit times the returned phase is the original unitary, gate by gate conventions included"""
    from importlib import import_module

    dec = import_module(sq.__name__ + ".decomposition") if hasattr(sq, "__name__") else None
    replace = dec.replace_trivial_CRY_gates
    n = 4
    c = H.adaptive_circuit(n, 2)
    P = c.get_Parameter_Num()
    rng = np.random.default_rng(5)
    x = rng.random(P) * 2 * np.pi
    # the adaptive parameters in application order: every block [U3, U3, adaptive] has it at offset 6
    d, _ = c.descriptors()
    ada = [int(r["param_start"]) for r in d if int(r["type"]) == sq.abi.ADAPTIVE]
    assert len(ada) == 12
    x[ada[0]] = np.pi / 2          # half turn, sin > 0
    x[ada[1]] = -np.pi / 2         # half turn, sin < 0
    x[ada[2]] = 0.0                # identity
    x[ada[3]] = 2 * np.pi          # identity (cos = 1)
    x[ada[4]] = np.pi / 2 + 2e-4   # inside the reference's tolerance: still a half turn
    x[ada[5]] = 3 * np.pi / 2      # half turn, sin < 0
    c2, x2, phase = replace(c, x)
    assert abs(abs(phase) - 1) < 1e-15
    types = [int(r["type"]) for r in c2.descriptors()[0]]
    assert sq.abi.ADAPTIVE not in types
    assert types.count(sq.abi.CZ) == 4 and types.count(sq.abi.CNOT) == 2 * 6
    assert x2.size == c2.get_Parameter_Num() == P - 12 + 4 * 3 + 6 * 2
    exact = x.copy()
    exact[ada[4]] = np.pi / 2  # the rewritten circuit IS the exact half turn there
    M_ref = circuit_matrix(port, c, exact)
    M_new = phase * circuit_matrix(port, c2, x2)
    assert np.abs(M_new - M_ref).max() < 1e-12
    # and within the reference's tolerance of the original parameters
    assert np.abs(M_new - circuit_matrix(port, c, x)).max() < 1e-3
    with pytest.raises(Exception):
        replace(c, x[:-1])
    flat = c.get_Flat_Circuit()
    with pytest.raises(Exception):
        replace(flat, x)  # only block gates are accepted, as in the reference
