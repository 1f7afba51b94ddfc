"""CPU: the C oracle (oracle/sq_oracle.c) against the committed golden fixtures -- outputs of the reference's own
code on data/Umtx.mat, data/19CNOT.qasm and the seeded recipes of its tests (tests/golden/make_golden.py).
This is the pin that travels: it does not need /root/reference or oracle/_ref."""
import numpy as np
import pytest

import golden_cases as G
import helpers as H


@pytest.mark.parametrize("name", G.COST_CASES)
def test_cost_and_gradient_golden(port, name):
    c = G.load(name)
    prev = float(c.prev[0])
    for vi, v in enumerate(c.variants):
        for pi, p in enumerate(c.params):
            f, g = port.cost_grad(c.descs, c.P, p, c.U, c.n, int(v), c.trace_offset, prev, pool=c.pool)
            assert abs(f - c.cost[vi, pi]) < 1e-13
            assert np.abs(g - c.grad[vi, pi]).max() < 1e-13
            assert abs(port.cost(c.descs, p, c.U, c.n, int(v), c.trace_offset, prev, pool=c.pool) - c.cost[vi, pi]) < 1e-13


@pytest.mark.parametrize("name", G.MATRIX_CASES)
def test_matrices_golden(port, name):
    c = G.load(name)
    assert np.abs(port.apply_circuit(c.descs, c.params[0], c.U, c.pool) - c.applied).max() < 1e-13
    d = port.apply_derivate(c.descs, c.P, c.params[0], c.U, c.pool)
    assert np.abs(d[c.deriv_idx] - c.deriv).max() < 1e-13


def test_general_blocks_golden(port):
    c = G.load("GENERAL_n5")
    assert np.abs(port.apply_circuit(c.descs, c.params, c.state_in, c.pool) - c.state_out).max() < 1e-13
    assert np.abs(port.apply_circuit(c.descs, c.params, c.U, c.pool) - c.applied).max() < 1e-13


def test_known_answers():
    """known-answer facts of the fixtures themselves"""
    c = G.load("OFFSET_n6")
    assert np.abs(c.cost[:, 0]).max() < 1e-8  # identity-cost KAT with trace_offset (reference test :123-184)
    c2 = G.load("C2_19CNOT")
    assert len(c2.descs) == 109 and c2.P == 172  # data/19CNOT.qasm: 41 u, 30 rx, 4 ry, 15 rz, 15 cz, 4 cx
    t = c2.descs["type"]
    assert [(t == x).sum() for x in (H.abi.U3, H.abi.RX, H.abi.RY, H.abi.RZ, H.abi.CZ, H.abi.CNOT)] == [41, 30, 4, 15, 15, 4]


def test_qasm_importer_matches_fixture(sq):
    """the Qiskit-free importer reproduces the descriptor stream stored in the fixture from an equivalent source"""
    src = 'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[3];\nu(0.5,pi/2,-pi/4) q[1];\nrx(-pi/2) q[0];\ncz q[2],q[1];\ncx q[0],q[2];\nrz(0.25) q[2];\n'
    c, p = sq.qasm.loads(src)
    d, _ = c.descriptors()
    assert list(d["type"]) == [H.abi.U3, H.abi.RX, H.abi.CZ, H.abi.CNOT, H.abi.RZ]
    assert list(d["target"]) == [1, 0, 1, 2, 2] and list(d["control"]) == [-1, -1, 2, 0, -1]
    assert np.allclose(p, [0.25, np.pi / 2, -np.pi / 4, -np.pi / 4, 0.125])
    with pytest.raises(ValueError):
        sq.qasm.loads("qreg q[1];\nmeasure q[0];")
    with pytest.raises(ValueError):
        sq.qasm.loads("qreg q[1];\nrx(__import__('os')) q[0];")
