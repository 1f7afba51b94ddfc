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


# ---- round 2: the benchmarked configurations (outputs of the reference's own code, tests/golden/make_golden_r2.py) ----------

def test_c3_benchmark_configuration_golden(port):
    """n = 10, L = 4 (550 gates, P = 1290), 8-column slice with trace_offset = 80: cost and all 1290 gradient entries of the
    C port against the reference's optimization_problem_combined (which takes the zgemm suffix route at this size)"""
    circ, Us, params, off, variants, cost, grad = G.c3_n10_slice()
    d, pool = circ.descriptors()
    for vi, v in enumerate(variants):
        f, g = port.cost_grad(d, circ.get_Parameter_Num(), params, Us, 10, v, off)
        assert abs(f - cost[vi]) <= 1e-12 * max(1.0, abs(cost[vi]))
        assert np.abs(g - grad[vi]).max() <= 1e-12 * max(1.0, np.abs(grad[vi]).max())


def test_c5_recipe_golden(port):
    """Heisenberg VQE (C5 recipe): energy + gradient at n = 10 and energy at n = 16 of the C port against
    Variational_Quantum_Eigensolver_Base; the sampled-gradient entry point equals the full one"""
    n, circ, p, (ip, ix, dat), e_ref, g_ref = G.c5_vqe("C5_n10_vqe")
    d, pool = circ.descriptors()
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[0] = 1.0
    e, g = port.vqe_energy_grad(d, circ.get_Parameter_Num(), p, psi0, ip, ix, dat)
    assert abs(e - e_ref) < 1e-12 and np.abs(g - g_ref).max() < 1e-12
    sample = [0, 3, 17, 100, len(p) - 1]
    e2, gs = port.vqe_energy_grad_sampled(d, p, psi0, ip, ix, dat, sample)
    assert e2 == e and np.abs(gs - g[sample]).max() < 1e-14
    n, circ, p, (ip, ix, dat), e_ref, _ = G.c5_vqe("C5_n16_vqe")
    psi0 = np.zeros(1 << n, dtype=np.complex128)
    psi0[0] = 1.0
    assert abs(port.vqe_energy(circ.descriptors()[0], p, psi0, ip, ix, dat) - e_ref) < 1e-12


def test_fast_heisenberg_builder_is_identical():
    for n, deg in ((4, 3), (7, 2), (10, 3)):
        a, b = H.heisenberg_csr(n, degree=deg), H.heisenberg_csr_fast(n, degree=deg)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


# ---- N2: the reference's binary gate-list format ------------------------------------------------------------------------------

def _binary_case(sq):
    c = H.adaptive_circuit(4, 2)
    extra = sq.Circuit(4)
    extra.add_H(0); extra.add_CNOT(1, 0); extra.add_RZ(2); extra.add_CZ(3, 2); extra.add_SX(1); extra.add_U2(3); extra.add_CRY(0, 3)
    c.add_Circuit(extra)
    c.add_X(2)
    return c, H.random_params(c.get_Parameter_Num(), seed=77)


def test_binary_gate_list_reader_on_reference_file(sq, tmp_path):
    """tests/golden/reference_export_n4.binary was written by the reference's own export_gate_list_to_binary
    (Gates_block.cpp:4807-4920) through oracle/_ref/libsqref_gpu.so: gate_io reads it back -- nested blocks, every gate class
    the format knows, parameters -- and writes the same bytes"""
    import os

    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_export_n4.binary")
    c, p = _binary_case(sq)
    c2, p2 = sq.gate_io.import_gate_list_from_binary(golden)
    assert (p2 == p).all()
    assert c2.descriptors(nested=True)[0].tobytes() == c.descriptors(nested=True)[0].tobytes()
    ours = tmp_path / "ours.binary"
    sq.gate_io.export_gate_list_to_binary(p, c, str(ours))
    assert ours.read_bytes() == open(golden, "rb").read()
    # truncated file, unknown gate tag, unsupported gate on export
    bad = tmp_path / "bad.binary"
    bad.write_bytes(open(golden, "rb").read()[:-9])
    with pytest.raises(Exception, match="Corrupted"):
        sq.gate_io.import_gate_list_from_binary(str(bad))
    import struct
    bad.write_bytes(struct.pack("<iiii", 2, 0, 1, 999))
    with pytest.raises(Exception, match="unimplemented"):
        sq.gate_io.import_gate_list_from_binary(str(bad))
    cc = sq.Circuit(3)
    cc.add_CCX(0, [1, 2])
    with pytest.raises(Exception, match="unimplemented"):
        sq.gate_io.export_gate_list_to_binary(np.zeros(0), cc, str(bad))


def test_binary_gate_list_matches_live_reference_export(sq, tmp_path):
    """where the reference-built checker is available: its exporter on a fresh random structure == ours, byte for byte"""
    import os
    import pyoracle

    if not pyoracle.RefGpu.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libsqref_gpu.so not built and /root/reference absent")
    rg = pyoracle.RefGpu()
    names = ["U1", "U2", "U3", "RX", "RY", "RZ", "CRY", "CNOT", "CZ", "CH", "SYC", "X", "Y", "Z", "H", "S", "Sdg", "SX", "adaptive"]
    for seed in (1, 2):
        c = H.random_circuit(5, 50, seed=seed, names=names, nested=True)
        p = H.random_params(c.get_Parameter_Num(), seed=seed)
        f_ref, f_ours = tmp_path / "ref.binary", tmp_path / "ours.binary"
        rg.export_binary(5, c.descriptors(nested=True)[0], p, str(f_ref))
        sq.gate_io.export_gate_list_to_binary(p, c, str(f_ours))
        assert f_ref.read_bytes() == f_ours.read_bytes()
        c2, p2 = sq.gate_io.import_gate_list_from_binary(str(f_ref))
        assert (p2 == p).all() and c2.descriptors()[0].tobytes() == c.descriptors()[0].tobytes()
