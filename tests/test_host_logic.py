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


def test_second_renyi_entropy_known_answers():
    """second_renyi_entropy (Gates_block::get_second_Renyi_entropy, Gates_block.cpp:3625-3650): product state -> 0, a Bell pair cut
    in the middle -> log 2, GHZ -> log 2 for every proper subset, a subset and its complement agree, and the value equals
    -log Tr rho_A^2 of an explicitly traced-out density matrix for a random state"""
    from squander_b200 import circuit as C

    n = 4
    prod = np.zeros(1 << n, dtype=np.complex128)
    prod[5] = 1.0
    assert abs(C.second_renyi_entropy(prod, n, [0, 2])) < 1e-14
    bell = np.zeros(4, dtype=np.complex128)
    bell[0] = bell[3] = 1 / np.sqrt(2)
    assert abs(C.second_renyi_entropy(bell, 2, [0]) - np.log(2)) < 1e-14
    assert abs(C.second_renyi_entropy(bell, 2, [0, 1])) < 1e-14
    ghz = np.zeros(1 << n, dtype=np.complex128)
    ghz[0] = ghz[-1] = 1 / np.sqrt(2)
    for sub in ([0], [1, 3], [0, 1, 2]):
        assert abs(C.second_renyi_entropy(ghz, n, sub) - np.log(2)) < 1e-14
    rng = np.random.default_rng(3)
    n = 5
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    sub = [1, 4]
    # explicit partial trace: rho_A[a, a'] = sum_r psi[idx(a, r)] conj(psi[idx(a', r)]), qubit q = bit q of the index
    rest = [q for q in range(n) if q not in sub]
    def idx(a, r):
        i = 0
        for j, q in enumerate(sub):
            i |= ((a >> j) & 1) << q
        for j, q in enumerate(rest):
            i |= ((r >> j) & 1) << q
        return i
    rho = np.array([[sum(psi[idx(a, r)] * np.conj(psi[idx(b, r)]) for r in range(1 << len(rest))) for b in range(4)] for a in range(4)])
    want = -np.log(np.real(np.trace(rho @ rho)))
    assert abs(C.second_renyi_entropy(psi, n, sub) - want) < 1e-12
    assert abs(C.second_renyi_entropy(psi, n, rest) - want) < 1e-12


def test_wrapper_data_format_methods(tmp_path, capsys):
    """the data-format methods of the decomposition wrapper either side of the cost path (qgd_N_Qubit_Decompositions_Wrapper.cpp:
    3125-3248), none of which needs a device: unitary binary files (Decomposition_Base.cpp:1128-1177: int32 rows, int32 cols,
    complex128 data), gate structures from binary gate lists (set_ / add_Gate_Structure_From_Binary), project name prefix,
    gate listing"""
    import struct

    import helpers as H

    sq = H.sq
    n = 3
    U = H.random_unitary(1 << n, seed=4)
    dec = sq.N_Qubit_Decomposition_adaptive(U, level_limit_max=2, level_limit_min=1)
    fn = str(tmp_path / "umtx.binary")
    dec.export_Unitary(fn)
    raw = open(fn, "rb").read()
    assert struct.unpack("ii", raw[:8]) == (8, 8) and len(raw) == 8 + 64 * 16
    assert np.array_equal(np.frombuffer(raw[8:], dtype=np.complex128).reshape(8, 8), U)
    dec2 = sq.N_Qubit_Decomposition_adaptive(np.eye(8, dtype=np.complex128), level_limit_max=2, level_limit_min=1)
    dec2.set_Unitary_From_Binary(fn)
    assert np.array_equal(dec2.get_Unitary(), U)
    with open(fn, "wb") as f:
        f.write(raw[:-16])
    with pytest.raises(Exception):
        dec2.set_Unitary_From_Binary(fn)
    with pytest.raises(Exception):
        dec2.set_Unitary(np.eye(4))
    # gate structures through the binary gate-list format
    c = H.adaptive_circuit(n, 1)
    P = c.get_Parameter_Num()
    x = H.random_params(P, seed=2)
    gl = str(tmp_path / "circuit.binary")
    sq.gate_io.export_gate_list_to_binary(x, c, gl)
    dec.set_Gate_Structure_From_Binary(gl)
    assert dec.get_Parameter_Num() == P and np.array_equal(dec.get_Optimized_Parameters(), x)
    assert dec.get_Gate_Num() == c.get_Gate_Num()
    d0 = dec.get_Circuit().descriptors()[0]
    dec.add_Gate_Structure_From_Binary(gl)
    assert dec.get_Parameter_Num() == 2 * P and np.array_equal(dec.get_Optimized_Parameters(), np.concatenate([x, x]))
    d1 = dec.get_Circuit().descriptors()[0]
    flat = lambda d: [(int(r["type"]), int(r["target"]), int(r["control"])) for r in d if int(r["type"]) not in (sq.abi.BLOCK_BEGIN, sq.abi.BLOCK_END)]
    assert flat(d1) == flat(d0) + flat(d0)
    # project name prefixes the file names, as in the reference
    dec.set_Project_Name(str(tmp_path / "proj"))
    assert dec.get_Project_Name().endswith("proj")
    dec.export_Unitary("u.binary")
    assert (tmp_path / "proj_u.binary").exists()
    # OpenQASM out of the wrapper (adaptive gates written as the CRY they are); a Qiskit object only where Qiskit exists
    dec.set_Optimized_Parameters(np.concatenate([x, x]))
    c_back, x_back = sq.qasm.loads(dec.get_QASM())
    assert np.array_equal(x_back, np.concatenate([x, x])) and c_back.get_Gate_Num() == len(flat(d1)) and "Adaptive" not in c_back.get_Gate_Nums()
    try:
        import qiskit  # noqa: F401
    except ImportError:
        with pytest.raises(Exception, match="Qiskit is not installed"):
            dec.get_Qiskit_Circuit()
    dec.set_Max_Iterations(17)
    assert dec.config["max_inner_iterations"] == 17
    dec.set_Verbose(0)
    dec.set_Debugfile("x.log")
    dec.List_Gates()
    out = capsys.readouterr().out.splitlines()
    assert len(out) == len(d1) and "U3" in out[0] + out[1] + out[2]


def test_circuit_queries_and_remap():
    """host-side queries of the circuit wrapper (qgd_Circuit_Wrapper.cpp:3109-3302): gate counts, involved qubits, remapping onto
    another register -- the remapped structure equals the one built directly on the new qubits"""
    import helpers as H

    sq = H.sq
    c = sq.Circuit(4)
    c.add_U3(0)
    c.add_CNOT(1, 0)
    inner = sq.Circuit(4)
    inner.add_RY(1)
    inner.add_CRY(0, 1)
    c.add_Circuit(inner)
    c.add_U3(1)
    assert c.get_Gate_Nums() == {"U3": 2, "CNOT": 1, "RY": 1, "CRY": 1}
    assert c.get_Qbits() == [0, 1] and c.get_Gate_Num() == 4 and len(c.get_Gates()) == 4 and c.get_Gate(2) is inner
    r = c.Remap_Qbits({0: 3, 1: 2}, 5)
    want = sq.Circuit(5)
    want.add_U3(3)
    want.add_CNOT(2, 3)
    wi = sq.Circuit(5)
    wi.add_RY(2)
    wi.add_CRY(3, 2)
    want.add_Circuit(wi)
    want.add_U3(2)
    assert r.get_Qbit_Num() == 5 and r.get_Qbits() == [2, 3] and r.get_Parameter_Num() == c.get_Parameter_Num()
    key = lambda circ: [tuple(int(x[f]) for f in ("type", "target", "control", "param_start", "n_params")) for x in circ.descriptors(nested=True)[0]]
    assert key(r) == key(want)
    # dependency queries (Gates_block::determine_parents / determine_children): U3(0) -> CNOT(1,0) -> [RY(1) CRY(0,1)] -> U3(1)
    assert c.get_Parents(0) == [] and c.get_Children(0) == [1]
    assert c.get_Parents(1) == [0] and c.get_Children(1) == [2]
    assert c.get_Parents(2) == [1] and c.get_Children(2) == [3] and c.get_Parents(inner) == [1]
    assert c.get_Parents(3) == [2] and c.get_Children(3) == []
    wide = sq.Circuit(4)
    wide.add_H(0)
    wide.add_H(2)
    wide.add_CNOT(2, 0)
    wide.add_X(3)
    assert wide.get_Parents(2) == [0, 1] and wide.get_Children(0) == [2] and wide.get_Parents(3) == [] and wide.get_Children(2) == []
    x = np.arange(c.get_Parameter_Num(), dtype=np.float64)
    assert c.get_Parameter_Start_Index(2) == 3 and list(c.Extract_Parameters(x, 2)) == [3.0, 4.0] and list(c.Extract_Parameters(x, 3)) == [5.0, 6.0, 7.0]
    assert list(c.Extract_Parameters(x, 1)) == [] and np.array_equal(c.Extract_Parameters(x), x)
    with pytest.raises(Exception):
        c.Extract_Parameters(x[:-1])
    import pickle

    c2 = pickle.loads(pickle.dumps(c))
    assert key(c2) == key(c) and c2.get_Gate_Nums() == c.get_Gate_Nums() and c2._engine is None
    c2.set_Qbit_Num(6)
    assert c2.get_Qbit_Num() == 6 and c2.get_Gate(2).get_Qbit_Num() == 6 and key(c2) == key(c)
    with pytest.raises(Exception):
        c2.set_Qbit_Num(1)
    with pytest.raises(Exception):
        c.Remap_Qbits({0: 1})  # target == control after the map
    with pytest.raises(Exception):
        c.Remap_Qbits({0: 7})


def test_import_qiskit_circuit_without_qiskit(tmp_path, port):
    """import_Qiskit_Circuit of the decomposition wrapper over the Qiskit-free QASM importer: source text, a file path and an
    object with qasm() give the same structure and parameters (theta / 2 convention of Qiskit_IO.py:330-461), and the oracle's
    matrix of that structure is the circuit's unitary (checked on a Bell-pair preparation)"""
    import helpers as H

    sq = H.sq
    src = 'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[2];\nh q[0];\ncx q[0],q[1];\nry(0.6) q[1];\nu3(0.2,0.4,-0.3) q[0];\n'
    fn = tmp_path / "c.qasm"
    fn.write_text(src)

    class Qc:
        def qasm(self):
            return src

    got = []
    for arg in (src, str(fn), Qc()):
        dec = sq.N_Qubit_Decomposition_custom(np.eye(4, dtype=np.complex128))
        dec.import_Qiskit_Circuit(arg)
        d, pool = dec.get_Circuit().descriptors()
        got.append(([(int(r["type"]), int(r["target"]), int(r["control"])) for r in d], dec.get_Optimized_Parameters()))
    assert got[0][0] == got[1][0] == got[2][0] == [(sq.abi.H, 0, -1), (sq.abi.CNOT, 1, 0), (sq.abi.RY, 1, -1), (sq.abi.U3, 0, -1)]
    assert all(np.array_equal(g[1], got[0][1]) for g in got) and np.allclose(got[0][1], [0.3, 0.1, 0.4, -0.3])
    psi = np.zeros(4, dtype=np.complex128)
    psi[0] = 1
    bell = port.apply_circuit(d[:2], [], psi, pool)
    assert np.allclose(np.abs(bell) ** 2, [0.5, 0, 0, 0.5])
    with pytest.raises(Exception):
        sq.N_Qubit_Decomposition_custom(np.eye(8, dtype=np.complex128)).import_Qiskit_Circuit(src)
    with pytest.raises(Exception):
        dec.import_Qiskit_Circuit(42)


def test_qasm_export_roundtrip(port):
    """qasm.dumps is the inverse of qasm.loads on every gate family of qelib1 (angle doubling, operand order of controlled,
    two-target and three-qubit gates): the structure and the parameters read back unchanged, and the oracle's matrices of the
    original and of the re-imported circuit are the same"""
    import helpers as H

    sq = H.sq
    names = [x for x in H.ONE_Q + H.CTRL + H.TWO_T + ["CCX", "CSWAP"] if x not in ("CROT", "CR", "SYC", "adaptive")]
    c = H.random_circuit(5, 120, seed=8, names=names, nested=True)
    x = H.random_params(c.get_Parameter_Num(), seed=4)
    text = sq.qasm.dumps(c, x)
    c2, x2 = sq.qasm.loads(text)
    key = lambda circ: [tuple(int(r[f]) for f in ("type", "target", "control", "target2", "control2", "param_start", "n_params")) for r in circ.descriptors()[0]]
    assert key(c2) == key(c.get_Flat_Circuit()) and np.array_equal(x2, x) and len(c.get_Gate_Nums()) >= 20
    I = np.eye(32, dtype=np.complex128)
    d1, p1 = c.descriptors()
    d2, p2 = c2.descriptors()
    assert np.abs(port.apply_circuit(d1, x, I, p1) - port.apply_circuit(d2, x2, I, p2)).max() < 1e-14
    a = H.adaptive_circuit(3, 1)
    xa = H.random_params(a.get_Parameter_Num(), seed=1)
    with pytest.raises(ValueError):
        sq.qasm.dumps(a, xa)
    ca, xb = sq.qasm.loads(sq.qasm.dumps(a, xa, adaptive_as_cry=True))
    assert np.array_equal(xa, xb) and "CRY" in ca.get_Gate_Nums() and "Adaptive" not in ca.get_Gate_Nums()
    da, pa = a.descriptors()
    db, pb = ca.descriptors()
    I8 = np.eye(8, dtype=np.complex128)
    assert np.abs(port.apply_circuit(da, xa, I8, pa) - port.apply_circuit(db, xb, I8, pb)).max() < 1e-14
    with pytest.raises(ValueError):
        sq.qasm.dumps(c, x[:-1])


def test_inverse_structure_against_oracle(port):
    """Circuit.get_Inverse (the host half of apply_from_right, Gates_block.cpp:717-760): for a random nested circuit over EVERY
    gate family (GENERAL and SYC included) the oracle's matrix of the inverse structure at the mapped parameters times the
    matrix of the circuit is the identity; the inverse of the inverse reproduces the circuit's matrix; cached per structure"""
    import helpers as H

    n = 5
    c = H.random_circuit(n, 150, seed=12, general_k=(1, 2, 3), nested=True)
    assert len(c.get_Gate_Nums()) >= 30
    x = H.random_params(c.get_Parameter_Num(), seed=7)
    I = np.eye(1 << n, dtype=np.complex128)
    d, pool = c.descriptors()
    M = port.apply_circuit(d, x, I, pool)
    inv, pmap = c.get_Inverse()
    di, pooli = inv.descriptors()
    Mi = port.apply_circuit(di, pmap(x), I, pooli)
    assert np.abs(Mi @ M - I).max() < 1e-12 and np.abs(M @ Mi - I).max() < 1e-12
    inv2, pmap2 = inv.get_Inverse()
    d2, pool2 = inv2.descriptors()
    assert np.abs(port.apply_circuit(d2, pmap2(pmap(x)), I, pool2) - M).max() < 1e-12
    assert c.get_Inverse()[0] is inv
    c.add_H(0)
    assert c.get_Inverse()[0] is not inv
    with pytest.raises(Exception):
        pmap(x[:-1])


def test_reorder_qubits_keeps_the_cost(port):
    """Reorder_Qubits of the decomposition wrapper (Decomposition_Base.cpp:910-950): gates and unitary are relabelled together,
    so the oracle's cost and gradient of a parameter vector are what they were; the permutation follows the reference's rule
    (new qubit idx = old qubit qbit_list[idx])"""
    import helpers as H

    sq = H.sq
    n = 3
    U = H.random_unitary(1 << n, seed=6)
    dec = sq.N_Qubit_Decomposition_adaptive(U, level_limit_max=2, level_limit_min=1)
    dec.set_Gate_Structure(H.random_circuit(n, 40, seed=5, names=H.ONE_Q + H.CTRL + H.TWO_T + ["CCX", "CSWAP"]))
    P = dec.get_Parameter_Num()
    x = H.random_params(P, seed=9)
    d0, p0 = dec.get_Circuit().descriptors()
    f0, g0 = port.cost_grad(d0, P, x, U, n, 0, pool=p0)
    dec.Reorder_Qubits([2, 0, 1])
    d1, p1 = dec.get_Circuit().descriptors()
    U1 = dec.get_Unitary()
    f1, g1 = port.cost_grad(d1, P, x, U1, n, 0, pool=p1)
    assert abs(f1 - f0) < 1e-13 and np.abs(g1 - g0).max() < 1e-13 and not np.array_equal(U1, U)
    # the reference's index rule on one element: old index 0b011 (qubits 0, 1 set) -> new qubits 1, 2 set = 0b110
    assert U1[0b110, 0] == U[0b011, 0]
    # a gate on old qubit 2 now sits on new qubit 0
    c = sq.Circuit(n)
    c.add_RX(2)
    dec.set_Gate_Structure(c)
    dec.Reorder_Qubits([2, 0, 1])
    assert int(dec.get_Circuit().descriptors()[0][0]["target"]) == 0
    with pytest.raises(Exception):
        dec.Reorder_Qubits([0, 1])
