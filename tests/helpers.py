"""Shared deterministic inputs for the parity tests (recipes from the reference's own tests)."""
import numpy as np

import squander_b200 as sq

abi = sq.abi

ONE_Q = ["U1", "U2", "U3", "RX", "RY", "RZ", "R", "H", "X", "Y", "Z", "S", "Sdg", "T", "Tdg", "SX", "SXdg"]
CTRL = ["CNOT", "CZ", "CH", "CU", "CRY", "CRX", "CRZ", "CP", "CR", "adaptive", "CROT", "SYC"]
TWO_T = ["RXX", "RYY", "RZZ", "SWAP"]


def random_unitary(dim, seed=123):
    """QR recipe of tests/gates/test_circuit.py:93-102 (reference)."""
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    q, r = np.linalg.qr(a)
    return np.ascontiguousarray(q * (np.diag(r) / np.abs(np.diag(r))))


def random_state(dim, seed=7):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    return np.ascontiguousarray(v / np.linalg.norm(v))


def add_named(c, name, qs):
    """add gate `name` on the leading entries of the qubit permutation qs"""
    if name in ONE_Q:
        getattr(c, "add_" + name)(qs[0])
    elif name in CTRL:
        getattr(c, "add_" + name)(qs[0], qs[1])
    elif name == "CCX":
        c.add_CCX(qs[0], [qs[1], qs[2]])
    elif name == "CSWAP":
        c.add_CSWAP([qs[0], qs[1]], [qs[2]])
    elif name in TWO_T:
        getattr(c, "add_" + name)([qs[0], qs[1]])
    else:
        raise KeyError(name)


def random_circuit(n, n_gates, seed, names=None, general_k=(), nested=False):
    """random flat (or 2-level nested) circuit over the supported gate set (+ GENERAL blocks of the given sizes)"""
    rng = np.random.default_rng(seed)
    names = list(names or (ONE_Q + CTRL + TWO_T + (["CCX", "CSWAP"] if n >= 3 else [])))
    if nested:
        # the reference cannot nest CRX/CRZ: Gates_block::set_qbit_num has no case for them (Gates_block.cpp:3287-3309)
        names = [x for x in names if x not in ("CRX", "CRZ")]
    c = sq.Circuit(n)
    cur = c
    for i in range(n_gates):
        if nested and i % 5 == 0:
            cur = sq.Circuit(n)
            c.add_Circuit(cur)
            # a parameter-free nested block makes the reference's derivative route call the unimplemented
            # Gates_block::gate_kernel (build_forward_inputs -> Gate::apply_to(Matrix&)); keep one parameter inside
            cur.add_RZ(int(rng.integers(0, n)))
        r = rng.integers(0, len(names) + len(general_k))
        qs = [int(q) for q in rng.permutation(n)]
        if r >= len(names):
            k = general_k[r - len(names)]
            cur.add_GENERAL(random_unitary(1 << k, seed=1000 + i), qs[:k])
        else:
            add_named(cur, names[r], qs)
    return c


def adaptive_circuit(n, levels, topology=None):
    """gate structure of N_Qubit_Decomposition_adaptive: `levels` x add_Adaptive_Layers + finalizing layer"""
    c = sq.Circuit(n)
    pairs = [(t, cq) for t in range(n) for cq in range(t + 1, n)] if not topology else [(t, cq) for (cq, t) in topology]
    for _ in range(levels):
        for t, cq in pairs:
            layer = sq.Circuit(n)
            layer.add_U3(t)
            layer.add_U3(cq)
            layer.add_adaptive(t, cq)
            c.add_Circuit(layer)
    fin = sq.Circuit(n)
    for q in range(n):
        fin.add_U3(q)
    c.add_Circuit(fin)
    return c


def random_params(n_params, seed=42, batch=None):
    rng = np.random.default_rng(seed)
    if batch is None:
        return rng.random(n_params) * 2 * np.pi
    return rng.random((batch, n_params)) * 2 * np.pi


def heisenberg_csr(n, seed=31415, degree=3):
    """sum_(i,j) (XX+YY+ZZ) on a random regular graph + sum_i Z_i as CSR complex128
    (examples/VQE/Heisenberg_VQE.py:43-52, tests/VQE/test_VQE.py:26-74; scipy only)."""
    import scipy.sparse as sp

    assert (n * degree) % 2 == 0, "a regular graph needs n * degree even"
    rng = np.random.default_rng(seed)
    # simple deterministic pseudo-random regular graph: pairing model with retries
    while True:
        stubs = np.repeat(np.arange(n), degree)
        rng.shuffle(stubs)
        edges = set()
        ok = True
        for a, b in stubs.reshape(-1, 2):
            if a == b or (min(a, b), max(a, b)) in edges:
                ok = False
                break
            edges.add((int(min(a, b)), int(max(a, b))))
        if ok:
            break
    X = sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=np.complex128))
    Y = sp.csr_matrix(np.array([[0, -1j], [1j, 0]], dtype=np.complex128))
    Z = sp.csr_matrix(np.array([[1, 0], [0, -1]], dtype=np.complex128))
    I = sp.identity(2, dtype=np.complex128, format="csr")

    def op(single, q):
        # qubit q <-> bit q of the basis index: kron with the highest qubit leftmost
        m = sp.identity(1, dtype=np.complex128, format="csr")
        for k in range(n - 1, -1, -1):
            m = sp.kron(m, single if k == q else I, format="csr")
        return m

    Hm = sp.csr_matrix((1 << n, 1 << n), dtype=np.complex128)
    for a, b in sorted(edges):
        for P in (X, Y, Z):
            Hm = Hm + op(P, a) @ op(P, b)
    for q in range(n):
        Hm = Hm + op(Z, q)
    Hm = sp.csr_matrix(Hm)
    Hm.sort_indices()
    return Hm.indptr.astype(np.int32), Hm.indices.astype(np.int32), Hm.data.astype(np.complex128)


def _regular_graph_edges(n, seed, degree):
    rng = np.random.default_rng(seed)
    while True:
        stubs = np.repeat(np.arange(n), degree)
        rng.shuffle(stubs)
        edges = set()
        ok = True
        for a, b in stubs.reshape(-1, 2):
            if a == b or (min(a, b), max(a, b)) in edges:
                ok = False
                break
            edges.add((int(min(a, b)), int(max(a, b))))
        if ok:
            return sorted(edges)


def heisenberg_csr_fast(n, seed=31415, degree=3):
    """the same matrix as heisenberg_csr (identical indptr / indices / data; checked in tests/test_oracle_golden.py), built
    from bit arithmetic instead of 2^n-dimensional Kronecker products (n = 20: 1 s instead of 30 s).
    Row i: diagonal sum_edges (+1 if bits equal else -1) + sum_q (1 - 2 bit_q); for every edge whose bits differ an entry 2
    at column i ^ (1 << a | 1 << b) (XX + YY = 2 there, 0 where the bits are equal); zeros are not stored (scipy prunes them)."""
    edges = _regular_graph_edges(n, seed, degree)
    dim = 1 << n
    idx = np.arange(dim, dtype=np.int64)
    diag = np.zeros(dim, dtype=np.float64)
    for q in range(n):
        diag += 1.0 - 2.0 * ((idx >> q) & 1)
    cols = [idx]
    vals = [diag]
    for a, b in edges:
        differ = ((idx >> a) ^ (idx >> b)) & 1
        diag += 1.0 - 2.0 * differ
        cols.append(np.where(differ == 1, idx ^ ((1 << a) | (1 << b)), -1))
        vals.append(np.full(dim, 2.0))
    vals[0] = diag
    cols[0] = np.where(diag != 0.0, idx, -1)
    C = np.stack(cols, axis=1)
    V = np.stack(vals, axis=1)
    order = np.argsort(np.where(C < 0, np.int64(1) << 40, C), axis=1, kind="stable")
    C = np.take_along_axis(C, order, axis=1)
    V = np.take_along_axis(V, order, axis=1)
    keep = C >= 0
    indptr = np.zeros(dim + 1, dtype=np.int64)
    np.cumsum(keep.sum(axis=1), out=indptr[1:])
    return indptr.astype(np.int32), C[keep].astype(np.int32), V[keep].astype(np.complex128)


def hea_zyz_circuit(n, layers, inner_blocks=1):
    """HEA_ZYZ ansatz exactly as generate_initial_circuit builds it
    (Variational_Quantum_Eigensolver_Base.cpp:1358-1416): blocks [RZ,RY,RZ] on both qubits of a pair + CNOT,
    pairs (1,0), then for odd control c: (c+2, c+1) if it exists, then (c+1, c)."""
    c = sq.Circuit(n)

    def zyz(q):
        b = sq.Circuit(n)
        b.add_RZ(q)
        b.add_RY(q)
        b.add_RZ(q)
        c.add_Circuit(b)

    def pair(first, second, tgt, ctl):
        for _ in range(inner_blocks):
            zyz(first)
            zyz(second)
            c.add_CNOT(tgt, ctl)

    for _ in range(layers):
        pair(1, 0, 1, 0)
        for cq in range(1, n - 1, 2):
            if cq + 2 < n:
                pair(cq + 1, cq + 2, cq + 2, cq + 1)
            pair(cq + 1, cq, cq + 1, cq)
    return c


def pauli_exponent_circuit(alpha=0.6217 * np.pi):
    """the 5-qubit target circuit of the reference's tests/decomposition/test_parametric_circuit.py:50-115, built with
    the Qiskit -> SQUANDER conventions of Qiskit_IO.py (cx(a, b): control a, target b; rotation angles stored halved).
    Returns (circuit, parameters)."""
    c = sq.Circuit(5)
    p = []

    def h(q):
        c.add_H(q)

    def cx(a, b):
        c.add_CNOT(b, a)

    def rx(t, q):
        c.add_RX(q)
        p.append(t / 2)

    def rz(t, q):
        c.add_RZ(q)
        p.append(t / 2)

    h(1); cx(1, 2)
    rx(np.pi / 2, 0); rx(np.pi / 2, 1); cx(2, 4); cx(0, 1)
    rx(np.pi / 2, 0); h(2); cx(0, 2)
    rx(np.pi / 2, 0); h(3); rz(alpha, 4); cx(0, 3)
    h(0); rz(-alpha, 1); cx(2, 4)
    cx(2, 1); rz(-alpha, 4); cx(3, 1)
    rz(alpha, 1); cx(0, 1); cx(3, 1); cx(4, 1)
    rz(-alpha, 1); cx(2, 1)
    rz(alpha, 1); cx(3, 1); cx(4, 1)
    rz(alpha, 1); cx(2, 4); cx(0, 1)
    h(0); cx(3, 1); cx(0, 3)
    rx(-np.pi / 2, 0); h(3); cx(0, 2)
    rx(-np.pi / 2, 0); h(2); cx(0, 1)
    rx(-np.pi / 2, 0); rx(-np.pi / 2, 1); cx(2, 4); cx(1, 2); h(1)
    return c, np.array(p, dtype=np.float64)


CONST_NAMES = ["CNOT", "CZ", "CH", "H", "X", "Y", "Z", "S", "Sdg", "T", "Tdg", "SX", "SXdg", "SWAP", "CCX", "CSWAP", "SYC"]


def const_heavy_circuit(n, stretches, gates_per_stretch, seed, support=4, general_k=()):
    """U3 layers with parameter-free stretches in between: every stretch is `gates_per_stretch` random constant gates on a
    random set of `support` qubits (N3: the planner multiplies such sub-circuits out into dense kernels)"""
    rng = np.random.default_rng(seed)
    c = sq.Circuit(n)
    for s in range(stretches):
        for q in range(n):
            c.add_U3(q)
        sub = sorted(int(q) for q in rng.choice(n, support, replace=False))
        for g in range(gates_per_stretch):
            qs = [sub[i] for i in rng.permutation(support)]
            r = int(rng.integers(0, len(CONST_NAMES) + len(general_k)))
            if r >= len(CONST_NAMES):
                k = general_k[r - len(CONST_NAMES)]
                c.add_GENERAL(random_unitary(1 << k, seed=5000 + 37 * s + g), qs[:k])
            else:
                add_named(c, CONST_NAMES[r], qs)
    for q in range(n):
        c.add_RY(q)
    return c
