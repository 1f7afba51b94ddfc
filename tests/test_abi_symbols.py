"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/sqgpu.h declares,
the ctypes struct matches the C struct, and -- on a box without a GPU -- fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H

abi = H.abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sqgpu.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sqgpu_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return abi.load_library()


def test_header_symbols_are_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libsqgpu.so does not export %s" % s
        assert s in abi.PROTOTYPES, "abi.PROTOTYPES has no prototype for %s" % s
    assert sorted(abi.PROTOTYPES) == syms
    assert lib.sqgpu_abi_version() == abi.ABI_VERSION == 2


def test_library_does_not_read_the_environment():
    """round 1 read SQGPU_* variables with getenv on every plan; the switches are per-handle options now
    (sqgpu_set_option): none of the old variable names is left in the binary (the statically linked CUDA runtime still
    imports getenv for its own CUDA_* variables, so the symbol itself cannot be the check)"""
    blob = open(abi.LIB_PATH, "rb").read()
    for name in (b"SQGPU_NO_FUSE", b"SQGPU_WINDOW", b"SQGPU_FORCE_STREAM", b"SQGPU_VQE_STREAM", b"SQGPU_SPLIT", b"SQGPU_THREADS",
                 b"SQGPU_CTAS_PER_SM", b"SQGPU_FUSE_CONSECUTIVE", b"SQGPU_MAX_FUSE_QUBITS", b"SQGPU_VERBOSE"):
        assert name not in blob, name
    src = "".join(open(os.path.join(os.path.dirname(abi.LIB_PATH), f)).read() for f in os.listdir(os.path.dirname(abi.LIB_PATH))
                  if f.endswith((".cu", ".cuh")))
    assert "getenv" not in src


def test_descriptor_struct_layout_matches_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu\\n", sizeof(sqgpu_gate_desc),'
        " offsetof(sqgpu_gate_desc, n_qubits), offsetof(sqgpu_gate_desc, qubits), offsetof(sqgpu_gate_desc, matrix_off));return 0;}\n"
        % HEADER
    )
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    size, o_nq, o_q, o_m = map(int, subprocess.check_output([str(exe)]).split())
    assert size == C.sizeof(abi.GateDesc) == abi.GATE_DESC_DTYPE.itemsize
    assert o_nq == abi.GateDesc.n_qubits.offset == abi.GATE_DESC_DTYPE.fields["n_qubits"][1]
    assert o_q == abi.GateDesc.qubits.offset == abi.GATE_DESC_DTYPE.fields["qubits"][1]
    assert o_m == abi.GateDesc.matrix_off.offset == abi.GATE_DESC_DTYPE.fields["matrix_off"][1]


def test_gate_and_cost_enums_match_header():
    src = open(HEADER).read()
    for name, val in re.findall(r"SQGPU_([A-Z0-9_]+)\s*=\s*(-?\d+)", src):
        if hasattr(abi, name):
            assert getattr(abi, name) == int(val), name
        elif hasattr(abi, "ERR_" + name[4:]) and name.startswith("ERR_"):
            assert getattr(abi, name) == int(val), name


def test_no_device_fails_loudly(lib):
    """no CPU fallback: without a GPU the engine refuses to create a context (and says why)"""
    n = C.c_int(-1)
    rc = lib.sqgpu_device_count(C.byref(n))
    if rc == abi.OK and n.value > 0:
        pytest.skip("a GPU is visible; covered by the -m gpu tests")
    h = abi._handle()
    rc = lib.sqgpu_create(0, C.byref(h))
    assert rc == abi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.sqgpu_last_error() or b"device" in lib.sqgpu_last_error()
    with pytest.raises(abi.SqgpuError):
        H.sq.Engine(0)


def test_circuit_descriptors_and_parameter_layout():
    """Gates_block::add_gate parameter layout (Gates_block.cpp:2500-2525): consecutive slices in insertion order"""
    c = H.adaptive_circuit(4, 5)
    assert c.get_Parameter_Num() == 7 * 6 * 5 + 3 * 4  # 222, SURVEY.md §8
    d, pool = c.descriptors()
    assert len(d) == 3 * 6 * 5 + 4  # 94 gates
    assert (np.cumsum(d["n_params"]) - d["n_params"] == d["param_start"]).all()
    dn, _ = c.descriptors(nested=True)
    assert (dn["type"] == abi.BLOCK_BEGIN).sum() == 6 * 5 + 1
    assert [x for x in dn["type"] if x < 1000] == list(d["type"])
    c10 = H.adaptive_circuit(10, 4)
    assert (len(c10.descriptors()[0]), c10.get_Parameter_Num()) == (550, 1290)


def test_plain_c_client_compiles_and_links(tmp_path):
    """tests/c_abi/abi_client.c is a C11 program that uses nothing but include/sqgpu.h: the boundary carries no C++ or
    torch types (it is what a reference-side shim would be built against)"""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "sequential-quantum-gate-decomposer_b200", "csrc")
    exe = str(tmp_path / "abi_client")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), "-o", exe,
                           os.path.join(root, "tests", "c_abi", "abi_client.c"), "-L" + libdir, "-lsqgpu", "-Wl,-rpath," + libdir])
    assert os.path.exists(exe)


def test_adaptive_wrapper_structure_without_gpu():
    """the gate-structure half of the N_Qubit_Decomposition_adaptive mirror needs no device: same sub-block order, qubit roles
    and parameter layout as the reference's add_adaptive_layers / add_finalyzing_layer
    (N_Qubit_Decomposition_adaptive.cpp:1840-1881, 1947-1966) -- helpers.adaptive_circuit is the structure the oracle tests
    pin against the reference -- and P = 7 n (n - 1) / 2 L + 3 n"""
    import numpy as np

    import helpers as H

    sq = H.sq
    for n, levels, topology in ((4, 2, None), (5, 1, None), (4, 3, [(0, 1), (1, 2), (2, 3)])):
        dec = sq.N_Qubit_Decomposition_adaptive(np.eye(1 << n, dtype=np.complex128), level_limit_max=5, level_limit_min=0,
                                                topology=topology, accelerator_num=1)
        for _ in range(levels):
            dec.add_Adaptive_Layers()
        dec.add_Finalyzing_Layer_To_Gate_Structure()
        want = H.adaptive_circuit(n, levels, topology)
        d0, _ = dec.get_Circuit().descriptors()
        d1, _ = want.descriptors()
        assert d0.tobytes() == d1.tobytes()
        pairs = len(topology) if topology else n * (n - 1) // 2
        assert dec.get_Parameter_Num() == 7 * pairs * levels + 3 * n
    with pytest.raises(Exception):
        sq.N_Qubit_Decomposition_adaptive(np.eye(4, dtype=np.complex128), accelerator_num=0)  # no CPU path in this package


def test_evaluation_without_device_fails_loudly():
    """no CPU fallback: on a box without a CUDA device the first evaluation raises, it does not compute on the host"""
    import ctypes

    import numpy as np

    import helpers as H

    sq = H.sq
    lib = sq.abi.load_library()
    count = ctypes.c_int(0)
    rc = lib.sqgpu_device_count(ctypes.byref(count))
    if rc == 0 and count.value > 0:
        pytest.skip("a CUDA device is present")
    dec = sq.N_Qubit_Decomposition_custom(np.eye(4, dtype=np.complex128))
    c = sq.Circuit(2)
    c.add_U3(0)
    dec.set_Gate_Structure(c)
    with pytest.raises(sq.abi.SqgpuError):
        dec.Optimization_Problem(np.zeros(3))
    with pytest.raises(sq.abi.SqgpuError):
        c.apply_to(np.zeros(3), np.eye(4, dtype=np.complex128))


def test_host_planner_without_gpu(monkeypatch):
    """sqgpu_plan_stats runs the library's own lowering, block fusion and window scheduling without a device: the numbers
    DESIGN.md quotes, and the validation errors of sqgpu_set_circuit"""
    import numpy as np

    import helpers as H

    sq = H.sq
    st = sq.abi.plan_stats(H.adaptive_circuit(10, 4))  # C3: 550 gates, P = 1290
    assert st["ops_plan2"] == 180 and st["ops_plan3"] == 84 and st["block_members"] == 550 and st["dense_ops"] == 0
    st_c = sq.abi.plan_stats(H.adaptive_circuit(10, 4), fuse_consecutive=1)  # runs of consecutive gates only: more, emptier blocks
    assert st_c["ops_plan2"] == 185 and st_c["ops_plan3"] == 100 and st_c["kern_total"] > st["kern_total"]
    assert st["w_total"] == st["kern_total"]  # every op of this structure carries parameters: one W accumulator per kernel
    assert st["segments"] == 1 and st["window"] == 10
    st = sq.abi.plan_stats(H.hea_zyz_circuit(20, 10))  # C5: 1330 gates, default window of 11 qubits
    assert st["block_members"] == 1330 and st["window"] == 11 and st["segments"] == 9 and st["ops_plan3"] == 99
    # a narrower window needs more segments (first fit alone: 20)
    assert 9 < sq.abi.plan_stats(H.hea_zyz_circuit(20, 10), window=10)["segments"] <= 16
    c = H.random_circuit(7, 80, seed=5, general_k=(2, 3))
    st = sq.abi.plan_stats(c, window=4)
    assert st["segments"] > 1 and st["max_segment_ops"] >= 1 and st["dense_ops"] == 3
    st = sq.abi.plan_stats(c, window=4, no_fuse=1)
    assert st["ops_plan2"] == st["ops_plan3"] == len(c.descriptors()[0]) and st["block_members"] == 0
    with pytest.raises(sq.abi.SqgpuError):
        sq.abi.plan_stats(c, no_such_option=1)
    with pytest.raises(sq.abi.SqgpuError):
        sq.abi.plan_stats(c, window=99)
    # the library no longer reads the process environment: an exported variable of round 1 changes nothing
    monkeypatch.setenv("SQGPU_NO_FUSE", "1")
    assert sq.abi.plan_stats(c, window=4) == sq.abi.plan_stats(c, window=4, no_fuse=0)
    # validation: the same errors sqgpu_set_circuit raises
    bad = sq.Circuit(3)
    bad.add_U3(0)
    d, pool = bad.descriptors()
    lib = sq.abi.load_library()
    import ctypes as C

    out = (C.c_int64 * 10)()
    d2 = d.copy()
    d2["target"][0] = 5  # qubit out of range
    assert lib.sqgpu_plan_stats(d2.ctypes.data_as(C.POINTER(sq.abi.GateDesc)), 1, 3, 3, None, 0, out, 10) == sq.abi.ERR_INVALID
    assert b"out of range" in lib.sqgpu_last_error()
    assert lib.sqgpu_plan_stats(d.ctypes.data_as(C.POINTER(sq.abi.GateDesc)), 1, 4, 3, None, 0, out, 10) == sq.abi.ERR_INVALID
    assert b"not used by any gate" in lib.sqgpu_last_error()


def test_structure_key_tracks_nested_mutations():
    """the device plan is keyed on a recursive structure key (a sub-circuit mutated after add_Circuit, with the
    parameter count unchanged, used to keep a stale plan)"""
    sq = H.sq
    inner = sq.Circuit(3)
    inner.add_U3(0)
    outer = sq.Circuit(3)
    outer.add_Circuit(inner)
    outer.add_CNOT(1, 0)
    k0 = outer.structure_key()
    assert outer.structure_key() == k0
    inner.add_CNOT(2, 1)  # no new parameters, same number of top-level items
    k1 = outer.structure_key()
    assert k1 != k0
    dec = sq.N_Qubit_Decomposition_custom(np.eye(8, dtype=np.complex128))
    dec.set_Gate_Structure(outer)
    k2 = dec.get_Circuit().structure_key()
    inner.add_H(0)  # shared nested block, mutated through the caller's reference
    assert dec.get_Circuit().structure_key() != k2
    dec.get_Circuit().add_X(1)  # the live object handed out by get_Circuit
    assert len(dec.get_Circuit().descriptors()[0]) == 5


def test_vqe_wrapper_structure_without_gpu():
    """Generate_Circuit reproduces generate_circuit (Variational_Quantum_Eigensolver_Base.cpp:1299-1437): the HEA_ZYZ
    structure equals helpers.hea_zyz_circuit, which the oracle tests pin against the reference class itself"""
    sq = H.sq
    n = 6
    Hm = H.heisenberg_csr(n)
    vqe = sq.Variational_Quantum_Eigensolver(Hm, n, accelerator_num=1)
    vqe.set_Ansatz("HEA_ZYZ")
    vqe.Generate_Circuit(3, 2)
    assert vqe.get_Circuit().descriptors()[0].tobytes() == H.hea_zyz_circuit(n, 3, 2).descriptors()[0].tobytes()
    vqe.set_Ansatz("HEA")
    vqe.Generate_Circuit(2, 1)
    d = vqe.get_Circuit().descriptors()[0]
    assert list(d["type"][:3]) == [sq.abi.U3, sq.abi.U3, sq.abi.CNOT] and vqe.get_Parameter_Num() == 6 * 5 * 2
    with pytest.raises(Exception):
        vqe.set_Ansatz("UCC")
    with pytest.raises(Exception):
        sq.Variational_Quantum_Eigensolver(Hm, n, accelerator_num=0)
    with pytest.raises(Exception):
        sq.Variational_Quantum_Eigensolver(Hm, n + 1)


def test_multi_handle_argument_checks(lib):
    """sqgpu_create_multi validates before it touches a device; without one it fails like sqgpu_create (no CPU fallback)"""
    h = abi._handle()
    assert lib.sqgpu_create_multi(0, None, abi.SHARD_AUTO, C.byref(h)) == abi.ERR_INVALID
    assert lib.sqgpu_create_multi(2, None, 7, C.byref(h)) == abi.ERR_INVALID and b"sharding mode" in lib.sqgpu_last_error()
    twice = (C.c_int * 2)(0, 0)
    assert lib.sqgpu_create_multi(2, twice, abi.SHARD_BATCH, C.byref(h)) == abi.ERR_INVALID and b"twice" in lib.sqgpu_last_error()
    n = C.c_int(-1)
    if lib.sqgpu_device_count(C.byref(n)) == abi.OK and n.value > 0:
        pytest.skip("a GPU is visible; the multi-device handle is covered by the -m gpu tests")
    assert lib.sqgpu_create_multi(2, None, abi.SHARD_AUTO, C.byref(h)) == abi.ERR_NO_DEVICE
    with pytest.raises(abi.SqgpuError):
        H.sq.Engine(devices=2)


def test_lbfgs_host_loop_with_oracle_callables(port):
    """optimize.lbfgs is host logic over two callables (cost+gradient, batched line search); with oracle-backed callables it
    decomposes a 3-qubit unitary made of the very structure it optimises: cost -> 0"""
    sq = H.sq
    n = 3
    c = H.adaptive_circuit(n, 2)
    d, pool = c.descriptors()
    P = c.get_Parameter_Num()
    U = np.ascontiguousarray(port.apply_circuit(d, H.random_params(P, seed=5), np.eye(1 << n, dtype=np.complex128)).conj().T)
    cg = lambda x: port.cost_grad(d, P, x, U, n, 0)

    def ls(x, dd, al):
        r = [port.cost_grad(d, P, x + a * dd, U, n, 0) for a in al]
        return np.array([q[0] for q in r]), np.array([q[1] @ dd for q in r])

    x, f, it, ne = sq.optimize.multistart_lbfgs(lambda X: [port.cost(d, xx, U, n, 0) for xx in X], cg, ls, P, np.random.default_rng(1),
                                                starts=8, keep=3, max_iter=300, tol=1e-9)
    assert f < 1e-4 and ne > it > 0  # far below the starting cost (~1); the decomposition tolerance of the reference is 1e-4
    dec = sq.N_Qubit_Decomposition_adaptive(U, level_limit_max=3, level_limit_min=1)
    with pytest.raises(Exception):
        dec.set_Optimizer("BAYES_OPT")
    for name in ("COSINE", "AGENTS", "AGENTS_COMBINED", "GRAD_DESCEND", "GRAD_DESCEND_PARAMETER_SHIFT_RULE", "ADAM", "BFGS"):
        dec.set_Optimizer(name)
    with pytest.raises(Exception):
        dec.get_Optimized_Parameters()


def test_cosine_engine_with_oracle_callables(port):
    """optimize.cosine (the reference's COSINE engine, optimization_engines/COSINE.cpp:60-657, with its evaluations arranged as
    device batches) is host logic over one callable. With the oracle as that callable:
    * the three-point rule lands on the exact minimum of the cost along one parameter -- Frobenius cost (period 2 pi) and the
      VQE energy with the doubled period -- checked against a dense scan of the oracle;
    * the engine's cost never increases and drops by 5x in 150 iterations on a 3-qubit unitary built from the structure it optimises."""
    sq = H.sq
    n = 3
    c = H.adaptive_circuit(n, 2)
    d, pool = c.descriptors()
    P = c.get_Parameter_Num()
    U = np.ascontiguousarray(port.apply_circuit(d, H.random_params(P, seed=5), np.eye(1 << n, dtype=np.complex128)).conj().T)
    x = H.random_params(P, seed=9)
    scan = np.linspace(-np.pi, np.pi, 721)
    for variant, dbl in ((0, False),):
        cost = lambda v: port.cost(d, v, U, n, variant)
        s = np.pi / 4 if dbl else np.pi / 2
        for i in (0, 7, P - 1):
            e = np.zeros(P)
            e[i] = 1.0
            upd = float(sq.optimize.cosine_updates(cost(x), cost(x + s * e), cost(x + 2 * s * e), dbl))
            f_scan = min(cost(x + t * e) for t in scan)
            assert cost(x + upd * e) <= f_scan + 1e-12 and abs(upd) <= (np.pi / 2 if dbl else np.pi) + 1e-12
    # VQE energy, doubled period (HEA_ZYZ: every parameter is a rotation angle in the theta / 2 convention)
    nv = 4
    ip, ix, dat = H.heisenberg_csr_fast(nv)
    cv = H.hea_zyz_circuit(nv, 2)
    dv, _ = cv.descriptors()
    Pv = cv.get_Parameter_Num()
    psi0 = np.zeros(1 << nv, dtype=np.complex128)
    psi0[0] = 1
    energy = lambda v: port.vqe_energy(dv, v, psi0, ip, ix, dat)
    xv = H.random_params(Pv, seed=2)
    for i in (1, Pv // 2):
        e = np.zeros(Pv)
        e[i] = 1.0
        upd = float(sq.optimize.cosine_updates(energy(xv), energy(xv + np.pi / 4 * e), energy(xv + np.pi / 2 * e), True))
        assert energy(xv + upd * e) <= min(energy(xv + t * e) for t in scan[::4]) + 1e-12
    # the engine
    calls, trace = [], []
    def cost_batched(X):
        calls.append(len(X))
        return np.array([port.cost(d, v, U, n, 0) for v in X])
    xs, f, it, ne = sq.optimize.cosine(cost_batched, x, np.random.default_rng(3), batch_size=16, max_iter=150, tol=1e-8,
                                        callback=lambda k, xx, ff: trace.append(ff))
    assert all(b <= a for a, b in zip(trace, trace[1:])) and trace[0] <= port.cost(d, x, U, n, 0)
    # (a coordinate-wise method: slow but steady; the starting cost is ~0.95)
    assert f < 0.2 and abs(f - port.cost(d, xs, U, n, 0)) < 1e-12
    # two device round trips per iteration: 2 x batch_size shifted sets, then the line-search grid
    assert calls[0] == 1 and set(calls[1::2]) == {32} and set(calls[2::2]) == {16} and ne == sum(calls)
    ev, fv, _, _ = sq.optimize.cosine(lambda X: np.array([energy(v) for v in X]), xv, np.random.default_rng(1), batch_size=8, max_iter=40,
                                      tol=-np.inf, double_period=True)
    assert fv < energy(xv) - 0.5
    with pytest.raises(Exception):
        sq.optimize.cosine(cost_batched, x, np.random.default_rng(3), batch_size=P + 1)
    # AGENTS (AGENTS.cpp:333-415, 700-870): independent walkers, each jumps to the exact minimum along one random parameter per
    # iteration, so the cost it PREDICTS (offset - amplitude) is the cost the oracle evaluates; census every `agent_lifetime`
    calls.clear()
    seen = []
    def spy(X):
        v = cost_batched(X)
        seen.append((np.array(X), v))
        return v
    xa, fa, ita, nea = sq.optimize.agents(spy, x, np.random.default_rng(5), agent_num=8, max_iter=120, tol=1e-8, agent_lifetime=30)
    assert fa < 0.4 and abs(fa - port.cost(d, xa, U, n, 0)) < 1e-12 and nea == sum(calls) and ita == 120
    # one batch of 2 x agent_num shifted sets per iteration (+ a census of agent_num every 30 iterations)
    assert calls.count(16) == 120 and calls.count(8) >= 1 + 4
    # iteration k's shifted batch starts from iteration k-1's updated agents: the cost at the un-shifted point, recovered from the
    # next census, equals the prediction chain -- checked through the census right after iteration 30
    census = [i for i, (X, v) in enumerate(seen) if len(X) == 8][2]
    Xc, vc = seen[census]
    Xs, vs = seen[census - 1]  # the shifted batch of the same iteration: rows 0..7 are +pi/2 of the state BEFORE the update
    assert np.abs(vc - np.array([min(port.cost(d, Xs[a] + t * (Xs[8 + a] - Xs[a]) * 2 / np.pi, U, n, 0) for t in scan) for a in range(8)])).max() < 1e-4
    assert (vc <= np.array([port.cost(d, 2 * Xs[a] - Xs[8 + a], U, n, 0) for a in range(8)]) + 1e-12).all()
    # GRAD_DESCEND_PARAMETER_SHIFT_RULE (…SHIFT_RULE.cpp:249-325): its gradient component f(+pi/4) - f(-pi/4) is sqrt(2) times the
    # oracle's derivative for the Frobenius cost, the line search (fraction 0 included) never lets the cost rise
    g_ref = port.cost_grad(d, P, x, U, n, 0)[1]
    for i in (0, 7, P - 1):
        e = np.zeros(P)
        e[i] = 1.0
        assert abs((cost(x + np.pi / 4 * e) - cost(x - np.pi / 4 * e)) / np.sqrt(2) - g_ref[i]) < 1e-12
    trace = []
    xp, fp, itp, nep = sq.optimize.grad_descend_shift_rule(cost_batched, x, np.random.default_rng(3), batch_size=16, max_iter=40, tol=1e-8, eta=1.0,
                                                           line_points=32, callback=lambda k, xx, ff: trace.append(ff))
    assert all(b <= a for a, b in zip(trace, trace[1:])) and fp < 0.3 and abs(fp - cost(xp)) < 1e-12 and nep == 1 + 40 * (32 + 32)
    xq, fq, _, _ = sq.optimize.grad_descend_shift_rule(cost_batched, x, np.random.default_rng(3), batch_size=16, max_iter=5, tol=1e-8, eta=0.05, use_line_search=False)
    assert fq < cost(x)
    # the five-point rule for the Hilbert-Schmidt test (AGENTS.cpp:536-660): along one parameter the cost is
    # kappa sin(2 p + xi) + gamma sin(p + varphi) + offset; the minimum of the fitted curve is the minimum of the oracle's cost
    hs = lambda v: port.cost(d, v, U, n, 3)
    for i in (0, 1, 7, P - 1):
        e = np.zeros(P)
        e[i] = 1.0
        upd, pred = sq.optimize.five_point_updates(x[i], *[hs(x + s * e) for s in (0, np.pi / 4, np.pi / 2, np.pi, 3 * np.pi / 2)])
        assert abs(hs(x + float(upd) * e) - float(pred)) < 1e-12 and float(pred) <= min(hs(x + t * e) for t in scan) + 1e-12
    trace = []
    xh, fh, _, _ = sq.optimize.agents(lambda X: np.array([hs(v) for v in X]), x, np.random.default_rng(3), agent_num=8, max_iter=120, tol=1e-8,
                                      agent_lifetime=30, five_point=True, callback=lambda k, xx, ff: trace.append(ff))
    assert fh < 0.5 * hs(x) and abs(fh - hs(xh)) < 1e-12 and all(b <= a for a, b in zip(trace, trace[1:]))


def test_constant_subcircuits_become_dense_kernels():
    """N3: parameter-free stretches whose support fits 4 qubits are multiplied out on the host into one dense kernel when
    that needs fewer flops than the 3-qubit blocks they would occupy; the parametric decomposition structures are left alone.
    Planner only (no device)."""
    import helpers as H

    sq = H.sq
    c = H.const_heavy_circuit(8, 3, 40, seed=11)
    ops = sq.abi.plan_ops(c, which=3)
    dense = [o for o in ops if o[0] == 16 and o[2] == 0 and o[3] == 0]
    assert len(dense) == 3 and all(len(o[1]) == 4 for o in dense)
    ops_off = sq.abi.plan_ops(c, which=3, const_fuse_qubits=0)
    assert not [o for o in ops_off if o[0] == 16] and len(ops_off) > len(ops)
    # the <= 2-qubit plan of the streaming fallback keeps the gates as they are
    assert sq.abi.plan_ops(c, which=2) == sq.abi.plan_ops(c, which=2, const_fuse_qubits=0)
    # up to five qubits on request
    c5 = H.const_heavy_circuit(8, 2, 80, seed=12, support=5)
    assert [o for o in sq.abi.plan_ops(c5, which=3, const_fuse_qubits=5) if o[0] == 32]
    assert not [o for o in sq.abi.plan_ops(c5, which=3) if o[0] == 32]
    # a short constant stretch is cheaper inside the neighbouring blocks: nothing changes
    short = H.const_heavy_circuit(8, 3, 4, seed=13)
    assert sq.abi.plan_ops(short, which=3) == sq.abi.plan_ops(short, which=3, const_fuse_qubits=0)
    # C3 / C5 structures: every CNOT-like gate sits between parametric ones, the plans are what they were
    assert sq.abi.plan_stats(H.adaptive_circuit(10, 4))["ops_plan3"] == 84
    assert sq.abi.plan_stats(H.hea_zyz_circuit(20, 10)) == sq.abi.plan_stats(H.hea_zyz_circuit(20, 10), const_fuse_qubits=0)


def test_cluster_plan_without_gpu():
    """the cluster planner (columns shared by 2 / 4 / 8 CTAs): RESPLIT ops are inserted so that no op touches a split qubit,
    qubits are rewritten to row-bit positions, and the exchange count stays small (Belady choice of the qubit to split on).
    The test replays the plan's bookkeeping: positions form a permutation and every op acts on local row bits only."""
    import helpers as H

    sq = H.sq
    for n, levels, rho in ((13, 1, 1), (13, 1, 2), (12, 2, 1), (14, 1, 3)):
        c = H.adaptive_circuit(n, levels)
        base = sq.abi.plan_ops(c, which=3)
        ops = sq.abi.plan_ops(c, which=10 + rho)
        L = n - rho
        resplits = [o for o in ops if o[0] == 0]
        work = [o for o in ops if o[0] != 0]
        assert len(work) == len(base) and [o[0] for o in work] == [o[0] for o in base]
        assert 0 < len(resplits) <= len(base) // 3
        for dim, qs, n_par, n_mem in work:
            assert all(0 <= q < L for q in qs) and len(set(qs)) == len(qs)
        for _, (j, i), _, _ in [(o[0], o[1][:2], o[2], o[3]) for o in resplits]:
            assert 0 <= j < L and 0 <= i < rho
        # the same logical qubits, only relabelled: supports have the same sizes op by op
        assert [len(o[1]) for o in work] == [len(o[1]) for o in base]
    # circuits with raw dense ops have no cluster plan; small circuits neither
    g = sq.Circuit(12)
    g.add_GENERAL(H.random_unitary(16, seed=1), [0, 3, 5, 9])
    g.add_U3(2)
    assert sq.abi.plan_ops(g, which=11) == []
    assert sq.abi.plan_ops(H.adaptive_circuit(10, 1), which=11) == []
