"""The drop-in for real: the reference's OWN decomposition class and OWN optimizers
(BFGS_Powell, ADAM), compiled from /root/reference together with the shim of integration/ (oracle/_ref/libsqref_gpu.so), with
every cost / gradient evaluation served by libsqgpu.so through the C-ABI.

CPU part: the shim's Gates_block -> descriptor flattening, the CPU flavour of the harness, and the loud failure of the GPU
flavour without a device. GPU part: the teacher-forced trajectory check of the north star -- each iterate the reference's
optimizer visits while driven by the GPU is re-evaluated by the reference's CPU cost path: f and grad within 1e-10 -- and the
free-running comparison of the two trajectories."""
import os

import numpy as np
import pytest

import golden_cases as G
import helpers as H

REL_TOL = 1e-10


@pytest.fixture(scope="module")
def refgpu():
    import pyoracle

    if not pyoracle.RefGpu.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libsqref_gpu.so not built and /root/reference absent")
    return pyoracle.RefGpu()


def close_rel(a, b, tol=REL_TOL):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def c1_case(levels=3):
    """BASELINE configs[0]: data/Umtx.mat (stored in the golden fixture), adaptive structure"""
    U = G.load("C1_L3").U  # the same matrix in every C1 case
    circ = H.adaptive_circuit(4, levels)
    return 4, circ, U


def c2_case():
    """BASELINE configs[1]: the 19-CNOT circuit of data/19CNOT.qasm against its Pauli-exponent target"""
    g = G.load("C2_19CNOT")
    return g


@pytest.mark.parametrize("seed", [3, 11])
def test_shim_flattening_matches_python_descriptors(refgpu, seed):
    """integration/common_GPU.cpp: to_gpu_gates walks the reference's (nested) Gates_block; its output must be the flat
    descriptor stream our own Circuit mirror sends to the engine -- every gate class, qubit role and parameter slot"""
    n = 5
    c = H.random_circuit(n, 60, seed=seed, nested=True)
    dn, pool = c.descriptors(nested=True)
    df, pool_f = c.descriptors()
    got, got_pool = refgpu.flatten(n, dn, pool)
    assert len(got) == len(df)
    for k in ("type", "param_start", "n_params"):
        assert (got[k] == df[k]).all(), k

    def roles(d):
        # the reference stores the qubit lists of its two-target / two-control gates sorted (RXX, RYY, RZZ, SWAP, CSWAP, CCX are
        # symmetric in them): compare as sets
        return [(frozenset((int(r["target"]), int(r["target2"]))), frozenset((int(r["control"]), int(r["control2"])))) for r in d]

    assert roles(got) == roles(df)
    c = H.adaptive_circuit(4, 2)
    got, _ = refgpu.flatten(4, c.descriptors(nested=True)[0])
    assert got.tobytes() == c.descriptors()[0].tobytes()


def test_cpu_flavour_runs_the_reference_optimizers(refgpu):
    n, circ, U = c1_case(2)
    d, pool = circ.descriptors(nested=True)
    x0 = H.random_params(circ.get_Parameter_Num(), seed=1)
    for alg, iters in ((refgpu_alg("BFGS"), 20), (refgpu_alg("ADAM"), 20)):
        s = refgpu.session(False, U, n, d, pool)
        x, f, log = s.optimize(alg, x0, iters)
        assert len(log["cost"]) >= iters and np.allclose(log["params"][0], x0) and f <= log["cost"][0]
        assert s.gpu_evaluations() == 0


def refgpu_alg(name):
    import pyoracle

    return getattr(pyoracle.RefGpuSession, name)


def test_gpu_flavour_fails_loudly_without_device(refgpu):
    if refgpu.available_gpus() > 0:
        pytest.skip("a GPU is visible; covered by the -m gpu tests")
    n, circ, U = c1_case(1)
    d, pool = circ.descriptors(nested=True)
    s = refgpu.session(True, U, n, d, pool)
    with pytest.raises(Exception, match="no CPU fallback|CUDA|device"):
        s.cost_grad(H.random_params(circ.get_Parameter_Num()))


# ---- on the GPU ------------------------------------------------------------------------------------------------------------

def _teacher_forced(refgpu, n, descs_nested, pool, U, x0, alg, iters, variant, eta=1e-3, prev=1.0):
    gpu = refgpu.session(True, U, n, descs_nested, pool)
    cpu = refgpu.session(False, U, n, descs_nested, pool)
    for s in (gpu, cpu):
        s.set_cost(variant, 0, prev)
    xg, fg, lg = gpu.optimize(alg, x0, iters, eta)
    assert gpu.gpu_evaluations() >= len(lg["cost"]) > 0  # the optimizer's evaluations really went through libsqgpu.so
    # teacher-forced: the CPU path of the reference at every iterate of the GPU-driven run
    worst_f = worst_g = 0.0
    check = cpu if alg == refgpu_alg("BFGS") else None
    fresh = refgpu.session(False, U, n, descs_nested, pool)
    for k in range(len(lg["cost"])):
        # ADAM rewrites prev_cost_fnv_val = f0 after every evaluation (ADAM.cpp:201); only the correction variants read it
        fresh.set_cost(variant, 0, prev if (k == 0 or alg == refgpu_alg("BFGS")) else lg["cost"][k - 1])
        f_ref, g_ref = fresh.cost_grad(lg["params"][k])
        worst_f = max(worst_f, abs(lg["cost"][k] - f_ref) / max(1.0, abs(f_ref)))
        worst_g = max(worst_g, np.abs(lg["grad"][k] - g_ref).max() / max(1.0, np.abs(g_ref).max()))
    assert worst_f <= REL_TOL and worst_g <= REL_TOL, (worst_f, worst_g)
    # free-running: the reference's optimizer on its own CPU path from the same start
    xc, fc, lc = cpu.optimize(alg, x0, iters, eta)
    m = min(len(lc["cost"]), len(lg["cost"]))
    drift = np.abs(lc["params"][:m] - lg["params"][:m]).max(axis=1)
    return {"evals": len(lg["cost"]), "worst_f": worst_f, "worst_g": worst_g, "drift": drift, "f_gpu": fg, "f_cpu": fc,
            "log_gpu": lg, "log_cpu": lc}


@pytest.mark.gpu
@pytest.mark.parametrize("alg,iters", [("BFGS", 60), ("ADAM", 200)])
@pytest.mark.parametrize("variant", [0, 3])
def test_reference_optimizer_over_gpu_c1(refgpu, alg, iters, variant):
    """config 1 structure (n = 4, data/Umtx.mat, adaptive L = 3, P = 138): the reference's BFGS / ADAM with the GPU cost path"""
    n, circ, U = c1_case(3)
    d, pool = circ.descriptors(nested=True)
    x0 = H.random_params(circ.get_Parameter_Num(), seed=5)
    r = _teacher_forced(refgpu, n, d, pool, U, x0, refgpu_alg(alg), iters, variant)
    assert r["evals"] >= iters
    # the two free-running trajectories stay together: same start, the first iterates agree to rounding, and the optimizer
    # reaches the same cost
    assert r["drift"][0] == 0.0 and r["drift"][: min(10, len(r["drift"]))].max() < 1e-8
    assert abs(r["f_gpu"] - r["f_cpu"]) < 1e-6 * max(1.0, abs(r["f_cpu"]))


@pytest.mark.gpu
@pytest.mark.parametrize("alg,iters", [("BFGS", 40), ("ADAM", 100)])
def test_reference_optimizer_over_gpu_c2(refgpu, alg, iters):
    """config 2: re-optimisation of the 19-CNOT circuit (109 gates, 172 parameters, Hilbert-Schmidt test cost) from a perturbed
    start, reference optimizer over the GPU cost path"""
    g = c2_case()
    rng = np.random.default_rng(2)
    x0 = g.params[0] + 0.05 * rng.standard_normal(g.P)
    r = _teacher_forced(refgpu, g.n, g.descs, g.pool, g.U, x0, refgpu_alg(alg), iters, 3)
    assert r["evals"] >= iters
    assert r["log_gpu"]["cost"][-1] < r["log_gpu"]["cost"][0]  # it optimises
    assert r["drift"][: min(10, len(r["drift"]))].max() < 1e-8


@pytest.mark.gpu
def test_gpu_flavour_batched_and_scalar_hooks(refgpu):
    """the other hooks: scalar optimization_problem (virtual) and the body of the non-virtual batched hook"""
    n, circ, U = c1_case(3)
    d, pool = circ.descriptors(nested=True)
    P = circ.get_Parameter_Num()
    ps = H.random_params(P, seed=9, batch=7)
    gpu = refgpu.session(True, U, n, d, pool)
    cpu = refgpu.session(False, U, n, d, pool)
    for variant in (0, 1, 2, 3, 4, 5, 9):
        for s in (gpu, cpu):
            s.set_cost(variant, 0, 0.37)
        fb_g, fb_c = gpu.cost_batched(ps), cpu.cost_batched(ps)
        assert close_rel(fb_g, fb_c)
        assert close_rel(gpu.cost(ps[0]), cpu.cost(ps[0]))
        fg, gg = gpu.cost_grad(ps[1])
        fc, gc = cpu.cost_grad(ps[1])
        assert close_rel(fg, fc) and close_rel(gg, gc)


def test_adam_mirror_matches_reference_class():
    """oracle/sq_oracle.c: sqo_adam_update (the host mirror of the device-resident ADAM loop) against the reference's own Adam
    class (common/Adam.cpp) over 300 updates of a quadratic: parameters within 1e-14, identical status flags. Run in a
    subprocess with SQREF_SERIAL=1: the reference advances its bias-correction members inside a TBB parallel_for, which is
    deterministic only on one thread."""
    import subprocess
    import sys

    import pyoracle

    if not pyoracle.RefGpu.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libsqref_gpu.so not built and /root/reference absent")
    code = r'''
import sys
sys.path[:0] = [%r, %r, %r]
import numpy as np, pyoracle
port, rg = pyoracle.Port(), pyoracle.RefGpu()
n = 37
rng = np.random.default_rng(0)
a, b = port.adam(n), rg.adam(n)
x1 = rng.normal(size=n); x2 = x1.copy()
A = rng.normal(size=(n, n)); A = A @ A.T / n
worst = 0.0
for it in range(300):
    s1 = a.update(x1, A @ x1, float(0.5 * x1 @ A @ x1))
    s2 = b.update(x2, A @ x2, float(0.5 * x2 @ A @ x2))
    assert s1 == s2, (it, s1, s2)
    worst = max(worst, float(np.abs(x1 - x2).max()))
assert worst < 1e-14, worst
print("ADAM_MIRROR_OK", worst)
''' % tuple(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d) for d in ("", "oracle", "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, SQREF_SERIAL="1"), timeout=300)
    assert r.returncode == 0 and "ADAM_MIRROR_OK" in r.stdout, r.stdout + r.stderr
