"""Compare builds of the library (kernel-experiment variants made with OUT=... build.sh -D...) on the GPU:
    python profiles/variants.py [lib.so ...]          (default: every csrc/var_*.so and libsqgpu.so)
Each variant runs in its own process (SQGPU_LIB): C3 cost+grad batch 256 (device-timed), C4 cost batch 64, C5 energy+grad
64 sets. One JSON line per variant."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sequential-quantum-gate-decomposer_b200", "csrc")

WORKER = r'''
import sys, os, json, time
sys.path[:0] = [%(root)r, os.path.join(%(root)r, "tests")]
import numpy as np, torch, helpers as H, squander_b200 as sq
which = %(which)r
out = {"lib": os.path.basename(os.environ["SQGPU_LIB"])}
st = torch.cuda.current_stream()
def timed(fn, reps=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
if "c3" in which:
    n, L, B = 10, 4, 256
    c = H.adaptive_circuit(n, L); P = c.get_Parameter_Num()
    e = sq.Engine(0); e.upload_matrix(np.ascontiguousarray(H.random_unitary(1 << n, seed=123).conj().T)); e.set_circuit(c); e.set_cost(0, 0)
    p = torch.from_numpy(np.random.default_rng(42).random((B, P)) * 2 * np.pi).cuda()
    o = torch.zeros(B * (1 + P), dtype=torch.float64, device="cuda")
    ms = timed(lambda: e.cost_grad_batched_dev(p.data_ptr(), B, o.data_ptr(), o.data_ptr() + 8 * B, st.cuda_stream))
    out["c3_evals_per_s"] = round(B / ms * 1e3, 1); out["c3_cost0"] = float(o[0].item()); out["c3_gradsum"] = float(o[B:].sum().item())
    ms1 = timed(lambda: e.cost_grad_batched_dev(p.data_ptr(), 1, o.data_ptr(), o.data_ptr() + 8 * B, st.cuda_stream), reps=20)
    out["c3_batch1_ms"] = round(ms1, 3)
    e.close()
if "c4" in which:
    n, M, B = 12, 64, 64
    rng = np.random.default_rng(7); c = sq.Circuit(n)
    for m in range(M):
        c.add_GENERAL(H.random_unitary(16, seed=1000 + m), sorted(int(q) for q in rng.choice(n, 4, replace=False)))
        if m %% 8 == 7:
            for q in range(n): c.add_U3(q)
    e = sq.Engine(0)
    rs = np.random.default_rng(1); U = (rs.normal(size=(1 << n, 1 << n)) + 1j * rs.normal(size=(1 << n, 1 << n))) / np.sqrt(2 << n)
    e.upload_matrix(U); e.set_circuit(c); e.set_cost(0, 0)
    p = torch.from_numpy(H.random_params(c.get_Parameter_Num(), batch=B)).cuda(); o = torch.zeros(B, dtype=torch.float64, device="cuda")
    ms = timed(lambda: e.cost_batched_dev(p.data_ptr(), B, o.data_ptr(), st.cuda_stream), reps=2, warm=1)
    out["c4_evals_per_s"] = round(B / ms * 1e3, 1); out["c4_cost0"] = float(o[0].item())
    e.close()
if "c5" in which:
    n, layers, B = 20, 10, 64
    ip, ix, dat = H.heisenberg_csr_fast(n); c = H.hea_zyz_circuit(n, layers); P = c.get_Parameter_Num()
    psi0 = np.zeros(1 << n, dtype=np.complex128); psi0[0] = 1
    e = sq.Engine(0); e.upload_matrix(psi0); e.set_circuit(c); e.set_hamiltonian_csr(ip, ix, dat)
    p = torch.from_numpy(H.random_params(P, seed=5, batch=B)).cuda(); o = torch.zeros(B * (1 + P), dtype=torch.float64, device="cuda")
    ms = timed(lambda: e.vqe_energy_grad_batched_dev(p.data_ptr(), B, o.data_ptr(), o.data_ptr() + 8 * B, st.cuda_stream), reps=2, warm=1)
    out["c5_evals_per_s"] = round(B / ms * 1e3, 1); out["c5_e0"] = float(o[0].item()); out["c5_gradsum"] = float(o[B:].sum().item())
    out["c5_fwd_ms"] = round(e.kernel_time("fused_exec<WINDOW_FWD>")[0], 3); out["c5_bwd_ms"] = round(e.kernel_time("fused_exec<WINDOW_BWD>")[0], 3)
    mse = timed(lambda: e.vqe_energy_batched_dev(p.data_ptr(), B, o.data_ptr(), st.cuda_stream), reps=2, warm=1)
    out["c5_energy_evals_per_s"] = round(B / mse * 1e3, 1)
    e.close()
print(json.dumps(out))
'''

if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    which = [a[2:] for a in sys.argv[1:] if a.startswith("--")] or ["c3", "c4", "c5"]
    libs = args or sorted(glob.glob(os.path.join(CSRC, "var_*.so"))) + [os.path.join(CSRC, "libsqgpu.so")]
    for lib in libs:
        env = dict(os.environ, SQGPU_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT, "which": which}], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"lib": os.path.basename(lib), "error": r.stderr[-400:]}), flush=True)
