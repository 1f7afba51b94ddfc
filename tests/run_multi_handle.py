"""The in-library multi-device handle (sqgpu_create_multi) on hardware, one process, G >= 2 GPUs:

    python tests/run_multi_handle.py [G]

cost and cost+gradient through ONE handle over G devices, batch- and column-sharded (one ncclAllReduce of the raw traces per
evaluation, two for the Hilbert-Schmidt correction variants), every device cost variant, against the CPU ORACLE at 1e-10 and
against a single-device handle; VQE parameter sets sharded over the devices; N_Qubit_Decomposition_custom(accelerator_num = G).
tests/test_gpu_parity.py::test_multi_device_handle spawns it when two devices are visible."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np

import helpers as H
import pyoracle
import squander_b200 as sq

abi = sq.abi
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
port = pyoracle.Port()
ok = True


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max()))


n = 6
circ = H.adaptive_circuit(n, 2)
P = circ.get_Parameter_Num()
descs, pool = circ.descriptors()
U = H.random_unitary(1 << n).conj().T.copy()
params = H.random_params(P, batch=7)  # not a multiple of G: uneven batch shards
single = sq.Engine(0)
single.upload_matrix(U)
single.set_circuit(circ)
for mode, mode_name in ((abi.SHARD_BATCH, "batch"), (abi.SHARD_COLUMNS, "columns")):
    multi = sq.Engine(devices=G, shard_mode=mode)
    multi.upload_matrix(U)
    multi.set_circuit(circ)
    assert multi.multi_info() == (G, mode)
    for variant in (0, 1, 2, 3, 4, 5, 6, 9):
        for e in (single, multi):
            e.set_cost(variant, 0, 0.37)
        c1, g1 = single.cost_grad_batched(params)
        cm, gm = multi.cost_grad_batched(params)
        cc = multi.cost_batched(params)
        orc = [port.cost_grad(descs, P, params[b], U, n, variant, 0, 0.37) for b in range(len(params))]
        e_orc = max(rel_err(cm, [o[0] for o in orc]), rel_err(gm, [o[1] for o in orc]), rel_err(cc, [o[0] for o in orc]))
        e_one = max(np.abs(cm - c1).max(), np.abs(gm - g1).max())
        good = e_orc <= 1e-10 and (e_one == 0.0 if mode == abi.SHARD_BATCH else e_one <= 1e-12)
        ok = ok and good
        print("multi handle G=%d %-7s variant %d: vs oracle (rel) %.2e, vs one device %.2e %s" % (G, mode_name, variant, e_orc, e_one, "ok" if good else "FAIL"), flush=True)
    # trace offset + rectangular matrix through the sharded handle (Frobenius family)
    Ur = np.ascontiguousarray(U[:, 8:8 + 4 * G])
    multi.upload_matrix(Ur)
    multi.set_cost(0, 8)
    cm, gm = multi.cost_grad_batched(params[:2])
    orc = [port.cost_grad(descs, P, params[b], Ur, n, 0, 8) for b in range(2)]
    e_orc = max(rel_err(cm, [o[0] for o in orc]), rel_err(gm, [o[1] for o in orc]))
    ok = ok and e_orc <= 1e-10
    print("multi handle G=%d %-7s rectangular + trace_offset: vs oracle %.2e %s" % (G, mode_name, e_orc, "ok" if e_orc <= 1e-10 else "FAIL"), flush=True)
    # the per-device entry points refuse a multi handle, loudly
    try:
        multi.apply(params[0], np.eye(1 << n, dtype=np.complex128))
        ok = False
    except abi.SqgpuError as ex:
        assert ex.status == abi.ERR_UNSUPPORTED
    multi.close()
single.close()

# AUTO: tall matrices by columns, small ones by batch
auto = sq.Engine(devices=G)
auto.upload_matrix(np.eye(64, dtype=np.complex128))
assert auto.multi_info()[1] == abi.SHARD_BATCH
auto.upload_matrix(np.zeros((1 << 11, 1 << 11), dtype=np.complex128))
assert auto.multi_info()[1] == abi.SHARD_COLUMNS
auto.close()

# VQE: parameter sets over the devices
nv = 10
vc = H.hea_zyz_circuit(nv, 2)
psi0 = np.zeros(1 << nv, dtype=np.complex128)
psi0[0] = 1.0
ip, ix, dv = H.heisenberg_csr_fast(nv)
vp = H.random_params(vc.get_Parameter_Num(), seed=3, batch=5)
mv = sq.Engine(devices=G)
mv.upload_matrix(psi0)
mv.set_circuit(vc)
mv.set_hamiltonian_csr(ip, ix, dv)
en, gr = mv.vqe_energy_grad_batched(vp)
vd = vc.descriptors()[0]
vo = [port.vqe_energy_grad(vd, vc.get_Parameter_Num(), vp[b], psi0, ip, ix, dv) for b in range(len(vp))]
e_orc = max(rel_err(en, [o[0] for o in vo]), rel_err(gr, [o[1] for o in vo]), rel_err(mv.vqe_energy_batched(vp), [o[0] for o in vo]))
ok = ok and e_orc <= 1e-10
print("multi handle G=%d VQE: vs oracle (rel) %.2e %s" % (G, e_orc, "ok" if e_orc <= 1e-10 else "FAIL"), flush=True)
mv.close()

# the wrapper class with accelerator_num = G
dec = sq.N_Qubit_Decomposition_custom(U, accelerator_num=G)
dec.set_Gate_Structure(circ)
dec.set_Cost_Function_Variant(3)
f, g = dec.Optimization_Problem_Combined(params[0])
f_ref, g_ref = port.cost_grad(descs, P, params[0], U, n, 3)
e_orc = max(rel_err(f, f_ref), rel_err(g, g_ref), rel_err(dec.get_Matrix(params[0]), port.apply_circuit(descs, params[0], np.eye(1 << n, dtype=np.complex128))))
ok = ok and e_orc <= 1e-10
print("N_Qubit_Decomposition_custom(accelerator_num=%d): vs oracle (rel) %.2e %s" % (G, e_orc, "ok" if e_orc <= 1e-10 else "FAIL"), flush=True)
print("MULTI_HANDLE_OK" if ok else "MULTI_HANDLE_FAIL", flush=True)
sys.exit(0 if ok else 1)
