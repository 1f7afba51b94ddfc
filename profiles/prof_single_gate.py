"""ncu driver for the HBM-bound single-gate streaming kernels: U3 / CRY / CNOT / a 2-qubit RXX on a 2^12 x 2^12 matrix (256 MiB,
larger than L2) resident on the device, through sqgpu_apply_gate_dev."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import helpers as H
import squander_b200 as sq

abi = sq.abi
n = 12
rows = cols = 1 << n
eng = sq.Engine(0)
buf = torch.zeros(rows * cols * 2, dtype=torch.float64, device="cuda")
buf[0::2] = 1.0 / np.sqrt(rows)
stream = torch.cuda.current_stream().cuda_stream
for name, qs in (("U3", [n - 1, 0, 1]), ("U3", [0, 1, 2]), ("CRY", [1, n - 1, 0]), ("CNOT", [0, n // 2, 1]), ("RXX", [2, 9, 0])):
    circ = sq.Circuit(n)
    H.add_named(circ, name, qs)
    d, pool = circ.descriptors()
    gp = np.array([0.3, 0.7, 1.1][: int(d[0]["n_params"])], dtype=np.float64)
    dd = np.ascontiguousarray(d[:1])
    args = (eng._h, dd.ctypes.data_as(C.POINTER(abi.GateDesc)), abi.as_dp(gp) if gp.size else None, None, -1, buf.data_ptr(), rows, cols, cols, stream)
    for _ in range(3):
        abi.check(eng.lib, eng.lib.sqgpu_apply_gate_dev(*args))
    torch.cuda.synchronize()
    print(name, qs[:2], eng.last_kernel_time())
