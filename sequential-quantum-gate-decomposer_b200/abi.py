"""ctypes view of include/sqgpu.h: the gate descriptor, the enums and the loader of libsqgpu.so.

The constants are numerically identical to the reference enums
(squander/src-cpp/gates/include/Gate.h:39-79, decomposition/include/Optimization_Interface.h:43-45).
There is no fallback: if the CUDA library is not built, ``load_library`` raises.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# SQGPU_LIB selects another build of the same library (kernel-experiment variants, profiles/variants.py); host-side only
LIB_PATH = os.environ.get("SQGPU_LIB") or os.path.join(PKG_DIR, "csrc", "libsqgpu.so")

# ---- enum sqgpu_gate_type -------------------------------------------------------------------------------------
GENERAL = 1
CZ = 4
CNOT = 5
CH = 6
U3 = 7
RY = 8
RX = 9
RZ = 10
X = 12
SX = 13
CRY = 14
SYC = 15
BLOCK = 16
ADAPTIVE = 18
Y = 23
Z = 24
H = 25
CROT = 27
R = 28
T = 29
TDG = 30
U1 = 31
U2 = 32
CR = 33
S = 34
SDG = 35
CU = 36
CP = 38
CRX = 39
CRZ = 40
CCX = 41
SWAP = 42
CSWAP = 43
RXX = 44
RYY = 45
RZZ = 46
SXDG = 47
BLOCK_BEGIN = 1001
BLOCK_END = 1002

GATE_NAMES = {
    GENERAL: "GENERAL", CZ: "CZ", CNOT: "CNOT", CH: "CH", U3: "U3", RY: "RY", RX: "RX", RZ: "RZ", X: "X", SX: "SX",
    CRY: "CRY", SYC: "SYC", ADAPTIVE: "Adaptive", Y: "Y", Z: "Z", H: "H", CROT: "CROT", R: "R", T: "T", TDG: "Tdg",
    U1: "U1", U2: "U2", CR: "CR", S: "S", SDG: "Sdg", CU: "CU", CP: "CP", CRX: "CRX", CRZ: "CRZ", CCX: "CCX",
    SWAP: "SWAP", CSWAP: "CSWAP", RXX: "RXX", RYY: "RYY", RZZ: "RZZ", SXDG: "SXdg",
}

# parameter_num of every gate class (grep `parameter_num =` in squander/src-cpp/gates/*.cpp)
PARAM_COUNT = {
    U3: 3, CU: 4, U2: 2, R: 2, CR: 2, CROT: 2,
    RX: 1, RY: 1, RZ: 1, U1: 1, CRY: 1, CRX: 1, CRZ: 1, CP: 1, ADAPTIVE: 1, RXX: 1, RYY: 1, RZZ: 1,
    GENERAL: 0, CZ: 0, CNOT: 0, CH: 0, X: 0, Y: 0, Z: 0, H: 0, S: 0, SDG: 0, T: 0, TDG: 0, SX: 0, SXDG: 0, SYC: 0,
    CCX: 0, SWAP: 0, CSWAP: 0,
}

# ---- enum sqgpu_cost_variant ----------------------------------------------------------------------------------
FROBENIUS_NORM = 0
FROBENIUS_NORM_CORRECTION1 = 1
FROBENIUS_NORM_CORRECTION2 = 2
HILBERT_SCHMIDT_TEST = 3
HILBERT_SCHMIDT_TEST_CORRECTION1 = 4
HILBERT_SCHMIDT_TEST_CORRECTION2 = 5
SUM_OF_SQUARES = 6
INFIDELITY = 9

# ---- enum sqgpu_shard_mode ------------------------------------------------------------------------------------
SHARD_AUTO = 0
SHARD_BATCH = 1
SHARD_COLUMNS = 2

# ---- enum sqgpu_status ----------------------------------------------------------------------------------------
OK = 0
ERR_NO_DEVICE = -1
ERR_INVALID = -2
ERR_STATE = -3
ERR_CUDA = -4
ERR_UNSUPPORTED = -5
ERR_NOMEM = -6


class GateDesc(C.Structure):
    """struct sqgpu_gate_desc (include/sqgpu.h)."""

    _fields_ = [
        ("type", C.c_int32),
        ("target", C.c_int32),
        ("control", C.c_int32),
        ("target2", C.c_int32),
        ("control2", C.c_int32),
        ("param_start", C.c_int32),
        ("n_params", C.c_int32),
        ("n_qubits", C.c_int32),
        ("qubits", C.c_int32 * 8),
        ("matrix_off", C.c_int64),
    ]


# numpy mirror of the same struct, for bulk construction
GATE_DESC_DTYPE = np.dtype(
    [
        ("type", "<i4"), ("target", "<i4"), ("control", "<i4"), ("target2", "<i4"), ("control2", "<i4"),
        ("param_start", "<i4"), ("n_params", "<i4"), ("n_qubits", "<i4"), ("qubits", "<i4", (8,)),
        ("matrix_off", "<i8"),
    ],
    align=True,
)
assert GATE_DESC_DTYPE.itemsize == C.sizeof(GateDesc) == 72


class SqgpuError(Exception):
    """Raised for every non-zero sqgpu_status; mirrors the reference wrappers turning std::string into a Python
    Exception (qgd_N_Qubit_Decompositions_Wrapper.cpp:1471-1483)."""

    def __init__(self, status, text):
        super().__init__("sqgpu status %d: %s" % (status, text))
        self.status = status
        self.text = text


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_handle = C.c_void_p

# name -> (restype, argtypes); must list EVERY symbol include/sqgpu.h declares (tests/test_abi_symbols.py checks)
PROTOTYPES = {
    "sqgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sqgpu_create": (C.c_int, [C.c_int, C.POINTER(_handle)]),
    "sqgpu_destroy": (C.c_int, [_handle]),
    "sqgpu_create_multi": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(_handle)]),
    "sqgpu_multi_info": (C.c_int, [_handle, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sqgpu_set_shard": (C.c_int, [_handle, C.c_int, C.c_int]),
    "sqgpu_grad_traces_with_global_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqgpu_last_error": (C.c_char_p, []),
    "sqgpu_abi_version": (C.c_int, []),
    "sqgpu_upload_matrix": (C.c_int, [_handle, _dp, C.c_int, C.c_int, C.c_int]),
    "sqgpu_set_circuit": (C.c_int, [_handle, C.POINTER(GateDesc), C.c_int, C.c_int, C.c_int, _dp, C.c_int64]),
    "sqgpu_plan_stats": (C.c_int, [C.POINTER(GateDesc), C.c_int, C.c_int, C.c_int, _dp, C.c_int64, C.POINTER(C.c_int64), C.c_int]),
    "sqgpu_plan_stats_opt": (C.c_int, [C.POINTER(GateDesc), C.c_int, C.c_int, C.c_int, _dp, C.c_int64, C.c_char_p, C.POINTER(C.c_int64), C.c_int]),
    "sqgpu_plan_ops": (C.c_int, [C.POINTER(GateDesc), C.c_int, C.c_int, C.c_int, _dp, C.c_int64, C.c_char_p, C.c_int, _ip, C.c_int,
                                 C.POINTER(C.c_int)]),
    "sqgpu_set_option": (C.c_int, [_handle, C.c_char_p, C.c_int64]),
    "sqgpu_get_option": (C.c_int, [_handle, C.c_char_p, C.POINTER(C.c_int64)]),
    "sqgpu_set_cost": (C.c_int, [_handle, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "sqgpu_cost_batched": (C.c_int, [_handle, _dp, C.c_int, _dp]),
    "sqgpu_cost_grad_batched": (C.c_int, [_handle, _dp, C.c_int, _dp, _dp]),
    "sqgpu_traces_batched": (C.c_int, [_handle, _dp, C.c_int, C.c_int, _dp]),
    "sqgpu_cost_from_traces": (C.c_int, [_handle, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "sqgpu_apply": (C.c_int, [_handle, _dp, _dp, C.c_int, C.c_int, C.c_int]),
    "sqgpu_apply_derivative": (C.c_int, [_handle, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp]),
    "sqgpu_apply_gate": (C.c_int, [_handle, C.POINTER(GateDesc), _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int]),
    "sqgpu_set_hamiltonian_csr": (C.c_int, [_handle, C.c_int, C.c_int64, _ip, _ip, _dp]),
    "sqgpu_vqe_energy_batched": (C.c_int, [_handle, _dp, C.c_int, _dp]),
    "sqgpu_vqe_energy_grad_batched": (C.c_int, [_handle, _dp, C.c_int, _dp, _dp]),
    "sqgpu_cost_batched_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sqgpu_cost_grad_batched_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqgpu_traces_batched_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sqgpu_cost_from_traces_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "sqgpu_apply_gate_dev": (C.c_int, [_handle, C.POINTER(GateDesc), _dp, _dp, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_void_p]),
    "sqgpu_vqe_energy_batched_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "sqgpu_vqe_energy_grad_batched_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqgpu_kernel_time": (C.c_int, [_handle, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "sqgpu_last_launch_shape": (C.c_int, [_handle, C.POINTER(C.c_int), C.c_int]),
    "sqgpu_last_exec_flops": (C.c_int, [_handle, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "sqgpu_adam_init": (C.c_int, [_handle, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "sqgpu_adam_steps": (C.c_int, [_handle, C.c_int, _dp]),
    "sqgpu_adam_get": (C.c_int, [_handle, _dp, _dp, _dp, C.POINTER(C.c_int)]),
    "sqgpu_line_search_batched": (C.c_int, [_handle, _dp, _dp, _dp, C.c_int, _dp, _dp]),
    "sqgpu_cost_shifted_batched": (C.c_int, [_handle, _dp, C.c_int, _dp, C.c_int, _dp, _dp]),
    "sqgpu_cost_shifted_batched_dev": (C.c_int, [_handle, C.c_void_p, C.c_int, _dp, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sqgpu_launch_count": (C.c_int, [_handle, C.POINTER(C.c_int64)]),
    "sqgpu_last_kernel_time": (C.c_int, [_handle, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "sqgpu_fp64_fma_peak": (C.c_int, [_handle, C.POINTER(C.c_double)]),
}

_lib = None
ABI_VERSION = 2


def load_library(path=None):
    """dlopen libsqgpu.so (built in-tree by __graft_entry__.build()). Raises if it is missing: the product has no
    CPU path to fall back to."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s is not built; run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU fallback for the sqgpu engine." % p
        )
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.sqgpu_abi_version() != ABI_VERSION:
        raise RuntimeError("libsqgpu.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(lib, status):
    if status != OK:
        raise SqgpuError(status, lib.sqgpu_last_error().decode("utf-8", "replace"))


def as_dp(a):
    return a.ctypes.data_as(_dp)


def as_ip(a):
    return a.ctypes.data_as(_ip)


PLAN_STATS = ("ops_plan2", "ops_plan3", "segments", "max_segment_ops", "window", "kern_total", "dkern_total", "w_total",
              "dense_ops", "block_members")


def plan_stats(circuit, **options):
    """What the host planner of libsqgpu.so makes of a Circuit (sqgpu_plan_stats_opt): needs no CUDA device.
    Keyword arguments are planner options (names of sqgpu_set_option), e.g. ``plan_stats(c, window=10)``."""
    import numpy as np

    lib = load_library()
    descs, pool = circuit.descriptors()
    descs = np.ascontiguousarray(descs, dtype=GATE_DESC_DTYPE)
    pool = np.ascontiguousarray(pool, dtype=np.complex128)
    out = (C.c_int64 * len(PLAN_STATS))()
    opts = ",".join("%s=%d" % (k, int(v)) for k, v in options.items()).encode() or None
    check(lib, lib.sqgpu_plan_stats_opt(descs.ctypes.data_as(C.POINTER(GateDesc)), len(descs), circuit.get_Parameter_Num(),
                                        circuit.qbit_num, as_dp(pool.view(np.float64)) if pool.size else None, pool.size,
                                        opts, out, len(PLAN_STATS)))
    return dict(zip(PLAN_STATS, (int(v) for v in out)))


def plan_ops(circuit, which=3, **options):
    """[(dim, [qubits], n_params, n_members)] of the planner's op list (sqgpu_plan_ops); which = 2, 3 or 0 (window plan)"""
    lib = load_library()
    descs, pool = circuit.descriptors()
    descs = np.ascontiguousarray(descs, dtype=GATE_DESC_DTYPE)
    pool = np.ascontiguousarray(pool, dtype=np.complex128)
    cap = 2 * len(descs) + 1  # (cluster plans add RESPLIT ops)
    out = np.zeros((cap, 8), dtype=np.int32)
    n = C.c_int(0)
    opts = ",".join("%s=%d" % (k, int(v)) for k, v in options.items()).encode() or None
    check(lib, lib.sqgpu_plan_ops(descs.ctypes.data_as(C.POINTER(GateDesc)), len(descs), circuit.get_Parameter_Num(),
                                  circuit.qbit_num, as_dp(pool.view(np.float64)) if pool.size else None, pool.size, opts,
                                  int(which), as_ip(out), cap, C.byref(n)))
    return [(int(r[0]), [int(q) for q in r[1:6] if q >= 0], int(r[6]), int(r[7])) for r in out[: n.value]]
