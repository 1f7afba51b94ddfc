"""ctypes bindings of the two CPU checkers. TEST INFRASTRUCTURE ONLY.

  * ``Port``  -- oracle/libsqoracle.so, the plain-C restatement (sq_oracle.c); always available (gcc).
  * ``Ref``   -- oracle/_ref/libsqref.so, the reference's OWN translation units behind ref_harness.cpp; built in the
                 container that has /root/reference, shipped prebuilt to the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module; the
product package never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import importlib

abi = importlib.import_module("sequential-quantum-gate-decomposer_b200.abi")

PORT_PATH = os.path.join(HERE, "libsqoracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libsqref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_gp = C.POINTER(abi.GateDesc)


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def build_ref():
    """Only possible where /root/reference exists."""
    subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def _c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dptr(a):
    return a.ctypes.data_as(_dp) if a is not None and a.size else None


def _descs(d):
    d = np.ascontiguousarray(d, dtype=abi.GATE_DESC_DTYPE)
    return d, d.ctypes.data_as(_gp)


class Port:
    """sq_oracle.c"""

    def __init__(self):
        if not os.path.exists(PORT_PATH) or os.path.getmtime(PORT_PATH) < os.path.getmtime(os.path.join(HERE, "sq_oracle.c")):
            build_port()
        L = self.lib = C.CDLL(PORT_PATH)
        L.sqo_gate_kernel.argtypes = [C.c_int, _dp, _dp]
        L.sqo_gate_derivative_kernel.argtypes = [C.c_int, _dp, C.c_int, _dp]
        L.sqo_apply_gate.argtypes = [_gp, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int]
        L.sqo_apply_circuit.argtypes = [_gp, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int]
        L.sqo_apply_derivate.argtypes = [_gp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp]
        L.sqo_traces.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.sqo_cost_from_traces.argtypes = [C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.sqo_cost_from_traces.restype = C.c_double
        L.sqo_grad_from_traces.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.sqo_grad_from_traces.restype = C.c_double
        L.sqo_cost.argtypes = [_gp, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_double, C.c_double, C.c_double, _dp]
        L.sqo_cost_grad.argtypes = [_gp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.sqo_csr_matvec.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.sqo_vqe_energy.argtypes = [_gp, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, _dp, _dp]
        L.sqo_vqe_energy_grad.argtypes = [_gp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.sqo_vqe_energy_grad_sampled.argtypes = [_gp, C.c_int, _dp, _dp, _dp, C.c_int, _ip, _ip, _dp, _ip, C.c_int, _dp, _dp]

    def gate_kernel(self, type_, gate_params):
        k = np.zeros(16, dtype=np.complex128)
        p = _f64(gate_params)
        dim = self.lib.sqo_gate_kernel(type_, _dptr(p), _dptr(k.view(np.float64)))
        if dim < 0:
            raise Exception("port: no kernel for gate type %d" % type_)
        return k[: dim * dim].reshape(dim, dim).copy()

    def gate_derivative_kernel(self, type_, gate_params, pidx):
        k = np.zeros(16, dtype=np.complex128)
        p = _f64(gate_params)
        dim = self.lib.sqo_gate_derivative_kernel(type_, _dptr(p), pidx, _dptr(k.view(np.float64)))
        if dim < 0:
            raise Exception("port: no derivative kernel for gate type %d" % type_)
        return k[: dim * dim].reshape(dim, dim).copy()

    def apply_gate(self, desc_row, params, mtx, pool=None, deriv_param=-1):
        """returns a transformed copy; params = the whole circuit parameter vector (desc.param_start indexes it)"""
        m = np.array(mtx, dtype=np.complex128, order="C", copy=True)
        m2 = m.reshape(m.shape[0], -1)
        d, dptr = _descs(np.asarray(desc_row).reshape(1))
        p = _f64(params)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        rc = self.lib.sqo_apply_gate(dptr, _dptr(p), _dptr(pl.view(np.float64)), deriv_param,
                                     _dptr(m2.view(np.float64)), m2.shape[0], m2.shape[1], m2.shape[1])
        if rc:
            raise Exception("port: apply_gate failed")
        return m

    def apply_circuit(self, descs, params, mtx, pool=None):
        m = np.array(mtx, dtype=np.complex128, order="C", copy=True)
        m2 = m.reshape(m.shape[0], -1)
        d, dptr = _descs(descs)
        p = _f64(params)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        rc = self.lib.sqo_apply_circuit(dptr, len(d), _dptr(p), _dptr(pl.view(np.float64)), _dptr(m2.view(np.float64)),
                                        m2.shape[0], m2.shape[1], m2.shape[1])
        if rc:
            raise Exception("port: apply_circuit failed")
        return m

    def apply_derivate(self, descs, n_params, params, mtx, pool=None):
        m = _c128(mtx)
        m2 = m.reshape(m.shape[0], -1)
        d, dptr = _descs(descs)
        p = _f64(params)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        out = np.zeros((max(n_params, 1),) + m2.shape, dtype=np.complex128)
        rc = self.lib.sqo_apply_derivate(dptr, len(d), n_params, _dptr(p), _dptr(pl.view(np.float64)),
                                         _dptr(m2.view(np.float64)), m2.shape[0], m2.shape[1], m2.shape[1],
                                         _dptr(out.view(np.float64)))
        if rc:
            raise Exception("port: apply_derivate failed")
        return out[:n_params].reshape((n_params,) + m.shape)

    def traces(self, mtx, qbit_num, trace_offset=0):
        m = _c128(mtx)
        out = np.zeros(6)
        self.lib.sqo_traces(_dptr(m.view(np.float64)), m.shape[0], m.shape[1], m.shape[1], qbit_num, trace_offset,
                            _dptr(out))
        return out

    def cost_from_traces(self, variant, tr6, cols, prev=1.0, c1=1 / 1.7, c2=0.5):
        t = _f64(tr6)
        return self.lib.sqo_cost_from_traces(variant, _dptr(t), cols, prev, c1, c2)

    def cost(self, descs, params, umtx, qbit_num, variant=0, trace_offset=0, prev=1.0, c1=1 / 1.7, c2=0.5, pool=None):
        U = _c128(umtx)
        d, dptr = _descs(descs)
        p = _f64(params)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        out = np.zeros(1)
        rc = self.lib.sqo_cost(dptr, len(d), _dptr(p), _dptr(pl.view(np.float64)), _dptr(U.view(np.float64)),
                               U.shape[0], U.shape[1], U.shape[1], qbit_num, variant, trace_offset, prev, c1, c2,
                               _dptr(out))
        if rc:
            raise Exception("port: cost failed")
        return float(out[0])

    def cost_grad(self, descs, n_params, params, umtx, qbit_num, variant=0, trace_offset=0, prev=1.0, c1=1 / 1.7,
                  c2=0.5, pool=None):
        U = _c128(umtx)
        d, dptr = _descs(descs)
        p = _f64(params)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        out = np.zeros(1)
        grad = np.zeros(max(n_params, 1))
        rc = self.lib.sqo_cost_grad(dptr, len(d), n_params, _dptr(p), _dptr(pl.view(np.float64)),
                                    _dptr(U.view(np.float64)), U.shape[0], U.shape[1], U.shape[1], qbit_num, variant,
                                    trace_offset, prev, c1, c2, _dptr(out), _dptr(grad))
        if rc:
            raise Exception("port: cost_grad failed")
        return float(out[0]), grad[:n_params]

    def vqe_energy(self, descs, params, state0, indptr, indices, data, pool=None):
        d, dptr = _descs(descs)
        p = _f64(params)
        s0 = _c128(state0).reshape(-1)
        ip = np.ascontiguousarray(indptr, dtype=np.int32)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        v = _c128(data)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        out = np.zeros(1)
        rc = self.lib.sqo_vqe_energy(dptr, len(d), _dptr(p), _dptr(pl.view(np.float64)), _dptr(s0.view(np.float64)),
                                     s0.size, ip.ctypes.data_as(_ip), ix.ctypes.data_as(_ip),
                                     _dptr(v.view(np.float64)), _dptr(out))
        if rc:
            raise Exception("port: vqe_energy failed")
        return float(out[0])

    def vqe_energy_grad(self, descs, n_params, params, state0, indptr, indices, data, pool=None):
        d, dptr = _descs(descs)
        p = _f64(params)
        s0 = _c128(state0).reshape(-1)
        ip = np.ascontiguousarray(indptr, dtype=np.int32)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        v = _c128(data)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        out = np.zeros(1)
        grad = np.zeros(max(n_params, 1))
        rc = self.lib.sqo_vqe_energy_grad(dptr, len(d), n_params, _dptr(p), _dptr(pl.view(np.float64)),
                                          _dptr(s0.view(np.float64)), s0.size, ip.ctypes.data_as(_ip),
                                          ix.ctypes.data_as(_ip), _dptr(v.view(np.float64)), _dptr(out), _dptr(grad))
        if rc:
            raise Exception("port: vqe_energy_grad failed")
        return float(out[0]), grad[:n_params]


    def adam(self, n_params, eta=1e-3, beta1=0.68, beta2=0.8, epsilon=1e-4):
        """host mirror of Adam::update (sequential semantics): object with .update(params, grad, f0) -> status, in place"""
        return PortAdam(self.lib, n_params, eta, beta1, beta2, epsilon)

    def vqe_energy_grad_sampled(self, descs, params, state0, indptr, indices, data, sample, pool=None):
        """(energy, grad[sample]): only the listed parameters' derivative states are formed"""
        d, dptr = _descs(descs)
        p = _f64(params)
        s0 = _c128(state0).reshape(-1)
        ip = np.ascontiguousarray(indptr, dtype=np.int32)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        v = _c128(data)
        sm = np.ascontiguousarray(sample, dtype=np.int32)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        out = np.zeros(1)
        grad = np.zeros(max(sm.size, 1))
        rc = self.lib.sqo_vqe_energy_grad_sampled(dptr, len(d), _dptr(p), _dptr(pl.view(np.float64)),
                                                  _dptr(s0.view(np.float64)), s0.size, ip.ctypes.data_as(_ip),
                                                  ix.ctypes.data_as(_ip), _dptr(v.view(np.float64)),
                                                  sm.ctypes.data_as(_ip), sm.size, _dptr(out), _dptr(grad))
        if rc:
            raise Exception("port: vqe_energy_grad_sampled failed")
        return float(out[0]), grad[: sm.size]


class PortAdam:
    class State(C.Structure):
        _fields_ = [("beta1_t", C.c_double), ("beta2_t", C.c_double), ("f0_mean", C.c_double), ("decreasing_test", C.c_double),
                    ("f0_prev", C.c_double), ("f0_idx", C.c_int), ("decreasing_idx", C.c_int), ("iter_t", C.c_int),
                    ("f0_vec", C.c_double * 100), ("decreasing_vec", C.c_int * 20)]

    def __init__(self, lib, n, eta, beta1, beta2, epsilon):
        self.lib, self.n, self.cfg = lib, n, (eta, beta1, beta2, epsilon)
        self.state = PortAdam.State()
        lib.sqo_adam_reset.argtypes = [C.POINTER(PortAdam.State)]
        lib.sqo_adam_update.argtypes = [C.POINTER(PortAdam.State), _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double]
        lib.sqo_adam_reset(C.byref(self.state))
        self.mom = np.zeros(n)
        self.var = np.zeros(n)

    def update(self, params, grad, f0):
        g = _f64(grad)
        assert params.dtype == np.float64 and params.flags["C_CONTIGUOUS"] and params.size == self.n
        return self.lib.sqo_adam_update(C.byref(self.state), _dptr(params), _dptr(g), _dptr(self.mom), _dptr(self.var), self.n,
                                        float(f0), *self.cfg)


class Ref:
    """the reference's own code (oracle/_ref/libsqref.so)"""

    @staticmethod
    def available():
        return os.path.exists(REF_PATH)

    def __init__(self):
        if not os.path.exists(REF_PATH):
            if os.path.isdir("/root/reference"):
                build_ref()
            else:
                raise FileNotFoundError(REF_PATH + " missing and /root/reference not present to build it")
        L = self.lib = C.CDLL(REF_PATH)
        L.sqref_last_error.restype = C.c_char_p
        L.sqref_circuit_create.restype = C.c_void_p
        L.sqref_circuit_create.argtypes = [C.c_int, _gp, C.c_int, _dp]
        L.sqref_circuit_free.argtypes = [C.c_void_p]
        L.sqref_circuit_param_num.argtypes = [C.c_void_p]
        L.sqref_circuit_set_min_fusion.argtypes = [C.c_void_p, C.c_int]
        L.sqref_circuit_apply.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int]
        L.sqref_circuit_apply_derivate.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp]
        L.sqref_decomp_create.restype = C.c_void_p
        L.sqref_decomp_create.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _gp, C.c_int, _dp]
        L.sqref_decomp_free.argtypes = [C.c_void_p]
        L.sqref_decomp_param_num.argtypes = [C.c_void_p]
        L.sqref_decomp_set_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.sqref_decomp_set_parallel.argtypes = [C.c_void_p, C.c_int]
        L.sqref_decomp_cost.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
        L.sqref_decomp_cost_batched.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp]
        L.sqref_decomp_cost_grad.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp]
        L.sqref_traces.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.sqref_vqe_create.restype = C.c_void_p
        L.sqref_vqe_create.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, C.c_int, C.c_int, _gp, C.c_int]
        L.sqref_vqe_free.argtypes = [C.c_void_p]
        L.sqref_vqe_param_num.argtypes = [C.c_void_p]
        L.sqref_vqe_energy.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
        L.sqref_vqe_energy_grad.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp]

    def _err(self):
        return self.lib.sqref_last_error().decode("utf-8", "replace")

    def circuit(self, qbit_num, descs, pool=None):
        return RefCircuit(self, qbit_num, descs, pool)

    def decomp(self, umtx, qbit_num, descs, pool=None):
        return RefDecomp(self, umtx, qbit_num, descs, pool)

    def vqe(self, qbit_num, indptr, indices, data, ansatz="HEA_ZYZ", layers=1, inner_blocks=1, descs=None):
        return RefVQE(self, qbit_num, indptr, indices, data, ansatz, layers, inner_blocks, descs)


class RefCircuit:
    def __init__(self, ref, qbit_num, descs, pool=None):
        self.ref = ref
        d, dptr = _descs(descs)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        self.h = ref.lib.sqref_circuit_create(qbit_num, dptr, len(d), _dptr(pl.view(np.float64)))
        if not self.h:
            raise Exception("ref: circuit_create failed: " + ref._err())
        self.n_params = ref.lib.sqref_circuit_param_num(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.ref.lib.sqref_circuit_free(self.h)
            self.h = None

    def set_min_fusion(self, mf):
        self.ref.lib.sqref_circuit_set_min_fusion(self.h, mf)

    def apply(self, params, mtx, parallel=0):
        m = np.array(mtx, dtype=np.complex128, order="C", copy=True)
        m2 = m.reshape(m.shape[0], -1)
        p = _f64(params)
        rc = self.ref.lib.sqref_circuit_apply(self.h, _dptr(p), p.size, _dptr(m2.view(np.float64)), m2.shape[0],
                                              m2.shape[1], parallel)
        if rc:
            raise Exception("ref: apply failed: " + self.ref._err())
        return m

    def apply_derivate(self, params, mtx, parallel=0):
        m = _c128(mtx)
        m2 = m.reshape(m.shape[0], -1)
        p = _f64(params)
        out = np.zeros((max(self.n_params, 1),) + m2.shape, dtype=np.complex128)
        rc = self.ref.lib.sqref_circuit_apply_derivate(self.h, _dptr(p), p.size, _dptr(m2.view(np.float64)),
                                                       m2.shape[0], m2.shape[1], parallel,
                                                       _dptr(out.view(np.float64)))
        if rc:
            raise Exception("ref: apply_derivate failed: " + self.ref._err())
        return out[: self.n_params].reshape((self.n_params,) + m.shape)


class RefDecomp:
    def __init__(self, ref, umtx, qbit_num, descs, pool=None):
        self.ref = ref
        U = _c128(umtx)
        d, dptr = _descs(descs)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        self.h = ref.lib.sqref_decomp_create(_dptr(U.view(np.float64)), U.shape[0], U.shape[1], qbit_num, dptr, len(d),
                                             _dptr(pl.view(np.float64)))
        if not self.h:
            raise Exception("ref: decomp_create failed: " + ref._err())
        self.n_params = ref.lib.sqref_decomp_param_num(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.ref.lib.sqref_decomp_free(self.h)
            self.h = None

    def set_cost(self, variant=0, trace_offset=0, prev=1.0, c1=1 / 1.7, c2=0.5):
        if self.ref.lib.sqref_decomp_set_cost(self.h, variant, trace_offset, prev, c1, c2):
            raise Exception("ref: set_cost failed: " + self.ref._err())

    def set_parallel(self, parallel):
        if self.ref.lib.sqref_decomp_set_parallel(self.h, parallel):
            raise Exception("ref: set_parallel failed: " + self.ref._err())

    def cost(self, params):
        p = _f64(params)
        out = np.zeros(1)
        if self.ref.lib.sqref_decomp_cost(self.h, _dptr(p), p.size, _dptr(out)):
            raise Exception("ref: cost failed: " + self.ref._err())
        return float(out[0])

    def cost_batched(self, params):
        p = _f64(params)
        out = np.zeros(p.shape[0])
        if self.ref.lib.sqref_decomp_cost_batched(self.h, _dptr(p), p.shape[1], p.shape[0], _dptr(out)):
            raise Exception("ref: cost_batched failed: " + self.ref._err())
        return out

    def cost_grad(self, params):
        p = _f64(params)
        out = np.zeros(1)
        g = np.zeros(max(p.size, 1))
        if self.ref.lib.sqref_decomp_cost_grad(self.h, _dptr(p), p.size, _dptr(out), _dptr(g)):
            raise Exception("ref: cost_grad failed: " + self.ref._err())
        return float(out[0]), g[: p.size]


class RefVQE:
    """Variational_Quantum_Eigensolver_Base with its own generated ansatz (or a custom gate structure)"""

    def __init__(self, ref, qbit_num, indptr, indices, data, ansatz, layers, inner_blocks, descs):
        self.ref = ref
        ip = np.ascontiguousarray(indptr, dtype=np.int32)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        v = _c128(data)
        if descs is not None:
            d, dptr = _descs(descs)
            nd = len(d)
        else:
            dptr, nd = None, 0
        self.h = ref.lib.sqref_vqe_create(qbit_num, len(ip) - 1, v.size, ip.ctypes.data_as(_ip), ix.ctypes.data_as(_ip),
                                          _dptr(v.view(np.float64)), 1 if ansatz == "HEA_ZYZ" else 0, layers,
                                          inner_blocks, dptr, nd)
        if not self.h:
            raise Exception("ref: vqe_create failed: " + ref._err())
        self.n_params = ref.lib.sqref_vqe_param_num(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.ref.lib.sqref_vqe_free(self.h)
            self.h = None

    def energy(self, params):
        p = _f64(params)
        out = np.zeros(1)
        if self.ref.lib.sqref_vqe_energy(self.h, _dptr(p), p.size, _dptr(out)):
            raise Exception("ref: vqe_energy failed: " + self.ref._err())
        return float(out[0])

    def energy_grad(self, params):
        p = _f64(params)
        out = np.zeros(1)
        g = np.zeros(max(p.size, 1))
        if self.ref.lib.sqref_vqe_energy_grad(self.h, _dptr(p), p.size, _dptr(out), _dptr(g)):
            raise Exception("ref: vqe_energy_grad failed: " + self.ref._err())
        return float(out[0]), g[: p.size]


REF_GPU_PATH = os.path.join(HERE, "_ref", "libsqref_gpu.so")


class RefGpu:
    """oracle/_ref/libsqref_gpu.so: the reference's own classes and optimizers with the drop-in of integration/ compiled in
    (oracle/ref_gpu_harness.cpp). ``session(use_gpu=...)`` gives the reference's N_Qubit_Decomposition_custom either as it is
    (CPU cost path) or as With_GPU_Cost_Path<N_Qubit_Decomposition_custom> (every cost / gradient call served by libsqgpu.so)."""

    @staticmethod
    def available():
        return os.path.exists(REF_GPU_PATH)

    def __init__(self):
        if not os.path.exists(REF_GPU_PATH):
            if os.path.isdir("/root/reference"):
                subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref_gpu"])
            else:
                raise FileNotFoundError(REF_GPU_PATH + " missing and /root/reference not present to build it")
        L = self.lib = C.CDLL(REF_GPU_PATH)
        L.sqrefgpu_last_error.restype = C.c_char_p
        L.sqrefgpu_set_library_path.argtypes = [C.c_char_p]
        L.sqrefgpu_session_create.restype = C.c_void_p
        L.sqrefgpu_session_create.argtypes = [C.c_int, _dp, C.c_int, C.c_int, C.c_int, _gp, C.c_int, _dp]
        L.sqrefgpu_session_free.argtypes = [C.c_void_p]
        L.sqrefgpu_param_num.argtypes = [C.c_void_p]
        L.sqrefgpu_set_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.sqrefgpu_cost_grad.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp]
        L.sqrefgpu_cost.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
        L.sqrefgpu_cost_batched.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp]
        L.sqrefgpu_optimize.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_longlong, C.c_double, _dp, _dp]
        L.sqrefgpu_get_log.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.sqrefgpu_gpu_evaluations.argtypes = [C.c_void_p]
        L.sqrefgpu_gpu_evaluations.restype = C.c_longlong
        L.sqrefgpu_flatten.argtypes = [C.c_int, _gp, C.c_int, _dp, _gp, C.c_int, _dp, C.c_longlong, C.POINTER(C.c_longlong)]
        L.sqrefgpu_adam_create.restype = C.c_void_p
        L.sqrefgpu_adam_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        L.sqrefgpu_adam_free.argtypes = [C.c_void_p]
        L.sqrefgpu_adam_update.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double]
        L.sqrefgpu_export_binary.argtypes = [C.c_int, _gp, C.c_int, _dp, C.c_int, C.c_char_p]
        L.sqrefgpu_set_library_path(abi.LIB_PATH.encode())

    def _err(self):
        return self.lib.sqrefgpu_last_error().decode("utf-8", "replace")

    def available_gpus(self):
        return int(self.lib.sqrefgpu_available_gpus())

    def adam(self, n_params, eta=1e-3, beta1=0.68, beta2=0.8, epsilon=1e-4):
        """the reference's own Adam object: .update(params, grad, f0) -> status, params updated in place"""
        ref = self

        class _A:
            def __init__(self):
                self.h = ref.lib.sqrefgpu_adam_create(beta1, beta2, epsilon, eta, n_params)

            def update(self, params, grad, f0):
                g = _f64(grad)
                return ref.lib.sqrefgpu_adam_update(self.h, _dptr(params), _dptr(g), n_params, float(f0))

            def __del__(self):
                if self.h:
                    ref.lib.sqrefgpu_adam_free(self.h)
                    self.h = None

        return _A()

    def export_binary(self, qbit_num, descs_nested, params, filename):
        """export_gate_list_to_binary of the reference on the structure given by a nested descriptor stream"""
        d, dptr = _descs(descs_nested)
        p = _f64(params)
        if self.lib.sqrefgpu_export_binary(qbit_num, dptr, len(d), _dptr(p), p.size, str(filename).encode()):
            raise Exception("ref_gpu: export_binary failed: " + self._err())

    def flatten(self, qbit_num, descs_nested, pool=None):
        """(flat descs, pool) as integration/common_GPU.cpp: to_gpu_gates makes them from the reference's Gates_block"""
        d, dptr = _descs(descs_nested)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        out = np.zeros(len(d) + 1, dtype=abi.GATE_DESC_DTYPE)
        pool_out = np.zeros(max(pl.size, 1), dtype=np.complex128)
        plen = C.c_longlong(0)
        n = self.lib.sqrefgpu_flatten(qbit_num, dptr, len(d), _dptr(pl.view(np.float64)), out.ctypes.data_as(_gp), len(out),
                                      _dptr(pool_out.view(np.float64)), pool_out.size, C.byref(plen))
        if n < 0:
            raise Exception("ref_gpu: flatten failed: " + self._err())
        return out[:n].copy(), pool_out[: plen.value].copy()

    def session(self, use_gpu, umtx, qbit_num, descs_nested, pool=None):
        return RefGpuSession(self, use_gpu, umtx, qbit_num, descs_nested, pool)


class RefGpuSession:
    ADAM, BFGS = 0, 1  # enum optimization_aglorithms (Optimization_Interface.h:50)

    def __init__(self, ref, use_gpu, umtx, qbit_num, descs, pool=None):
        self.ref = ref
        U = _c128(umtx)
        d, dptr = _descs(descs)
        pl = _c128(pool) if pool is not None else np.zeros(0, dtype=np.complex128)
        self.h = ref.lib.sqrefgpu_session_create(1 if use_gpu else 0, _dptr(U.view(np.float64)), U.shape[0], U.shape[1], qbit_num,
                                                 dptr, len(d), _dptr(pl.view(np.float64)))
        if not self.h:
            raise Exception("ref_gpu: session_create failed: " + ref._err())
        self.n_params = ref.lib.sqrefgpu_param_num(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.ref.lib.sqrefgpu_session_free(self.h)
            self.h = None

    def set_cost(self, variant=0, trace_offset=0, prev=1.0, c1=1 / 1.7, c2=0.5):
        if self.ref.lib.sqrefgpu_set_cost(self.h, variant, trace_offset, prev, c1, c2):
            raise Exception("ref_gpu: set_cost failed: " + self.ref._err())

    def cost_grad(self, params):
        p = _f64(params)
        out = np.zeros(1)
        g = np.zeros(max(p.size, 1))
        if self.ref.lib.sqrefgpu_cost_grad(self.h, _dptr(p), p.size, _dptr(out), _dptr(g)):
            raise Exception("ref_gpu: cost_grad failed: " + self.ref._err())
        return float(out[0]), g[: p.size]

    def cost(self, params):
        p = _f64(params)
        out = np.zeros(1)
        if self.ref.lib.sqrefgpu_cost(self.h, _dptr(p), p.size, _dptr(out)):
            raise Exception("ref_gpu: cost failed: " + self.ref._err())
        return float(out[0])

    def cost_batched(self, params):
        p = _f64(params)
        out = np.zeros(p.shape[0])
        if self.ref.lib.sqrefgpu_cost_batched(self.h, _dptr(p), p.shape[1], p.shape[0], _dptr(out)):
            raise Exception("ref_gpu: cost_batched failed: " + self.ref._err())
        return out

    def optimize(self, alg, x0, max_inner_iterations, eta=1e-3):
        """run the reference's optimizer; returns (x_final, f_min, log) with log = dict(params [n, P], cost [n], grad [n, P]) of
        every cost+gradient evaluation it made, in order"""
        x0 = _f64(x0)
        x = np.zeros_like(x0)
        f = np.zeros(1)
        n = self.ref.lib.sqrefgpu_optimize(self.h, int(alg), _dptr(x0), x0.size, int(max_inner_iterations), float(eta), _dptr(x), _dptr(f))
        if n < 0:
            raise Exception("ref_gpu: optimize failed: " + self.ref._err())
        P = x0.size
        lp, lc, lg = np.zeros((n, P)), np.zeros(n), np.zeros((n, P))
        if n and self.ref.lib.sqrefgpu_get_log(self.h, _dptr(lp), _dptr(lc), _dptr(lg)):
            raise Exception("ref_gpu: get_log failed: " + self.ref._err())
        return x, float(f[0]), {"params": lp, "cost": lc, "grad": lg}

    def gpu_evaluations(self):
        return int(self.ref.lib.sqrefgpu_gpu_evaluations(self.h))
