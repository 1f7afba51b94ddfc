"""Thin object wrapper over the C-ABI handle (include/sqgpu.h). All numerics happen in libsqgpu.so on the GPU."""
import ctypes as C

import numpy as np

from . import abi


def _c128(a, copy=False):
    return np.array(a, dtype=np.complex128, order="C", copy=copy) if copy else np.ascontiguousarray(a, dtype=np.complex128)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    """One context on one CUDA device (sqgpu_create / sqgpu_destroy)."""

    # options every new Engine starts with (sqgpu_set_option names); the parity tests set e.g. {"no_fuse": 1} here so that
    # engines created deep inside the wrapper classes pick the executor path under test. Empty = product configuration.
    default_options = {}

    def __init__(self, device=0, options=None, devices=None, shard_mode=abi.SHARD_AUTO):
        """``devices`` (a list of CUDA device indices, or an int n for devices 0..n-1) makes ONE handle over several GPUs
        (sqgpu_create_multi): the batch or the columns of the matrix are sharded inside the library (``shard_mode``)."""
        self.lib = abi.load_library()
        self._h = abi._handle()
        if devices is not None:
            devs = list(range(devices)) if isinstance(devices, int) else [int(d) for d in devices]
            arr = (C.c_int * len(devs))(*devs)
            abi.check(self.lib, self.lib.sqgpu_create_multi(len(devs), arr, int(shard_mode), C.byref(self._h)))
            self.devices = devs
            device = devs[0] if devs else 0
        else:
            abi.check(self.lib, self.lib.sqgpu_create(int(device), C.byref(self._h)))
            self.devices = [int(device)]
        self.device = int(device)
        for k, v in dict(Engine.default_options, **(options or {})).items():
            self.set_option(k, v)
        self.n_params = 0
        self.n_gates = 0
        self.qbit_num = 0
        self.rows = 0
        self.cols = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.sqgpu_destroy(self._h)
            self._h = abi._handle()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        """sqgpu_set_option: planner options take effect with the next set_circuit"""
        abi.check(self.lib, self.lib.sqgpu_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int64(0)
        abi.check(self.lib, self.lib.sqgpu_get_option(self._h, name.encode(), C.byref(v)))
        return v.value

    # ---- inputs ---------------------------------------------------------------------------------------------
    def upload_matrix(self, umtx):
        m = _c128(umtx)
        if m.ndim == 1:
            m = m.reshape(-1, 1)
        rows, cols = m.shape
        abi.check(self.lib, self.lib.sqgpu_upload_matrix(self._h, abi.as_dp(m.view(np.float64)), rows, cols, cols))
        self.rows, self.cols = rows, cols

    def set_circuit(self, circuit):
        descs, pool = circuit.descriptors(nested=False)
        self.set_circuit_raw(descs, pool, circuit.get_Parameter_Num(), circuit.qbit_num)

    def set_circuit_raw(self, descs, pool, n_params, qbit_num):
        descs = np.ascontiguousarray(descs, dtype=abi.GATE_DESC_DTYPE)
        pool = _c128(pool)
        pd = abi.as_dp(pool.view(np.float64)) if pool.size else None
        abi.check(
            self.lib,
            self.lib.sqgpu_set_circuit(self._h, descs.ctypes.data_as(C.POINTER(abi.GateDesc)), len(descs),
                                       int(n_params), int(qbit_num), pd, pool.size),
        )
        self.n_params, self.n_gates, self.qbit_num = int(n_params), len(descs), int(qbit_num)

    def multi_info(self):
        """(number of devices behind the handle, sharding mode in force)"""
        n, m = C.c_int(0), C.c_int(0)
        abi.check(self.lib, self.lib.sqgpu_multi_info(self._h, C.byref(n), C.byref(m)))
        return n.value, m.value

    def set_shard(self, col_begin, cols_total):
        """sqgpu_set_shard: the resident matrix is U[:, col_begin : col_begin + cols) of a cols_total-column matrix"""
        abi.check(self.lib, self.lib.sqgpu_set_shard(self._h, int(col_begin), int(cols_total)))

    def set_cost(self, variant=abi.FROBENIUS_NORM, trace_offset=0, prev_cost=1.0, c1=1 / 1.7, c2=1 / 2.0):
        # defaults: Optimization_Interface.cpp:74-76
        abi.check(self.lib, self.lib.sqgpu_set_cost(self._h, int(variant), int(trace_offset), float(prev_cost),
                                                    float(c1), float(c2)))

    def _params(self, params):
        p = _f64(params)
        if p.ndim == 1:
            p = p.reshape(1, -1)
        if p.ndim != 2 or p.shape[1] != self.n_params:
            raise Exception("Number of free parameters should be %d, but got %d" % (self.n_params, p.shape[-1]))
        return p

    # ---- hot path -------------------------------------------------------------------------------------------
    def cost_batched(self, params):
        p = self._params(params)
        out = np.empty(p.shape[0], dtype=np.float64)
        abi.check(self.lib, self.lib.sqgpu_cost_batched(self._h, abi.as_dp(p), p.shape[0], abi.as_dp(out)))
        return out

    def cost_grad_batched(self, params):
        p = self._params(params)
        cost = np.empty(p.shape[0], dtype=np.float64)
        grad = np.empty_like(p)
        abi.check(self.lib, self.lib.sqgpu_cost_grad_batched(self._h, abi.as_dp(p), p.shape[0], abi.as_dp(cost),
                                                             abi.as_dp(grad)))
        return cost, grad

    def cost_shifted_batched(self, params, shift):
        """(cost[b], shifted[b, p] = cost(params_b + shift e_p) for EVERY parameter p) from one adjoint sweep per set
        (sqgpu_cost_shifted_batched): the shift batches of the COSINE engine without one forward pass per shifted parameter.
        ``shift``: a number, or a sequence of shifts -> shifted[s, b, p] (one sweep serves them all)"""
        p = self._params(params)
        scalar = np.ndim(shift) == 0
        sh = _f64(np.atleast_1d(shift)).reshape(-1)
        cost = np.empty(p.shape[0], dtype=np.float64)
        shifted = np.empty((sh.size,) + p.shape, dtype=np.float64)
        abi.check(self.lib, self.lib.sqgpu_cost_shifted_batched(self._h, abi.as_dp(p), p.shape[0], abi.as_dp(sh), sh.size, abi.as_dp(cost),
                                                                abi.as_dp(shifted)))
        return cost, (shifted[0] if scalar else shifted)

    def traces_batched(self, params, with_grad):
        p = self._params(params)
        k = 1 + (self.n_params if with_grad else 0)
        out = np.empty((p.shape[0], k, 3, 2), dtype=np.float64)
        abi.check(self.lib, self.lib.sqgpu_traces_batched(self._h, abi.as_dp(p), p.shape[0], int(bool(with_grad)),
                                                          abi.as_dp(out)))
        return out

    def cost_from_traces(self, traces, with_grad, cols_total):
        t = _f64(traces)
        batch = t.shape[0]
        cost = np.empty(batch, dtype=np.float64)
        grad = np.empty((batch, self.n_params), dtype=np.float64) if with_grad else None
        abi.check(self.lib, self.lib.sqgpu_cost_from_traces(self._h, abi.as_dp(t), batch, int(bool(with_grad)),
                                                            int(cols_total), abi.as_dp(cost),
                                                            abi.as_dp(grad) if with_grad else None))
        return (cost, grad) if with_grad else cost

    def apply(self, params, inout):
        """in place on a C-contiguous complex128 ndarray (1-D state vector or 2-D matrix)"""
        if not (isinstance(inout, np.ndarray) and inout.dtype == np.complex128 and inout.flags["C_CONTIGUOUS"]
                and inout.flags["WRITEABLE"]):
            raise Exception("apply: input should be a writeable C-contiguous complex128 numpy array")
        p = _f64(params).reshape(-1)
        if p.size != self.n_params:
            raise Exception("Number of free parameters should be %d, but got %d" % (self.n_params, p.size))
        rows = inout.shape[0]
        cols = 1 if inout.ndim == 1 else inout.shape[1]
        abi.check(self.lib, self.lib.sqgpu_apply(self._h, abi.as_dp(p), abi.as_dp(inout.view(np.float64)), rows, cols,
                                                 cols))

    def apply_derivative(self, params, inp):
        m = _c128(inp)
        p = _f64(params).reshape(-1)
        if p.size != self.n_params:
            raise Exception("Number of free parameters should be %d, but got %d" % (self.n_params, p.size))
        rows = m.shape[0]
        cols = 1 if m.ndim == 1 else m.shape[1]
        out = np.empty((self.n_params,) + m.shape, dtype=np.complex128)
        abi.check(self.lib, self.lib.sqgpu_apply_derivative(self._h, abi.as_dp(p), abi.as_dp(m.view(np.float64)), rows,
                                                            cols, cols, abi.as_dp(out.view(np.float64))))
        return [out[i] for i in range(self.n_params)]

    def apply_gate(self, desc_row, gate_params, inout, pool=None, deriv_param=-1):
        d = np.ascontiguousarray(desc_row, dtype=abi.GATE_DESC_DTYPE).reshape(1)
        gp = _f64(gate_params).reshape(-1)
        rows = inout.shape[0]
        cols = 1 if inout.ndim == 1 else inout.shape[1]
        pl = _c128(pool) if pool is not None and len(pool) else None
        abi.check(self.lib, self.lib.sqgpu_apply_gate(
            self._h, d.ctypes.data_as(C.POINTER(abi.GateDesc)), abi.as_dp(gp) if gp.size else None,
            abi.as_dp(pl.view(np.float64)) if pl is not None else None, int(deriv_param),
            abi.as_dp(inout.view(np.float64)), rows, cols, cols))

    # ---- VQE ------------------------------------------------------------------------------------------------
    def set_hamiltonian_csr(self, indptr, indices, data):
        ip = np.ascontiguousarray(indptr, dtype=np.int32)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        v = _c128(data)
        abi.check(self.lib, self.lib.sqgpu_set_hamiltonian_csr(self._h, len(ip) - 1, v.size, abi.as_ip(ip),
                                                               abi.as_ip(ix), abi.as_dp(v.view(np.float64))))

    def vqe_energy_batched(self, params):
        p = self._params(params)
        out = np.empty(p.shape[0], dtype=np.float64)
        abi.check(self.lib, self.lib.sqgpu_vqe_energy_batched(self._h, abi.as_dp(p), p.shape[0], abi.as_dp(out)))
        return out

    def vqe_energy_grad_batched(self, params):
        p = self._params(params)
        e = np.empty(p.shape[0], dtype=np.float64)
        g = np.empty_like(p)
        abi.check(self.lib, self.lib.sqgpu_vqe_energy_grad_batched(self._h, abi.as_dp(p), p.shape[0], abi.as_dp(e),
                                                                   abi.as_dp(g)))
        return e, g

    # ---- device-resident entry points (pointers are ints: torch tensor.data_ptr()) ---------------------------
    def cost_batched_dev(self, d_params, batch, d_cost, stream=0):
        abi.check(self.lib, self.lib.sqgpu_cost_batched_dev(self._h, d_params, int(batch), d_cost, stream))

    def cost_grad_batched_dev(self, d_params, batch, d_cost, d_grad, stream=0):
        abi.check(self.lib, self.lib.sqgpu_cost_grad_batched_dev(self._h, d_params, int(batch), d_cost, d_grad, stream))

    def cost_shifted_batched_dev(self, d_params, batch, shifts, d_cost, d_shifted, stream=0):
        sh = _f64(np.atleast_1d(shifts)).reshape(-1)
        abi.check(self.lib, self.lib.sqgpu_cost_shifted_batched_dev(self._h, d_params, int(batch), abi.as_dp(sh), sh.size, d_cost, d_shifted, stream))

    def traces_batched_dev(self, d_params, batch, with_grad, d_traces, stream=0):
        abi.check(self.lib, self.lib.sqgpu_traces_batched_dev(self._h, d_params, int(batch), int(bool(with_grad)),
                                                              d_traces, stream))

    def grad_traces_with_global_dev(self, d_params, batch, d_global_traces0, d_traces, stream=0):
        abi.check(self.lib, self.lib.sqgpu_grad_traces_with_global_dev(self._h, d_params, int(batch), d_global_traces0, d_traces,
                                                                       stream))

    def cost_from_traces_dev(self, d_traces, batch, with_grad, cols_total, d_cost, d_grad, stream=0):
        abi.check(self.lib, self.lib.sqgpu_cost_from_traces_dev(self._h, d_traces, int(batch), int(bool(with_grad)),
                                                                int(cols_total), d_cost, d_grad, stream))

    def vqe_energy_batched_dev(self, d_params, batch, d_energy, stream=0):
        abi.check(self.lib, self.lib.sqgpu_vqe_energy_batched_dev(self._h, d_params, int(batch), d_energy, stream))

    def vqe_energy_grad_batched_dev(self, d_params, batch, d_energy, d_grad, stream=0):
        abi.check(self.lib, self.lib.sqgpu_vqe_energy_grad_batched_dev(self._h, d_params, int(batch), d_energy, d_grad,
                                                                       stream))

    # ---- device-resident optimizer loops (N1) -----------------------------------------------------------------
    def adam_init(self, theta0, eta=1e-3, beta1=0.68, beta2=0.8, epsilon=1e-4):
        """``theta0`` [batch, P] (or [P]): independent ADAM trajectories kept on the device (sqgpu_adam_init)"""
        t = self._params(theta0)
        self._adam_batch = t.shape[0]
        abi.check(self.lib, self.lib.sqgpu_adam_init(self._h, abi.as_dp(t), t.shape[0], float(eta), float(beta1), float(beta2),
                                                     float(epsilon)))

    def adam_steps(self, n_steps):
        """run n_steps on the device; returns the cost history [n_steps, batch] (cost before each step's update)"""
        hist = np.empty((int(n_steps), self._adam_batch), dtype=np.float64)
        abi.check(self.lib, self.lib.sqgpu_adam_steps(self._h, int(n_steps), abi.as_dp(hist) if hist.size else None))
        return hist

    def adam_get(self):
        """(theta, best_cost, best_theta, status)"""
        B, P = self._adam_batch, self.n_params
        theta, best = np.empty((B, P)), np.empty((B, P))
        bc = np.empty(B)
        st = np.zeros(B, dtype=np.int32)
        abi.check(self.lib, self.lib.sqgpu_adam_get(self._h, abi.as_dp(theta), abi.as_dp(bc), abi.as_dp(best),
                                                    st.ctypes.data_as(C.POINTER(C.c_int))))
        return theta, bc, best, st

    def line_search_batched(self, x, direction, alphas, with_derivative=True):
        """cost (and directional derivative) at x + alpha_j * direction for all alphas as one batch (sqgpu_line_search_batched)"""
        x, d, a = _f64(x).reshape(-1), _f64(direction).reshape(-1), _f64(alphas).reshape(-1)
        if x.size != self.n_params or d.size != self.n_params:
            raise Exception("Number of free parameters should be %d" % self.n_params)
        cost = np.empty(a.size)
        dphi = np.empty(a.size) if with_derivative else None
        abi.check(self.lib, self.lib.sqgpu_line_search_batched(self._h, abi.as_dp(x), abi.as_dp(d), abi.as_dp(a), a.size,
                                                               abi.as_dp(cost), abi.as_dp(dphi) if with_derivative else None))
        return (cost, dphi) if with_derivative else cost

    # ---- introspection --------------------------------------------------------------------------------------
    def launch_count(self):
        n = C.c_int64(0)
        abi.check(self.lib, self.lib.sqgpu_launch_count(self._h, C.byref(n)))
        return n.value

    def last_kernel_time(self):
        buf = C.create_string_buffer(128)
        ms = C.c_double(0)
        n = C.c_int(0)
        abi.check(self.lib, self.lib.sqgpu_last_kernel_time(self._h, buf, 128, C.byref(ms), C.byref(n)))
        return buf.value.decode(), ms.value, n.value

    def last_launch_shape(self):
        """dict of the fused executor's last launch geometry (sqgpu_last_launch_shape)"""
        a = (C.c_int * 6)()
        abi.check(self.lib, self.lib.sqgpu_last_launch_shape(self._h, a, 6))
        return dict(zip(("log_ct", "threads", "chunks", "tiles_per_cta", "smem", "cluster"), (int(v) for v in a)))

    def last_exec_flops(self):
        """(tensor, scalar) FP64 flops issued by the fused executor's last cost / gradient launch"""
        t, sc = C.c_double(0), C.c_double(0)
        abi.check(self.lib, self.lib.sqgpu_last_exec_flops(self._h, C.byref(t), C.byref(sc)))
        return t.value, sc.value

    def kernel_time(self, name):
        """(ms, launches) of one kernel by name since its ring was last reset (sqgpu_kernel_time)"""
        ms = C.c_double(0)
        n = C.c_int(0)
        abi.check(self.lib, self.lib.sqgpu_kernel_time(self._h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def fp64_fma_peak(self):
        t = C.c_double(0)
        abi.check(self.lib, self.lib.sqgpu_fp64_fma_peak(self._h, C.byref(t)))
        return t.value
