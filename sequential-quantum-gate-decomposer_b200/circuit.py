"""Host-side mirror of the reference's ``Circuit`` (= C++ ``Gates_block``) for the hot path.

Same method names and argument meaning as the CPython wrapper squander/gates/qgd_Circuit_Wrapper.cpp:3109-3302
(``add_U3(target_qbit)``, ``add_CNOT(target_qbit, control_qbit)``, ``add_Circuit``, ``get_Parameter_Num``,
``apply_to(parameters, unitary)`` in place, ``apply_derivate_to``, ``get_Matrix`` ...). The class only keeps the
gate structure; every numerical call goes to the CUDA engine through the C-ABI (engine.Engine) -- there is no
CPU evaluation here.

Parameter layout follows Gates_block::add_gate (Gates_block.cpp:2500-2525): gates own consecutive slices of one
flat parameter vector in insertion order, nested circuits included.
"""
import numpy as np

from . import abi


class _Gate:
    __slots__ = ("type", "target", "control", "target2", "control2", "qubits", "matrix")

    def __init__(self, type_, target=-1, control=-1, target2=-1, control2=-1, qubits=None, matrix=None):
        self.type = type_
        self.target = target
        self.control = control
        self.target2 = target2
        self.control2 = control2
        self.qubits = qubits
        self.matrix = matrix

    @property
    def n_params(self):
        return abi.PARAM_COUNT[self.type]


class Circuit:
    """Ordered gate structure over ``qbit_num`` qubits; gate 0 is applied first (Gates_block.cpp:683-708)."""

    def __init__(self, qbit_num, device=0):
        qbit_num = int(qbit_num)
        if qbit_num < 1 or qbit_num > 30:
            raise Exception("Circuit: number of qubits should be between 1 and 30")
        self.qbit_num = qbit_num
        self._items = []  # _Gate or Circuit
        self._min_fusion = 14  # Gates_block.cpp:91,113 (kept for API parity; the engine plans its own windows)
        self._device = int(device)  # CUDA device of the numerical calls (apply_to, get_Matrix, ...)
        self._engine = None
        self._engine_key = None
        self._version = 0  # bumped by every structural change of THIS block

    # ---- structure ------------------------------------------------------------------------------------------
    def _check_q(self, *qs):
        for q in qs:
            if q < 0 or q >= self.qbit_num:
                raise Exception("Circuit: qubit index %d out of range for %d qubits" % (q, self.qbit_num))
        if len(set(qs)) != len(qs):
            raise Exception("Circuit: target and control qubits must differ")

    def _add(self, gate):
        self._items.append(gate)
        self._version += 1

    def structure_key(self):
        """Changes whenever the gate structure does, nested blocks included (they stay shared with whoever added them, as
        in Gates_block::add_gate, which stores the pointer): the versions and identities of all blocks, depth first. The
        device plan of a circuit is rebuilt when this key differs from the one it was built for."""
        key = [id(self), self._version, len(self._items)]
        for it in self._items:
            if isinstance(it, Circuit):
                key.append(it.structure_key())
        return tuple(key)

    def _add_1q(self, type_, target_qbit):
        self._check_q(int(target_qbit))
        self._add(_Gate(type_, target=int(target_qbit)))

    def _add_c1q(self, type_, target_qbit, control_qbit):
        self._check_q(int(target_qbit), int(control_qbit))
        self._add(_Gate(type_, target=int(target_qbit), control=int(control_qbit)))

    def _add_2t(self, type_, target_qbits):
        t = [int(q) for q in target_qbits]
        if len(t) != 2:
            raise Exception("gate requires exactly 2 target qubits")
        self._check_q(*t)
        self._add(_Gate(type_, target=t[0], target2=t[1]))

    def add_U1(self, target_qbit): self._add_1q(abi.U1, target_qbit)
    def add_U2(self, target_qbit): self._add_1q(abi.U2, target_qbit)
    def add_U3(self, target_qbit): self._add_1q(abi.U3, target_qbit)
    def add_RX(self, target_qbit): self._add_1q(abi.RX, target_qbit)
    def add_RY(self, target_qbit): self._add_1q(abi.RY, target_qbit)
    def add_RZ(self, target_qbit): self._add_1q(abi.RZ, target_qbit)
    def add_R(self, target_qbit): self._add_1q(abi.R, target_qbit)
    def add_H(self, target_qbit): self._add_1q(abi.H, target_qbit)
    def add_X(self, target_qbit): self._add_1q(abi.X, target_qbit)
    def add_Y(self, target_qbit): self._add_1q(abi.Y, target_qbit)
    def add_Z(self, target_qbit): self._add_1q(abi.Z, target_qbit)
    def add_S(self, target_qbit): self._add_1q(abi.S, target_qbit)
    def add_Sdg(self, target_qbit): self._add_1q(abi.SDG, target_qbit)
    def add_T(self, target_qbit): self._add_1q(abi.T, target_qbit)
    def add_Tdg(self, target_qbit): self._add_1q(abi.TDG, target_qbit)
    def add_SX(self, target_qbit): self._add_1q(abi.SX, target_qbit)
    def add_SXdg(self, target_qbit): self._add_1q(abi.SXDG, target_qbit)
    def add_CNOT(self, target_qbit, control_qbit): self._add_c1q(abi.CNOT, target_qbit, control_qbit)
    def add_CZ(self, target_qbit, control_qbit): self._add_c1q(abi.CZ, target_qbit, control_qbit)
    def add_CH(self, target_qbit, control_qbit): self._add_c1q(abi.CH, target_qbit, control_qbit)
    def add_CU(self, target_qbit, control_qbit): self._add_c1q(abi.CU, target_qbit, control_qbit)
    def add_CRY(self, target_qbit, control_qbit): self._add_c1q(abi.CRY, target_qbit, control_qbit)
    def add_CRX(self, target_qbit, control_qbit): self._add_c1q(abi.CRX, target_qbit, control_qbit)
    def add_CRZ(self, target_qbit, control_qbit): self._add_c1q(abi.CRZ, target_qbit, control_qbit)
    def add_CP(self, target_qbit, control_qbit): self._add_c1q(abi.CP, target_qbit, control_qbit)
    def add_CR(self, target_qbit, control_qbit): self._add_c1q(abi.CR, target_qbit, control_qbit)
    def add_adaptive(self, target_qbit, control_qbit): self._add_c1q(abi.ADAPTIVE, target_qbit, control_qbit)
    def add_CROT(self, target_qbit, control_qbit): self._add_c1q(abi.CROT, target_qbit, control_qbit)
    def add_SYC(self, target_qbit, control_qbit): self._add_c1q(abi.SYC, target_qbit, control_qbit)
    def add_RXX(self, target_qbits): self._add_2t(abi.RXX, target_qbits)
    def add_RYY(self, target_qbits): self._add_2t(abi.RYY, target_qbits)
    def add_RZZ(self, target_qbits): self._add_2t(abi.RZZ, target_qbits)
    def add_SWAP(self, target_qbits): self._add_2t(abi.SWAP, target_qbits)

    def add_CCX(self, target_qbit, control_qbits):
        c = [int(q) for q in control_qbits]
        if len(c) != 2:
            raise Exception("CCX requires exactly 2 control qubits")
        self._check_q(int(target_qbit), *c)
        self._add(_Gate(abi.CCX, target=int(target_qbit), control=c[0], control2=c[1]))

    def add_CSWAP(self, target_qbits, control_qbits):
        t = [int(q) for q in target_qbits]
        c = [int(q) for q in control_qbits]
        if len(t) != 2 or len(c) != 1:
            raise Exception("CSWAP requires 2 target qubits and 1 control qubit")
        self._check_q(*t, *c)
        self._add(_Gate(abi.CSWAP, target=t[0], target2=t[1], control=c[0]))

    def add_GENERAL(self, operation_mtx, target_qbits, control_qbits=None):
        """Constant dense block on ``target_qbits`` (a GENERAL_OPERATION with a local 2^k x 2^k matrix,
        Gate.cpp:1586-1660). Local index bit j belongs to the j-th *ascending* target qubit
        (apply_large_kernel_to_input.cpp:160-169)."""
        if control_qbits:
            raise Exception("add_GENERAL: controlled general gates are not supported on the device path")
        q = sorted(int(x) for x in target_qbits)
        self._check_q(*q)
        k = len(q)
        if k < 1 or k > 5:
            raise Exception("add_GENERAL: 1..5 target qubits supported")
        m = np.ascontiguousarray(operation_mtx, dtype=np.complex128)
        if m.shape != (1 << k, 1 << k):
            raise Exception("add_GENERAL: operation matrix has invalid size")
        self._add(_Gate(abi.GENERAL, qubits=q, matrix=m))

    def add_Circuit(self, circuit):
        if circuit.qbit_num != self.qbit_num:
            raise Exception("add_Circuit: qubit count mismatch")
        self._add(circuit)

    # ---- queries --------------------------------------------------------------------------------------------
    def get_Qbit_Num(self):
        return self.qbit_num

    def get_Parameter_Num(self):
        return sum(it.get_Parameter_Num() if isinstance(it, Circuit) else it.n_params for it in self._items)

    def get_Gate_Num(self):
        return len(self._items)

    def set_min_fusion(self, min_fusion):
        self._min_fusion = int(min_fusion)

    def get_Gates(self):
        """the gates and sub-circuits of this block in application order (Gates_block::get_gates)"""
        return list(self._items)

    def get_Gate(self, idx):
        return self._items[idx]

    def get_Gate_Nums(self):
        """{gate name: count} over the whole (nested) structure (Gates_block::get_gate_nums, Gates_block.cpp:2630-2637)"""
        out = {}
        for g in self._flat_gates():
            name = abi.GATE_NAMES[g.type]
            out[name] = out.get(name, 0) + 1
        return out

    @staticmethod
    def _gate_qubits(g):
        return [q for q in ([g.target, g.control, g.target2, g.control2] + list(g.qubits or [])) if q is not None and q >= 0]

    def _item_qubits(self, it):
        return set(it.get_Qbits()) if isinstance(it, Circuit) else {int(q) for q in self._gate_qubits(it)}

    def set_Qbit_Num(self, qbit_num):
        """Gates_block::set_qbit_num: the register grows or shrinks; every gate must still fit"""
        qbit_num = int(qbit_num)
        used = self.get_Qbits()
        if qbit_num < 1 or qbit_num > 30 or (used and used[-1] >= qbit_num):
            raise Exception("set_Qbit_Num: a gate acts on a qubit outside the new register")
        self.qbit_num = qbit_num
        for it in self._items:
            if isinstance(it, Circuit):
                it.set_Qbit_Num(qbit_num)
        self._version += 1
        self._engine = None
        self._engine_key = None

    def __getstate__(self):
        """pickling (qgd_Circuit_Wrapper __getstate__ / __setstate__): the gate structure travels, the device handle does not"""
        st = dict(self.__dict__)
        st["_engine"] = None
        st["_engine_key"] = None
        st.pop("_inverse_cache", None)
        return st

    def __setstate__(self, state):
        self.__dict__.update(state)

    def get_Parents(self, gate):
        """indices (in this block) of the gates that must run before gate ``gate``: for each of its qubits the closest earlier
        gate on that qubit (Gates_block::determine_parents, Gates_block.cpp:3668-3720; qgd_Circuit.get_Parents)"""
        idx = gate if isinstance(gate, (int, np.integer)) else self._items.index(gate)
        need = self._item_qubits(self._items[idx])
        out = []
        for j in range(idx - 1, -1, -1):
            hit = need & self._item_qubits(self._items[j])
            if hit:
                out.append(j)
                need -= hit
            if not need:
                break
        return sorted(out)

    def get_Children(self, gate):
        """indices of the gates that wait for gate ``gate``: for each of its qubits the closest later gate on that qubit
        (Gates_block::determine_children)"""
        idx = gate if isinstance(gate, (int, np.integer)) else self._items.index(gate)
        need = self._item_qubits(self._items[idx])
        out = []
        for j in range(idx + 1, len(self._items)):
            hit = need & self._item_qubits(self._items[j])
            if hit:
                out.append(j)
                need -= hit
            if not need:
                break
        return sorted(out)

    def get_Parameter_Start_Index(self, gate=0):
        """index of the first parameter of item ``gate`` of this block in the block's parameter vector"""
        idx = gate if isinstance(gate, (int, np.integer)) else self._items.index(gate)
        return sum(it.get_Parameter_Num() if isinstance(it, Circuit) else it.n_params for it in self._items[:idx])

    def Extract_Parameters(self, parameters, gate=None):
        """the slice of ``parameters`` that belongs to item ``gate`` of this block (Gate::extract_parameters); without ``gate``:
        the block's own parameters, checked for their number"""
        q = np.asarray(parameters, dtype=np.float64).reshape(-1)
        if q.size != self.get_Parameter_Num():
            raise Exception("Number of free parameters should be %d, but got %d" % (self.get_Parameter_Num(), q.size))
        if gate is None:
            return q.copy()
        idx = gate if isinstance(gate, (int, np.integer)) else self._items.index(gate)
        it = self._items[idx]
        st = self.get_Parameter_Start_Index(idx)
        return q[st:st + (it.get_Parameter_Num() if isinstance(it, Circuit) else it.n_params)].copy()

    def get_Qbits(self):
        """sorted list of the qubits the circuit acts on (Gates_block::get_involved_qubits)"""
        return sorted({int(q) for g in self._flat_gates() for q in self._gate_qubits(g)})

    def Remap_Qbits(self, qbit_map, qbit_num=None):
        """a new circuit with qubit q replaced by qbit_map[q] (Gates_block::create_remapped_circuit, qgd_Circuit.py:709-731);
        the register may change its size; qubits without an entry keep their index"""
        n = self.qbit_num if qbit_num is None else int(qbit_num)
        m = lambda q: q if q is None or q < 0 else int(qbit_map.get(q, q))
        out = Circuit(n, self._device)
        for it in self._items:
            if isinstance(it, Circuit):
                out._add(it.Remap_Qbits(qbit_map, n))
                continue
            g = _Gate(it.type, m(it.target), m(it.control), m(it.target2), m(it.control2),
                      None if it.qubits is None else [m(q) for q in it.qubits], it.matrix)
            if it.qubits is not None and sorted(g.qubits) != list(g.qubits):
                raise Exception("Remap_Qbits: the qubits of a GENERAL gate must stay in ascending order")
            out._check_q(*self._gate_qubits(g))
            out._add(g)
        return out

    # ---- the inverse structure and multiplication from the right ---------------------------------------------------------
    # parameter maps of the inverse gates, each checked against the oracle's kernels (tests/test_host_logic.py): for the gate
    # at parameters p the inverse is the SAME gate type at sign * p[perm] + offset
    _INV_RULES = {
        abi.U3: ([0, 2, 1], [-1, -1, -1], [0, 0, 0]), abi.CU: ([0, 2, 1, 3], [-1, -1, -1, -1], [0, 0, 0, 0]),
        abi.U2: ([1, 0], [-1, -1], [-np.pi, np.pi]), abi.R: ([0, 1], [-1, 1], [0, 0]), abi.CR: ([0, 1], [-1, 1], [0, 0]),
        abi.CROT: ([0, 1], [-1, 1], [0, 0]),
    }
    _INV_TYPE = {abi.S: abi.SDG, abi.SDG: abi.S, abi.T: abi.TDG, abi.TDG: abi.T, abi.SX: abi.SXDG, abi.SXDG: abi.SX}
    # Sycamore gate fSim(pi/2, pi/6) (kernels/apply_dedicated_gate_kernel_to_input.cpp:582-640), symmetric in its two qubits
    _SYC = np.array([[1, 0, 0, 0], [0, 0, -1j, 0], [0, -1j, 0, 0], [0, 0, 0, np.exp(-1j * np.pi / 6)]], dtype=np.complex128)

    def get_Inverse(self):
        """(inverse circuit, map) with inverse(map(p)) @ self(p) = identity for every parameter vector p: the gates in reverse
        order, each replaced by its inverse -- the same gate type at mapped parameters (negated angles, U3's phi and lambda
        exchanged, ...), S / T / SX exchanged with their daggers, GENERAL and SYC as GENERAL gates with the adjoint matrix.
        ``map(p)`` returns the parameter vector of the inverse circuit. Cached per structure."""
        key = self.structure_key()
        if getattr(self, "_inverse_cache", None) is not None and self._inverse_cache[0] == key:
            return self._inverse_cache[1], self._inverse_cache[2]
        gates, starts, pos = list(self._flat_gates()), [], 0
        for g in gates:
            starts.append(pos)
            pos += g.n_params
        inv = Circuit(self.qbit_num, self._device)
        src, sign, off = [], [], []
        for g, st in zip(reversed(gates), reversed(starts)):
            if g.type == abi.GENERAL:
                inv._add(_Gate(abi.GENERAL, qubits=list(g.qubits), matrix=np.ascontiguousarray(g.matrix.conj().T)))
                continue
            if g.type == abi.SYC:
                inv._add(_Gate(abi.GENERAL, qubits=sorted([g.target, g.control]), matrix=np.ascontiguousarray(self._SYC.conj().T)))
                continue
            inv._add(_Gate(self._INV_TYPE.get(g.type, g.type), g.target, g.control, g.target2, g.control2, g.qubits, g.matrix))
            n = g.n_params
            perm, sg, of = self._INV_RULES.get(g.type, (list(range(n)), [-1] * n, [0] * n))
            src += [st + q for q in perm]
            sign += sg
            off += of
        src, sign, off = np.array(src, dtype=np.int64), np.array(sign, dtype=np.float64), np.array(off, dtype=np.float64)

        def pmap(parameters):
            q = np.asarray(parameters, dtype=np.float64).reshape(-1)
            if q.size != pos:
                raise Exception("Number of free parameters should be %d, but got %d" % (pos, q.size))
            return sign * q[src] + off if pos else np.zeros(0)

        self._inverse_cache = (key, inv, pmap)
        return inv, pmap

    def apply_from_right(self, parameters, unitary):
        """In place ``unitary <- unitary @ C(parameters)`` for a rows x 2^n complex128 array (Gates_block::apply_from_right,
        Gates_block.cpp:717-760). On the device: U C = (C^dagger U^dagger)^dagger, and C^dagger = C^-1 is the inverse structure
        of get_Inverse applied from the left to U^dagger (the gates must be unitary, as for the gradient)."""
        u = np.asarray(unitary)
        if u.dtype != np.complex128 or u.ndim != 2 or u.shape[1] != (1 << self.qbit_num):
            raise Exception("apply_from_right: expected a rows x 2^qbit_num complex128 array")
        for g in self._flat_gates():
            if g.type == abi.GENERAL and np.abs(g.matrix.conj().T @ g.matrix - np.eye(g.matrix.shape[0])).max() > 1e-10:
                raise Exception("apply_from_right: a GENERAL gate is not unitary")
        inv, pmap = self.get_Inverse()
        a = np.ascontiguousarray(u.conj().T)
        inv.apply_to(pmap(parameters), a)
        u[...] = a.conj().T

    def apply_to_list(self, inputs, parameters, parallel=1, is_f32=False):
        """apply_to on every array of ``inputs`` in place (Gates_block::apply_to_list, Gates_block.cpp:575-600): the gate
        structure and the kernel tables stay on the device between the inputs"""
        for m in inputs:
            self.apply_to(parameters, m, parallel, is_f32)

    def _flat_gates(self):
        for it in self._items:
            if isinstance(it, Circuit):
                yield from it._flat_gates()
            else:
                yield it

    def get_Flat_Circuit(self):
        """Un-nested copy (Gates_block::get_flat_circuit, Gates_block.cpp:3827-3856)."""
        c = Circuit(self.qbit_num, self._device)
        c._items = list(self._flat_gates())
        c._version = 1
        return c

    def descriptors(self, nested=False):
        """(descs, pool): the sqgpu_gate_desc array in application order and the constant-kernel pool.

        nested=True keeps the block structure as BLOCK_BEGIN/BLOCK_END markers (only the oracle harness reads
        those); the engine always gets nested=False."""
        rows = []
        pool = []
        state = {"p": 0, "off": 0}

        def emit(c):
            for it in c._items:
                if isinstance(it, Circuit):
                    if nested:
                        rows.append((abi.BLOCK_BEGIN, -1, -1, -1, -1, state["p"], 0, 0, (0,) * 8, 0))
                    emit(it)
                    if nested:
                        rows.append((abi.BLOCK_END, -1, -1, -1, -1, state["p"], 0, 0, (0,) * 8, 0))
                    continue
                q = tuple(it.qubits) + (0,) * (8 - len(it.qubits)) if it.qubits else (0,) * 8
                nq = len(it.qubits) if it.qubits else 0
                off = 0
                if it.matrix is not None:
                    off = state["off"]
                    pool.append(it.matrix.reshape(-1))
                    state["off"] += it.matrix.size
                rows.append((it.type, it.target, it.control, it.target2, it.control2, state["p"], it.n_params, nq, q,
                             off))
                state["p"] += it.n_params

        emit(self)
        descs = np.array(rows, dtype=abi.GATE_DESC_DTYPE) if rows else np.zeros(0, dtype=abi.GATE_DESC_DTYPE)
        pool_arr = np.concatenate(pool) if pool else np.zeros(0, dtype=np.complex128)
        return descs, np.ascontiguousarray(pool_arr, dtype=np.complex128)

    # ---- numerics (all on the GPU through the C-ABI) --------------------------------------------------------
    def _get_engine(self):
        from .engine import Engine

        key = self.structure_key()
        if self._engine is None:
            self._engine = Engine(self._device)
        if self._engine_key != key:
            self._engine.set_circuit(self)
            self._engine_key = key
        return self._engine

    def apply_to(self, parameters, unitary, parallel=1, is_f32=False):
        """In place ``unitary <- C(parameters) @ unitary`` for a 2^n x cols complex128 array (cols = 1 or a 1-D
        array: state vector). Mirrors qgd_Circuit_Wrapper_apply_to (qgd_Circuit_Wrapper.cpp:861-995)."""
        if is_f32:
            raise Exception("apply_to: the device path is fp64 only")
        self._get_engine().apply(parameters, unitary)

    def apply_derivate_to(self, parameters, unitary, parallel=1, is_f32=False):
        """List of P arrays d(C @ unitary)/d parameters[i] (Gates_block::apply_derivate_to)."""
        if is_f32:
            raise Exception("apply_derivate_to: the device path is fp64 only")
        return self._get_engine().apply_derivative(parameters, unitary)

    def apply_to_combined(self, parameters, unitary, parallel=1, is_f32=False):
        """[C @ unitary, d_0, ..., d_{P-1}] (Gates_block::apply_to_combined, Gates_block.cpp:1320-1364)."""
        out = np.array(unitary, dtype=np.complex128, copy=True)
        eng = self._get_engine()
        derivs = eng.apply_derivative(parameters, unitary)
        eng.apply(parameters, out)
        return [out] + derivs

    def get_Second_Renyi_Entropy(self, parameters=None, input_state=None, qubit_list=None):
        """-log Tr rho_A^2 of the state C(parameters) |input_state> reduced to the qubits of ``qubit_list``
        (Gates_block::get_second_Renyi_entropy, Gates_block.cpp:3625-3650; wrapper qgd_Circuit_Wrapper.cpp). The circuit runs on
        the device; the reduction of the 2^n amplitudes is host-side numpy (second_renyi_entropy below)."""
        if parameters is None:
            raise Exception("get_Second_Renyi_entropy: array of input parameters is None")
        n = self.qbit_num
        if qubit_list is None:
            qubit_list = list(range(n))
        if any(not isinstance(q, (int, np.integer)) or q < 0 or q >= n for q in qubit_list):
            raise Exception("Elements of qbit_list should be integers in [0, qbit_num)")
        if input_state is None:
            state = np.zeros(1 << n, dtype=np.complex128)
            state[0] = 1.0
        else:
            state = np.array(input_state, dtype=np.complex128).reshape(-1)
            if state.size != (1 << n):
                raise Exception("input state should have 2^qbit_num elements")
        self._get_engine().apply(parameters, state)
        return second_renyi_entropy(state, n, sorted(set(int(q) for q in qubit_list)))

    def get_Matrix(self, parameters=None, is_f32=False):
        """C(parameters) as a dense 2^n x 2^n matrix (apply_to on the identity, Gates_block::get_matrix)."""
        if parameters is None:
            parameters = np.zeros(0)
        m = np.eye(1 << self.qbit_num, dtype=np.complex128)
        self._get_engine().apply(parameters, m)
        return m


def second_renyi_entropy(state, qbit_num, qubits):
    """-log Tr rho_A^2 for the pure state ``state`` (2^n amplitudes, qubit q = bit q of the index) and the subsystem A = ``qubits``
    (Gates_block::get_reduced_density_matrix + get_second_Renyi_entropy, Gates_block.cpp:3480-3650). With M the amplitudes as an
    |A| x |rest| matrix, rho_A = M M^dagger and Tr rho_A^2 = || M M^dagger ||_F^2; for a pure state the complement has the same
    purity, so the smaller side is the one that is squared."""
    n = int(qbit_num)
    qubits = list(qubits)
    rest = [q for q in range(n) if q not in qubits]
    if len(qubits) > len(rest):
        qubits, rest = rest, qubits
    psi = np.asarray(state, dtype=np.complex128).reshape((2,) * n)  # axis a <-> qubit n - 1 - a
    m = np.transpose(psi, [n - 1 - q for q in qubits] + [n - 1 - q for q in rest]).reshape(1 << len(qubits), -1)
    rho = m @ m.conj().T
    return float(-np.log(np.sum(rho.real ** 2 + rho.imag ** 2)))
