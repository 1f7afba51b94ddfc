"""Reader / writer of the reference's binary gate-list format.

The format is the one ``export_gate_list_to_binary`` / ``import_gate_list_from_binary`` of the reference define
(squander/src-cpp/gates/Gates_block.cpp:4807-4920 and :4929-5330): little-endian, no padding,

    block  :=  int32 qbit_num, int32 parameter_num, int32 gates_num, gate * gates_num
    gate   :=  int32 gate_type (enum gate_type, Gate.h:39-79), then by type
               CNOT, CZ, CH, SYC            int32 target, int32 control
               U1 / U2 / U3                 int32 target, float64 * (1 / 2 / 3)
               RX, RY, RZ                   int32 target, float64
               CRY, ADAPTIVE                int32 target, int32 control, float64
               X, Y, Z, H, S, SDG, SX, T, TDG   int32 target
               BLOCK                        a nested block (its own header follows the type tag)

Parameters are stored gate by gate, so the flat parameter vector of the circuit (Gates_block::add_gate layout) is their
concatenation in file order. The reference writes this file as its checkpoint (ADAM.cpp:241-247,
``export_circuit_2_binary``) and reads it in ``set_Gate_Structure_From_Binary``; with this module those files feed the GPU
engine directly, without the reference or Qiskit.

Reference quirk kept out: its importer adds a T gate for a TDG tag (Gates_block.cpp:5223-5231); here TDG imports as Tdg.
"""
import struct

import numpy as np

from . import abi
from .circuit import Circuit

_CTRL = {abi.CNOT: "add_CNOT", abi.CZ: "add_CZ", abi.CH: "add_CH", abi.SYC: "add_SYC"}
_PARAM_1Q = {abi.U1: ("add_U1", 1), abi.U2: ("add_U2", 2), abi.U3: ("add_U3", 3), abi.RX: ("add_RX", 1), abi.RY: ("add_RY", 1),
             abi.RZ: ("add_RZ", 1)}
_PARAM_CTRL = {abi.CRY: "add_CRY", abi.ADAPTIVE: "add_adaptive"}
_FIXED_1Q = {abi.X: "add_X", abi.Y: "add_Y", abi.Z: "add_Z", abi.H: "add_H", abi.S: "add_S", abi.SDG: "add_Sdg", abi.SX: "add_SX",
             abi.T: "add_T", abi.TDG: "add_Tdg"}
# what export_gate_list_to_binary can write (Gates_block.cpp:4869-4913); T / TDG are import-only in the reference
_EXPORTABLE = set(_CTRL) | set(_PARAM_1Q) | set(_PARAM_CTRL) | {abi.X, abi.Y, abi.Z, abi.H, abi.S, abi.SDG, abi.SX}


class _Reader:
    def __init__(self, data):
        self.data = data
        self.pos = 0

    def i32(self):
        if self.pos + 4 > len(self.data):
            raise Exception("Corrupted input file, reached end of the file before contructing the whole gate structure")
        v = struct.unpack_from("<i", self.data, self.pos)[0]
        self.pos += 4
        return v

    def f64(self, n):
        if self.pos + 8 * n > len(self.data):
            raise Exception("Corrupted input file, reached end of the file before contructing the whole gate structure")
        v = struct.unpack_from("<%dd" % n, self.data, self.pos)
        self.pos += 8 * n
        return v


def _read_block(r, params, device):
    qbit_num, parameter_num, gates_num = r.i32(), r.i32(), r.i32()
    if qbit_num < 1 or qbit_num > 30 or parameter_num < 0 or gates_num < 0:
        raise Exception("import_gate_list_from_binary: implausible block header")
    c = Circuit(qbit_num, device)
    p0 = len(params)
    for _ in range(gates_num):
        t = r.i32()
        if t in _CTRL:
            target, control = r.i32(), r.i32()
            getattr(c, _CTRL[t])(target, control)
        elif t in _PARAM_1Q:
            name, n = _PARAM_1Q[t]
            target = r.i32()
            params.extend(r.f64(n))
            getattr(c, name)(target)
        elif t in _PARAM_CTRL:
            target, control = r.i32(), r.i32()
            params.extend(r.f64(1))
            getattr(c, _PARAM_CTRL[t])(target, control)
        elif t in _FIXED_1Q:
            getattr(c, _FIXED_1Q[t])(r.i32())
        elif t == abi.BLOCK:
            c.add_Circuit(_read_block(r, params, device))
        else:
            raise Exception("import_gate_list_from_binary: unimplemented gate")
    if len(params) - p0 != parameter_num or c.get_Parameter_Num() != parameter_num:
        raise Exception("import_gate_list_from_binary: parameter count of a block does not match its header")
    return c


def import_gate_list_from_binary(filename, device=0):
    """(Circuit, parameters) of a file written by the reference's export_gate_list_to_binary (or by the function below)"""
    with open(filename, "rb") as f:
        r = _Reader(f.read())
    params = []
    c = _read_block(r, params, device)
    return c, np.array(params, dtype=np.float64)


def _write_block(out, c, params, p):
    out.append(struct.pack("<iii", c.qbit_num, c.get_Parameter_Num(), c.get_Gate_Num()))
    for it in c._items:
        if isinstance(it, Circuit):
            out.append(struct.pack("<i", abi.BLOCK))
            p = _write_block(out, it, params, p)
            continue
        t = it.type
        if t not in _EXPORTABLE:
            raise Exception("export_gate_list_to_binary: unimplemented gate")
        out.append(struct.pack("<i", t))
        if t in _CTRL or t in _PARAM_CTRL:
            out.append(struct.pack("<ii", it.target, it.control))
        else:
            out.append(struct.pack("<i", it.target))
        n = it.n_params
        if n:
            out.append(struct.pack("<%dd" % n, *params[p:p + n]))
        p += n
    return p


def export_gate_list_to_binary(parameters, circuit, filename):
    """write ``circuit`` with ``parameters`` in the reference's binary gate-list format"""
    params = np.ascontiguousarray(parameters, dtype=np.float64).reshape(-1)
    if params.size != circuit.get_Parameter_Num():
        raise Exception("export_gate_list_to_binary: wrong number of parameters")
    out = []
    _write_block(out, circuit, params, 0)
    with open(filename, "wb") as f:
        f.write(b"".join(out))
