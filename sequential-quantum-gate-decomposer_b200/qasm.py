"""OpenQASM 2 -> (Circuit, parameters) without Qiskit.

The reference imports circuits through Qiskit (squander/IO_interfaces/Qiskit_IO.py:278-560 convert_Qiskit_to_Squander,
used by tests/decomposition/test_parametric_circuit.py:140-143); Qiskit is not part of this image, so this module
parses the subset of qelib1 that path understands and applies the SAME conventions:

  * two-qubit gates "g q[a],q[b]": a is the control, b the target -> add_G(target_qbit=b, control_qbit=a)
    (Qiskit_IO.py:351-356, 415-420);
  * rotation angles of u/u3, cu, rx, ry, rz, r, crx, cry, crz are stored halved ("SQUANDER works with theta/2",
    Qiskit_IO.py:330, 365, 379, 393, 432-461); u1/p, u2, cp/cu1 parameters are stored as they are;
  * parameters are appended in gate order, so they line up with Gates_block's parameter layout.
"""
import ast
import math
import operator
import re

import numpy as np

from .circuit import Circuit

_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv, ast.Pow: operator.pow}
_UN = {ast.USub: operator.neg, ast.UAdd: operator.pos}
_FN = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log, "sqrt": math.sqrt}


def _eval(node):
    if isinstance(node, ast.Expression):
        return _eval(node.body)
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return float(node.value)
    if isinstance(node, ast.Name) and node.id == "pi":
        return math.pi
    if isinstance(node, ast.BinOp) and type(node.op) in _BIN:
        return _BIN[type(node.op)](_eval(node.left), _eval(node.right))
    if isinstance(node, ast.UnaryOp) and type(node.op) in _UN:
        return _UN[type(node.op)](_eval(node.operand))
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FN and len(node.args) == 1:
        return _FN[node.func.id](_eval(node.args[0]))
    raise ValueError("unsupported expression in QASM parameter")


def eval_param(text):
    """arithmetic over numbers and pi only (no names, no attribute access)"""
    return _eval(ast.parse(text.strip().replace("^", "**"), mode="eval"))


# name -> (adder, number of qubits, indices of the parameters that are stored halved)
_ONE = {
    "u": ("add_U3", {0}), "u3": ("add_U3", {0}), "u2": ("add_U2", set()), "u1": ("add_U1", set()), "p": ("add_U1", set()),
    "rx": ("add_RX", {0}), "ry": ("add_RY", {0}), "rz": ("add_RZ", {0}), "r": ("add_R", {0}),
    "h": ("add_H", set()), "x": ("add_X", set()), "y": ("add_Y", set()), "z": ("add_Z", set()), "s": ("add_S", set()),
    "sdg": ("add_Sdg", set()), "t": ("add_T", set()), "tdg": ("add_Tdg", set()), "sx": ("add_SX", set()),
    "sxdg": ("add_SXdg", set()),
}
_CTRL = {
    "cx": ("add_CNOT", set()), "cz": ("add_CZ", set()), "ch": ("add_CH", set()), "cu": ("add_CU", {0}),
    "cry": ("add_CRY", {0}), "crx": ("add_CRX", {0}), "crz": ("add_CRZ", {0}), "cp": ("add_CP", set()),
    "cu1": ("add_CP", set()),
}
_TWO = {"swap": "add_SWAP", "rxx": "add_RXX", "ryy": "add_RYY", "rzz": "add_RZZ"}
_HALVED_TWO = {"rxx", "ryy", "rzz"}

_STMT = re.compile(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:\((.*)\))?\s+(.*)$", re.S)


def _split_args(text):
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def loads(text):
    """parse QASM source; returns (Circuit, numpy parameter vector)"""
    text = re.sub(r"//[^\n]*", "", text)
    regs = {}  # name -> (offset, size)
    total = 0
    stmts = [s.strip() for s in text.split(";") if s.strip()]
    body = []
    for st in stmts:
        if st.startswith("OPENQASM") or st.startswith("include") or st.startswith("creg") or st.startswith("barrier"):
            continue
        m = re.match(r"qreg\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*(\d+)\s*\]", st)
        if m:
            regs[m.group(1)] = (total, int(m.group(2)))
            total += int(m.group(2))
            continue
        if st.startswith("measure") or st.startswith("reset"):
            raise ValueError("non-unitary QASM statement: " + st)
        body.append(st)
    if total == 0:
        raise ValueError("no qreg declaration")
    circ = Circuit(total)
    params = []

    def qubit(tok):
        m = re.match(r"\s*([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*(\d+)\s*\]\s*$", tok)
        if not m or m.group(1) not in regs:
            raise ValueError("bad qubit reference: " + tok)
        off, size = regs[m.group(1)]
        idx = int(m.group(2))
        if idx >= size:
            raise ValueError("qubit index out of range: " + tok)
        return off + idx

    for st in body:
        m = _STMT.match(st)
        if not m:
            raise ValueError("cannot parse QASM statement: " + st)
        name, ptxt, qtxt = m.group(1).lower(), m.group(2), m.group(3)
        pv = [eval_param(a) for a in _split_args(ptxt)] if ptxt else []
        qs = [qubit(t) for t in qtxt.split(",")]
        if name in _ONE:
            adder, halved = _ONE[name]
            getattr(circ, adder)(qs[0])
        elif name in _CTRL:
            adder, halved = _CTRL[name]
            getattr(circ, adder)(qs[1], qs[0])  # target = second operand, control = first
        elif name in _TWO:
            getattr(circ, _TWO[name])([qs[1], qs[0]])
            halved = {0} if name in _HALVED_TWO else set()
        elif name == "ccx":
            circ.add_CCX(qs[2], [qs[1], qs[0]])
            halved = set()
        elif name == "cswap":
            circ.add_CSWAP([qs[2], qs[1]], [qs[0]])
            halved = set()
        else:
            raise ValueError("unsupported QASM gate: " + name)
        for i, v in enumerate(pv):
            params.append(v / 2 if i in halved else v)
    if len(params) != circ.get_Parameter_Num():
        raise ValueError("parameter count mismatch while importing QASM")
    return circ, np.array(params, dtype=np.float64)


def load(path):
    with open(path) as f:
        return loads(f.read())


# ---- the way out: (Circuit, parameters) -> OpenQASM 2 ---------------------------------------------------------------------
# (the reference exports through Qiskit, Qiskit_IO.get_Qiskit_Circuit, squander/IO_interfaces/Qiskit_IO.py:40-275; same
# conventions as above, inverted: halved angles are doubled, "g q[control],q[target]")
def _export_tables():
    from . import abi

    one = {abi.U3: ("u3", {0}), abi.U2: ("u2", set()), abi.U1: ("u1", set()), abi.RX: ("rx", {0}), abi.RY: ("ry", {0}), abi.RZ: ("rz", {0}),
           abi.R: ("r", {0}), abi.H: ("h", set()), abi.X: ("x", set()), abi.Y: ("y", set()), abi.Z: ("z", set()), abi.S: ("s", set()),
           abi.SDG: ("sdg", set()), abi.T: ("t", set()), abi.TDG: ("tdg", set()), abi.SX: ("sx", set()), abi.SXDG: ("sxdg", set())}
    ctrl = {abi.CNOT: ("cx", set()), abi.CZ: ("cz", set()), abi.CH: ("ch", set()), abi.CU: ("cu", {0}), abi.CRY: ("cry", {0}),
            abi.CRX: ("crx", {0}), abi.CRZ: ("crz", {0}), abi.CP: ("cp", set())}
    two = {abi.SWAP: ("swap", set()), abi.RXX: ("rxx", {0}), abi.RYY: ("ryy", {0}), abi.RZZ: ("rzz", {0})}
    return one, ctrl, two


def dumps(circuit, parameters, adaptive_as_cry=False):
    """OpenQASM 2 source of ``circuit`` at ``parameters`` (nested blocks are flattened). ``loads(dumps(c, p))`` returns the flat
    structure of c and p again. Gates outside qelib1 (GENERAL, CROT, CR, SYC) raise ValueError; the adaptive gate is a CRY
    (gates/Adaptive.cpp) and is written as one with ``adaptive_as_cry`` (the structure then reads back with CRY in its place)."""
    from . import abi

    one, ctrl, two = _export_tables()
    p = np.asarray(parameters, dtype=np.float64).reshape(-1)
    if p.size != circuit.get_Parameter_Num():
        raise ValueError("Number of free parameters should be %d, but got %d" % (circuit.get_Parameter_Num(), p.size))
    lines = ["OPENQASM 2.0;", 'include "qelib1.inc";', "qreg q[%d];" % circuit.qbit_num]
    pos = 0

    def args(n, halved):
        nonlocal pos
        vals = [repr(float(p[pos + i] * 2 if i in halved else p[pos + i])) for i in range(n)]
        pos += n
        return "(" + ",".join(vals) + ")" if n else ""

    for g in circuit._flat_gates():
        t = g.type
        if t == abi.ADAPTIVE and adaptive_as_cry:
            t = abi.CRY
        n = abi.PARAM_COUNT[g.type]
        if t in one:
            name, halved = one[t]
            lines.append("%s%s q[%d];" % (name, args(n, halved), g.target))
        elif t in ctrl:
            name, halved = ctrl[t]
            lines.append("%s%s q[%d],q[%d];" % (name, args(n, halved), g.control, g.target))
        elif t in two:
            name, halved = two[t]
            lines.append("%s%s q[%d],q[%d];" % (name, args(n, halved), g.target2, g.target))
        elif t == abi.CCX:
            lines.append("ccx q[%d],q[%d],q[%d];" % (g.control2, g.control, g.target))
        elif t == abi.CSWAP:
            lines.append("cswap q[%d],q[%d],q[%d];" % (g.control, g.target2, g.target))
        else:
            raise ValueError("gate %s has no OpenQASM 2 (qelib1) counterpart" % abi.GATE_NAMES.get(g.type, g.type))
    return "\n".join(lines) + "\n"


def dump(circuit, parameters, path, **kw):
    with open(path, "w") as f:
        f.write(dumps(circuit, parameters, **kw))
