"""B200-native engine for SQUANDER's decomposition hot path.

Import with ``importlib.import_module("sequential-quantum-gate-decomposer_b200")`` or through the root-level
alias module ``squander_b200`` (the directory name has hyphens, so a plain ``import`` statement cannot name it).

Public surface = the reference's own names for this path (squander/__init__.py:1-60):
``Circuit``, ``N_Qubit_Decomposition_adaptive``, ``N_Qubit_Decomposition_custom``, ``N_Qubit_State_Preparation_adaptive``, plus ``Engine`` (the C-ABI
handle) and ``Variational_Quantum_Eigensolver`` (state-vector cost path).
"""
from . import abi
from . import qasm
from . import dist
from . import gate_io
from . import optimize
from .circuit import Circuit
from .engine import Engine
from .decomposition import N_Qubit_Decomposition_adaptive, N_Qubit_Decomposition_custom, N_Qubit_State_Preparation_adaptive
from .vqe import Variational_Quantum_Eigensolver

# the reference exports the circuit class under both names (squander/__init__.py)
qgd_Circuit = Circuit
qgd_Variational_Quantum_Eigensolver_Base = Variational_Quantum_Eigensolver

__all__ = [
    "abi", "qasm", "dist", "gate_io", "optimize", "Circuit", "qgd_Circuit", "Engine", "N_Qubit_Decomposition_adaptive", "N_Qubit_Decomposition_custom",
    "N_Qubit_State_Preparation_adaptive", "Variational_Quantum_Eigensolver", "qgd_Variational_Quantum_Eigensolver_Base",
]
