"""Host-side mirror of the cost path of the reference's VQE class.

``Variational_Quantum_Eigensolver`` keeps the constructor and method names of
``qgd_Variational_Quantum_Eigensolver_Base`` (squander/VQA/qgd_Variational_Quantum_Eigensolver_Base.py:145-420) for the calls
on the hot path: ``set_Ansatz``, ``Generate_Circuit``, ``set_Gate_Structure``, ``set_Initial_State``, ``get_Parameter_Num``,
``Optimization_Problem``, ``Optimization_Problem_Grad``, ``Optimization_Problem_Combined``, ``Optimization_Problem_Batch``,
``apply_to``, ``get_Second_Renyi_Entropy``. Every evaluation runs on the GPU through the C-ABI (sqgpu_vqe_energy[_grad]_batched): there is no CPU path.

``Start_Optimization`` (with ``set_Optimizer``, ``set_Optimized_Parameters``, ``get_Optimized_Parameters``) is the thin N1 layer
over that path: "COSINE" and "AGENTS" (the reference's parameter-shift engines, their shift batches as device batches),
and "BFGS" (L-BFGS, every line search one device batch).

The state-vector backend only; the density-matrix backend, the reference's other optimizers and the entropy helpers are outside
the hot path and run over this path through the drop-in of integration/.
"""
import numpy as np

from . import abi
from .circuit import Circuit
from .engine import Engine


class Variational_Quantum_Eigensolver:
    """E(theta) = Re <psi(theta)| H |psi(theta)>, psi = C(theta) |initial state>
    (Variational_Quantum_Eigensolver_Base::optimization_problem, ...Base.cpp:1088-1121)."""

    def __init__(self, Hamiltonian, qbit_num, config=None, accelerator_num=1, device=0, backend=None):
        if backend not in (None, "state_vector"):
            raise Exception("Unsupported backend '%s': the device path implements the state-vector backend" % backend)
        if accelerator_num < 1:
            raise Exception("accelerator_num should be >= 1: this package only provides the GPU path")
        self.qbit_num = int(qbit_num)
        rows = 1 << self.qbit_num
        # a scipy.sparse CSR matrix (the reference takes .data / .indices / .indptr of one) or an (indptr, indices, data) triple
        if hasattr(Hamiltonian, "indptr"):
            H = Hamiltonian.tocsr() if hasattr(Hamiltonian, "tocsr") else Hamiltonian
            indptr, indices, data = H.indptr, H.indices, H.data
        else:
            indptr, indices, data = Hamiltonian
        self._indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self._indices = np.ascontiguousarray(indices, dtype=np.int32)
        self._data = np.ascontiguousarray(data, dtype=np.complex128)
        if len(self._indptr) != rows + 1:
            raise Exception("Hamiltonian should be a 2^qbit_num x 2^qbit_num sparse matrix")
        self.config = dict(config or {})
        self.accelerator_num = int(accelerator_num)
        self.backend = "state_vector"
        self._device = int(device)
        self._ansatz = "HEA"  # Variational_Quantum_Eigensolver_Base.cpp: ansatz = HEA by default
        self._circuit = Circuit(self.qbit_num, device)
        self._state0 = np.zeros(rows, dtype=np.complex128)
        self._state0[0] = 1.0  # initialize_zero_state
        self._engine_obj = None
        self._circuit_key = None
        self._state_dirty = True

    # ---- structure -------------------------------------------------------------------------------------------------------
    def set_Ansatz(self, ansatz_new):
        if ansatz_new not in ("HEA", "HEA_ZYZ"):
            raise Exception("Variational_Quantum_Eigensolver: ansatz not implemented")
        self._ansatz = ansatz_new

    def Generate_Circuit(self, layers, inner_blocks=1):
        """generate_circuit (...Base.cpp:1299-1437): HEA = [U3, U3, CNOT] per pair, HEA_ZYZ = [RZ RY RZ] blocks + CNOT;
        pairs (1, 0), then for odd control c: (c + 2, c + 1) if it exists, then (c + 1, c)."""
        n = self.qbit_num
        c = Circuit(n, self._device)
        zyz = self._ansatz == "HEA_ZYZ"
        if n < (2 if zyz else 1):
            raise Exception("Variational_Quantum_Eigensolver_Base::generate_initial_circuit: number of qubits should be at least %d" % (2 if zyz else 1))

        def single(q):
            if zyz:
                b = Circuit(n, self._device)
                b.add_RZ(q)
                b.add_RY(q)
                b.add_RZ(q)
                c.add_Circuit(b)
            else:
                c.add_U3(q)

        def pair(first, second, tgt, ctl):
            for _ in range(inner_blocks):
                single(first)
                single(second)
                c.add_CNOT(tgt, ctl)

        for _ in range(layers):
            if n == 1:
                for _ in range(inner_blocks):
                    c.add_U3(0)
                continue
            pair(1, 0, 1, 0)
            for cq in range(1, n - 1, 2):
                if cq + 2 < n:
                    pair(cq + 1, cq + 2, cq + 2, cq + 1)
                pair(cq + 1, cq, cq + 1, cq)
        self._circuit = c

    def set_Gate_Structure(self, Gate_structure):
        if Gate_structure.qbit_num != self.qbit_num:
            raise Exception("set_Gate_Structure: qubit count mismatch")
        self._circuit = Circuit(self.qbit_num, self._device)
        self._circuit._items = list(Gate_structure._items)
        self._circuit._version = 1

    def set_Gate_Structure_from_Binary(self, filename):
        from . import gate_io

        circ, params = gate_io.import_gate_list_from_binary(filename)
        self.set_Gate_Structure(circ)
        self._optimized_parameters = params

    def set_Initial_State(self, initial_state):
        s = np.ascontiguousarray(initial_state, dtype=np.complex128).reshape(-1)
        if s.size != (1 << self.qbit_num):
            raise Exception("Initial state should have 2^qbit_num elements")
        self._state0 = s
        self._state_dirty = True

    def get_Circuit(self):
        return self._circuit

    def get_Qbit_Num(self):
        return self.qbit_num

    def get_Parameter_Num(self):
        return self._circuit.get_Parameter_Num()

    # ---- engine ----------------------------------------------------------------------------------------------------------
    @property
    def _engine(self):
        if self._engine_obj is None:
            self._engine_obj = Engine(self._device)
            self._engine_obj.set_hamiltonian_csr(self._indptr, self._indices, self._data)
        return self._engine_obj

    def _sync(self):
        eng = self._engine
        if self._state_dirty:
            eng.upload_matrix(self._state0)
            self._state_dirty = False
        key = self._circuit.structure_key()
        if key != self._circuit_key:
            eng.set_circuit(self._circuit)
            self._circuit_key = key
        return eng

    # ---- the hot path ----------------------------------------------------------------------------------------------------
    def Optimization_Problem(self, parameters):
        return float(self._sync().vqe_energy_batched(np.asarray(parameters, dtype=np.float64).reshape(1, -1))[0])

    def Optimization_Problem_Batch(self, parameters):
        p = np.asarray(parameters, dtype=np.float64)
        if p.ndim != 2:
            raise Exception("Optimization_Problem_Batch: parameters should be a 2 dimensional array")
        return self._sync().vqe_energy_batched(p)

    def Optimization_Problem_Combined(self, parameters):
        """optimization_problem_combined_non_static (...Base.cpp:1131-1199): (E, grad)"""
        e, g = self._sync().vqe_energy_grad_batched(np.asarray(parameters, dtype=np.float64).reshape(1, -1))
        return float(e[0]), g[0]

    def Optimization_Problem_Combined_Batch(self, parameters):
        return self._sync().vqe_energy_grad_batched(np.asarray(parameters, dtype=np.float64))

    def Optimization_Problem_Grad(self, parameters):
        return self.Optimization_Problem_Combined(parameters)[1]

    # ---- optimisation over the hot path --------------------------------------------------------------
    def set_Optimizer(self, alg="COSINE"):
        if alg not in ("COSINE", "AGENTS", "BFGS", "GRAD_DESCEND", "AGENTS_COMBINED", "GRAD_DESCEND_PARAMETER_SHIFT_RULE"):
            raise Exception("set_Optimizer: '%s' is not provided by this package (COSINE, AGENTS, BFGS, GRAD_DESCEND, AGENTS_COMBINED, GRAD_DESCEND_PARAMETER_SHIFT_RULE); use the reference's "
                            "engines over the GPU energy path through the drop-in of integration/" % alg)
        self._optimizer = alg

    def set_Optimized_Parameters(self, parameters):
        p = np.ascontiguousarray(parameters, dtype=np.float64).reshape(-1)
        if p.size != self.get_Parameter_Num():
            raise Exception("Number of free parameters should be %d, but got %d" % (self.get_Parameter_Num(), p.size))
        self._optimized_parameters = p.copy()

    def get_Optimized_Parameters(self):
        if getattr(self, "_optimized_parameters", None) is None:
            raise Exception("get_Optimized_Parameters: no parameters have been set or optimised")
        return self._optimized_parameters.copy()

    def Start_Optimization(self):
        """start_optimization (...Base.cpp:100-160): minimise the energy from the stored parameters (set_Optimized_Parameters;
        random in [0, 2 pi) otherwise, as the reference's engines draw them) with the engine chosen by set_Optimizer. Config keys
        as in the reference: max_inner_iterations[_cosine], batch_size[_cosine], check_for_convergence, seed.
        Returns the energy; the parameters are in get_Optimized_Parameters()."""
        from . import optimize

        eng = self._sync()
        P = self.get_Parameter_Num()
        cfg = self.config
        rng = np.random.default_rng(int(cfg.get("seed", 0)))
        x0 = getattr(self, "_optimized_parameters", None)
        if x0 is None or x0.size != P:
            x0 = rng.random(P) * 2 * np.pi
        alg = getattr(self, "_optimizer", "COSINE")
        max_iter = int(cfg.get("max_inner_iterations", 1000))

        def energy_grad(x):
            e, g = eng.vqe_energy_grad_batched(x.reshape(1, -1))
            return float(e[0]), g[0]

        ne = 0
        if alg == "COSINE":
            # COSINE.cpp:226-228: cost_fnc == VQE selects the three-point rule with the doubled period (shifts pi/4, pi/2)
            x, f, it, ne = optimize.cosine(eng.vqe_energy_batched, x0, rng,
                                           batch_size=min(P, int(cfg.get("batch_size_cosine", cfg.get("batch_size", min(64, P))))),
                                           max_iter=int(cfg.get("max_inner_iterations_cosine", max_iter)), tol=-np.inf, double_period=True,
                                           check_for_convergence=bool(cfg.get("check_for_convergence", 1)))

        if alg == "GRAD_DESCEND_PARAMETER_SHIFT_RULE":  # …SHIFT_RULE.cpp:249-325 over the batched energy
            x, f, it, ne = optimize.grad_descend_shift_rule(
                eng.vqe_energy_batched, x0, rng, batch_size=min(P, int(cfg.get("batch_size_grad_descend_shift_rule", cfg.get("batch_size", min(64, P))))),
                max_iter=int(cfg.get("max_inner_iterations_grad_descend_shift_rule", max_iter)), tol=-np.inf,
                eta=float(cfg.get("eta_grad_descend_shift_rule", cfg.get("eta", 1e-3))), use_line_search=bool(int(cfg.get("use_line_search", 1))))

        def line_search(x, d, alphas):  # all trial step lengths of an iteration: one batched energy+gradient call
            e, g = eng.vqe_energy_grad_batched(x[None, :] + np.asarray(alphas)[:, None] * d[None, :])
            return e, g @ d

        gtol = float(cfg.get("gradient_tolerance", 1e-8))
        if alg in ("AGENTS", "AGENTS_COMBINED"):
            # AGENTS.cpp:334: cost_fnc == VQE with linesearch_points == 3 (the default) is the doubled-period three-point rule;
            # randomize_parameters does not scale the radius by the cost for the VQE (Optimization_Interface.cpp:596)
            x, f, it, ne = optimize.agents(eng.vqe_energy_batched, x0, rng, agent_num=int(cfg.get("agent_num_agent", cfg.get("agent_num", 64))),
                                           max_iter=int(cfg.get("max_inner_iterations_agent", max_iter)), tol=-np.inf, double_period=True,
                                           agent_lifetime=int(cfg.get("agent_lifetime_agent", cfg.get("agent_lifetime", 1000))),
                                           exploration_rate=float(cfg.get("agent_exploration_rate", 0.2)),
                                           agent_randomization_rate=float(cfg.get("agent_randomization_rate", 0.2)),
                                           radius=float(cfg.get("Randomized_Radius", 1.0)),
                                           convergence_length=int(cfg.get("convergence_length_agent", cfg.get("convergence_length", 20))), scale_by_cost=False,
                                           five_point=int(cfg.get("linesearch_points_agent", cfg.get("linesearch_points", 3))) == 5)  # AGENTS.cpp:211-221, 335
            x0 = x
        if alg in ("BFGS", "GRAD_DESCEND", "AGENTS_COMBINED"):
            # GRAD_DESCEND / the second stage of AGENTS_COMBINED (AGENTS.cpp:914-933): steepest descent, same batched line search
            x, f, it, ne2 = optimize.lbfgs(energy_grad, line_search, x0, max_iter=int(cfg.get("max_inner_iterations_grad_descend", max_iter)) if alg != "BFGS" else max_iter,
                                           tol=-np.inf, gtol=gtol, **({} if alg == "BFGS" else {"history": 0}))
            ne += ne2
        self._optimized_parameters = np.asarray(x, dtype=np.float64).copy()
        self._num_evaluations = getattr(self, "_num_evaluations", 0) + ne
        self._current_minimum = float(f)
        return float(f)

    def set_Optimization_Tolerance(self, tolerance):
        self.config["gradient_tolerance"] = float(tolerance)

    def set_Project_Name(self, project_name):
        self.project_name = str(project_name)

    def get_Second_Renyi_Entropy(self, parameters=None, input_state=None, qubit_list=None):
        """qgd_Variational_Quantum_Eigensolver_Base.get_Second_Renyi_Entropy (…Base.py:288-323): the entropy of the ansatz state
        on a subset of the qubits (default input: |0...0>)"""
        return self._circuit.get_Second_Renyi_Entropy(parameters, input_state, qubit_list)

    def apply_to(self, parameters_mtx, state_to_be_transformed):
        """in place: state <- C(parameters) state"""
        self._circuit.apply_to(parameters_mtx, state_to_be_transformed)
