"""Host-side mirror of the cost path of the reference decomposition classes.

``N_Qubit_Decomposition_adaptive`` keeps the constructor and method names of the CPython wrapper
(squander/decomposition/qgd_N_Qubit_Decompositions_Wrapper.cpp:281-291, 3125-3248) for the calls that sit on the
hot path: ``add_Adaptive_Layers``, ``add_Finalyzing_Layer_To_Gate_Structure``, ``set_Gate_Structure``,
``get_Parameter_Num``, ``set_Cost_Function_Variant``, ``set_Trace_Offset``, ``Optimization_Problem``,
``Optimization_Problem_Grad``, ``Optimization_Problem_Combined``, ``Optimization_Problem_Batch``,
``Optimization_Problem_Combined_Unitary``, ``get_Matrix``. ``accelerator_num`` is the number of GPUs, exactly the
kwarg the reference reserves for accelerators (Optimization_Interface.cpp:189-197); it must be >= 1 here --
the CPU path belongs to the reference, not to this package.

``Start_Decomposition`` / ``set_Optimizer`` / ``get_Optimized_Parameters`` are a THIN control loop over that cost path
(SURVEY.md §8f N4): the level search of determine_initial_gate_structure (N_Qubit_Decomposition_adaptive.cpp:786-1037) with
this package's own optimizer loops (optimize.py: L-BFGS with a device-batched line search, device-resident ADAM). The
reference's full control plane -- its optimizer engines, compression, CRY -> CNOT finalisation -- is not
mirrored here; it runs unchanged over the GPU cost path through the drop-in of integration/.
"""
import numpy as np

from . import abi
from .circuit import Circuit
from .engine import Engine


class N_Qubit_Decomposition_custom:
    """Cost path of Optimization_Interface over a user-supplied gate structure."""

    def __init__(self, Umtx, qbit_num=-1, config=None, accelerator_num=1, device=0):
        U = np.ascontiguousarray(Umtx, dtype=np.complex128)
        if U.ndim != 2:
            raise Exception("Umtx should be a 2 dimensional complex array")
        rows = U.shape[0]
        n = int(round(np.log2(rows)))
        if (1 << n) != rows:
            raise Exception("Umtx should have 2^n rows")
        if qbit_num not in (-1, n):
            raise Exception("qbit_num does not match the size of Umtx")
        if U.shape[1] > rows:
            raise Exception("Umtx cannot have more columns than rows")
        if accelerator_num < 1:
            raise Exception("accelerator_num should be >= 1: this package only provides the GPU path")
        self.qbit_num = n
        self.Umtx = U
        self.config = dict(config or {})
        self.accelerator_num = int(accelerator_num)
        self._circuit = Circuit(n, device)
        self._circuit_key = None
        self._variant = abi.FROBENIUS_NORM
        self._trace_offset = 0
        self._prev_cost = 1.0  # Optimization_Interface.cpp:74-76
        self._c1 = 1 / 1.7
        self._c2 = 1 / 2.0
        # the device context is created with the first evaluation (the gate-structure methods need no GPU); there is
        # no CPU evaluation path: without a CUDA device that first evaluation raises SqgpuError
        self._device = int(device)
        self._engine_obj = None
        self._dirty = True
        self._optimizer = "BFGS"  # Optimization_Interface: set_optimizer(BFGS) is the constructors' default for small problems
        self._optimization_tolerance = float(self.config.get("optimization_tolerance", 1e-4))  # ..._adaptive.cpp:539-544
        self._optimized_parameters = None
        self._current_minimum = None
        self._num_evaluations = 0

    @property
    def _engine(self):
        if self._engine_obj is None:
            # accelerator_num = G > 1: one handle over GPUs device .. device + G - 1, sharded inside the library
            # (the reference splits its batched cost path over accelerators the same way, Optimization_Interface.cpp:806-832)
            if self.accelerator_num > 1:
                self._engine_obj = Engine(devices=list(range(self._device, self._device + self.accelerator_num)))
            else:
                self._engine_obj = Engine(self._device)
            self._engine_obj.upload_matrix(self.Umtx)
        return self._engine_obj

    # ---- gate structure -------------------------------------------------------------------------------------
    def set_Gate_Structure(self, circuit):
        """Optimization_Interface::set_custom_gate_structure (Optimization_Interface.cpp:1770-1778)."""
        if circuit.qbit_num != self.qbit_num:
            raise Exception("set_Gate_Structure: qubit count mismatch")
        self._circuit = Circuit(self.qbit_num, self._device)
        self._circuit._items = list(circuit._items)  # release_gates(); combine(gate_structure_in)
        self._circuit._version = 1
        self._dirty = True

    def get_Circuit(self):
        return self._circuit

    # ---- the wrapper's data-format methods either side of the path (qgd_N_Qubit_Decompositions_Wrapper.cpp:3125-3248) -----
    def get_Gate_Num(self):
        return self._circuit.get_Gate_Num()

    def get_Unitary(self):
        return self.Umtx.copy()

    def set_Unitary(self, Umtx):
        """Decomposition_Base::set_unitary: a new matrix of the same register for the same gate structure; it is uploaded
        with the next evaluation (or at once by Upload_Umtx_to_DFE)"""
        U = np.ascontiguousarray(Umtx, dtype=np.complex128)
        if U.ndim != 2 or U.shape[0] != (1 << self.qbit_num) or U.shape[1] > U.shape[0]:
            raise Exception("set_Unitary: Umtx should be a 2^qbit_num x cols complex array, cols <= rows")
        self.Umtx = U
        if self._engine_obj is not None:
            self._engine_obj.upload_matrix(self.Umtx)
        self._dirty = True

    def Upload_Umtx_to_DFE(self):
        """upload_Umtx_to_DFE (Optimization_Interface.cpp:1819-1824), the hook the optimizers call before they start: the
        matrix (and the gate structure) become resident on the device now instead of with the first evaluation"""
        self._sync()

    def export_Unitary(self, filename):
        """Decomposition_Base::export_unitary (Decomposition_Base.cpp:1128-1146): int32 rows, int32 cols, rows x cols complex128"""
        if getattr(self, "project_name", ""):
            filename = self.project_name + "_" + filename
        with open(filename, "wb") as f:
            f.write(np.array(self.Umtx.shape, dtype=np.int32).tobytes())
            f.write(np.ascontiguousarray(self.Umtx).tobytes())

    def set_Unitary_From_Binary(self, filename):
        """Decomposition_Base::import_unitary_from_binary (Decomposition_Base.cpp:1154-1177)"""
        if getattr(self, "project_name", ""):
            filename = self.project_name + "_" + filename
        with open(filename, "rb") as f:
            rows, cols = np.frombuffer(f.read(8), dtype=np.int32)
            data = np.frombuffer(f.read(), dtype=np.complex128)
        if rows <= 0 or cols <= 0 or data.size != int(rows) * int(cols):
            raise Exception("set_Unitary_From_Binary: truncated or malformed file")
        self.set_Unitary(data.reshape(int(rows), int(cols)).copy())

    def set_Gate_Structure_From_Binary(self, filename):
        """set_adaptive_gate_structure(filename) (N_Qubit_Decomposition_adaptive.cpp:1398-1420): gate structure and parameters
        from a file written by export_gate_list_to_binary"""
        from . import gate_io

        circ, params = gate_io.import_gate_list_from_binary(filename, self._device)
        self.set_Gate_Structure(circ)
        self._optimized_parameters = np.asarray(params, dtype=np.float64).copy()

    def add_Gate_Structure_From_Binary(self, filename):
        """add_adaptive_gate_structure(filename) (N_Qubit_Decomposition_adaptive.cpp:1430-1460): the stored gates are applied
        AFTER the current structure, their parameters follow the current ones"""
        from . import gate_io

        circ, params = gate_io.import_gate_list_from_binary(filename, self._device)
        if circ.qbit_num != self.qbit_num:
            raise Exception("add_Gate_Structure_From_Binary: qubit count mismatch")
        old = self._optimized_parameters if self._optimized_parameters is not None else np.zeros(self.get_Parameter_Num())
        self._circuit.add_Circuit(circ)
        self._optimized_parameters = np.concatenate([old, np.asarray(params, dtype=np.float64)])
        self._dirty = True

    def import_Qiskit_Circuit(self, qc_in):
        """import_Qiskit_Circuit of the reference wrapper (Qiskit_IO.convert_Qiskit_to_Squander, then set_Gate_Structure +
        set_Optimized_Parameters) without Qiskit: ``qc_in`` is OpenQASM 2 source, the path of a .qasm file, or any object whose
        ``qasm()`` method returns the source (a Qiskit QuantumCircuit up to 0.46; newer ones: pass qiskit.qasm2.dumps(qc))"""
        import os

        from . import qasm

        if hasattr(qc_in, "qasm") and callable(qc_in.qasm):
            qc_in = qc_in.qasm()
        if not isinstance(qc_in, str):
            raise Exception("import_Qiskit_Circuit: expected OpenQASM 2 source, a .qasm path or an object with qasm()")
        circ, params = qasm.load(qc_in) if ("OPENQASM" not in qc_in and os.path.exists(qc_in)) else qasm.loads(qc_in)
        if circ.qbit_num != self.qbit_num:
            raise Exception("import_Qiskit_Circuit: the circuit has %d qubits, the decomposition %d" % (circ.qbit_num, self.qbit_num))
        self.set_Gate_Structure(circ)
        self._optimized_parameters = np.asarray(params, dtype=np.float64).copy()

    def Reorder_Qubits(self, qbit_list):
        """Decomposition_Base::reorder_qubits (Decomposition_Base.cpp:910-950, Gate.cpp:1151-1200): the new qubit ``idx`` is the
        old qubit ``qbit_list[idx]`` -- in every gate and in the rows and columns of the unitary, so the cost of a parameter
        vector does not change"""
        n = self.qbit_num
        ql = [int(q) for q in qbit_list]
        if sorted(ql) != list(range(n)):
            raise Exception("Reorder_Qubits: Wrong number of qubits.")
        if self.Umtx.shape[0] != self.Umtx.shape[1]:
            raise Exception("Reorder_Qubits: the unitary should be square")
        circ = self._circuit.Remap_Qbits({q: idx for idx, q in enumerate(ql)})
        idx = np.arange(1 << n)
        perm = np.zeros(1 << n, dtype=np.int64)
        for new_bit, old_bit in enumerate(ql):
            perm |= ((idx >> old_bit) & 1) << new_bit
        U = np.empty_like(self.Umtx)
        U[np.ix_(perm, perm)] = self.Umtx
        self.set_Gate_Structure(circ)
        self.set_Unitary(np.ascontiguousarray(U))

    def get_QASM(self, adaptive_as_cry=True):
        """the circuit at the optimised parameters as OpenQASM 2 source (the reference hands out a Qiskit circuit,
        get_Qiskit_Circuit; this is the Qiskit-free counterpart, readable by qiskit.QuantumCircuit.from_qasm_str)"""
        from . import qasm

        return qasm.dumps(self._circuit, self.get_Optimized_Parameters(), adaptive_as_cry=adaptive_as_cry)

    def get_Qiskit_Circuit(self):
        """the decomposition as a Qiskit QuantumCircuit (the reference goes through Qiskit_IO.get_Qiskit_Circuit): built from
        get_QASM() when Qiskit is installed; this image has none, then the QASM source is what there is to hand out"""
        try:
            from qiskit import QuantumCircuit
        except ImportError:
            raise Exception("get_Qiskit_Circuit: Qiskit is not installed; get_QASM() returns the same circuit as OpenQASM 2 source")
        return QuantumCircuit.from_qasm_str(self.get_QASM())

    def get_Project_Name(self):
        return getattr(self, "project_name", "")

    def set_Project_Name(self, project_name):
        self.project_name = str(project_name)

    def set_Max_Iterations(self, max_iterations):
        """Decomposition_Base::set_max_inner_iterations"""
        self.config["max_inner_iterations"] = int(max_iterations)

    def set_Verbose(self, verbose):
        self.verbose = int(verbose)

    def set_Debugfile(self, debugfile):
        self.debugfile = str(debugfile)

    def List_Gates(self):
        """Gates_block::list_gates: one line per gate, in application order"""
        names = {v: k for k, v in vars(abi).items() if isinstance(v, int) and k.isupper() and not k.startswith(("ERR", "SHARD", "OK"))}
        descs = self._circuit.descriptors()[0]
        for i, r in enumerate(descs):
            print("%d: %s target %d control %d parameters %d" % (i, names.get(int(r["type"]), str(int(r["type"]))), int(r["target"]), int(r["control"]), int(r["n_params"])))

    def get_Second_Renyi_Entropy(self, parameters=None, input_state=None, qubit_list=None):
        return self._circuit.get_Second_Renyi_Entropy(parameters, input_state, qubit_list)

    def get_Parameter_Num(self):
        return self._circuit.get_Parameter_Num()

    def get_Qbit_Num(self):
        return self.qbit_num

    # ---- optimisation over the GPU cost path (thin; see the module docstring) ------------------------------------
    def set_Optimizer(self, optimizer="BFGS"):
        """"BFGS": L-BFGS with a device-batched line search; "ADAM": device-resident ADAM trajectories; "COSINE": the reference's
        parameter-shift engine with its shift batches and its line search as device batches (optimize.cosine); "AGENTS": the
        reference's independent walkers, all their shifted parameter sets one device batch per iteration (optimize.agents);
        "GRAD_DESCEND": steepest descent with the batched line search (Grad_Descend, common/grad_descend.cpp:459-480: d = -g);
        "AGENTS_COMBINED": AGENTS, then GRAD_DESCEND from its result (AGENTS.cpp:914-933). The
        reference's other engines (BAYES_OPT, BFGS2 ...) stay with the reference: they run over this cost path through integration/."""
        if optimizer not in ("BFGS", "ADAM", "COSINE", "AGENTS", "GRAD_DESCEND", "AGENTS_COMBINED", "GRAD_DESCEND_PARAMETER_SHIFT_RULE"):
            raise Exception("set_Optimizer: '%s' is not provided by this package (BFGS, ADAM, COSINE, AGENTS, GRAD_DESCEND, AGENTS_COMBINED, "
                            "GRAD_DESCEND_PARAMETER_SHIFT_RULE); use the reference's engines "
                            "over the GPU cost path through the drop-in of integration/" % optimizer)
        self._optimizer = optimizer

    def set_Optimization_Tolerance(self, tolerance):
        self._optimization_tolerance = float(tolerance)

    def get_Optimized_Parameters(self):
        if self._optimized_parameters is None:
            raise Exception("get_Optimized_Parameters: no optimisation has been run")
        return self._optimized_parameters.copy()

    def set_Optimized_Parameters(self, parameters):
        p = np.ascontiguousarray(parameters, dtype=np.float64).reshape(-1)
        if p.size != self.get_Parameter_Num():
            raise Exception("Number of free parameters should be %d, but got %d" % (self.get_Parameter_Num(), p.size))
        self._optimized_parameters = p.copy()

    def get_Decomposition_Error(self):
        return self._current_minimum

    def get_Num_of_Iters(self):
        return self._num_evaluations

    def _optimize_structure(self, rng, x0=None):
        """minimise the cost over the current gate structure; returns (parameters, cost). x0: a start to polish (the
        final optimisation of finalize_circuit) instead of random starts."""
        from . import optimize

        eng = self._sync()
        P = self.get_Parameter_Num()
        tol = self._optimization_tolerance
        if P == 0:
            return np.zeros(0), float(eng.cost_batched(np.zeros((1, 0)))[0])
        if self._optimizer == "ADAM":
            starts = int(self.config.get("adam_trajectories", 16)) if x0 is None else 1
            steps_max = int(self.config.get("max_inner_iterations", 4000))
            X0 = rng.random((starts, P)) * 2 * np.pi
            X0[0] = 0.0 if x0 is None else np.asarray(x0, dtype=np.float64)
            eng.adam_init(X0, eta=float(self.config.get("eta", 1e-2)))
            done = 0
            while done < steps_max:
                hist = eng.adam_steps(200)
                done += 200
                self._num_evaluations += 200 * starts
                if hist.min() < tol:
                    break
            _, best_cost, best_theta, _ = eng.adam_get()
            b = int(np.argmin(best_cost))
            return best_theta[b], float(best_cost[b])

        if self._optimizer in ("COSINE", "AGENTS", "AGENTS_COMBINED"):
            # COSINE.cpp:226-228, 411-414 / AGENTS.cpp:333-335: the three-point rule is defined for the Frobenius cost (a sinusoid
            # of period 2 pi in every parameter); COSINE throws for the other variants, AGENTS takes its five-point rule for the
            # Hilbert-Schmidt test
            five_point = self._optimizer != "COSINE" and self._variant == abi.HILBERT_SCHMIDT_TEST  # AGENTS.cpp:335, 536-660
            if self._variant != abi.FROBENIUS_NORM and not five_point:
                raise Exception("solve_layer_optimization_problem_%s: Not implemented method." % self._optimizer)
            cfg = self.config
            if self._optimizer in ("AGENTS", "AGENTS_COMBINED"):
                x, f, _, ne = optimize.agents(
                    eng.cost_batched, rng.random(P) * 2 * np.pi if x0 is None else x0, rng,
                    agent_num=int(cfg.get("agent_num_agent", cfg.get("agent_num", 64))),
                    max_iter=int(cfg.get("max_inner_iterations_agent", cfg.get("max_inner_iterations", 10000))),
                    tol=float(cfg.get("optimization_tolerance_agent", tol)),
                    agent_lifetime=int(cfg.get("agent_lifetime_agent", cfg.get("agent_lifetime", 1000))),
                    exploration_rate=float(cfg.get("agent_exploration_rate_agent", cfg.get("agent_exploration_rate", 0.2))),
                    agent_randomization_rate=float(cfg.get("agent_randomization_rate", 0.2)), radius=float(cfg.get("Randomized_Radius", 1.0)),
                    convergence_length=int(cfg.get("convergence_length_agent", cfg.get("convergence_length", 20))), five_point=five_point)
                self._num_evaluations += ne
                if self._optimizer == "AGENTS_COMBINED":  # AGENTS.cpp:914-933: gradient descent from the agents' result
                    x, f, _, ne = optimize.lbfgs(lambda v: tuple(a[0] for a in eng.cost_grad_batched(v.reshape(1, -1))), eng.line_search_batched, x,
                                                 max_iter=int(cfg.get("max_inner_iterations_grad_descend", cfg.get("max_inner_iterations", 2000))),
                                                 tol=tol * 1e-2, history=0)
                    self._num_evaluations += ne
                return x, f
            x, f, _, ne = optimize.cosine(
                eng.cost_batched, rng.random(P) * 2 * np.pi if x0 is None else x0, rng,
                batch_size=min(P, int(cfg.get("batch_size_cosine", cfg.get("batch_size", min(64, P))))),
                max_iter=int(cfg.get("max_inner_iterations_cosine", cfg.get("max_inner_iterations", 2000))),
                tol=float(cfg.get("optimization_tolerance_cosine", tol)), double_period=False,
                check_for_convergence=bool(cfg.get("check_for_convergence", cfg.get("check_for_convergence_cosine", 1))),
                # the shift batch from one adjoint sweep instead of 2 x batch_size forward passes (sqgpu_cost_shifted_batched):
                # measured 6.1x at n = 10, 1.9x at n = 8, 0.6x at n = 4 (profiles/r2_shift_engines.jsonl) -- from 7 qubits on
                cost_shifted=eng.cost_shifted_batched if self.accelerator_num == 1 and int(cfg.get("cosine_shift_sweep", self.qbit_num >= 7)) else None)
            self._num_evaluations += ne
            return x, f

        def cost_grad(x):
            f, g = eng.cost_grad_batched(x.reshape(1, -1))
            return float(f[0]), g[0]

        if self._optimizer == "GRAD_DESCEND_PARAMETER_SHIFT_RULE":
            cfg = self.config
            sweep_ok = self._variant in (0, 1, 2, 3, 9) and self.accelerator_num == 1
            x, f, _, ne = optimize.grad_descend_shift_rule(
                eng.cost_batched, rng.random(P) * 2 * np.pi if x0 is None else x0, rng,
                batch_size=min(P, int(cfg.get("batch_size_grad_descend_shift_rule", cfg.get("batch_size", min(64, P))))),
                max_iter=int(cfg.get("max_inner_iterations_grad_descend_shift_rule", cfg.get("max_inner_iterations", 2000))),
                tol=float(cfg.get("optimization_tolerance_grad_descend_shift_rule", tol)),
                eta=float(cfg.get("eta_grad_descend_shift_rule", cfg.get("eta", 1e-3))), use_line_search=bool(int(cfg.get("use_line_search", 1))),
                cost_shifted=eng.cost_shifted_batched if sweep_ok and int(cfg.get("cosine_shift_sweep", self.qbit_num >= 7)) else None)
            self._num_evaluations += ne
            return x, f
        if self._optimizer == "GRAD_DESCEND":  # GRAD_DESCEND.cpp:126-150: Grad_Descend from the guess
            x, f, _, ne = optimize.lbfgs(cost_grad, eng.line_search_batched, rng.random(P) * 2 * np.pi if x0 is None else np.asarray(x0, dtype=np.float64),
                                         max_iter=int(self.config.get("max_inner_iterations_grad_descend", self.config.get("max_inner_iterations", 2000))),
                                         tol=tol * 1e-2, history=0)
            self._num_evaluations += ne
            return x, f
        if x0 is not None:
            x, f, _, ne = optimize.lbfgs(cost_grad, eng.line_search_batched, np.asarray(x0, dtype=np.float64),
                                         max_iter=int(self.config.get("max_inner_iterations_final", self.config.get("max_inner_iterations", 2000))), tol=tol * 1e-2)
            self._num_evaluations += ne
            return x, f
        x, f, _, ne = optimize.multistart_lbfgs(eng.cost_batched, cost_grad, eng.line_search_batched, P, rng,
                                                starts=int(self.config.get("initial_points", 64)), keep=int(self.config.get("restarts", 4)),
                                                max_iter=int(self.config.get("max_inner_iterations", 2000)), tol=tol * 1e-2)
        self._num_evaluations += ne
        return x, f

    # ---- cost configuration ---------------------------------------------------------------------------------
    def set_Cost_Function_Variant(self, costfnc=0):
        self._variant = int(costfnc)
        self._dirty = True

    def set_Trace_Offset(self, trace_offset=0):
        self._trace_offset = int(trace_offset)
        self._dirty = True

    def get_Trace_Offset(self):
        return self._trace_offset

    def set_Previous_Cost_Function_Value(self, value):
        """prev_cost_fnv_val, written by the ADAM engines (optimization_engines/ADAM.cpp:201)."""
        self._prev_cost = float(value)
        self._dirty = True

    def _sync(self):
        # the structure can also change behind our back -- get_Circuit() hands out the live object and set_Gate_Structure
        # shares nested blocks with the caller -- so the plan is keyed on the recursive structure key, not on a flag
        key = self._circuit.structure_key()
        if key != self._circuit_key:
            self._engine.set_circuit(self._circuit)
            self._circuit_key = key
        if self._dirty:
            self._engine.set_cost(self._variant, self._trace_offset, self._prev_cost, self._c1, self._c2)
            self._dirty = False
        return self._engine

    def _apply_engine(self):
        """engine of the matrix-valued calls (apply_to, derivative matrices): they run on one device; a multi-device handle
        only shards the cost path, so those calls go through the circuit's own single-device engine"""
        if self.accelerator_num > 1:
            return self._circuit._get_engine()
        return self._sync()

    # ---- the hot path ---------------------------------------------------------------------------------------
    def Optimization_Problem(self, parameters):
        """Optimization_Interface::optimization_problem (Optimization_Interface.cpp:634-668)."""
        return float(self._sync().cost_batched(np.asarray(parameters, dtype=np.float64).reshape(1, -1))[0])

    def Optimization_Problem_Batch(self, parameters):
        """optimization_problem_batched (Optimization_Interface.cpp:939-1033): rows of a 2-D array."""
        p = np.asarray(parameters, dtype=np.float64)
        if p.ndim != 2:
            raise Exception("Optimization_Problem_Batch: parameters should be a 2 dimensional array")
        return self._sync().cost_batched(p)

    def Optimization_Problem_Combined(self, parameters):
        """optimization_problem_combined (Optimization_Interface.cpp:1145-1490): (f0, grad)."""
        c, g = self._sync().cost_grad_batched(np.asarray(parameters, dtype=np.float64).reshape(1, -1))
        return float(c[0]), g[0]

    def Optimization_Problem_Combined_Batch(self, parameters):
        """Batched (f, grad): the entry the GPU engine adds; row b equals
        Optimization_Problem_Combined(parameters[b])."""
        return self._sync().cost_grad_batched(np.asarray(parameters, dtype=np.float64))

    def Optimization_Problem_Grad(self, parameters):
        """optimization_problem_grad (Optimization_Interface.cpp:1124-1133)."""
        return self.Optimization_Problem_Combined(parameters)[1]

    def Optimization_Problem_Combined_Unitary(self, parameters):
        """optimization_problem_combined_unitary (Optimization_Interface.cpp:1525-1547): (C Umtx, [d_i C Umtx])."""
        eng = self._apply_engine()
        derivs = eng.apply_derivative(parameters, self.Umtx)
        m = self.Umtx.copy()
        eng.apply(parameters, m)
        return m, derivs

    def get_Matrix(self, parameters):
        """Unitary of the gate structure itself (Gates_block::get_matrix)."""
        m = np.eye(1 << self.qbit_num, dtype=np.complex128)
        self._apply_engine().apply(parameters, m)
        return m


def replace_trivial_CRY_gates(circuit, parameters):
    """N_Qubit_Decomposition_adaptive::replace_trivial_CRY_gates (N_Qubit_Decomposition_adaptive.cpp:1398-1590): rewrite the
    adaptive (controlled-RY) gates of an optimised structure with their final parameters:

    * |sin p| > 0.999 and |cos p| < 1e-3 (a half turn):  RX(target, -pi/4), CZ(target, control), RX(target, +pi/4),
      RZ(control, +-pi/4) and the global phase exp(+-i pi/4) that the reference moves onto Umtx (:1443-1497);
    * |sin p| < 1e-3 and |1 - cos p| < 1e-3 (the identity): the gate is dropped (:1503-1513);
    * otherwise the standard CRY = RY(p/2) CNOT RY(-p/2) CNOT (:1515-1547).

    `circuit` is a Circuit whose top-level items are blocks (as the adaptive builder makes them); parameters are the stored
    (half-angle) values in application order. Returns (new_circuit, new_parameters, global_phase) with
    global_phase * new_circuit(new_parameters) == circuit(parameters) as matrices."""
    import cmath

    params = np.asarray(parameters, dtype=np.float64).reshape(-1)
    if params.size != circuit.get_Parameter_Num():
        raise Exception("replace_trivial_CRY_gates: %d parameters for a circuit with %d" % (params.size, circuit.get_Parameter_Num()))
    out = Circuit(circuit.qbit_num, circuit._device)
    new_params = []
    phase = 1.0 + 0.0j
    idx = 0
    for layer in circuit._items:
        if not isinstance(layer, Circuit):
            raise Exception("replace_trivial_adaptive_gates: Only block gates are accepted in this conversion.")
        new_layer = Circuit(circuit.qbit_num, circuit._device)
        for g in layer._flat_gates():
            n_p = g.n_params
            if g.type != abi.ADAPTIVE:
                new_layer._add(g)
                new_params.extend(params[idx:idx + n_p])
                idx += n_p
                continue
            par = float(params[idx])  # activation_function(parameter, 1) is the identity (common/common.cpp:35-38)
            idx += 1
            sp, cp = np.sin(par), np.cos(par)
            t, c = g.target, g.control
            sub = Circuit(circuit.qbit_num, circuit._device)
            if abs(sp) > 0.999 and abs(cp) < 1e-3:
                sub.add_RX(t)
                sub.add_CZ(t, c)
                sub.add_RX(t)
                sub.add_RZ(c)
                sgn = -1.0 if sp < 0 else 1.0
                new_params.extend([-np.pi / 4, np.pi / 4, sgn * np.pi / 4])
                phase *= cmath.exp(1j * sgn * np.pi / 4)
                new_layer.add_Circuit(sub)
            elif abs(sp) < 1e-3 and abs(1 - cp) < 1e-3:
                pass  # trivial gate released
            else:
                sub.add_RY(t)
                sub.add_CNOT(t, c)
                sub.add_RY(t)
                sub.add_CNOT(t, c)
                new_params.extend([par / 2, -par / 2])
                new_layer.add_Circuit(sub)
        out.add_Circuit(new_layer)
    return out, np.asarray(new_params, dtype=np.float64), phase


class N_Qubit_Decomposition_adaptive(N_Qubit_Decomposition_custom):
    """Adaptive gate structure builder + cost path (N_Qubit_Decomposition_adaptive.cpp:1820-1966)."""

    def __init__(self, Umtx, qbit_num=-1, level_limit_max=8, level_limit_min=0, topology=None, config=None,
                 accelerator_num=1, device=0):
        super().__init__(Umtx, qbit_num=qbit_num, config=config, accelerator_num=accelerator_num, device=device)
        self.level_limit = int(level_limit_max)
        self.level_limit_min = int(level_limit_min)
        self.topology = [tuple(int(q) for q in pair) for pair in (topology or [])]
        for pair in self.topology:
            if len(pair) != 2:
                raise Exception("The connectivity data should contains two qubits.")
            if max(pair) >= self.qbit_num:
                raise Exception("Label of control/target qubit should be less than the number of qubits in the register.")

    def add_Adaptive_Layers(self):
        """One level: for every pair a sub-block [U3(target), U3(control), Adaptive(target, control)]
        (N_Qubit_Decomposition_adaptive.cpp:1840-1881), appended to the top-level structure one by one
        (gate_structure->combine(layer), :1806-1811). Deterministic order only (the randomized order needs the
        reference's RNG stream and sits above the hot path)."""
        if self.topology:
            pairs = [(t, c) for (c, t) in self.topology]  # (*it)[0] is the control, [1] the target (:1849-1850)
        else:
            pairs = [(t, c) for t in range(self.qbit_num) for c in range(t + 1, self.qbit_num)]
        for t, c in pairs:
            layer = Circuit(self.qbit_num)
            layer.add_U3(t)
            layer.add_U3(c)
            layer.add_adaptive(t, c)
            self._circuit.add_Circuit(layer)
        self._dirty = True

    def add_Finalyzing_Layer_To_Gate_Structure(self):
        """U3 on every qubit (N_Qubit_Decomposition_adaptive.cpp:1947-1966)."""
        block = Circuit(self.qbit_num)
        for q in range(self.qbit_num):
            block.add_U3(q)
        self._circuit.add_Circuit(block)
        self._dirty = True

    def Start_Decomposition(self):
        """The level search of determine_initial_gate_structure (N_Qubit_Decomposition_adaptive.cpp:786-1037): for
        level = level_limit_min .. level_limit_max build `level` adaptive layers + the finalizing layer, minimise the cost from
        random starts, stop at the first level whose minimum is below the optimization tolerance, keep the best level otherwise.
        Then Compress_Circuit() (:372-520; config["compress"], default 1) and Finalize_Circuit() (the CRY -> CNOT / CZ
        finalisation, :530-640; config["finalize"], default 1), as the reference's start_decomposition does (:255-280).
        Afterwards: get_Circuit(), get_Optimized_Parameters(), get_Decomposition_Error(), get_CNOT_Count()."""
        self.get_Initial_Circuit()
        if int(self.config.get("compress", 1)) and self._current_minimum < self._optimization_tolerance:
            self.Compress_Circuit()
        if int(self.config.get("finalize", 1)):
            self.Finalize_Circuit()
        return self._current_minimum

    def get_Initial_Circuit(self):
        """get_initial_circuit (N_Qubit_Decomposition_adaptive.cpp:290-365): the first of the three phases of the decomposition,
        callable on its own like Compress_Circuit and Finalize_Circuit -- the level search; returns the cost reached"""
        rng = np.random.default_rng(int(self.config.get("seed", 0)))
        best = None
        for level in range(self.level_limit_min, self.level_limit + 1):
            self._circuit = Circuit(self.qbit_num, self._device)
            for _ in range(level):
                self.add_Adaptive_Layers()
            self.add_Finalyzing_Layer_To_Gate_Structure()
            x, f = self._optimize_structure(rng)
            if best is None or f < best[2]:
                best = (self._circuit, x, f, level)
            if f < self._optimization_tolerance:
                break
        self._circuit, self._optimized_parameters, self._current_minimum, self.decomposition_level = best
        self._dirty = True
        return self._current_minimum

    def Compress_Circuit(self):
        """compress_circuit / compress_gate_structure (N_Qubit_Decomposition_adaptive.cpp:372-520, 1046-1335), thin form: as
        long as some decomposing layer can go, the layers are tried for removal in the order of the rotation left in their
        adaptive gate (|sin| of the stored parameter, smallest first, at most five candidates per round, :1067-1083); a candidate
        is accepted when the reduced structure, re-optimised from the reduced parameters (create_reduced_parameters, :1337-1395;
        config max_inner_iterations_compression), still reaches the optimization tolerance. The finalizing U3 layer is never
        removed (:1051). Returns the number of layers removed."""
        from . import optimize

        if self._optimized_parameters is None:
            raise Exception("Compress_Circuit: no optimised parameters (run Start_Decomposition or set_Optimized_Parameters)")
        tol = self._optimization_tolerance
        max_it = int(self.config.get("max_inner_iterations_compression", 200))
        removed = 0
        for _ in range(25):  # :420
            layers = self._circuit._items
            if len(layers) <= 1 or not all(isinstance(it, Circuit) for it in layers):
                break
            x = np.asarray(self._optimized_parameters, dtype=np.float64)
            starts = np.cumsum([0] + [it.get_Parameter_Num() for it in layers])
            theta = []
            for li, it in enumerate(layers[:-1]):
                off, val = starts[li], 15.0  # (the reference's default for layers without an adaptive gate)
                for g in it._flat_gates():
                    if g.type == abi.ADAPTIVE:
                        val = abs(np.sin(x[off]))
                    off += g.n_params
                theta.append(val)
            accepted = False
            for li in np.argsort(theta, kind="stable")[:5]:
                trial = Circuit(self.qbit_num, self._device)
                for lj, it in enumerate(layers):
                    if lj != li:
                        trial.add_Circuit(it)
                x_red = np.concatenate([x[: starts[li]], x[starts[li + 1]:]])
                keep_c, keep_x, keep_f = self._circuit, self._optimized_parameters, self._current_minimum
                self._circuit = trial
                eng = self._sync()

                def cost_grad(v):
                    f, g = eng.cost_grad_batched(v.reshape(1, -1))
                    return float(f[0]), g[0]

                xr, fr, _, ne = optimize.lbfgs(cost_grad, eng.line_search_batched, x_red, max_iter=max_it, tol=tol * 1e-2)
                self._num_evaluations += ne
                if fr < tol:
                    self._optimized_parameters, self._current_minimum = np.asarray(xr, dtype=np.float64), float(fr)
                    removed += 1
                    accepted = True
                    break
                self._circuit, self._optimized_parameters, self._current_minimum = keep_c, keep_x, keep_f
            if not accepted:
                break
        self._sync()
        return removed

    def Finalize_Circuit(self):
        """finalize_circuit (N_Qubit_Decomposition_adaptive.cpp:530-640): the adaptive gates are replaced by CZ / CNOT
        constructions or dropped according to their optimised parameters (replace_trivial_CRY_gates), the global phase moves
        onto Umtx, and the parameters are polished by one more optimisation of the new, CRY-free structure started from the
        converted parameters."""
        if self._optimized_parameters is None:
            raise Exception("Finalize_Circuit: no optimised parameters (run Start_Decomposition or set_Optimized_Parameters)")
        circ, x0, phase = replace_trivial_CRY_gates(self._circuit, self._optimized_parameters)
        # C(x) U = phase C'(x') U = C'(x') (phase U): the reference moves the phase onto Umtx (apply_global_phase_factor,
        # :1478-1495); the Frobenius-family costs take Re tr, so the matrix really has to carry it
        self.Umtx = np.ascontiguousarray(self.Umtx * phase)
        if self._engine_obj is not None:
            self._engine_obj.upload_matrix(self.Umtx)
        self._circuit = circ
        self._dirty = True
        self._optimized_parameters = np.asarray(x0, dtype=np.float64)
        f0 = float(self.Optimization_Problem(self._optimized_parameters))
        x, f = self._optimize_structure(np.random.default_rng(int(self.config.get("seed", 0)) + 1), x0=self._optimized_parameters)
        if not (f <= f0):
            x, f = self._optimized_parameters, f0
        self._optimized_parameters, self._current_minimum = np.asarray(x, dtype=np.float64), float(f)
        return self._current_minimum

    def get_CNOT_Count(self):
        """two-qubit gates (CNOT + CZ) of the current structure, the figure of merit of the decomposition"""
        return sum(1 for g in self._circuit._flat_gates() if g.type in (abi.CNOT, abi.CZ))


class N_Qubit_State_Preparation_adaptive(N_Qubit_Decomposition_adaptive):
    """qgd_N_Qubit_State_Preparation_adaptive (squander/decomposition/qgd_N_Qubit_State_Preparation_adaptive.py:35-62): the adaptive
    decomposition with a 2^n x 1 column as "Umtx" -- the cost 1 - Re (C state)[0] is minimal when the circuit maps the state
    onto |0...0>, so the inverse of the circuit found prepares the state. Same checks as the reference's constructor."""

    def __init__(self, State, level_limit_max=8, level_limit_min=0, topology=None, config=None, accelerator_num=1, device=0):
        if not isinstance(State, np.ndarray):
            raise Exception("Initial state should be a numpy array")
        if State.dtype != np.complex128:
            raise Exception("Initial state should be made of complex values")
        if not State.data.c_contiguous:
            raise Exception("Initial state should be contiguous in memory")
        if State.ndim == 1:
            State = State.reshape((State.size, 1))
        if not (State.ndim == 2 and State.shape[1] == 1):
            raise Exception("Initial state not properly formatted. Input state must be a column vector")
        super().__init__(State, level_limit_max=level_limit_max, level_limit_min=level_limit_min, topology=topology, config=config,
                         accelerator_num=accelerator_num, device=device)
