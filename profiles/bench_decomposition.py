"""Config 1 (data/Umtx.mat, 4 qubits; the matrix is stored in the golden fixture) through every optimisation engine of this
package at a fixed structure (3 adaptive levels + finalizing layer, P = 138): final cost, cost evaluations, wall time.
    python profiles/bench_decomposition.py   -> one JSON line per engine, then the full Start_Decomposition (level search,
                                                compression, finalisation) with the default engine"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import golden_cases as G
import squander_b200 as sq

Uct = G.load("C1_L3").U
ENGINES = [("BFGS", {}), ("ADAM", {}), ("COSINE", {"max_inner_iterations": 1500}), ("AGENTS", {"max_inner_iterations": 3000, "agent_lifetime": 100}),
           ("AGENTS_COMBINED", {"max_inner_iterations_agent": 1000, "agent_lifetime": 100, "max_inner_iterations_grad_descend": 500}),
           ("GRAD_DESCEND", {"max_inner_iterations": 1500}), ("GRAD_DESCEND_PARAMETER_SHIFT_RULE", {"max_inner_iterations": 1500, "eta": 1.0})]
for name, cfg in ENGINES:
    dec = sq.N_Qubit_Decomposition_adaptive(Uct, level_limit_max=3, level_limit_min=3, config=dict(cfg, optimization_tolerance=1e-6, compress=0, finalize=0))
    dec.set_Optimizer(name)
    t0 = time.perf_counter()
    err = dec.Start_Decomposition()
    dt = time.perf_counter() - t0
    print(json.dumps({"engine": name, "config": cfg, "levels": 3, "P": dec.get_Parameter_Num(), "final_cost": err, "cost_evaluations": dec.get_Num_of_Iters(),
                      "wall_s": round(dt, 3)}), flush=True)
dec = sq.N_Qubit_Decomposition_adaptive(Uct, level_limit_max=5, level_limit_min=1, config={"optimization_tolerance": 1e-6})
t0 = time.perf_counter()
err = dec.Start_Decomposition()
print(json.dumps({"engine": "BFGS, full Start_Decomposition (level search + Compress_Circuit + Finalize_Circuit)", "final_cost": err, "level": dec.decomposition_level,
                  "two_qubit_gates": dec.get_CNOT_Count(), "gates": dec.get_Gate_Num(), "wall_s": round(time.perf_counter() - t0, 3)}), flush=True)
