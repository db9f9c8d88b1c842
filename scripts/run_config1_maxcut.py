#!/usr/bin/env python
"""BASELINE.json configs[0]: random 3-regular MaxCut, 1000 qubits, max_bond_dim 16 (shape of
benchmarks_against_mqlib/random_3_regular_maxcut_1000.py:10-44: zero fields, unit couplings, damping 0.5,
eps 1e-5, 250 BP iterations, total_time 200 / 1000 steps, then "measure"), on one GPU with the generic kernels.
The reference needs 6-9 h on 8 host cores for the anneal alone.  Usage:
    python scripts/run_config1_maxcut.py [STEPS=1000] [--measure]      one JSON line on stdout"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import logging  # noqa: E402

logging.disable(logging.WARNING)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph, ising_energy  # noqa: E402
from bqa_b200.config import config_to_context  # noqa: E402
from bqa_b200.engine import Engine  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000
    nodes, edges = generate_qubo_on_random_regular_graph(1000, 3, seed=42, node_ampl_func=lambda *_: 0.0,
                                                         edge_ampl_func=lambda *_: 1.0)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 16, "measurement_threshold": 0.99, "damping": 0.5,
           "bp_eps": 1e-5, "pinv_eps": 1e-5, "max_bp_iter_number": 250,
           "schedule": {"total_time": 200.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 1000, "final_mixing": 0.0}]}}
    ctx = config_to_context(cfg)
    eng = Engine(ctx, precision="single")
    if os.environ.get("BQA_B200_KERNEL_MODE"):           # 1: generic kernels only, 2: round-1 n = 8 canonicalizer
        eng.lib.set_kernel_mode(int(os.environ["BQA_B200_KERNEL_MODE"]))
    layers = [i for i in ctx.instructions if isinstance(i, dict)][:steps]
    t0 = time.perf_counter()
    marks = {}
    for k, ins in enumerate(layers):
        eng.run_layer(ins["xtime"], ins["ztime"])
        if eng.D not in marks:
            marks[eng.D] = k
    torch.cuda.synchronize()
    t_anneal = time.perf_counter() - t0
    out = {"config": "configs[0] random 3-regular MaxCut, 1000 qubits, max_bond_dim 16, complex64, generic kernels",
           "steps": len(layers), "anneal_s": t_anneal, "first_step_with_bond_dim": marks,
           "bp_sweeps_last10": eng.stats["bp_sweeps"][-10:], "bp_sweeps_total": int(sum(eng.stats["bp_sweeps"])),
           "bp_sweeps_by_50_steps": [int(sum(eng.stats["bp_sweeps"][i:i + 50])) for i in range(0, len(layers), 50)],
           "final_bond_dim": eng.D,
           "switches": {k: os.environ.get(k, "default") for k in ("BQA_B200_FAST_GRAM", "BQA_B200_ROUND_ROBIN",
                                                                    "BQA_B200_KERNEL_MODE")}}
    if "--measure" in sys.argv:
        t0 = time.perf_counter()
        s = eng.measure()
        out["measure_s"] = time.perf_counter() - t0
        out["cut_energy"] = ising_energy(edges, nodes, s)
    else:
        b = eng.bloch_vectors()
        out["energy_of_sign_z"] = ising_energy(edges, nodes, np.where(b[:, 2] > 0, 1, -1))
        out["bloch_abs_mean_xyz"] = [float(v) for v in np.abs(b).mean(0)]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
