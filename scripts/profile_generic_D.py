#!/usr/bin/env python
"""Where the shape-generic kernels stand at bond dimension 8 and 16 (BASELINE configs[0] / [2] live there): a MaxCut-shaped
3-regular instance (zero fields, unit couplings, damping 0.5, eps 1e-5 like benchmarks_against_mqlib/
random_3_regular_maxcut_1000.py:10-44) large enough to fill the GPU is annealed until the bond dimension reaches
--dmax, then --steps steps are timed per entry point with CUDA events.  One JSON line.  Run it under ncu for the
pipe utilisation of the two node kernels and the canonicalizer (profiles/r2_generic_D*.md)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=20_000)
    ap.add_argument("--dmax", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--precision", default="single")
    args = ap.parse_args()
    import logging
    logging.disable(logging.WARNING)
    from bqa_b200 import _lib
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    nodes, edges = generate_qubo_on_random_regular_graph(args.qubits, 3, seed=42, node_ampl_func=lambda *_: 0.0,
                                                         edge_ampl_func=lambda *_: 1.0)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": args.dmax, "measurement_threshold": 0.99, "damping": 0.5,
           "bp_eps": 1e-5, "pinv_eps": 1e-5, "max_bp_iter_number": 250,
           "schedule": {"total_time": 200.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 1000, "final_mixing": 0.0}]}}
    ctx = config_to_context(cfg)
    lib = _lib.load_library()
    eng = Engine(ctx, precision=args.precision)
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    k = 0
    while eng.D < args.dmax and k < 400:
        eng.run_layer(layers[k]["xtime"], layers[k]["ztime"])
        k += 1
    names = ["ext_msgs", "ext_msgs_classes", "canonicalize", "apply_update", "apply_update_classes", "bp_sweep", "bp_run",
             "bp_run_classes", "gauge_msgs"]
    events = {n: [] for n in names}
    orig = {n: getattr(lib, n) for n in names}

    def wrap(n):
        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig[n](*a)
            e1.record()
            events[n].append((e0, e1))
            return r
        return call
    for n in names:
        setattr(lib, n, wrap(n))
    n0 = len(eng.stats["bp_sweeps"])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()       # ncu --profile-from-start off: only the steps at the final D are captured
    for ins in layers[k:k + args.steps]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    for n in names:
        setattr(lib, n, orig[n])
    out = {"qubits": args.qubits, "D": eng.D, "precision": args.precision, "steps_to_reach_D": k, "timed_steps": args.steps,
           "bp_sweeps_per_step": float(np.mean(eng.stats["bp_sweeps"][n0:])),
           "ms_per_step": {n: sum(a.elapsed_time(b) for a, b in events[n]) / args.steps for n in names if events[n]}}
    out["ms_per_step_total"] = sum(out["ms_per_step"].values())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
