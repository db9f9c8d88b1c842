#!/usr/bin/env python
"""BASELINE.json configs[1] and [2] end to end on one GPU, with the oracle (numpy restatement of the reference)
timed beside them on the host.  These graphs (400 / 127 qubits) are launch-latency bound: the numbers document
where the generic kernels stand, the 100k-qubit bench is the throughput metric.  One JSON line per config."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import logging  # noqa: E402

logging.disable(logging.WARNING)
import torch  # noqa: E402
from bqa_b200 import run_qa  # noqa: E402
from bqa_b200.benchmarking import generate_qubo_on_2d_grid, heavy_hex_127  # noqa: E402
from oracle import bqa_oracle as O  # noqa: E402


def anneal(total_time, steps, tail):
    return {"total_time": total_time, "starting_mixing": 1.0,
            "actions": [{"weight": 1.0, "steps_number": steps, "final_mixing": 0.0}, *tail]}


def main():
    with_oracle = "--no-oracle" not in sys.argv
    nodes, edges = generate_qubo_on_2d_grid(20, 20, seed=42)
    grid = {"nodes": nodes, "edges": edges, "max_bond_dim": 4, "schedule": anneal(10.0, 100, ["get_bloch_vectors"])}
    nodes, edges = heavy_hex_127(seed=42)
    hexa = {"nodes": nodes, "edges": edges, "max_bond_dim": 8, "schedule": anneal(10.0, 10, ["get_bloch_vectors", "measure"])}
    for name, cfg in (("configs[1] 2D grid 20x20, 100 steps, D<=4 (examples/2d_greed.py, Bloch vectors)", grid),
                      ("configs[2] heavy-hex 127, 10 steps, D<=8, measure (examples/full_size_ibm_heavy_hex.py)", hexa)):
        out = {"config": name}
        for prec in ("double", "single"):
            run_qa(cfg, precision=prec)                                   # warm-up (module load, allocator)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = dict(run_qa(cfg, precision=prec))
            torch.cuda.synchronize()
            out[f"gpu_{prec}_s"] = time.perf_counter() - t0
            out[f"res_{prec}"] = res
            # the annealing steps alone: launches and blocking host reads per step
            from bqa_b200 import _lib
            from bqa_b200.config import config_to_context
            from bqa_b200.engine import Engine
            lib = _lib.load_library()
            ctx = config_to_context(cfg)
            layers = [i for i in ctx.instructions if isinstance(i, dict)]
            for multi in (True, False):
                eng = Engine(ctx, precision=prec)
                eng._multiclass = multi
                l0 = lib.launch_count()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for ins in layers:
                    eng.run_layer(ins["xtime"], ins["ztime"])
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                tag = f"{prec}_{'one_launch_for_all_classes' if multi else 'per_class_launches'}"
                out[f"steps_per_s_{tag}"] = len(layers) / dt
                out[f"launches_per_step_{tag}"] = (lib.launch_count() - l0) / len(layers)
                out[f"host_reads_per_step_{tag}"] = eng.host_reads / len(layers)
                t0 = time.perf_counter()
                if "measure" in cfg["schedule"]["actions"]:
                    eng.measure()
                    out[f"measure_s_{tag}"] = time.perf_counter() - t0
        if with_oracle:
            t0 = time.perf_counter()
            want = dict(O.run_qa(cfg))
            out["oracle_cpu_s"] = time.perf_counter() - t0
            out["cpu_cores"] = os.cpu_count()
            for prec in ("double", "single"):
                d = np.abs(np.array(out[f"res_{prec}"]["bloch_vectors"]) - np.array(want["bloch_vectors"]))
                out[f"bloch_max_abs_diff_{prec}"] = float(d.max())
                if "measurement_outcomes" in want:
                    out[f"bitstring_equal_{prec}"] = out[f"res_{prec}"]["measurement_outcomes"] == want["measurement_outcomes"]
        for prec in ("double", "single"):
            del out[f"res_{prec}"]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
