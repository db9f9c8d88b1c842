#!/usr/bin/env python
"""BASELINE.json configs[4]: random 3-regular QUBO, 500,000 qubits, sharded across the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_config5_500k.py [N] [STEPS]

Runs the first STEPS steps of the dt = 0.2 schedule (benchmarks_against_mqlib/random_3_regular_qubo_500000.py shape,
max_bond_dim 4, complex64) on the node-partitioned engine, reports steps/s of the last 10 steps, and checks the
per-qubit marginals against a single-GPU run of the same instance on rank 0 (the 500k state is 2.3 GB: it fits one
B200; the CPU reference needs ~4 min per step at this size).  One JSON line on stdout."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import logging  # noqa: E402

logging.disable(logging.WARNING)
from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph, ising_energy  # noqa: E402
from bqa_b200.config import config_to_context  # noqa: E402
from bqa_b200.engine import Engine  # noqa: E402
from bqa_b200.partitioned import PartitionedEngine, cut_fraction  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    t0 = time.perf_counter()
    nodes, edges = generate_qubo_on_random_regular_graph(n, 3, seed=42)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 4, "bp_eps": 1e-6, "pinv_eps": 1e-6,
           "schedule": {"total_time": 0.2 * 100, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 100, "final_mixing": 0.0}]}}
    ctx = config_to_context(cfg)
    layers = [i for i in ctx.instructions if isinstance(i, dict)][:steps]
    t_setup = time.perf_counter() - t0
    eng = PartitionedEngine(ctx, precision="single", device=dev)
    timed = min(10, steps)
    for ins in layers[:-timed]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for ins in layers[-timed:]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t1
    bloch = eng.bloch_vectors()
    out = None
    if rank == 0:
        se = Engine(ctx, precision="single", device=dev)
        for ins in layers:
            se.run_layer(ins["xtime"], ins["ztime"])
        ref = se.bloch_vectors()
        diff = np.abs(bloch - ref)
        s_p, s_1 = np.where(bloch[:, 2] > 0, 1, -1), np.where(ref[:, 2] > 0, 1, -1)
        e_p, e_1 = ising_energy(edges, nodes, s_p), ising_energy(edges, nodes, s_1)
        out = {"config": f"random 3-regular QUBO, {n} qubits, {world} GPUs, complex64, D = {eng.D}", "steps": steps,
               "steps_per_s_last10": timed / dt, "sweeps_per_step_last10": float(np.mean(eng.stats["bp_sweeps"][-timed:])),
               "cut_fraction": cut_fraction(eng.part, np.asarray(ctx.edges)), "setup_s": t_setup,
               "bloch_max_abs_diff_vs_1gpu": float(diff.max()), "bloch_mean_abs_diff_vs_1gpu": float(diff.mean()),
               "sign_flips_vs_1gpu": int((s_p != s_1).sum()), "energy_partitioned": e_p, "energy_1gpu": e_1,
               "bond_dims_equal": eng.stats["bond_dims"] == se.stats["bond_dims"],
               "ok": bool(diff.max() < 5e-3 and diff.mean() < 1e-4 and abs(e_p - e_1) <= 1e-4 * abs(e_1))}
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not out["ok"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
