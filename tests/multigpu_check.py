"""Launched under torchrun on N GPUs (tests/test_gpu_parity.py::test_partitioned_nccl, or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_check.py
The node-partitioned engine over NCCL must reproduce the single-GPU engine on the same instance: complex128 to
rounding (same bond dimensions and sweep counts), complex64 within the stated fp32 tolerance."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import logging  # noqa: E402

logging.disable(logging.WARNING)
from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph  # noqa: E402
from bqa_b200.config import config_to_context  # noqa: E402
from bqa_b200.engine import Engine  # noqa: E402
from bqa_b200.partitioned import PartitionedEngine  # noqa: E402


def run(eng, ctx):
    for ins in ctx.instructions:
        if isinstance(ins, dict):
            eng.run_layer(ins["xtime"], ins["ztime"])
    return eng.bloch_vectors()


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    nodes, edges = generate_qubo_on_random_regular_graph(n, 3, seed=42)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 4,
           "schedule": {"total_time": 6.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 30, "final_mixing": 0.0}]}}
    ctx = config_to_context(cfg)
    ok = True
    for precision, tol in (("double", 1e-9), ("single", 1e-6)):   # designed to be bit-identical; the tolerances are slack
        pe = PartitionedEngine(ctx, precision=precision, device=dev)
        got = run(pe, ctx)
        if rank == 0:
            se = Engine(ctx, precision=precision, device=dev)
            want = run(se, ctx)
            err = float(np.abs(got - want).max())
            same = pe.stats["bond_dims"] == se.stats["bond_dims"]
            sweeps = np.abs(np.array(pe.stats["bp_sweeps"]) - np.array(se.stats["bp_sweeps"])).max()
            print(f"{precision}: world {dist.get_world_size()} max |bloch - single GPU| = {err:.3e}, bond dims equal: {same}, "
                  f"max sweep-count difference {sweeps}, boundary bytes sent by rank 0: {pe.comm_bytes}", flush=True)
            ok = ok and err < tol and same and sweeps == 0
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if not int(flag.item()):
        raise SystemExit(1)
    if rank == 0:
        print("multigpu_check ok", flush=True)


if __name__ == "__main__":
    main()
