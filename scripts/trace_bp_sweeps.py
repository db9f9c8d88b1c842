"""Where a BP sweep's time goes inside the single-launch run, per rank (launch under torchrun for N > 1):
one GPU: sweep | grid barrier | residual test; N GPUs: boundary groups | interior groups + grid barrier | residual line
sent | wait for the peers' halo-ready lines | lagged residual test.  Uses bqa_b200_set_bp_trace
(%globaltimer stamps of CTA 0).  Prints one JSON line per rank."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import logging
    logging.disable(logging.WARNING)
    import instances
    from bqa_b200 import _lib
    from bqa_b200.config import config_to_context
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    ctx = config_to_context(instances.bench_config(n))
    if world > 1:
        from bqa_b200.partitioned import PartitionedEngine
        eng = PartitionedEngine(ctx, precision="single", device=dev)
    else:
        from bqa_b200.engine import Engine
        eng = Engine(ctx, precision="single", device=dev)
    lib = _lib.load_library()
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    for ins in layers[:40]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    trace = torch.zeros(5 * eng.max_iters, dtype=torch.int64, device=dev)
    lib.set_bp_trace(trace.data_ptr())
    rows = []
    for ins in layers[40:50]:
        trace.zero_()
        eng.run_layer(ins["xtime"], ins["ztime"])
        torch.cuda.synchronize()
        t = trace.cpu().numpy().reshape(-1, 5)[: eng.stats["bp_sweeps"][-1]].astype(np.float64)
        rows.append(np.stack([t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3],
                              np.append(t[1:, 0] - t[:-1, 4], np.nan)], 1))
    lib.set_bp_trace(None)
    r = np.concatenate(rows)
    us = np.nanmean(r, 0) / 1e3
    print(json.dumps({"rank": rank, "world": world, "qubits": n, "owned_nodes": int(eng.N), "sweeps": int(r.shape[0]),
                      "us_per_sweep": ({"sweep_cta0": us[0], "grid_barrier": us[1], "residual_test_to_next_sweep": us[4]}
                                       if world == 1 else
                                       {"boundary_groups_cta0": us[0], "interior_groups_and_grid_barrier": us[1],
                                        "resid_line_sent": us[2], "wait_for_peers_halo_lines": us[3],
                                        "lagged_residual_test_to_next_sweep": us[4]}) | {"total": float(np.nansum(us))}}),
          flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
