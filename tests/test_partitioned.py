"""Node-partitioned engine (bqa_b200/partitioned.py): partitioner, local contexts / exchange plan, and the
world_size-2 and -3 runs over gloo on CPU (kernels = the test-only host emulation) against the single-process
engine on the same instance: identical bond dimensions and sweep counts, Bloch vectors and bitstrings equal."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import instances
from bqa_b200 import _lib
from bqa_b200.config import config_to_context
from bqa_b200.engine import Engine
from bqa_b200.partitioned import PartitionedEngine, build_local_context, cut_fraction, partition_nodes
from hostemu.build import build as build_hostemu


def _rr(n, seed=42):
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    return generate_qubo_on_random_regular_graph(n, 3, seed=seed)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_partitioner_is_balanced_deterministic_and_beats_random(P):
    nodes, edges = _rr(4000)
    E = np.array(list(edges.keys()))
    part = partition_nodes(4000, E, P, seed=0)
    assert part.min() == 0 and part.max() == P - 1
    sizes = np.bincount(part, minlength=P)
    assert sizes.max() - sizes.min() <= 1
    assert np.array_equal(part, partition_nodes(4000, E, P, seed=0))
    rnd = np.random.default_rng(0).integers(P, size=4000)
    assert cut_fraction(part, E) < 0.75 * cut_fraction(rnd, E)


def test_partitioner_handles_tiny_and_disconnected_graphs():
    E = np.array([(0, 1), (2, 3), (4, 5), (5, 6)])
    part = partition_nodes(8, E, 3, seed=1)                  # node 7 is isolated
    assert sorted(np.bincount(part, minlength=3).tolist()) == [2, 3, 3]


def test_local_contexts_cover_the_graph_and_plans_match():
    cfg = instances.cfg_grid4()
    ctx = config_to_context(cfg)
    P = 3
    part = partition_nodes(ctx.nodes_number, ctx.edges, P, seed=0)
    L = ctx.edges_number // 2
    locs = [build_local_context(ctx, part, r, P) for r in range(P)]
    assert sum(lc.nodes_number for lc, _ in locs) == ctx.nodes_number
    for r, (lc, plan) in enumerate(locs):
        Lr = lc.edges_number // 2
        for d, lay in lc.degree_to_layout.items():
            # reference slot convention holds locally: lambda slot = message slot mod L_r, in/out slots are mirror images
            assert np.array_equal(lay.lmbds_position, lay.input_msgs_position % Lr)
            assert np.array_equal((lay.input_msgs_position + Lr) % (2 * Lr), lay.output_msgs_position)
        for q, s_slots in plan.send_slots.items():
            # what r sends to q is, in the same order, what q expects from r (compare through global positions)
            r_edges = lc.edges[s_slots % Lr]
            q_lc, q_plan = locs[q]
            q_edges = q_lc.edges[q_plan.recv_slots[r] % (q_lc.edges_number // 2)]
            assert np.array_equal(r_edges, q_edges)
            assert np.array_equal(s_slots // Lr, q_plan.recv_slots[r] // (q_lc.edges_number // 2))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cfg, emu_path, precision, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import logging
        logging.disable(logging.WARNING)
        ctx = config_to_context(cfg)
        eng = PartitionedEngine(ctx, precision=precision, _testing_lib=_lib.bind(emu_path))
        res = {}
        for ins in ctx.instructions:
            if isinstance(ins, dict):
                eng.run_layer(ins["xtime"], ins["ztime"])
            elif ins == "get_bloch_vectors":
                res["bloch"] = eng.bloch_vectors()
            elif ins == "measure":
                res["outcomes"] = eng.measure()
        res["bond_dims"] = eng.stats["bond_dims"]
        res["bp_sweeps"] = eng.stats["bp_sweeps"]
        res["comm_bytes"] = eng.comm_bytes
        out[rank] = res
    finally:
        dist.destroy_process_group()


def _run_partitioned(cfg, world, precision="double"):
    emu_path = build_hostemu()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), cfg, emu_path, precision, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


def _run_single(cfg, precision="double"):
    emu = _lib.bind(build_hostemu())
    ctx = config_to_context(cfg)
    eng = Engine(ctx, precision=precision, _testing_lib=emu)
    res = {}
    for ins in ctx.instructions:
        if isinstance(ins, dict):
            eng.run_layer(ins["xtime"], ins["ztime"])
        elif ins == "get_bloch_vectors":
            res["bloch"] = eng.bloch_vectors()
        elif ins == "measure":
            res["outcomes"] = eng.measure()
    res["bond_dims"] = eng.stats["bond_dims"]
    res["bp_sweeps"] = eng.stats["bp_sweeps"]
    return res


@pytest.mark.parametrize("name,world", [("ring24", 2), ("grid4", 3), ("comb", 2)])
def test_partitioned_run_reproduces_single_process_run(name, world):
    cfg = instances.GOLDEN_CONFIGS[name]()
    want = _run_single(cfg)
    got = _run_partitioned(cfg, world)
    for r in range(world):
        assert got[r]["bond_dims"] == want["bond_dims"]
        assert got[r]["bp_sweeps"] == want["bp_sweeps"]
        assert np.abs(got[r]["bloch"] - want["bloch"]).max() < 1e-12
        if "outcomes" in want:
            assert got[r]["outcomes"] == want["outcomes"]
    assert any(g["comm_bytes"] > 0 for g in got)


def test_partitioned_run_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ring24.npz"))
    got = _run_partitioned(instances.cfg_ring24(), 2)
    assert np.abs(got[0]["bloch"] - g["bloch"]).max() < 1e-8
    assert got[0]["outcomes"] == g["outcomes"].tolist()


def _worker_ckpt(rank, world, port, cfg, emu_path, path, crash_after, out):
    """run_context on the partitioned engine with checkpoints; `crash_after` layers every rank raises."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import logging
        logging.disable(logging.WARNING)
        from bqa_b200.core import run_context

        class Flaky(PartitionedEngine):
            def run_layer(self, xtime, ztime, next_ztime=None):
                if crash_after is not None and len(self.stats["bond_dims"]) == crash_after:
                    raise KeyboardInterrupt("power cut")
                return super().run_layer(xtime, ztime, next_ztime=next_ztime)
        try:
            res = run_context(config_to_context(cfg), precision="double", engine_cls=Flaky, checkpoint=path,
                              checkpoint_every=4, resume=True, _testing_lib=_lib.bind(emu_path))
            out[rank] = res
        except KeyboardInterrupt:
            out[rank] = "crashed"
    finally:
        dist.destroy_process_group()


def test_partitioned_checkpoint_resume(tmp_path):
    """Per-rank checkpoint files; the resumed 2-rank run returns what the uninterrupted single-process run returns."""
    from bqa_b200.core import run_context
    cfg = instances.cfg_grid4()
    emu_path = build_hostemu()
    want = run_context(config_to_context(cfg), precision="double", _testing_lib=_lib.bind(emu_path))
    path = str(tmp_path / "anneal.npz")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_ckpt, args=(2, _free_port(), cfg, emu_path, path, 10, out), nprocs=2, join=True)
    assert out[0] == "crashed" and out[1] == "crashed"
    assert sorted(os.listdir(tmp_path)) == ["anneal.rank0of2.npz", "anneal.rank1of2.npz"]
    out2 = mgr.dict()
    mp.spawn(_worker_ckpt, args=(2, _free_port(), cfg, emu_path, path, None, out2), nprocs=2, join=True)
    for r in range(2):
        got = dict(out2[r])
        assert np.abs(np.array(got["bloch_vectors"]) - np.array(dict(want)["bloch_vectors"])).max() < 1e-12
        assert got["measurement_outcomes"] == dict(want)["measurement_outcomes"]
