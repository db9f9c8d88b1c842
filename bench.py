#!/usr/bin/env python
"""bench.py -- annealing steps/s and BP msg-updates/s on the 100k-qubit random 3-regular QUBO (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]              our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]        reference arm (the reference's numpy backend from
                                                                     baseline/_ref -- the oracle port if that install is
                                                                     absent -- on the host cores, at the SAME 100k config)

A "step" is one annealing step = one ``run_layer`` (reference src/bqa/state.py:315-321): simple update
(extended messages, canonicalizers, truncation, Rz/Rx, symmetric gauge) followed by BP to convergence.

Workload (BASELINE.md section 2 protocol): ``generate_qubo_on_random_regular_graph(100_000, 3, seed=42)``,
max_bond_dim 4, bp_eps = pinv_eps = 1e-6, damping 0, max_bp_iter_number 75, dt = 0.2 like
benchmarks_against_mqlib/random_3_regular_qubo_100000.py, schedule ``total_time = 0.2 S, steps_number = S``,
mixing 1 -> 0; the first RAMP steps (bond dimension 1 -> 4) are executed untimed, then W warm-up steps, then
exactly K timed steps of the same schedule.  complex64 (what the reference's GPU backend uses).

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_QUBITS = 100_000
SCHEDULE_STEPS = 100          # S
RAMP = 30                     # untimed steps that take the bond dimension from 1 to 4
DT = 0.2
CPU_REF_STEPS = 3             # steps the reference arm times at the full 100k config (about 25-45 s each)
# stated fp32 tolerances (BASELINE.md section 4 / DESIGN.md section 4) asserted on the parity block
TOL_BLOCH_MAX, TOL_BLOCH_MEAN, TOL_ENERGY_REL, TOL_SWEEPS = 5e-3, 1e-4, 1e-4, 1
METRIC = "annealing steps/s @100k-qubit random 3-regular QUBO (D=4, complex64)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_config(n: int, total_steps: int) -> dict:
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    nodes, edges = generate_qubo_on_random_regular_graph(n, 3, seed=42)
    return {"nodes": nodes, "edges": edges, "max_bond_dim": 4, "bp_eps": 1e-6, "pinv_eps": 1e-6,
            "damping": 0.0, "max_bp_iter_number": 75, "seed": 42, "default_field": 0.0,
            "measurement_threshold": 0.95, "backend": "b200",
            "schedule": {"total_time": DT * total_steps, "starting_mixing": 1.0,
                         "actions": [{"type": "real_time_evolution", "weight": 1.0, "steps_number": total_steps,
                                      "final_mixing": 0.0}]}}


def schedule_len(steps: int, warmup: int) -> int:
    return max(SCHEDULE_STEPS, RAMP + warmup + 2 * steps + 2)


# -------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md "clocks DURING the timed region")
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w": float(np.median(power)) if power else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------------------
# CPU legs: the reference's own numpy backend (baseline/_ref, unmodified) or, without that install, the oracle
# port, on the host cores, at the SAME 100k-qubit config -- no scaling from a smaller instance
# -------------------------------------------------------------------------------------------------
class CpuRunner:
    """Uniform driver of the two CPU implementations of the path.  kind = "reference": LuchnikovI/bqa installed
    unmodified into baseline/_ref by baseline/install_ref.py (numpy backend, complex128 default, src/bqa/state.py);
    kind = "port": oracle/bqa_oracle.py (numpy restatement pinned to the reference's outputs, tests/test_oracle.py)."""

    def __init__(self):
        os.environ["BQA_PRECISION"] = "double"               # the reference's default (src/bqa/utils.py:9-20)
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        self.kind = "port"
        try:
            import install_ref
            if install_ref.add_to_path():
                import bqa                                    # noqa: F401
                from bqa import state as rstate
                from bqa.backends import NumPyBackend
                from bqa.config.core import config_to_context as ref_compile
                self.rstate, self.NB, self.ref_compile = rstate, NumPyBackend, ref_compile
                self.kind = "reference"
        except Exception as e:                                # a broken install must not take the bench down
            log(f"baseline/_ref is not usable ({type(e).__name__}: {e}); the CPU legs use the oracle port")
        if self.kind == "port":
            from oracle import bqa_oracle as O
            self.O = O
        self._sweeps = 0

    def compile(self, cfg: dict):
        if self.kind == "reference":
            return self.ref_compile({**cfg, "backend": "numpy"})
        return self.O.compile_config(cfg)

    def layers(self, ctx) -> list:
        return [i for i in ctx.instructions if isinstance(i, dict)]

    def node_ids(self, ctx) -> dict:
        if self.kind == "reference":
            return {int(d): np.asarray(l.node_ids.numpy) for d, l in ctx.degree_to_layout.items()}
        return {int(d): np.asarray(l.node_ids) for d, l in ctx.layouts.items()}

    def init_state(self, ctx):
        return self.rstate._initialize_state(ctx) if self.kind == "reference" else self.O.init_state(ctx)

    def load(self, ctx, st, snap: dict) -> None:
        """D = 4 state handed over from the GPU run: tensors {degree: (B, 2, D..)}, msgs (2L, D, D), lmbds (L, D)."""
        c = np.complex128
        if self.kind == "reference":
            for d in list(st.degree_to_tensor):
                st.degree_to_tensor[d] = self.NB(np.ascontiguousarray(snap["tensors"][int(d)], dtype=c))
            st.msgs = self.NB(np.ascontiguousarray(snap["msgs"], dtype=c))
            st.lmbds = self.NB(np.ascontiguousarray(snap["lmbds"], dtype=c))
        else:
            st.tensors = {d: np.asarray(t, c) for d, t in snap["tensors"].items()}
            st.msgs = np.asarray(snap["msgs"], c)
            st.lmbds = np.asarray(snap["lmbds"], c)

    def bond_dim(self, st) -> int:
        return int(st.bond_dim)

    def run_layer(self, ctx, st, xtime: float, ztime: float) -> int:
        """one annealing step (src/bqa/state.py:315-321); returns the number of BP sweeps it ran"""
        if self.kind == "port":
            n0 = len(st.stats["bp_sweeps"])
            self.O.run_layer(ctx, st, xtime, ztime)
            return int(sum(st.stats["bp_sweeps"][n0:]))
        NB, counter = self.NB, {"n": 0}
        orig = NB.get_dist

        def counting(this, other):                            # one get_dist per sweep (state.py:113)
            counter["n"] += 1
            return orig(this, other)
        NB.get_dist = counting
        try:
            self.rstate.run_layer(ctx, xtime, ztime, st)
        finally:
            NB.get_dist = orig
        return counter["n"]

    def bloch(self, ctx, st) -> np.ndarray:
        rho = self.rstate.get_density_matrices(ctx, st) if self.kind == "reference" else self.O.density_matrices(ctx, st)
        return np.stack([(rho[:, 0, 1] + rho[:, 1, 0]).real, (rho[:, 1, 0] - rho[:, 0, 1]).imag,
                         (rho[:, 0, 0] - rho[:, 1, 1]).real], axis=1)          # src/bqa/utils.py:23-27

    def lmbds(self, st) -> np.ndarray:
        lm = st.lmbds.numpy if self.kind == "reference" else st.lmbds
        return np.real(np.asarray(lm))


def gpu_state_after_ramp(cfg: dict, n_steps: int, dev=None) -> dict:
    """The GPU engine (complex64, the arm being benchmarked) takes the schedule's first n_steps steps; its state is
    what the CPU legs continue from -- the CPU would need about 15 minutes for this ramp at 100k qubits."""
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    ctx = config_to_context(cfg)
    eng = Engine(ctx, precision="single", device=dev)
    for ins in [i for i in ctx.instructions if isinstance(i, dict)][:n_steps]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    snap = eng.state_to_host()
    snap["node_ids"] = {c.degree: c.node_ids_host for c in eng.classes}
    return snap


def cpu_run(runner: CpuRunner, cfg: dict, snap: dict | None, first: int, steps: int) -> dict:
    """`steps` annealing steps of the CPU implementation at the full config, starting at schedule position `first`
    from `snap` (None: ramped on the CPU from the initial state).  Wall clock per step, sweeps, final observables."""
    ctx = runner.compile(cfg)
    layers = runner.layers(ctx)
    st = runner.init_state(ctx)
    t_ramp = time.perf_counter()
    if snap is not None:
        ids = runner.node_ids(ctx)
        for d, mine in snap["node_ids"].items():              # both compile steps order the degree classes the same way
            assert np.array_equal(ids[int(d)], mine), "the CPU and GPU compile steps order the nodes differently"
        runner.load(ctx, st, snap)
        how = f"D=4 state after {first} steps handed over from the GPU run"
    else:
        for ins in layers[:first]:
            runner.run_layer(ctx, st, ins["xtime"], ins["ztime"])
        how = f"ramped on the CPU ({first} untimed steps)"
    t_ramp = time.perf_counter() - t_ramp
    assert runner.bond_dim(st) == 4, f"bond dimension {runner.bond_dim(st)} after the ramp"
    secs, sweeps = [], []
    for ins in layers[first:first + steps]:
        t0 = time.perf_counter()
        sweeps.append(runner.run_layer(ctx, st, ins["xtime"], ins["ztime"]))
        secs.append(time.perf_counter() - t0)
        log(f"[cpu {runner.kind}] step {len(secs)}/{steps}: {secs[-1]:.1f} s, {sweeps[-1]} sweeps")
    dt = float(np.sum(secs))
    return {"value": steps / dt, "unit": "steps/s", "cores": os.cpu_count(), "kind": runner.kind,
            "sample": (f"{steps} steady-state step(s) (D=4, complex128: the reference's default precision) of the full "
                       f"{N_QUBITS}-qubit config itself, {how}; {dt / steps:.1f} s/step measured, nothing scaled; "
                       f"numpy/OpenBLAS may use all {os.cpu_count()} host cores"),
            "s_per_step": dt / steps, "sweeps_per_step": float(np.mean(sweeps)), "sweeps": sweeps, "ramp_s": t_ramp,
            "bloch": runner.bloch(ctx, st), "lmbds": runner.lmbds(st)}


def parity_block(cfg: dict, bloch_gpu, lmbds_gpu, sweeps_gpu, cpu: dict) -> dict:
    """c64 GPU against the c128 CPU implementation after the same steps from the same state (gauge-invariant
    observables only); asserts the stated fp32 tolerances."""
    from bqa_b200.benchmarking import ising_energy
    d = np.abs(np.asarray(bloch_gpu) - cpu["bloch"])
    spins = lambda b: np.where(b[:, 2] > 0, 1.0, -1.0)
    e_gpu = ising_energy(cfg["edges"], cfg["nodes"], spins(np.asarray(bloch_gpu)))
    e_cpu = ising_energy(cfg["edges"], cfg["nodes"], spins(cpu["bloch"]))
    lm_g, lm_c = np.sort(lmbds_gpu, axis=1)[:, ::-1], np.sort(cpu["lmbds"], axis=1)[:, ::-1]
    par = {"max_abs": float(d.max()), "mean_abs": float(d.mean()), "energy_rel": abs(e_gpu - e_cpu) / max(abs(e_cpu), 1e-300),
           "sweeps_diff": int(np.max(np.abs(np.asarray(sweeps_gpu) - np.asarray(cpu["sweeps"])))),
           "lmbds_max_abs": float(np.abs(lm_g - lm_c).max()), "sign_flips": int(np.sum(spins(np.asarray(bloch_gpu)) != spins(cpu["bloch"]))),
           "qubits": int(d.shape[0]), "steps": len(cpu["sweeps"]), "against": cpu["kind"],
           "tolerance": {"max_abs": TOL_BLOCH_MAX, "mean_abs": TOL_BLOCH_MEAN, "energy_rel": TOL_ENERGY_REL,
                         "sweeps_diff": TOL_SWEEPS}}
    par["ok"] = bool(par["max_abs"] <= TOL_BLOCH_MAX and par["mean_abs"] <= TOL_BLOCH_MEAN
                     and par["energy_rel"] <= TOL_ENERGY_REL and par["sweeps_diff"] <= TOL_SWEEPS)
    return par


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, CPU_REF_STEPS))
    total = schedule_len(args.steps, max(args.warmup, 3))     # the schedule our arm runs for the same K, W
    cfg = make_config(N_QUBITS, total)
    first = RAMP + max(args.warmup, 3)                        # schedule position of our arm's first timed step
    runner = CpuRunner()
    snap = None
    try:
        import torch
        if torch.cuda.is_available():
            from bqa_b200.build import build
            build()
            snap = gpu_state_after_ramp(cfg, first, torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    except Exception as e:
        log(f"no GPU hand-over ({type(e).__name__}: {e}): the CPU ramps from the initial state (about 15 minutes)")
    r = cpu_run(runner, cfg, snap, first, steps)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 0, "requested": {"steps": args.steps, "warmup": args.warmup},
            "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": make_bench_config(args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "sweeps_per_step": r["sweeps_per_step"], "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
            "note": (f"the first {steps} of the K timed steps of the 100k config, each a full annealing step on the "
                     "host cores; the D=4 start state comes from the GPU run (untimed)")}
    emit(line)


def make_bench_config(n_gpus: int) -> dict:
    return {"workload": "random 3-regular QUBO, 100,000 qubits (BASELINE.json configs[3]), "
                        "generate_qubo_on_random_regular_graph(100000, 3, seed=42)",
            "max_bond_dim": 4, "bp_eps": 1e-6, "pinv_eps": 1e-6, "damping": 0.0, "dt": DT,
            "schedule": f"total_time=0.2*S, steps=S>={SCHEDULE_STEPS}, mixing 1->0; first {RAMP} steps untimed (D 1->4)",
            "partition": "none" if n_gpus == 1 else f"node-partitioned over {n_gpus} GPUs, halo exchange per BP sweep",
            "l2": "state (T 102 MB + 2 x 38 MB messages + ext/canon 2 x 154 MB) exceeds the 126 MB L2; no explicit flush"}


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), __file__,
                   "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
            raise SystemExit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} does not match WORLD_SIZE {world}")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from bqa_b200 import _lib
    from bqa_b200.build import build
    from bqa_b200.config import config_to_context
    if rank == 0:
        build()
    if world > 1:
        dist.barrier()
    lib = _lib.load_library()
    steps, warmup = args.steps, args.warmup
    total = schedule_len(steps, warmup)
    t_setup = time.perf_counter()
    cfg = make_config(N_QUBITS, total)
    ctx = config_to_context(cfg)
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    if world == 1:
        from bqa_b200.engine import Engine
        eng = Engine(ctx, precision="single", device=dev)
    else:
        from bqa_b200.partitioned import PartitionedEngine
        eng = PartitionedEngine(ctx, precision="single", device=dev)
    log(f"[rank {rank}] setup {time.perf_counter() - t_setup:.1f} s")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    k = 0
    for ins in layers[:RAMP]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    k = RAMP
    assert eng.D == 4, f"bond dimension {eng.D} after the ramp"
    for ins in layers[k:k + warmup]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    k += warmup
    snapshot = eng.state_to_host(pinned=True)

    # ---- timed region: exactly K steps, state resident in HBM --------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n0 = len(eng.stats["bp_sweeps"])
    launches0 = lib.launch_count()
    canon0 = lib.canon_stats()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    ev0.record()
    eng.run_layers(layers[k:k + steps])      # = run_layer per step, each announcing the next step's ztime
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    launches = lib.launch_count() - launches0
    torch.cuda.synchronize(dev)
    canon1 = lib.canon_stats()
    sweeps = eng.stats["bp_sweeps"][n0:]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    bloch_resident = eng.bloch_vectors()

    # ---- instrumented pass: per-launch duration of the dominant kernel (BP sweep) ------------
    roof = measure_roofline(eng, lib, layers[k + steps:k + 2 * steps], torch, dev) if rank == 0 or world > 1 else None

    # clocks under load: samples from the start of the timed region to the end of the instrumented replay of the same
    # workload (the timed region alone is a few tens of milliseconds: too short for nvidia-smi's sampling period)
    clocks = sampler.stop(tw0, time.perf_counter()) if sampler else None
    if clocks is not None:
        clocks["window"] = "timed region + instrumented replay"

    # ---- end-to-end through the public engine API with HOST buffers --------------------------
    e2e = measure_e2e(eng, snapshot, layers[k:k + steps], torch, dev, bloch_resident, world, dist)

    # ---- multi-GPU correctness: rank 0 replays ramp + warm-up + the K timed steps on a single-GPU engine ------
    par1 = None
    if world > 1 and rank == 0:
        from bqa_b200.engine import Engine
        one = Engine(ctx, precision="single", device=dev)
        for ins in layers[:k + steps]:
            one.run_layer(ins["xtime"], ins["ztime"])
        d1 = np.abs(one.bloch_vectors() - bloch_resident)
        par1 = {"max_abs": float(d1.max()), "expected": 0.0, "steps_compared": k + steps,
                "sweeps_equal": one.stats["bp_sweeps"][:k + steps] == eng.stats["bp_sweeps"][:k + steps],
                "what": "Bloch vectors of the partitioned run after the timed region vs a single-GPU replay on rank 0"}
        del one

    # ---- configs[4] (500k qubits) with the same kernels, at N = 8 only ------------------------------------------
    cfg5 = None
    if world == 8 and not args.no_config5:
        del eng
        torch.cuda.empty_cache()
        cfg5 = config5_block(steps, warmup, torch, dist, dev, rank, world)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    E2 = int(ctx.edges_number)
    value = steps / (ms * 1e-3)
    bp_updates = float(np.sum(sweeps)) * E2
    line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c64", "data": "synthetic", "config": make_bench_config(world),
            "sweeps_per_step": float(np.mean(sweeps)), "bp_sweep_msg_updates_per_step": bp_updates / steps,
            "clocks": clocks, "gpu_launches": int(launches), "wall_ms_per_step": (tw1 - tw0) * 1e3 / steps}
    if canon1[0] > canon0[0]:                                  # Jacobi sweeps per warp-level 8 x 8 problem in the timed region
        line["canon_jacobi_sweeps_per_problem"] = (canon1[1] - canon0[1]) / (canon1[0] - canon0[0])
    if roof:
        line["roofline"] = roof["roofline"]
        line["bp_msg_updates_per_s"] = roof["bp_msg_updates_per_s"]
        line["kernel_ms_per_step"] = roof["kernel_ms_per_step"]     # instrumented replay of the NEXT K steps of the schedule
        line["kernel_ms_per_step"]["replay_sweeps_per_step"] = roof["replay_sweeps_per_step"]
    if e2e:
        line["e2e"] = e2e
    if roof and world > 1:
        line["roofline"]["traffic"] = None                    # the ncu capture behind `traffic` is a 1-GPU launch
    if world == 1 and not args.no_cpu:
        # cpu_baseline + parity: ONE steady-state step of the full 100k config on the host cores (the reference's numpy
        # backend from baseline/_ref, else the oracle port) from the state the timed region started with, against the
        # same step on the GPU from the same state
        runner = CpuRunner()
        snap = {k: v for k, v in snapshot.items() if k != "_pinned"}
        snap["node_ids"] = {c.degree: c.node_ids_host for c in eng.classes}
        r = cpu_run(runner, cfg, snap, k, args.cpu_steps)
        eng.load_state(snapshot)
        n1 = len(eng.stats["bp_sweeps"])
        for ins in layers[k:k + args.cpu_steps]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        line["cpu_baseline"] = {kk: r[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        line["parity"] = parity_block(cfg, eng.bloch_vectors(), eng.lmbds_numpy(), eng.stats["bp_sweeps"][n1:], r)
    if par1 is not None:
        line["parity_vs_1gpu"] = par1
    if cfg5 is not None:
        line["config5_500k"] = cfg5
    emit(line)
    if line.get("parity") and not line["parity"]["ok"]:
        raise SystemExit(f"parity outside the stated fp32 tolerance: {line['parity']}")
    if par1 is not None and par1["max_abs"] != 0.0:
        raise SystemExit(f"the partitioned run differs from the single-GPU run: {par1}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config5_block(steps: int, warmup: int, torch, dist, dev, rank: int, world: int) -> dict | None:
    """BASELINE.json configs[4]: the 500 000-qubit random 3-regular QUBO (benchmarks_against_mqlib/
    random_3_regular_qubo_500000.py shape, dt = 0.2) sharded over the 8 GPUs with the same kernels: steps/s of K timed
    steps after the ramp, and the partitioned result against a single-GPU replay of the same steps on rank 0."""
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    from bqa_b200.partitioned import PartitionedEngine
    t0 = time.perf_counter()
    cfg = make_config(500_000, schedule_len(steps, warmup))
    ctx = config_to_context(cfg)
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    eng = PartitionedEngine(ctx, precision="single", device=dev)
    k = RAMP + warmup
    for ins in layers[:k]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    assert eng.D == 4
    dist.barrier()
    torch.cuda.synchronize(dev)
    n0 = len(eng.stats["bp_sweeps"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_layers(layers[k:k + steps])
    e1.record()
    dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    bloch = eng.bloch_vectors()
    sweeps = eng.stats["bp_sweeps"][n0:]
    out = None
    if rank == 0:
        one = Engine(ctx, precision="single", device=dev)
        for ins in layers[:k + steps]:
            one.run_layer(ins["xtime"], ins["ztime"])
        d1 = np.abs(one.bloch_vectors() - bloch)
        spins = np.where(bloch[:, 2] > 0, 1.0, -1.0)
        from bqa_b200.benchmarking import ising_energy
        out = {"workload": "random 3-regular QUBO, 500,000 qubits (BASELINE.json configs[4]), "
                           "generate_qubo_on_random_regular_graph(500000, 3, seed=42), node-partitioned over 8 GPUs",
               "value": steps / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms / steps, "steps": steps,
               "sweeps_per_step": float(np.mean(sweeps)),
               "bp_msg_updates_per_s": float(np.sum(sweeps)) * int(ctx.edges_number) / (ms * 1e-3),
               "parity_vs_1gpu": {"max_abs": float(d1.max()), "expected": 0.0, "steps_compared": k + steps,
                                  "sweeps_equal": one.stats["bp_sweeps"][:k + steps] == eng.stats["bp_sweeps"][:k + steps]},
               "energy_of_sign_z": ising_energy(cfg["edges"], cfg["nodes"], spins),
               "cut_fraction": None, "wall_s": None}
        del one
    del eng
    dist.barrier()
    if out is not None:
        out["wall_s"] = time.perf_counter() - t0
    return out


def measure_roofline(eng, lib, layers, torch, dev) -> dict:
    """Re-runs steps with CUDA events around every launch of the instrumented entry points (same stream the
    kernels are launched on).  roofline.achieved = algorithmic bytes of a BP sweep launch / its mean duration."""
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        with open(ppath) as f:
            peaks = json.load(f)
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")
    # entry point -> bucket (the partitioned engine calls the *_p2p variants)
    buckets = {"bp_sweep": "bp_sweep", "bp_sweep_p2p": "bp_sweep", "ext_msgs": "ext_msgs", "ext_msgs_p2p": "ext_msgs",
               "ext_msgs_after_run": "ext_msgs",
               "canonicalize": "canonicalize", "canonicalize_ordered": "canonicalize", "sort_edges_by_cost": "canonicalize",
               "canonicalize_p2p": "canonicalize",
               "apply_update": "apply_update", "sweep_sync": "sweep_sync",
               "gauge_msgs": "gauge_msgs", "bp_run": "bp_run"}
    names = sorted(set(buckets.values()))
    events = {n: [] for n in names}
    orig = {n: getattr(lib, n) for n in buckets}

    def wrap(name):
        fn = orig[name]

        def call(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ret = fn(*a)
            e1.record()
            events[buckets[name]].append((e0, e1, a))
            return ret
        return call
    for n in buckets:
        setattr(lib, n, wrap(n))
    n_runs0 = len(eng.stats["bp_sweeps"])
    try:
        eng.run_layers(layers)
        torch.cuda.synchronize(dev)
    finally:
        for n in buckets:
            setattr(lib, n, orig[n])
    nsteps = max(len(layers), 1)
    out = {"kernel_ms_per_step": {}}
    for n in names:
        out["kernel_ms_per_step"][n] = sum(e0.elapsed_time(e1) for e0, e1, _ in events[n]) / nsteps
    s = 8
    if events["bp_run"]:
        # single-launch BP runs (bqa_b200_bp_run): one launch = all sweeps of one BP run of the class; its
        # algorithmic bytes = bytes per sweep x sweeps it executed (engine statistics), args = (prec, degree, D, B, ...)
        a0 = events["bp_run"][0][2]
        degree, D, B = int(a0[1]), int(a0[2]), int(a0[3])
        bytes_per_node = s * (2 * D ** degree + 3 * degree * D * D)
        sweeps = np.array(eng.stats["bp_sweeps"][n_runs0:n_runs0 + len(events["bp_run"])], dtype=np.float64)
        durs = np.array([e0.elapsed_time(e1) for e0, e1, _ in events["bp_run"]])
        alg_bytes = bytes_per_node * B                                    # per sweep
        mean_ms = float(durs.sum() / sweeps.sum())                        # per sweep, barriers included
        launches_timed, noop = int(len(durs)), 0
        kernel = f"bp_run (degree {degree}, D={D}, c64): persistent kernel, per-sweep figures = launch / sweeps executed"
    else:
        # BP sweep launches of the 3-regular class: args = (prec, degree, D, B, ...)
        bp = [(e0.elapsed_time(e1), a) for e0, e1, a in events["bp_sweep"]]
        # converged sweeps exit early (device-side no-op): keep launches that did the work = all but the trailing
        # no-op launches of each BP run; identify them by duration (a no-op takes a few microseconds)
        durs = np.array([d for d, _ in bp])
        work = durs > 0.5 * np.median(durs)
        a0 = bp[0][1]
        degree, D, B = int(a0[1]), int(a0[2]), int(a0[3])
        bytes_per_node = s * (2 * D ** degree + 3 * degree * D * D)
        alg_bytes = bytes_per_node * B
        mean_ms = float(durs[work].mean())
        launches_timed, noop = int(work.sum()), int((~work).sum())
        kernel = f"bp_sweep (degree {degree}, D={D}, c64)"
    achieved = alg_bytes / (mean_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "bp_sweep_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": traffic, "kernel": kernel, "launch_ms": mean_ms,
                       "algorithmic_bytes_per_launch": alg_bytes, "launches_timed": launches_timed,
                       "noop_launches_skipped": noop, "peak_source": peak_src}
    out["bp_msg_updates_per_s"] = degree * B / (mean_ms * 1e-3)
    out["replay_sweeps_per_step"] = float(np.sum(eng.stats["bp_sweeps"][n_runs0:])) / nsteps
    return out


def measure_e2e(eng, snapshot, layers, torch, dev, bloch_resident, world, dist) -> dict:
    """Same K steps driven from HOST buffers through the public engine API: load_state (pinned host -> HBM; every rank
    its own shard), run_layers over the K steps (each reads its bond-dimension decision and BP control block back), bloch_vectors
    (HBM -> host, gathered over the ranks).  Wall clock between barriers, max over ranks; bytes summed over ranks."""
    h2d = sum(int(t.numel() * t.element_size()) for t in snapshot["_pinned"].values())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    n0 = len(eng.stats["bp_sweeps"])
    t0 = time.perf_counter()
    eng.load_state(snapshot)
    eng.run_layers(layers)
    b = eng.bloch_vectors()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    K = len(layers)
    reads = K + len(eng.stats["bp_sweeps"]) - n0              # column maxima per step + control block per BP read (>= 1 per run)
    d2h = b.shape[0] * 4 * 4 / world + K * eng.colmax_bytes + (reads - K) * eng.ctrl_bytes_per_bp_read
    if world > 1:
        t = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, h2d, d2h = float(tmax[0]), float(t[1]), float(t[2])
    assert np.abs(b - bloch_resident).max() < 1e-6, "end-to-end run disagrees with the HBM-resident run"
    return {"value": K / dt, "unit": "steps/s", "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
            "what": "Engine.load_state(pinned host snapshot) + Engine.run_layers(K steps) + Engine.bloch_vectors(), "
                    "wall clock incl. all copies"}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line, written to the process's original stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main() -> None:
    # libraries (NCCL's version banner, warnings) write to fd 1: keep stdout for the JSON line alone
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=1, help="steps of the cpu_baseline / parity leg (N=1 only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config5", action="store_true", help="N = 8: skip the 500k-qubit block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
