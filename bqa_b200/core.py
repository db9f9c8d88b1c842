"""``run_qa``: drop-in for ``bqa.run_qa`` (reference src/bqa/core.py:13-35) on the B200 engine, plus checkpoint /
resume of a running schedule (the reference keeps ``State`` only in memory, state.py:24-29: a 10 000-step anneal
cannot be resumed there)."""
from __future__ import annotations

import json
import logging
import os

import numpy as np

from .config import config_to_context
from .engine import Engine

log = logging.getLogger(__name__)

CHECKPOINT_VERSION = 2


def engine_fingerprint(engine) -> str:
    """Hash of what a checkpoint is only valid for: the compiled layouts (graph, slot numbering, amplitudes) and the
    numerical parameters -- a config of the same sizes but other couplings would load silently otherwise."""
    import hashlib
    h = hashlib.sha256()
    for c in engine.classes:
        h.update(repr((c.degree, c.B)).encode())
        h.update(np.ascontiguousarray(c.node_ids_host).tobytes())
        for t in (c.in_pos, c.out_pos, c.node_ampls, c.edge_ampls):
            h.update(t.cpu().numpy().tobytes())
    ctx = engine.ctx
    for name in ("nodes_number", "edges_number", "max_bond_dim", "max_bp_iters_number", "bp_eps", "pinv_eps", "damping",
                 "measurement_threshold", "seed"):
        h.update(repr((name, getattr(ctx, name, None))).encode())
    return h.hexdigest()


def save_checkpoint(path: str, engine, next_instruction: int, n_instructions: int, results: list) -> None:
    """Atomic dump of {tensors per degree, messages, lambdas, bond dimension, host RNG state, position in the
    instruction list, results so far}: everything ``run_context(resume=True)`` needs to continue bit for bit."""
    snap = engine.state_to_host()
    arrays = {f"tensors_{d}": t for d, t in snap["tensors"].items()}
    arrays.update(msgs=snap["msgs"], lmbds=snap["lmbds"])
    meta = {"version": CHECKPOINT_VERSION, "D": int(snap["D"]), "next": int(next_instruction), "n": int(n_instructions),
            "precision": engine.precision, "rng": engine.rng.bit_generator.state, "results": results,
            "degrees": sorted(int(d) for d in snap["tensors"]), "fingerprint": engine_fingerprint(engine)}
    path = engine.checkpoint_file(path)
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "wb") as f:
        np.savez(f, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    engine.checkpoint_barrier()                 # partitioned run: every rank has its file before any is published
    os.replace(tmp, path)


def load_checkpoint(path: str, engine, n_instructions: int):
    """Restores the engine from ``save_checkpoint``'s file; returns (next instruction index, results so far)."""
    with np.load(engine.checkpoint_file(path)) as z:
        meta = json.loads(bytes(z["meta"]).decode())
        if meta.get("version") != CHECKPOINT_VERSION:
            raise ValueError(f"{path}: checkpoint version {meta.get('version')} is not {CHECKPOINT_VERSION}")
        if meta["n"] != n_instructions:
            raise ValueError(f"{path}: written for a schedule of {meta['n']} instructions, this one has {n_instructions}")
        if meta.get("fingerprint") != engine_fingerprint(engine):
            raise ValueError(f"{path}: written for another config (graph, amplitudes or parameters differ)")
        if meta["precision"] != engine.precision:
            raise ValueError(f"{path}: written in {meta['precision']} precision, the engine runs in {engine.precision}")
        snap = {"D": meta["D"], "tensors": {d: z[f"tensors_{d}"] for d in meta["degrees"]}, "msgs": z["msgs"],
                "lmbds": z["lmbds"]}
        engine.load_state(snap)
    engine.rng.bit_generator.state = meta["rng"]
    engine.checkpoint_agree(int(meta["next"]), "schedule position")
    return int(meta["next"]), meta["results"]


def run_context(context, precision=None, device=None, engine_cls=Engine, checkpoint=None, checkpoint_every=0,
                resume=False, **engine_kwargs) -> list:
    """Interprets the instruction list of a compiled context (ours or bqa's own ``Context``).

    ``checkpoint``: path of an ``.npz`` file written every ``checkpoint_every`` instructions (0: never) and, with
    ``resume=True`` and the file present, read back first: the run continues after the last completed instruction
    and returns the same result list as an uninterrupted run.  A partitioned engine writes one file per rank
    (``<name>.rank<r>of<P>.npz``) and resumes only when every rank holds a file of the same save."""
    engine = engine_cls(context, precision=precision, device=device, **engine_kwargs)
    instructions = list(context.instructions)
    n = len(instructions)
    results = []
    start = 0
    if resume and checkpoint and engine.checkpoint_agree(os.path.exists(engine.checkpoint_file(checkpoint)), "presence"):
        start, results = load_checkpoint(checkpoint, engine, n)
        log.info(f"Resumed from {checkpoint} at instruction number {start} / {n}")
    for i in range(start, n):
        ins = instructions[i]
        log.info(f"Instruction number {i} / {n} started")
        if isinstance(ins, dict):
            nxt = instructions[i + 1] if i + 1 < n else None    # "type" is ignored like in the reference (core.py:23)
            engine.run_layer(ins["xtime"], ins["ztime"], next_ztime=nxt["ztime"] if isinstance(nxt, dict) else None)
        elif ins == "measure":
            results.append(["measurement_outcomes", engine.measure()])
        elif ins == "get_bloch_vectors":
            results.append(["bloch_vectors", [[float(x), float(y), float(z)] for x, y, z in engine.bloch_vectors()]])
        else:
            raise ValueError(f"Unknown instruction {ins}")
        if checkpoint and checkpoint_every > 0 and (i + 1) % checkpoint_every == 0 and i + 1 < n:
            save_checkpoint(checkpoint, engine, i + 1, n, results)
    return results


def run_qa(config, precision=None, device=None, checkpoint=None, checkpoint_every=0, resume=False) -> list:
    """Same input dict / JSON shape and same output list as ``bqa.run_qa``:
    ``[["bloch_vectors", [[x, y, z], ...]] | ["measurement_outcomes", [+1 | -1, ...]], ...]``."""
    return run_context(config_to_context(config), precision=precision, device=device, checkpoint=checkpoint,
                       checkpoint_every=checkpoint_every, resume=resume)
