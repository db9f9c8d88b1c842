"""``run_qa``: drop-in for ``bqa.run_qa`` (reference src/bqa/core.py:13-35) on the B200 engine."""
from __future__ import annotations

import logging

from .config import config_to_context
from .engine import Engine

log = logging.getLogger(__name__)


def run_context(context, precision=None, device=None, engine_cls=Engine, **engine_kwargs) -> list:
    """Interprets the instruction list of a compiled context (ours or bqa's own ``Context``)."""
    engine = engine_cls(context, precision=precision, device=device, **engine_kwargs)
    instructions = list(context.instructions)
    n = len(instructions)
    results = []
    for i, ins in enumerate(instructions):
        log.info(f"Instruction number {i} / {n} started")
        if isinstance(ins, dict):
            engine.run_layer(ins["xtime"], ins["ztime"])        # "type" is ignored like in the reference (core.py:23)
        elif ins == "measure":
            results.append(["measurement_outcomes", engine.measure()])
        elif ins == "get_bloch_vectors":
            results.append(["bloch_vectors", [[float(x), float(y), float(z)] for x, y, z in engine.bloch_vectors()]])
        else:
            raise ValueError(f"Unknown instruction {ins}")
    return results


def run_qa(config, precision=None, device=None) -> list:
    """Same input dict / JSON shape and same output list as ``bqa.run_qa``:
    ``[["bloch_vectors", [[x, y, z], ...]] | ["measurement_outcomes", [+1 | -1, ...]], ...]``."""
    return run_context(config_to_context(config), precision=precision, device=device)
