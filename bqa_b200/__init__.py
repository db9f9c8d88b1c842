"""bqa_b200: B200-native (sm_100a) belief-propagation engine behind bqa's backend interface.

Public surface (mirrors the reference package ``bqa``):

    run_qa(config)                      -- bqa.run_qa (src/bqa/core.py:13-35)
    config_to_context(config)           -- bqa.config.core.config_to_context
    Engine(context)                     -- run_layer / run_bp / measure / bloch_vectors (src/bqa/state.py)
    register_with_bqa()                 -- adds the "b200" backend to bqa's registry when bqa is installed

Importing the package does not need a GPU; constructing an Engine does (there is no CPU fallback).
"""
from .config import ConfigSyntaxError, Context, Layout, config_to_context  # noqa: F401
from .core import run_context, run_qa  # noqa: F401
from .engine import Engine  # noqa: F401
from .backend import register_with_bqa  # noqa: F401

__version__ = "0.1.0"
