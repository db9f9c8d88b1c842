"""Registration of the "b200" backend with an installed bqa (reference plugin boundary: the registry dict
``bqa.backends.BACKEND_STR_TO_BACKEND`` validated by ``config_syntax.py:63-70`` and resolved into
``Context.backend`` by ``config_canonicalization.py:211``).

bqa's compile step only needs the backend to build index tensors (``make_from_list`` / ``make_from_iter``,
``config_canonicalization.py:109-126, :205``); those stay host arrays, so ``B200Backend`` derives from bqa's
``NumPyBackend`` for them.  Execution does not go through the per-op ``Tensor`` methods: ``bqa.run_qa`` is
dispatched, for contexts whose backend is ``B200Backend``, to ``bqa_b200.run_context`` which drives the fused
CUDA kernels behind the C ABI (include/bqa_b200.h) from the very same ``Context``.  INTEGRATION.md shows the
three-line change a bqa maintainer would make instead of this monkey patch."""
from __future__ import annotations

_registered = None


def make_backend_class(numpy_backend_cls):
    """The marker class a bqa maintainer would register as "b200" (INTEGRATION.md): bqa's own numpy backend for
    the index tensors of the compile step; execution is dispatched on ``context.backend is B200Backend``."""

    class B200Backend(numpy_backend_cls):
        """numpy index tensors for the compile step, bqa_b200.Engine for execution."""

    return B200Backend


def register_with_bqa():
    """Idempotent.  Returns the ``B200Backend`` class now listed as ``"b200"`` in bqa's registry."""
    global _registered
    if _registered is not None:
        return _registered
    import bqa
    import bqa.backends as backends
    import bqa.core as core

    from .core import run_context

    B200Backend = make_backend_class(backends.NumPyBackend)
    backends.BACKEND_STR_TO_BACKEND["b200"] = B200Backend
    reference_run_qa = core.run_qa

    def run_qa(config, **engine_kwargs) -> list:
        context = core.config_to_context(config)
        if context.backend is B200Backend:
            return run_context(context, **engine_kwargs)
        return reference_run_qa(config)

    run_qa.__doc__ = reference_run_qa.__doc__
    core.run_qa = run_qa
    bqa.run_qa = run_qa
    _registered = B200Backend
    return B200Backend
