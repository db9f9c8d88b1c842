"""Registration of the "b200" backend with an installed bqa (reference plugin boundary: the registry dict
``bqa.backends.BACKEND_STR_TO_BACKEND`` validated by ``config_syntax.py:63-70`` and resolved into
``Context.backend`` by ``config_canonicalization.py:211``).

``B200Backend`` (bqa_b200/tensor_backend.py) is a real implementation of bqa's ``Tensor`` ABC on device arrays: all 36
abstract raw operations are CUDA kernels behind the C ABI and the hot composites call the fused kernels, so the
unmodified ``bqa.state`` engine runs with it op by op (``run_qa(config, fused=False)``).  The default dispatch for a
whole config stays the fused engine: ``bqa.run_qa`` hands contexts whose backend is ``B200Backend`` to
``bqa_b200.run_context``, which keeps the state resident in HBM, fuses across the ABC's op boundaries and runs the BP
loop on the device.  INTEGRATION.md shows the change a bqa maintainer would make instead of the patch below."""
from __future__ import annotations

_registered = None


def register_with_bqa():
    """Idempotent.  Returns the ``B200Backend`` class now listed as ``"b200"`` in bqa's registry."""
    global _registered
    if _registered is not None:
        return _registered
    import bqa
    import bqa.backends as backends
    import bqa.core as core

    from .core import run_context
    from .tensor_backend import make_backend_class

    B200Backend = make_backend_class(backends.Tensor)
    backends.BACKEND_STR_TO_BACKEND["b200"] = B200Backend
    reference_run_qa = core.run_qa

    def run_qa(config, fused: bool = True, **engine_kwargs) -> list:
        """``fused=True``: the compiled Context is executed by bqa_b200.Engine (fused kernels, device-resident BP loop);
        ``fused=False``: by the reference's own engine (src/bqa/state.py) through the Tensor methods of B200Backend."""
        context = core.config_to_context(config)
        if fused and context.backend is B200Backend:
            return run_context(context, **engine_kwargs)
        return reference_run_qa(config)

    run_qa.__doc__ = (reference_run_qa.__doc__ or "") + (run_qa.__doc__ or "")
    core.run_qa = run_qa
    bqa.run_qa = run_qa
    _registered = B200Backend
    return B200Backend
