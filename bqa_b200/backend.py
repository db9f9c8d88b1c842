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

# Tensor methods that do arithmetic on tensor data (reference src/bqa/backends.py:28-572): everything the engine in
# src/bqa/state.py calls.  Constructors, shape queries and ``numpy`` stay: the compile step needs them for index tensors.
NUMERICAL_METHODS = (
    "pass_msgs", "get_density_matrices", "get_dist", "make_inplace_damping_update", "apply_x_gates", "apply_z_gates",
    "apply_canonicalizers", "apply_canonicalizers_with_extensions", "apply_conditional_z_gates", "mul_by_lmbds",
    "decompose_iden_using_msgs", "truncate_lmbds", "batch_truncate_all_but", "batch_tensordot", "batch_matmul",
    "get_batch_svd", "batch_trace_normalize", "batch_normalize", "batched_diag", "measure", "sqrt", "pinv", "inv",
    "batched_svd", "batched_matmul", "batched_trace", "batched_l2_norm", "max_norm",
    "measure_raw_tensor_by_position_in_place", "apply_x_to_phys_dim_raw", "apply_z_to_phys_dim_raw",
    "make_inplace_damping_update_raw",
)


def make_backend_class(numpy_backend_cls):
    """The marker class a bqa maintainer would register as "b200" (INTEGRATION.md): bqa's own numpy backend for
    the index tensors of the compile step; execution is dispatched on ``context.backend is B200Backend``."""

    class B200Backend(numpy_backend_cls):
        """numpy index tensors for the compile step, bqa_b200.Engine for execution.

        The numerical methods of the Tensor interface are NOT inherited: calling one (e.g. by driving ``bqa.state``
        by hand with this backend) would silently compute on the host under the name "b200".  They raise instead."""

    def _refuse(name):
        def method(self, *args, **kwargs):
            raise RuntimeError(
                f"B200Backend.{name}: the b200 backend has no per-op host path; it executes a compiled Context on the "
                "GPU through bqa.run_qa (after bqa_b200.register_with_bqa()) or bqa_b200.run_context")
        method.__name__ = name
        return method

    for name in NUMERICAL_METHODS:
        if hasattr(numpy_backend_cls, name):
            setattr(B200Backend, name, _refuse(name))
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
