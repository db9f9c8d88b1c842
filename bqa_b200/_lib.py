"""ctypes binding of libbqa_b200.so (C ABI: include/bqa_b200.h).

The product path loads the CUDA library or raises -- there is no CPU fallback.  (The test-suite may bind a
host emulation of the same ABI through `bind(path)`; nothing in the package does.)"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libbqa_b200.so")

C64, C128 = 0, 1
_vp, _i, _ll, _d, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_size_t

# name -> argtypes, exactly as declared in include/bqa_b200.h
SIGNATURES = {
    "bqa_b200_bp_sweep_p2p": [_i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _d, _i, _d, _i, _vp, _vp, _vp, _sz, _vp, _vp, _vp],
    "bqa_b200_ext_msgs_p2p": [_i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _sz, _vp, _vp, _vp],
    "bqa_b200_sweep_sync": [_i, _i, _i, _vp, _i, _vp, C.c_uint, _vp, _vp],
    "bqa_b200_gauge_msgs": [_i, _i, _i, _ll, _vp, _vp, _vp],
    "bqa_b200_bp_run": [_i, _i, _i, _ll, _vp, _vp, _vp, _i, _vp, _vp, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp,
                        C.c_uint, _vp, _vp, _ll, _vp],
    "bqa_b200_bp_sweep": [_i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _d, _i, _d, _i, _vp, _vp, _vp, _sz, _vp],
    "bqa_b200_ext_msgs_after_run": [_i, _i, _i, _ll, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp,
                                    _vp],
    "bqa_b200_ext_msgs": [_i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _sz, _vp],
    "bqa_b200_canonicalize": [_i, _i, _ll, _vp, _vp, _vp, _vp, _d, _i, _vp],
    "bqa_b200_canonicalize_ordered": [_i, _i, _ll, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _vp],
    "bqa_b200_sort_edges_by_cost": [_ll, _vp, _vp, _vp],
    "bqa_b200_canonicalize_p2p": [_i, _i, _ll, _vp, _vp, _vp, _vp, _d, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp],
    "bqa_b200_apply_update": [_i, _i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _vp, _sz, _vp],
    "bqa_b200_density": [_i, _i, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "bqa_b200_argmax_unmeasured": [_i, _ll, _vp, _vp, _vp, _vp, _vp],
    "bqa_b200_project_node": [_i, _i, _i, _vp, _ll, _i, _vp],
    "bqa_b200_threshold_project": [_i, _i, _i, _ll, _vp, _vp, _vp, _vp, _d, _vp, _vp],
    # all degree classes in one launch (bqa_multiclass.cuh); the class table is a host array of ClassDesc
    "bqa_b200_ext_msgs_classes": [_i, _i, _vp, _i, _vp, _vp, _d, _vp, _sz, _vp],
    "bqa_b200_apply_update_classes": [_i, _i, _vp, _i, _i, _vp, _vp, _vp, _d, _d, _vp, _sz, _vp],
    "bqa_b200_bp_run_classes": [_i, _i, _vp, _i, _vp, _vp, _i, _d, _d, _i, _vp, _vp, _vp, _sz, _vp],
    # raw operations of the backend interface (bqa_tensor_ops.cu)
    "bqa_b200_t_unary": [_i, _i, _ll, _vp, _vp, _vp],
    "bqa_b200_t_binary": [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "bqa_b200_t_copy": [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "bqa_b200_t_rows": [_i, _i, _ll, _ll, _vp, _vp, _vp, _vp],
    "bqa_b200_t_fill": [_i, _ll, _vp, _d, _d, _vp],
    "bqa_b200_t_axpby": [_i, _ll, _vp, _vp, _d, _d, _vp],
    "bqa_b200_t_max_abs": [_i, _ll, _vp, _vp, _vp],
    "bqa_b200_t_col_max": [_i, _ll, _ll, _vp, _vp, _vp],
    "bqa_b200_t_batch_reduce": [_i, _i, _ll, _ll, _i, _vp, _vp, _vp],
    "bqa_b200_t_diag": [_i, _ll, _i, _vp, _vp, _vp],
    "bqa_b200_t_matmul": [_i, _ll, _i, _i, _i, _vp, _vp, _vp, _vp],
    "bqa_b200_t_svd": [_i, _ll, _i, _vp, _vp, _vp, _vp, _d, _vp, _sz, _vp],
    "bqa_b200_t_bloch_to_rho": [_i, _ll, _vp, _vp, _vp],
}
EXPORTS = list(SIGNATURES) + ["bqa_b200_last_error", "bqa_b200_version", "bqa_b200_launch_count",
                              "bqa_b200_workspace_bytes", "bqa_b200_set_kernel_mode", "bqa_b200_canon_stats",
                              "bqa_b200_t_svd_scratch_bytes", "bqa_b200_canon_stats_detail",
                              "bqa_b200_set_barrier_timeout", "bqa_b200_set_bp_trace", "bqa_b200_canon_span"]


_CUDA_ONLY = ("bqa_b200_canonicalize_ordered", "bqa_b200_sort_edges_by_cost", "bqa_b200_canonicalize_p2p",
              "bqa_b200_ext_msgs_after_run")


class ClassDesc(C.Structure):
    """bqa_b200_class of include/bqa_b200.h: one row of the degree-class table of the *_classes entry points"""
    _fields_ = [("degree", _i), ("B", _ll), ("T_in", _vp), ("T_out", _vp), ("in_pos", _vp), ("out_pos", _vp),
                ("lmbd_pos", _vp), ("node_ampls", _vp), ("edge_ampls", _vp)]


class Library:
    """Thin checked wrapper: every entry point returns 0 or raises RuntimeError(last_error)."""

    def __init__(self, path: str):
        self.path = path
        self._dll = C.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            try:
                fn = getattr(self._dll, name)
            except AttributeError:
                # the test-only host emulation covers the per-class fused entry points only
                if name.startswith("bqa_b200_t_") or name.endswith("_classes") or name in _CUDA_ONLY:
                    continue
                raise
            fn.argtypes = argtypes
            fn.restype = _i
            setattr(self, name[len("bqa_b200_"):], self._checked(name, fn))
        self._dll.bqa_b200_last_error.restype = C.c_char_p
        self._dll.bqa_b200_version.restype = _i
        self._dll.bqa_b200_launch_count.restype = _ll
        self._dll.bqa_b200_workspace_bytes.argtypes = [_i, _i, _i, _i]
        self._dll.bqa_b200_workspace_bytes.restype = _sz
        if hasattr(self._dll, "bqa_b200_t_svd_scratch_bytes"):
            self._dll.bqa_b200_t_svd_scratch_bytes.argtypes = [_i, _i]
            self._dll.bqa_b200_t_svd_scratch_bytes.restype = _sz

    def _checked(self, name, fn):
        def call(*args):
            rc = fn(*args)
            if rc == 2 and name in ("bqa_b200_bp_run", "bqa_b200_ext_msgs_after_run"):
                return False                  # no single-launch kernel for this shape: not an error
            if rc != 0:
                raise RuntimeError(f"{name}: {self._dll.bqa_b200_last_error().decode()}")
            return True
        call.__name__ = name
        return call

    def version(self) -> int:
        return int(self._dll.bqa_b200_version())

    def launch_count(self) -> int:
        return int(self._dll.bqa_b200_launch_count())

    def set_kernel_mode(self, mode: int) -> None:
        """0: specialised kernels where they exist (default); 1: generic kernels only; 2: like 0 with the first-design
        n = 8 canonicalizer (side-by-side measurements)."""
        if self._dll.bqa_b200_set_kernel_mode(int(mode)) != 0:
            raise RuntimeError(self._dll.bqa_b200_last_error().decode())

    def set_bp_trace(self, device_ptr) -> None:
        self._dll.bqa_b200_set_bp_trace.argtypes = [_vp]
        self._dll.bqa_b200_set_bp_trace(device_ptr)

    def set_barrier_timeout(self, seconds: float) -> None:
        self._dll.bqa_b200_set_barrier_timeout.argtypes = [_d]
        if self._dll.bqa_b200_set_barrier_timeout(float(seconds)) != 0:
            raise RuntimeError(self._dll.bqa_b200_last_error().decode())

    def canon_stats(self) -> tuple[int, int, int]:
        out = (C.c_ulonglong * 3)()
        self._dll.bqa_b200_canon_stats(out)
        return int(out[0]), int(out[1]), int(out[2])

    def svd_scratch_bytes(self, prec: int, n: int) -> int:
        return int(self._dll.bqa_b200_t_svd_scratch_bytes(prec, n))

    def canon_span(self) -> tuple[int, int]:
        out = (C.c_ulonglong * 2)()
        self._dll.bqa_b200_canon_span(out)
        return int(out[0]), int(out[1])

    def canon_stats_detail(self) -> list[int]:
        out = (C.c_ulonglong * 7)()
        self._dll.bqa_b200_canon_stats_detail(out)
        return [int(v) for v in out]

    def workspace_bytes(self, prec: int, degree: int, D: int, D_new: int) -> int:
        return int(self._dll.bqa_b200_workspace_bytes(prec, degree, D, D_new))

    @property
    def is_host_emulation(self) -> bool:
        return self.version() < 0


_cached: Library | None = None


def bind(path: str) -> Library:
    return Library(path)


def load_library() -> Library:
    """The CUDA library, built in-tree by `python -m bqa_b200.build` / `__graft_entry__.build()`."""
    global _cached
    if _cached is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m bqa_b200.build` (needs nvcc). "
                "bqa_b200 has no CPU fallback.")
        lib = Library(LIB_PATH)
        if lib.is_host_emulation:
            raise RuntimeError(f"{LIB_PATH} is not the CUDA build")
        _cached = lib
    return _cached
