"""``B200Backend``: bqa's backend interface (the ABC ``bqa.backends.Tensor``, reference src/bqa/backends.py:28-252)
implemented on device arrays.

Every one of the 36 abstract raw operations runs as a CUDA kernel behind the C ABI (``bqa_b200_t_*``,
bqa_b200/csrc/bqa_tensor_ops.cu); the hot composites are overridden with the fused entry points the engine uses
(precedent: the CuPy backend overrides composites, backends.py:890-938):

    pass_msgs                      -> bqa_b200_bp_sweep / bqa_b200_ext_msgs      (backends.py:381-408, :519-526)
    get_density_matrices           -> bqa_b200_density (+ t_bloch_to_rho)        (backends.py:440-448)

(the remaining composites -- get_dist, damping, the canonicalizer algebra of state.py:171-200 -- run through the ABC's
own definitions on the raw kernels).

With this class registered as ``"b200"`` the UNMODIFIED reference engine (``bqa.state.run_layer`` / ``_run_bp`` /
``measure``) runs on the GPU op by op; ``bqa_b200.Engine`` stays the default dispatch for whole configs because it fuses
across ops and keeps the BP loop on the device (bqa_b200/backend.py).  PyTorch owns the device memory; no arithmetic is
done by torch or numpy here (``numpy`` is the D2H accessor the interface requires).

The class is built against whichever ``Tensor`` ABC is importable: ``make_backend_class(bqa.backends.Tensor)``.
"""
from __future__ import annotations

import ctypes as C
from math import prod

import numpy as np
import torch

from . import _lib

_COMPLEX = {np.dtype(np.complex64): (torch.complex64, _lib.C64), np.dtype(np.complex128): (torch.complex128, _lib.C128)}
U_INV, U_PINV, U_SQRT, U_SIN, U_COS, U_CONJ = range(6)
B_MUL, B_ADD, B_SUB, B_DIV = range(4)


def _default_complex() -> np.dtype:
    """The reference's working dtype (BQA_PRECISION, src/bqa/utils.py:9-30) when bqa is importable, else complex128."""
    try:
        from bqa.utils import NP_DTYPE
        return np.dtype(NP_DTYPE)
    except Exception:
        return np.dtype(np.complex128)


class DeviceArray:
    """Raw tensor of the backend: a dense row-major device buffer + shape.  dtype is complex64 / complex128 (data) or
    int64 (index tensors, backends.py:583); ``real`` marks complex storage that holds a real result (norms, maxima)."""

    __slots__ = ("t", "shape", "real")

    def __init__(self, t: torch.Tensor, shape, real: bool = False):
        self.t = t                         # flat, contiguous
        self.shape = tuple(int(s) for s in shape)
        self.real = real

    @property
    def is_index(self) -> bool:
        return self.t.dtype == torch.int64

    @property
    def prec(self) -> int:
        return _lib.C64 if self.t.dtype == torch.complex64 else _lib.C128

    @property
    def esize(self) -> int:
        return self.t.element_size()

    @property
    def size(self) -> int:
        return prod(self.shape)

    def ptr(self, offset_elems: int = 0) -> int:
        return self.t.data_ptr() + offset_elems * self.esize

    def host(self) -> np.ndarray:
        a = self.t.cpu().numpy().reshape(self.shape)
        return a.real.copy() if self.real else a

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.t.dtype}, device={self.t.device})"


def _ll(vals):
    return (C.c_longlong * max(len(vals), 1))(*[int(v) for v in vals])


def _dense_strides(shape):
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= s
    return list(reversed(st))


class _Runtime:
    """Library handle, device and stream shared by every B200Backend tensor."""

    def __init__(self):
        self.lib = None
        self.dev = None
        self._svd_scratch = None

    def ensure(self):
        if self.lib is None:
            self.lib = _lib.load_library()
            if not torch.cuda.is_available():
                raise RuntimeError("the b200 backend needs a CUDA device (no CPU fallback)")
            self.dev = torch.device("cuda", torch.cuda.current_device())
        return self.lib

    def stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def empty(self, shape, dtype) -> DeviceArray:
        self.ensure()
        return DeviceArray(torch.empty(max(prod(shape), 1), dtype=dtype, device=self.dev), shape)

    def svd_scratch(self, nbytes: int) -> torch.Tensor:
        if self._svd_scratch is None or self._svd_scratch.numel() < nbytes:
            self._svd_scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        return self._svd_scratch


RT = _Runtime()


# -------------------------------------------------------------------------------------------------------------------
# raw operations
# -------------------------------------------------------------------------------------------------------------------
def _upload(arr: np.ndarray) -> DeviceArray:
    RT.ensure()
    arr = np.asarray(arr)
    if arr.dtype.kind in "iu" or arr.dtype == np.bool_:
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64).reshape(-1))
    else:
        if arr.dtype not in _COMPLEX:                     # real input: complex of the same precision
            arr = arr.astype(np.complex64 if arr.dtype == np.float32 else np.complex128)
        t = torch.from_numpy(np.ascontiguousarray(arr).reshape(-1))
    if t.numel() == 0:
        t = torch.zeros(1, dtype=t.dtype)
    return DeviceArray(t.to(RT.dev), arr.shape)


def _unary(a: DeviceArray, op: int) -> DeviceArray:
    out = RT.empty(a.shape, a.t.dtype)
    RT.lib.t_unary(a.prec, op, a.size, a.ptr(), out.ptr(), RT.stream())
    return out


def _same_precision(a: DeviceArray, b: DeviceArray):
    if a.t.dtype == b.t.dtype:
        return a, b
    # mixed precisions (a complex128 constant against complex64 data): the result has the higher one, like numpy
    hi = torch.complex128
    conv = lambda x: x if x.t.dtype == hi else _upload(x.host().astype(np.complex128))
    return conv(a), conv(b)


def _binary(a: DeviceArray, b: DeviceArray, op: int) -> DeviceArray:
    if a.is_index or b.is_index:
        raise TypeError("arithmetic on index tensors is not part of the backend interface")
    a, b = _same_precision(a, b)
    rank = max(len(a.shape), len(b.shape))
    sha = (1,) * (rank - len(a.shape)) + a.shape
    shb = (1,) * (rank - len(b.shape)) + b.shape
    shape = []
    for x, y in zip(sha, shb):
        if x != y and x != 1 and y != 1:
            raise ValueError(f"operands could not be broadcast together with shapes {a.shape} {b.shape}")
        shape.append(max(x, y) if min(x, y) > 0 else 0)
    sa = [0 if d == 1 else s for d, s in zip(sha, _dense_strides(sha))]
    sb = [0 if d == 1 else s for d, s in zip(shb, _dense_strides(shb))]
    out = RT.empty(shape, a.t.dtype)
    out.real = a.real and b.real
    RT.lib.t_binary(a.prec, op, rank, _ll(shape), _ll(sa), _ll(sb), a.ptr(), b.ptr(), out.ptr(), RT.stream())
    return out


def _copy(src: DeviceArray, shape, strides_in, offset_in: int = 0, dst: DeviceArray | None = None, strides_out=None,
          offset_out: int = 0) -> DeviceArray:
    if dst is None:
        dst = RT.empty(shape, src.t.dtype)
        dst.real = src.real
    so = strides_out if strides_out is not None else _dense_strides(shape)
    RT.lib.t_copy(src.esize, len(shape), _ll(shape), _ll(strides_in), _ll(so), src.ptr(offset_in), dst.ptr(offset_out),
                  RT.stream())
    return dst


def make_backend_class(tensor_abc):
    """``B200Backend(tensor_abc)``: the class a bqa maintainer registers as ``BACKEND_STR_TO_BACKEND["b200"]``."""

    class B200Backend(tensor_abc):

        def __init__(self, raw: DeviceArray):
            if not isinstance(raw, DeviceArray):
                raise TypeError(f"B200Backend wraps a DeviceArray, got {type(raw).__name__}")   # like backends.py:780
            self._tensor = raw

        # ---- constructors (backends.py:36-58) ----------------------------------------------------------------
        @classmethod
        def make_from_list(cls, lst: list):
            dtype = np.intp if lst and isinstance(lst[0], int) else _default_complex()
            return cls(_upload(np.array(lst, dtype=dtype)))

        @classmethod
        def make_from_numpy(cls, arr):
            return cls(_upload(arr))

        @classmethod
        def make_from_raw_tensor(cls, raw_tensor):
            return cls(raw_tensor)

        @classmethod
        def make_constant(cls, const):
            return cls(_upload(np.array(const, dtype=_default_complex())))

        @staticmethod
        def make_empty_raw_tensor(batch_size: int, shape):
            tdt = _COMPLEX[_default_complex()][0]
            return RT.empty((batch_size, *shape), tdt)

        # ---- getters ------------------------------------------------------------------------------------------
        @property
        def raw_tensor(self):
            return self._tensor

        @property
        def numpy(self):
            return self._tensor.host()

        @property
        def raw_shape(self):
            return self._tensor.shape

        # ---- element-wise -------------------------------------------------------------------------------------
        @staticmethod
        def inv_raw(raw):
            return _unary(raw, U_INV)

        @staticmethod
        def pinv_raw(raw):
            return _unary(raw, U_PINV)

        @staticmethod
        def sqrt_raw(raw):
            return _unary(raw, U_SQRT)

        @staticmethod
        def sin_raw(raw):
            return _unary(raw, U_SIN)

        @staticmethod
        def cos_raw(raw):
            return _unary(raw, U_COS)

        @staticmethod
        def conj_raw(raw):
            return _unary(raw, U_CONJ)

        @staticmethod
        def mul_raw(lhs, rhs):
            return _binary(lhs, rhs, B_MUL)

        @staticmethod
        def sum_raw(lhs, rhs):
            return _binary(lhs, rhs, B_ADD)

        @staticmethod
        def sub_raw(lhs, rhs):
            return _binary(lhs, rhs, B_SUB)

        @staticmethod
        def div_raw(lhs, rhs):
            return _binary(lhs, rhs, B_DIV)

        # ---- reshape / transpose ------------------------------------------------------------------------------
        @staticmethod
        def reshape_raw(raw, shape):
            shape = list(shape)
            if -1 in shape:
                known = prod(s for s in shape if s != -1)
                shape[shape.index(-1)] = raw.size // max(known, 1)
            if prod(shape) != raw.size:
                raise ValueError(f"cannot reshape array of size {raw.size} into shape {tuple(shape)}")
            return DeviceArray(raw.t, shape, raw.real)      # dense row-major: a view of the same buffer

        @staticmethod
        def transpose_raw(raw, index_order):
            st = _dense_strides(raw.shape)
            shape = [raw.shape[i] for i in index_order]
            return _copy(raw, shape, [st[i] for i in index_order])

        # ---- reductions ---------------------------------------------------------------------------------------
        @staticmethod
        def max_norm(raw, is_full: bool = True):
            if is_full:
                out = RT.empty((), raw.t.dtype)
                RT.lib.t_max_abs(raw.prec, raw.size, raw.ptr(), out.ptr(), RT.stream())
            else:
                inner = prod(raw.shape[1:])
                out = RT.empty(raw.shape[1:], raw.t.dtype)
                RT.lib.t_col_max(raw.prec, raw.shape[0], inner, raw.ptr(), out.ptr(), RT.stream())
            out.real = True
            return out

        @staticmethod
        def batched_l2_norm(raw):
            batch = raw.shape[0]
            out = RT.empty((batch,), raw.t.dtype)
            RT.lib.t_batch_reduce(raw.prec, 0, batch, prod(raw.shape[1:]), 0, raw.ptr(), out.ptr(), RT.stream())
            out.real = True
            return out

        @staticmethod
        def batched_trace(raw):
            n = raw.shape[-1]
            assert raw.shape[-2] == n, raw.shape
            lead = raw.shape[:-2]
            out = RT.empty(lead, raw.t.dtype)
            RT.lib.t_batch_reduce(raw.prec, 1, prod(lead), n * n, n, raw.ptr(), out.ptr(), RT.stream())
            return out

        def batched_diag(self):
            raw = self._tensor
            n = raw.shape[-1]
            out = RT.empty((*raw.shape, n), raw.t.dtype)
            RT.lib.t_diag(raw.prec, prod(raw.shape[:-1]), n, raw.ptr(), out.ptr(), RT.stream())
            return self.make_from_raw_tensor(out)

        # ---- batch axis ---------------------------------------------------------------------------------------
        @staticmethod
        def batched_gather(raw, indices):
            if not indices.is_index:
                raise TypeError("indices must be an index tensor")
            row = prod(raw.shape[1:])
            out = RT.empty((indices.size, *raw.shape[1:]), raw.t.dtype)
            out.real = raw.real
            RT.lib.t_rows(raw.esize, 0, indices.size, row, indices.ptr(), raw.ptr(), out.ptr(), RT.stream())
            return out

        @staticmethod
        def take_batch_slice(raw, start: int, end: int):
            start, end = max(start, 0), min(end, raw.shape[0])
            shape = (max(end - start, 0), *raw.shape[1:])
            st = _dense_strides(raw.shape)
            return _copy(raw, shape, st, offset_in=start * (st[0] if st else 1))

        @staticmethod
        def assign_at_batch_indices_raw(dst, src, indices):
            row = prod(dst.shape[1:])
            assert src.size == indices.size * row, (src.shape, indices.shape, dst.shape)
            RT.lib.t_rows(dst.esize, 1, indices.size, row, indices.ptr(), src.ptr(), dst.ptr(), RT.stream())
            return dst

        # ---- linear algebra -----------------------------------------------------------------------------------
        @staticmethod
        def batched_matmul(lhs, rhs):
            lhs, rhs = _same_precision(lhs, rhs)
            assert len(lhs.shape) == 3 and len(rhs.shape) == 3 and lhs.shape[0] == rhs.shape[0] \
                and lhs.shape[2] == rhs.shape[1], (lhs.shape, rhs.shape)
            b, m, k = lhs.shape
            n = rhs.shape[2]
            out = RT.empty((b, m, n), lhs.t.dtype)
            RT.lib.t_matmul(lhs.prec, b, m, k, n, lhs.ptr(), rhs.ptr(), out.ptr(), RT.stream())
            return out

        @staticmethod
        def batched_svd(raw, pinv_eps: float):
            assert len(raw.shape) == 3 and raw.shape[1] == raw.shape[2], f"square matrices only, got {raw.shape}"
            b, n, _ = raw.shape
            u, s, vh = RT.empty((b, n, n), raw.t.dtype), RT.empty((b, n), raw.t.dtype), RT.empty((b, n, n), raw.t.dtype)
            nbytes = RT.lib.svd_scratch_bytes(raw.prec, n)
            scratch = RT.svd_scratch(nbytes)
            RT.lib.t_svd(raw.prec, b, n, raw.ptr(), u.ptr(), s.ptr(), vh.ptr(), float(pinv_eps), scratch.data_ptr(), nbytes,
                         RT.stream())
            return u, s, vh

        # ---- physics ------------------------------------------------------------------------------------------
        @staticmethod
        def apply_x_to_phys_dim_raw(raw):
            st = _dense_strides(raw.shape)                   # raw[:, ::-1]: negative stride on the physical axis
            sin = list(st)
            sin[1] = -st[1]
            return _copy(raw, raw.shape, sin, offset_in=(raw.shape[1] - 1) * st[1])

        @staticmethod
        def apply_z_to_phys_dim_raw(raw):
            sign = _upload(np.array([1.0, -1.0], dtype=np.complex64 if raw.prec == _lib.C64 else np.complex128)
                           .reshape((1, 2) + (1,) * (len(raw.shape) - 2)))
            return _binary(raw, sign, B_MUL)

        @staticmethod
        def measure_raw_tensor_by_position_in_place(raw, position: int, outcome) -> None:
            half = prod(raw.shape[2:])
            per_node = 2 * half
            RT.lib.t_fill(raw.prec, half, raw.ptr(position * per_node + (1 - int(outcome)) * half), 0.0, 0.0, RT.stream())
            nrm = RT.empty((1,), raw.t.dtype)                # Frobenius norm of the WHOLE batch (backends.py:729-734)
            RT.lib.t_batch_reduce(raw.prec, 0, 1, raw.size, 0, raw.ptr(), nrm.ptr(), RT.stream())
            RT.lib.t_binary(raw.prec, B_DIV, 1, _ll([raw.size]), _ll([1]), _ll([0]), raw.ptr(), nrm.ptr(), raw.ptr(),
                            RT.stream())

        # ---- other --------------------------------------------------------------------------------------------
        @staticmethod
        def concatenate(lhs, rhs, axis: int):
            lhs, rhs = _same_precision(lhs, rhs)
            axis = axis % len(lhs.shape)
            shape = list(lhs.shape)
            shape[axis] += rhs.shape[axis]
            out = RT.empty(shape, lhs.t.dtype)
            so = _dense_strides(shape)
            _copy(lhs, lhs.shape, _dense_strides(lhs.shape), dst=out, strides_out=so)
            _copy(rhs, rhs.shape, _dense_strides(rhs.shape), dst=out, strides_out=so, offset_out=lhs.shape[axis] * so[axis])
            return out

        @staticmethod
        def truncate_raw_tensor(raw, dims):
            shape = [min(int(d), s) for d, s in zip(dims, raw.shape)]
            return _copy(raw, shape, _dense_strides(raw.shape))

        @staticmethod
        def compute_minimal_rank_from_raw_lmbds(lmbds, eps: float) -> int:
            # returns a host integer by contract (backends.py:753-754); not on the annealing path (state.py:160-168 is unused)
            return int(np.max((lmbds.host().real > eps).sum(-1)))

        @staticmethod
        def make_inplace_damping_update_raw(dst, src, alpha: float) -> None:
            RT.lib.t_axpby(dst.prec, dst.size, dst.ptr(), src.ptr(), float(alpha), 1.0 - float(alpha), RT.stream())

        # ---- fused composites ---------------------------------------------------------------------------------
        def _node_batch(self, msgs):
            """(degree, D, B, concatenated aligned messages (d B, D, D), identity positions (d, B))"""
            raw = self._tensor
            d = len(msgs)
            B, D = raw.shape[0], raw.shape[-1] if d else 1
            assert raw.shape == (B, 2) + (D,) * d, raw.shape
            cat = RT.empty((d * B, D, D), raw.t.dtype)
            for j, m in enumerate(msgs):
                mr = m.raw_tensor
                assert mr.shape == (B, D, D), (mr.shape, raw.shape)
                _copy(mr, mr.shape, _dense_strides(mr.shape), dst=cat, strides_out=_dense_strides(mr.shape), offset_out=j * B * D * D)
            pos = torch.arange(d * B, dtype=torch.int32, device=RT.dev)
            return d, D, B, cat, pos

        def pass_msgs(self, msgs, evolution_times=None):
            """All d outgoing messages of the degree class in ONE kernel (reference: d (d - 1) leg contractions through
            transposes, reshapes and batched matmuls, backends.py:381-408; with ``evolution_times`` the ZZ half gate on
            the output leg, :393-404, :519-526)."""
            raw = self._tensor
            lib = RT.lib
            d, D, B, cat, pos = self._node_batch(msgs)
            if d == 0 or D > 16 or d > 8:
                return super().pass_msgs(msgs, evolution_times)
            ws = torch.empty(max(lib.workspace_bytes(raw.prec, d, D, D), 16), dtype=torch.uint8, device=RT.dev)
            st = RT.stream()
            if evolution_times is None:
                out = RT.empty((d * B, D, D), raw.t.dtype)
                rdt = torch.float32 if raw.prec == _lib.C64 else torch.float64
                resid = torch.zeros(2, dtype=rdt, device=RT.dev)
                status = torch.zeros(4, dtype=torch.int32, device=RT.dev)
                lib.bp_sweep(raw.prec, d, D, B, raw.ptr(), cat.ptr(), out.ptr(), pos.data_ptr(), pos.data_ptr(), 0.0, 1, 0.0, 0,
                             resid.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), st)
                n = D
            else:
                out = RT.empty((d * B, 2 * D, 2 * D), raw.t.dtype)
                # couplings arrive as complex tensors holding real angles (backends.py:583): their real parts become
                # the (d, B) real array the kernel reads (a stride-2 copy in units of the real type, on the device)
                rdt = torch.float32 if raw.prec == _lib.C64 else torch.float64
                ea = torch.empty(d * B, dtype=rdt, device=RT.dev)
                for j, t in enumerate(evolution_times):
                    tr = t.raw_tensor
                    assert tr.size == B and tr.prec == raw.prec, (tr.shape, B)
                    lib.t_copy(ea.element_size(), 1, _ll([B]), _ll([2]), _ll([1]), tr.ptr(),
                               ea.data_ptr() + j * B * ea.element_size(), st)
                lib.ext_msgs(raw.prec, d, D, B, raw.ptr(), cat.ptr(), out.ptr(), pos.data_ptr(), pos.data_ptr(), ea.data_ptr(),
                             1.0, ws.data_ptr(), ws.numel(), st)
                n = 2 * D
            step = B * n * n
            return tuple(self.make_from_raw_tensor(DeviceArray(out.t[j * step:(j + 1) * step], (B, n, n))) for j in range(d))

        def get_density_matrices(self, msgs):
            """Single-qubit marginals of the degree class in one kernel (backends.py:440-448)."""
            raw = self._tensor
            lib = RT.lib
            d, D, B, cat, pos = self._node_batch(msgs)
            if d == 0 or D > 16 or d > 8:
                return super().get_density_matrices(msgs)
            rdt = torch.float32 if raw.prec == _lib.C64 else torch.float64
            bloch = torch.zeros(B * 4, dtype=rdt, device=RT.dev)
            ids = torch.arange(B, dtype=torch.int32, device=RT.dev)
            ws = torch.empty(max(lib.workspace_bytes(raw.prec, d, D, D), 16), dtype=torch.uint8, device=RT.dev)
            lib.density(raw.prec, d, D, B, raw.ptr(), cat.ptr(), pos.data_ptr(), ids.data_ptr(), bloch.data_ptr(),
                        ws.data_ptr(), ws.numel(), RT.stream())
            rho = RT.empty((B, 2, 2), raw.t.dtype)
            lib.t_bloch_to_rho(raw.prec, B, bloch.data_ptr(), rho.ptr(), RT.stream())
            return self.make_from_raw_tensor(rho)

    B200Backend.__name__ = "B200Backend"
    B200Backend.__qualname__ = "B200Backend"
    return B200Backend
