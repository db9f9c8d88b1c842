"""Device-resident annealing engine: the B200 replacement of the reference "VM" (src/bqa/state.py).

It consumes the same compiled ``Context`` (ours from ``bqa_b200.config`` or bqa's own from
``bqa.config.core.config_to_context``) and keeps the reference control flow:

    run_layer  = simple update -> Rz layer -> Rx layer -> symmetric gauge -> BP      (state.py:315-321)
    run_bp     = damped BP with the reference termination rules                      (state.py:97-124)
    measure    = sequential decimation sampler with the host numpy RNG stream        (state.py:250-312)
    bloch_vectors / density_matrices                                                 (state.py:77-94)

All state (node tensors per degree class, messages, lambdas) lives in HBM; every numerical stage is a
CUDA kernel reached through the C ABI (include/bqa_b200.h).  PyTorch is used only to own device memory and
streams.  Host synchronisations per annealing step: one (the global bond-dimension decision, the
reference's ``truncate_lmbds`` host round trip) plus one per *chunk* of BP sweeps instead of one per sweep:
sweep kernels test the previous sweep's residual on the device and turn into no-ops after convergence.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

log = logging.getLogger(__name__)

MAX_BOND_DIM = 16
MAX_DEGREE = 8


def default_precision() -> str:
    """``BQA_PRECISION`` = single | double like the reference (src/bqa/utils.py:9-20); unset means single
    (complex64, what the reference's GPU backend always uses, backends.py:772)."""
    p = (os.environ.get("BQA_PRECISION") or "single").lower()
    if p not in ("single", "double"):
        raise ValueError(f"Unknown value of BQA_PRECISION environment variable {p}")
    return p


def _on_device(method):
    """Runs a public Engine method with the engine's GPU as the current CUDA device: the raw kernel launches behind the
    C ABI (cooperative launches, function attributes, the SM count) act on cudaGetDevice(), not on the device of the
    pointers they are given."""
    import functools

    @functools.wraps(method)
    def call(self, *args, **kwargs):
        if not self.cuda:
            return method(self, *args, **kwargs)
        with torch.cuda.device(self.dev):
            return method(self, *args, **kwargs)
    return call


def _np(x) -> np.ndarray:
    """numpy view of a layout field: ours are arrays, bqa's are Tensor wrappers (.numpy) or lists of them."""
    if isinstance(x, (list, tuple)):
        return np.stack([_np(e) for e in x]) if len(x) else np.zeros((0, 0))
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.numpy
    return np.asarray(x)


@dataclass
class _DegreeClass:
    degree: int
    B: int
    node_ids: torch.Tensor
    in_pos: torch.Tensor
    out_pos: torch.Tensor
    lmbd_pos: torch.Tensor
    remote_pos: torch.Tensor     # peer-memory path only: (peer << 27 | slot on that peer) or -1, like out_pos
    node_ampls: torch.Tensor
    edge_ampls: torch.Tensor
    T: list            # two flat complex buffers (ping-pong across bond-dimension changes)
    cur: int = 0
    node_ids_host: np.ndarray = None


class Engine:
    def __init__(self, context, precision: str | None = None, device=None, _testing_lib=None):
        self.ctx = context
        self.precision = precision or default_precision()
        if self.precision not in ("single", "double"):
            raise ValueError(f"precision must be 'single' or 'double', got {self.precision}")
        if _testing_lib is not None:                       # tests only: host emulation of the kernels
            self.lib = _testing_lib
            self.dev = torch.device("cpu")
        else:
            self.lib = _lib.load_library()
            if not torch.cuda.is_available():
                raise RuntimeError("bqa_b200 needs a CUDA device (no CPU fallback)")
            self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            if self.dev.type != "cuda":
                raise RuntimeError(f"bqa_b200 runs on CUDA devices only, got {self.dev}")
        self.cuda = self.dev.type == "cuda"
        self.prec = _lib.C64 if self.precision == "single" else _lib.C128
        self.cdtype = torch.complex64 if self.precision == "single" else torch.complex128
        self.rdtype = torch.float32 if self.precision == "single" else torch.float64
        self.np_rdtype = np.float32 if self.precision == "single" else np.float64

        ctx = context
        self.N = int(ctx.nodes_number)
        self.E2 = int(ctx.edges_number)
        self.L = self.E2 // 2
        self.Dmax = int(ctx.max_bond_dim)
        if self.Dmax > MAX_BOND_DIM:
            raise ValueError(f"max_bond_dim {self.Dmax} exceeds the supported maximum {MAX_BOND_DIM}")
        self.max_iters = int(ctx.max_bp_iters_number)
        self.bp_eps = float(ctx.bp_eps)
        self.pinv_eps = float(ctx.pinv_eps)
        self.damping = float(ctx.damping)
        self.threshold = float(ctx.measurement_threshold)
        self.rng = np.random.default_rng(int(ctx.seed))
        self.D = 1
        self.stats = {"bp_sweeps": [], "bp_dist": [], "bond_dims": [], "trunc_err": []}
        self.host_reads = 0          # blocking device -> host reads so far (control block, column maxima, results)
        self._bp_chunk = 4
        self._single_launch_ok = self.cuda and os.environ.get("BQA_B200_SINGLE_LAUNCH_BP", "1") != "0"
        self._no_bp_run = {}         # bond dimension -> True once bqa_b200_bp_run reported "no kernel for this shape"
        # all degree classes in one launch (bqa_multiclass.cuh): single GPU, CUDA library only
        self._multiclass = (self.cuda and hasattr(self.lib, "bp_run_classes")
                            and os.environ.get("BQA_B200_MULTICLASS", "1") != "0")

        self.classes: list[_DegreeClass] = []
        self._node_class = np.zeros(self.N, np.int64)
        self._node_slot = np.zeros(self.N, np.int64)
        for ci, (degree, lay) in enumerate(ctx.degree_to_layout.items()):
            degree = int(degree)
            if degree > MAX_DEGREE:
                raise ValueError(f"node degree {degree} exceeds the supported maximum {MAX_DEGREE}")
            ids = _np(lay.node_ids).astype(np.int64)
            B = int(ids.shape[0])
            self._node_class[ids] = ci
            self._node_slot[ids] = np.arange(B)
            elems = B * 2 * self.Dmax ** degree
            if elems * (8 if self.precision == "single" else 16) > 64 * 2 ** 30:
                raise MemoryError(f"degree class {degree} with max_bond_dim {self.Dmax} needs more than 64 GiB")
            i32 = lambda a: torch.from_numpy(np.ascontiguousarray(_np(a).reshape(degree, B).astype(np.int32))).to(self.dev)
            real = lambda a, shp: torch.from_numpy(
                np.ascontiguousarray(np.real(_np(a)).reshape(shp).astype(self.np_rdtype))).to(self.dev)
            self.classes.append(_DegreeClass(
                degree=degree, B=B,
                node_ids=torch.from_numpy(ids.astype(np.int32)).to(self.dev),
                in_pos=i32(lay.input_msgs_position), out_pos=i32(lay.output_msgs_position),
                lmbd_pos=i32(lay.lmbds_position),
                remote_pos=i32(getattr(lay, "remote_msgs_position", None)) if getattr(lay, "remote_msgs_position", None) is not None
                else torch.full((max(degree, 1), B), -1, dtype=torch.int32, device=self.dev),
                node_ampls=real(lay.node_ampls, (B,)), edge_ampls=real(lay.edge_ampls, (degree, B)),
                T=[torch.zeros(elems, dtype=self.cdtype, device=self.dev) for _ in range(2)],
                node_ids_host=ids))
        Dm = self.Dmax
        # message buffers the BP sweeps rotate through: 2 (read one, write the other); 3 in the partitioned engine's
        # peer-memory mode, where the convergence test lags one sweep (bqa_b200_bp_run)
        self._nbuf = self._n_msg_buffers()
        self._msgs = [self._alloc_shared(self.E2 * Dm * Dm, self.cdtype, f"msgs{i}") for i in range(self._nbuf)]
        self._msgs_cur = 0
        self._ext = None         # allocated on first simple update (size depends on the largest D reached)
        self._canon = None
        self._lmbds = self._alloc_shared(self.L * 2 * Dm, self.rdtype, "lmbds")
        self._lmbds.fill_(1.0)
        self._lmbd_stride = 2      # row stride of the lambda array = 2 * D of the update that produced it
        # control block, read back with ONE copy per chunk of sweeps (one per step in the steady state):
        # [resid (max_iters, 2) reals | status int32 x4 | column maxima of the lambdas (2 Dmax reals)]
        rsize = 4 if self.precision == "single" else 8
        rbytes = (((max(self.max_iters, 1) + 2) * 2 * rsize) + 15) // 16 * 16      # + 2 rows: counters of bqa_b200_bp_run
        cbytes = (2 * Dm * rsize + 15) // 16 * 16
        self._ctrl = self._alloc_shared(rbytes + 16 + cbytes, torch.uint8, "ctrl")
        self._ctrl_rbytes = rbytes
        # the part a BP run resets: residuals + status[0..2]; status[3] (a grid / peer barrier timed out) is sticky, so a
        # timeout in any exchange of the step is still there at the step's host read
        self._ctrl_bp = self._ctrl[: rbytes + 12]
        self._resid = self._ctrl[:rbytes].view(self.rdtype)
        self._status = self._ctrl[rbytes: rbytes + 16].view(torch.int32)
        self._colmax = self._ctrl[rbytes + 16:].view(self.rdtype)[: 2 * Dm]
        self._ctrl_host = torch.empty(self._ctrl.numel(), dtype=torch.uint8, pin_memory=True) if self.cuda else None
        self._ctrl_last = None                               # host copy of the control block of the last read
        # Speculative truncation: at D == max_bond_dim the update keeps max_bond_dim columns unless the state's rank
        # collapses (backends.py:297-303), so the step is enqueued without waiting for the column maxima and they are
        # checked with the BP control block, in the step's only host read; a wrong guess redoes the update.
        self.speculate = os.environ.get("BQA_B200_SPECULATE", "1") != "0"
        # extended messages of the next step enqueued behind the BP run, before the host reads the run's outcome
        # (run_layer(next_ztime=...) / run_layers): the GPU never waits for the host between two steps
        self.ext_ahead = os.environ.get("BQA_B200_EXT_AHEAD", "1") != "0"
        self._next_ztime = None                 # lookahead of the run_layer call in progress
        self._ext_done_for = None               # (ztime, D) of extended messages already sitting in self._ext
        self._side_stream = torch.cuda.Stream(self.dev) if self.cuda else None
        self._bp_done_event = torch.cuda.Event() if self.cuda else None
        self._read_after_event = False
        # edge order of the n = 8 canonicalizer kernel (bqa_b200_canonicalize_ordered).  OFF by default: measured on the
        # 100k benchmark, regrouping the edges by last step's Jacobi sweep counts lowers the sweeps a warp runs only from
        # 5.34 to 5.19 (every 8 steps) / 5.09 (every step): the slow matrices of one step are not the slow ones of the
        # next, and the counting sort costs more than it saves (profiles/r2_canon_experiments.md)
        self._canon_order = self._canon_cost = None
        self._canon_resort_every = int(os.environ.get("BQA_B200_CANON_RESORT", "0"))
        self._canon_age = 1                                   # the first regrouping needs one step of recorded costs
        if self.cuda and self._canon_resort_every > 0 and hasattr(self.lib, "canonicalize_ordered") and self.L > 0:
            self._canon_order = torch.arange(self.L, dtype=torch.int32, device=self.dev)
            self._canon_cost = torch.zeros(self.L, dtype=torch.uint8, device=self.dev)
        self._bloch = torch.zeros(self.N * 4, dtype=self.rdtype, device=self.dev)
        self._ws = torch.zeros(16, dtype=torch.uint8, device=self.dev)
        # candidate of a sampling pass: (node, unmeasured count) int32 x2 | p0 real -- one buffer, one host read per pass
        self._argmax = torch.zeros(16, dtype=torch.uint8, device=self.dev)
        self._argmax_i = self._argmax[:8].view(torch.int32)
        self._argmax_p = self._argmax[8:8 + (4 if self.precision == "single" else 8)].view(self.rdtype)
        self._nproj = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._init_state()

    # ------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------
    def _n_msg_buffers(self) -> int:
        return 2

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream if self.cuda else 0

    def _to_host(self, t: torch.Tensor) -> np.ndarray:
        self.host_reads += 1
        return t.cpu().numpy()          # synchronises the current stream

    def _read_ctrl(self) -> np.ndarray:
        """Host copy of the whole control block (page-locked staging buffer: one asynchronous copy + stream sync)."""
        if self._ctrl_host is None:
            self._read_after_event = False
            ctrl = self._to_host(self._ctrl)
        elif self._read_after_event:
            # work of the next step is already queued behind the BP run: copy on a side stream that waits for the run only
            self._read_after_event = False
            self.host_reads += 1
            self._side_stream.wait_event(self._bp_done_event)
            with torch.cuda.stream(self._side_stream):
                self._ctrl_host.copy_(self._ctrl, non_blocking=True)
            self._side_stream.synchronize()
            ctrl = self._ctrl_host.numpy().copy()
        else:
            self.host_reads += 1
            self._ctrl_host.copy_(self._ctrl, non_blocking=True)
            torch.cuda.current_stream(self.dev).synchronize()
            ctrl = self._ctrl_host.numpy().copy()
        self._ctrl_last = ctrl
        return ctrl

    def _ensure_ws(self, D: int, Dn: int) -> None:
        need = max([self.lib.workspace_bytes(self.prec, c.degree, D, Dn) for c in self.classes] + [16])
        if self._ws.numel() < need:
            self._ws = torch.zeros(need, dtype=torch.uint8, device=self.dev)

    def _alloc_shared(self, numel: int, dtype, tag: str) -> torch.Tensor:
        """Buffers other GPUs write into on the peer-memory path (messages, extended messages, BP control
        block); plain device memory here, symmetric memory in PartitionedEngine."""
        return torch.zeros(numel, dtype=dtype, device=self.dev)

    def _ensure_edge_buffers(self, D: int) -> None:
        need = self.E2 * 4 * D * D
        if self._ext is None or self._ext.numel() < need:
            self._ext = self._alloc_shared(need, self.cdtype, "ext")
        if self._canon is None or self._canon.numel() < need:
            self._canon = self._alloc_shared(need, self.cdtype, "canon")

    @property
    def ctrl_bytes_per_bp_read(self) -> int:
        """bytes of one device -> host read of the BP control block (one per chunk of sweeps)"""
        return int(self._ctrl.numel())

    @property
    def colmax_bytes(self) -> int:
        """bytes of the per-step device -> host read behind the bond-dimension decision"""
        return int(self._colmax.numel() * self._colmax.element_size())

    @property
    def msgs_buffer(self) -> torch.Tensor:
        return self._msgs[self._msgs_cur]

    # ------------------------------------------------------------------------------------------
    # state
    # ------------------------------------------------------------------------------------------
    def _init_state(self) -> None:
        """|-> on every qubit, bond dimension 1, lambdas 1, messages 1 (state.py:21, :41-74)."""
        self.D = 1
        amp = np.sqrt(0.5)
        for c in self.classes:
            t = c.T[0][: c.B * 2].view(c.B, 2)
            t[:, 0] = amp
            t[:, 1] = -amp
            c.cur = 0
        self._msgs_cur = 0
        self._msgs[0][: self.E2] = 1.0
        self._lmbds[:] = 1.0
        self._lmbd_stride = 2

    def tensors_numpy(self) -> dict:
        """{degree: (B, 2, D, ..., D)} copies on the host."""
        D = self.D
        return {c.degree: self._to_host(c.T[c.cur][: c.B * 2 * D ** c.degree]).reshape((c.B, 2) + (D,) * c.degree)
                for c in self.classes}

    def msgs_numpy(self) -> np.ndarray:
        D = self.D
        return self._to_host(self.msgs_buffer[: self.E2 * D * D]).reshape(self.E2, D, D)

    def lmbds_numpy(self) -> np.ndarray:
        s = self._lmbd_stride
        return self._to_host(self._lmbds[: self.L * s]).reshape(self.L, s)[:, : self.D].copy()

    @_on_device
    def state_to_host(self, pinned: bool = False) -> dict:
        """Checkpoint of the run-time state (the reference has none).  With ``pinned`` the
        arrays are views of page-locked host buffers (listed under "_pinned"), which ``load_state`` uploads
        with asynchronous copies."""
        snap = {"D": self.D, "tensors": self.tensors_numpy(), "msgs": self.msgs_numpy(), "lmbds": self.lmbds_numpy()}
        if pinned and self.cuda:
            keep = {}

            def pin(key, a):
                t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                keep[key] = t
                return t.numpy()
            snap["tensors"] = {d: pin(("tensors", d), t) for d, t in snap["tensors"].items()}
            snap["msgs"] = pin("msgs", snap["msgs"])
            snap["lmbds"] = pin("lmbds", snap["lmbds"])
            snap["_pinned"] = keep          # page-locked torch tensors behind the numpy views above
        return snap

    # checkpoint hooks of core.save_checkpoint / load_checkpoint (the partitioned engine writes one file per rank)
    def checkpoint_file(self, path: str) -> str:
        return path

    def checkpoint_barrier(self) -> None:
        """All ranks have written their temporary file (single process: nothing to wait for)."""

    def checkpoint_agree(self, value: int, what: str) -> int:
        """The value every rank holds (single process: the value); raises when the ranks disagree."""
        return int(value)

    @_on_device
    def load_state(self, snap: dict) -> None:
        """Uploads a checkpoint / an oracle state: tensors {degree: (B, 2, D..)}, msgs (2L, D, D), lmbds (L, D)."""
        D = int(snap["D"])
        if not 1 <= D <= self.Dmax:
            raise ValueError(f"bond dimension {D} outside [1, {self.Dmax}]")
        self._ext_done_for = None
        np_c = np.complex64 if self.precision == "single" else np.complex128
        pinned = snap.get("_pinned", {})

        def host(key, arr):
            """flat host tensor in the working dtype; the page-locked original when there is one (asynchronous copy)"""
            t = pinned.get(key)
            if t is not None and t.dtype == self.cdtype:
                return t.view(-1)
            return torch.from_numpy(np.ascontiguousarray(np.asarray(arr), dtype=np_c).reshape(-1))
        for c in self.classes:
            t = host(("tensors", c.degree), snap["tensors"][c.degree])
            if t.shape[0] != c.B * 2 * D ** c.degree:
                raise ValueError(f"degree-{c.degree} tensor batch has {t.shape[0]} elements, expected {c.B * 2 * D ** c.degree}")
            c.cur = 0
            c.T[0][: t.shape[0]].copy_(t, non_blocking=True)
        m = host("msgs", snap["msgs"])
        if m.shape[0] != self.E2 * D * D:
            raise ValueError(f"message array has {m.shape[0]} elements, expected {self.E2 * D * D}")
        if np.asarray(snap["lmbds"]).shape != (self.L, D):
            raise ValueError(f"lambda array has shape {np.asarray(snap['lmbds']).shape}, expected {(self.L, D)}")
        self._msgs_cur = 0
        self._msgs[0][: m.shape[0]].copy_(m, non_blocking=True)
        lm = np.zeros((self.L, 2 * D), self.np_rdtype)
        lm[:, :D] = np.real(np.asarray(snap["lmbds"]))
        self._lmbd_stride = 2 * D
        self._lmbds[: lm.size].copy_(torch.from_numpy(lm.reshape(-1)))
        self.D = D

    # ------------------------------------------------------------------------------------------
    # BP  (state.py:97-124)
    # ------------------------------------------------------------------------------------------
    def _class_table(self, out_side: bool = False):
        """(ctypes array of bqa_b200_class, n): the degree classes as the *_classes entry points take them; with
        ``out_side`` the tensors' other ping-pong buffer is the output (apply update)"""
        rows = (_lib.ClassDesc * len(self.classes))()
        for r, c in zip(rows, self.classes):
            r.degree, r.B = c.degree, c.B
            r.T_in = c.T[c.cur].data_ptr()
            r.T_out = c.T[1 - c.cur].data_ptr() if out_side else None
            r.in_pos, r.out_pos, r.lmbd_pos = c.in_pos.data_ptr(), c.out_pos.data_ptr(), c.lmbd_pos.data_ptr()
            r.node_ampls, r.edge_ampls = c.node_ampls.data_ptr(), c.edge_ampls.data_ptr()
        return rows, len(self.classes)

    def _use_multiclass(self) -> bool:
        """Table-driven launches over all degree classes (one GPU): always for shapes without a specialised kernel;
        when a class has one (degree 3, D = 4, complex64), only while that class is small enough for the launches
        saved to outweigh the faster kernel."""
        if not self._multiclass or self._peer_targets("msgs") is not None:
            return False
        fast = [c for c in self.classes if self.precision == "single" and c.degree == 3 and self.D == 4 and c.B >= 4]
        if not fast:
            return True
        return len([c for c in self.classes if c.B > 0]) > 1 and max(c.B for c in fast) <= 8192

    def _enqueue_sweep(self, it: int, write_undamped: bool) -> None:
        D = self.D
        nb = self._nbuf
        cur = self._msgs[(self._msgs_cur + it) % nb]
        nxt = self._msgs[(self._msgs_cur + it + 1) % nb]
        st = self._stream()
        peers = self._peer_targets("msgs", (self._msgs_cur + it + 1) % nb)
        for c in self.classes:
            if c.degree == 0:
                continue
            args = (self.prec, c.degree, D, c.B, c.T[c.cur].data_ptr(), cur.data_ptr(), nxt.data_ptr(),
                    c.in_pos.data_ptr(), c.out_pos.data_ptr(), self.damping, int(write_undamped),
                    self.bp_eps, it, self._resid.data_ptr(), self._status.data_ptr(),
                    self._ws.data_ptr(), self._ws.numel())
            if peers is None:
                self.lib.bp_sweep(*args, st)
            else:
                self.lib.bp_sweep_p2p(*args, c.remote_pos.data_ptr(), peers, st)
        self._after_sweep(it, nxt)

    def _after_sweep(self, it: int, nxt: torch.Tensor) -> None:
        """Hook for the partitioned engine: halo exchange + residual all-reduce (no-op on one GPU)."""

    def _peer_targets(self, what: str, parity: int = 0):
        """Peer-memory path of the partitioned engine: host array of the peers' base pointers of the array the
        kernel is about to write (``None`` = everything stays on this GPU)."""
        return None

    def _before_bp(self) -> None:
        """Hook for the partitioned engine (orders the control-block reset against the peers' residual pushes)."""

    def _bp_run_peers(self):
        """(rank, world, peers0, peers1, peer_resid, peer_flags, seq_base, peers2, boundary nodes) of the single-launch BP
        run; one GPU here."""
        return 0, 1, None, None, None, None, 0, None, 0

    def _bp_run_done(self, sweeps: int) -> None:
        """Hook for the partitioned engine (advances the cross-GPU barrier sequence)."""

    def _single_launch_applies(self) -> bool:
        """True when run_bp will take the one-launch kernel of the headline shape (bqa_b200_bp_run)."""
        active = [c for c in self.classes if c.degree > 0 and c.B > 0]
        return (self._single_launch_ok and len(active) == 1 and not self._no_bp_run.get(self.D)
                and self.precision == "single" and active[0].degree == 3 and self.D == 4 and active[0].B >= 4)

    def _try_single_launch_bp(self):
        """The whole BP run in one cooperative launch (bqa_b200_bp_run) when every BP-active node sits in one degree
        class with a specialised kernel; returns (converged, sweeps, resid) or None."""
        active = [c for c in self.classes if c.degree > 0 and c.B > 0]
        if not self._single_launch_ok:
            return None
        if active and self._use_multiclass():
            rows, n = self._class_table()
            self.lib.bp_run_classes(self.prec, n, C.byref(rows), self.D, self._msgs[0].data_ptr(), self._msgs[1].data_ptr(),
                                    self._msgs_cur, self.damping, self.bp_eps, self.max_iters, self._resid.data_ptr(),
                                    self._status.data_ptr(), self._ws.data_ptr(), self._ws.numel(), self._stream())
            return self._read_bp_run()
        if len(active) != 1 or self._no_bp_run.get(self.D):
            return None
        c = active[0]
        peers = self._bp_run_peers()
        if peers is None:
            return None
        rank, world, p0, p1, presid, pflags, seq, p2, n_boundary = peers
        m2 = self._msgs[2].data_ptr() if self._nbuf > 2 else None
        ok = self.lib.bp_run(self.prec, c.degree, self.D, c.B, c.T[c.cur].data_ptr(), self._msgs[0].data_ptr(),
                             self._msgs[1].data_ptr(), self._msgs_cur, c.in_pos.data_ptr(), c.out_pos.data_ptr(),
                             self.damping, self.bp_eps, self.max_iters, self._resid.data_ptr(), self._status.data_ptr(),
                             c.remote_pos.data_ptr(), p0, p1, rank, world, presid, pflags, seq, m2, p2, n_boundary,
                             self._stream())
        if not ok:
            self._no_bp_run[self.D] = True
            return None
        if self._next_ztime is not None:
            self._enqueue_ext_ahead(c, self._next_ztime)
        return self._read_bp_run()

    def _enqueue_ext_ahead(self, c, ztime: float) -> None:
        """Extended messages of the next step behind the BP run just launched (bqa_b200_ext_msgs_after_run: the kernel
        picks the message buffer from the run's status words); the control-block read then waits for the run only."""
        st = torch.cuda.current_stream(self.dev)
        self._bp_done_event.record(st)
        m2 = self._msgs[2].data_ptr() if self._nbuf > 2 else None
        ok = self.lib.ext_msgs_after_run(self.prec, c.degree, self.D, c.B, c.T[c.cur].data_ptr(), self._msgs[0].data_ptr(),
                                         self._msgs[1].data_ptr(), m2, self._nbuf, self._msgs_cur, self.max_iters,
                                         self._status.data_ptr(), self._ext.data_ptr(), c.in_pos.data_ptr(),
                                         c.out_pos.data_ptr(), c.edge_ampls.data_ptr(), float(ztime),
                                         c.remote_pos.data_ptr() if self._peer_targets("ext") is not None else None,
                                         self._peer_targets("ext"), st.cuda_stream)
        if ok:
            self._ext_done_for = (float(ztime), self.D)
            self._read_after_event = True

    def _read_bp_run(self):
        """(converged, sweeps, residuals) of a single-launch BP run: the step's one read of the control block"""
        ctrl = self._read_ctrl()
        resid = ctrl[: self._ctrl_rbytes].view(self.np_rdtype).reshape(-1, 2)
        status = ctrl[self._ctrl_rbytes: self._ctrl_rbytes + 16].view(np.int32)
        if status[3]:
            raise RuntimeError("BP run aborted: a grid or peer barrier timed out (bqa_b200_bp_run)")
        self._bp_run_done(int(status[1]))
        return bool(status[0]), int(status[1]), resid

    @_on_device
    def run_bp(self) -> int:
        max_it = self.max_iters
        assert max_it > 0, "max_bp_iter_number must be positive"      # reference: assert best_msgs is not None
        self._ext_done_for = None                                     # the messages change
        self._ensure_ws(self.D, self.D)
        self._ctrl_bp.zero_()
        self._before_bp()
        eps = self.np_rdtype(self.bp_eps)
        single = self._try_single_launch_bp()
        if single is not None:
            done, sweeps, resid = single
            return self._finish_bp(done, sweeps, resid, max_it)
        it = 0
        done = False
        sweeps = max_it
        while it < max_it and not done:
            n = min(self._bp_chunk, max_it - it)
            for _ in range(n):
                self._enqueue_sweep(it, write_undamped=(it == max_it - 1))
                it += 1
            ctrl = self._read_ctrl()                         # one D2H + sync per chunk of sweeps
            resid = ctrl[: self._ctrl_rbytes].view(self.np_rdtype).reshape(-1, 2)
            status = ctrl[self._ctrl_rbytes: self._ctrl_rbytes + 16].view(np.int32)
            if status[3]:
                raise RuntimeError("a peer GPU never reached the sweep barrier (bqa_b200_sweep_sync timed out)")
            if status[0]:                                    # a later sweep saw convergence on the device
                done, sweeps = True, int(status[1])
            else:                                            # the last enqueued sweep is tested here
                num, den = resid[it - 1]
                with np.errstate(divide="ignore", invalid="ignore"):
                    if np.sqrt(num / den) < eps:
                        done, sweeps = True, it
        return self._finish_bp(done, sweeps, resid, max_it)

    def _finish_bp(self, done: bool, sweeps: int, resid, max_it: int) -> int:
        num, den = resid[sweeps - 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            dist = float(np.sqrt(num / den))
        if not np.isfinite(dist):
            # the reference fails loudly here too: np.abs().max() propagates NaN, `dist < best_dist` never holds and
            # `assert best_msgs is not None` fires (state.py:113-123)
            raise FloatingPointError(f"BP residual is not finite ({dist}) after {sweeps} sweeps: the state holds NaN/Inf")
        if done:
            # converged: keep the *input* of the converging sweep (state.py:118-120)
            self._msgs_cur = (self._msgs_cur + sweeps - 1) % self._nbuf
            log.debug(f"BP algorithm completed after {sweeps - 1} iterations")
        else:
            # cap reached: the last, undamped sweep output becomes the state (state.py:122-124)
            self._msgs_cur = (self._msgs_cur + max_it) % self._nbuf
            log.warning(f"BP algorithm exceeds iterations limit set to {max_it}, obtained bp_eps {dist}")
        self.stats["bp_sweeps"].append(sweeps)
        self.stats["bp_dist"].append(dist)
        # next run: enqueue about as many sweeps as this one needed before the first status read
        self._bp_chunk = int(min(max(sweeps + 1, 2), 64))
        return sweeps

    # ------------------------------------------------------------------------------------------
    # one annealing step  (state.py:230-247, :142-156, :219-227, :315-321)
    # ------------------------------------------------------------------------------------------
    def _reduce_colmax(self, colmax: torch.Tensor) -> None:
        """Hook for the partitioned engine (all-reduce max); no-op on one GPU."""

    def run_layers(self, layers) -> None:
        """A sequence of annealing steps ({"xtime": .., "ztime": ..} dicts, the reference's instruction format): every
        step but the last one announces the next step's ztime to run_layer."""
        layers = list(layers)
        for k, ins in enumerate(layers):
            nxt = layers[k + 1]["ztime"] if k + 1 < len(layers) else None
            self.run_layer(ins["xtime"], ins["ztime"], next_ztime=nxt)

    @_on_device
    def run_layer(self, xtime: float, ztime: float, next_ztime: float | None = None) -> None:
        """One annealing step.  ``next_ztime``: ztime of the step that follows, when the caller knows it -- its extended
        messages are then enqueued behind this step's BP run before the host reads the run's outcome (same kernels,
        same inputs, same results; the step after that finds them ready)."""
        D = self.D
        st = self._stream()
        self._ensure_ws(D, min(2 * D, self.Dmax))
        self._ensure_edge_buffers(D)
        ahead, self._ext_done_for = self._ext_done_for, None
        if ahead == (float(ztime), D):
            return self._rest_of_layer(D, st, xtime, ztime, next_ztime)
        cur = self.msgs_buffer
        peers = self._peer_targets("ext")
        multi = self._use_multiclass()
        if multi:
            rows, n = self._class_table()
            self.lib.ext_msgs_classes(self.prec, n, C.byref(rows), D, cur.data_ptr(), self._ext.data_ptr(), float(ztime),
                                      self._ws.data_ptr(), self._ws.numel(), st)
        for c in self.classes:
            if c.degree == 0 or multi:
                continue
            args = (self.prec, c.degree, D, c.B, c.T[c.cur].data_ptr(), cur.data_ptr(),
                    self._ext.data_ptr(), c.in_pos.data_ptr(), c.out_pos.data_ptr(),
                    c.edge_ampls.data_ptr(), float(ztime), self._ws.data_ptr(), self._ws.numel())
            if peers is None:
                self.lib.ext_msgs(*args, st)
            else:
                self.lib.ext_msgs_p2p(*args, c.remote_pos.data_ptr(), peers, st)
        self._rest_of_layer(D, st, xtime, ztime, next_ztime)

    def _rest_of_layer(self, D: int, st: int, xtime: float, ztime: float, next_ztime) -> None:
        self._exchange_ext()
        self._colmax.zero_()
        self._canonicalize(D, st)
        self._lmbd_stride = 2 * D
        self._reduce_colmax(self._colmax)
        speculative = self.speculate and self.cuda and D == self.Dmax
        if speculative:
            Dn = self.Dmax                                                        # checked after the BP run, see below
        else:
            colmax = self._to_host(self._colmax)[: 2 * D].astype(np.float64)      # host sync: the bond dimension changes
            Dn, err = self._truncation(colmax, D)
        n_bp = len(self.stats["bp_sweeps"])
        # lookahead only where a mis-speculated bond dimension cannot occur after the fact: the check below redoes the step
        self._next_ztime = next_ztime if (speculative and self.ext_ahead and next_ztime is not None) else None
        try:
            self._apply_and_bp(D, Dn, xtime, ztime)
        finally:
            self._next_ztime = None
        if speculative:
            rb = self._ctrl_rbytes + 16
            colmax = self._ctrl_last[rb: rb + 2 * D * self._colmax.element_size()].view(self.np_rdtype).astype(np.float64)
            Dn_true, err = self._truncation(colmax, D)
            if Dn_true != Dn:                                                     # the rank collapsed: redo with the right one
                for c in self.classes:
                    c.cur = 1 - c.cur
                del self.stats["bp_sweeps"][n_bp:], self.stats["bp_dist"][n_bp:]
                self.D, Dn = D, Dn_true
                self._ext_done_for = None                                         # computed from the discarded state
                self._apply_and_bp(D, Dn, xtime, ztime)
        self.stats["bond_dims"].append(Dn)
        self.stats["trunc_err"].append(err)
        log.info(f"Truncation performed, per edge error upper bound: {err}")
        log.info(f"Layer with ztime {ztime} and xtime {xtime} has been applied")

    def _canonicalize(self, D: int, st: int) -> None:
        """Canonicalizers and lambdas of every edge from the extended messages (state.py:171-200)."""
        if self._canon_order is not None and self.precision == "single" and D == 4:
            # n = 8 kernel: edges grouped by the Jacobi sweeps they needed (regrouped every few steps from the costs the
            # kernel records; a warp sweeps until its slowest matrix is done, results do not depend on the grouping)
            if self._canon_age >= self._canon_resort_every:
                self.lib.sort_edges_by_cost(self.L, self._canon_cost.data_ptr(), self._canon_order.data_ptr(), st)
                self._canon_age = 0
            self._canon_age += 1
            self.lib.canonicalize_ordered(self.prec, D, self.L, self._ext.data_ptr(), self._canon.data_ptr(),
                                          self._lmbds.data_ptr(), self._colmax.data_ptr(), self.pinv_eps,
                                          min(2 * D, self.Dmax), self._canon_order.data_ptr(),
                                          self._canon_cost.data_ptr(), st)
        else:
            self.lib.canonicalize(self.prec, D, self.L, self._ext.data_ptr(), self._canon.data_ptr(),
                                  self._lmbds.data_ptr(), self._colmax.data_ptr(), self.pinv_eps,
                                  min(2 * D, self.Dmax), st)

    def _truncation(self, colmax: np.ndarray, D: int):
        """(new bond dimension, per-edge error bound) from the column maxima of the lambdas (backends.py:297-303)."""
        rank = 2 * D - int(np.sum(colmax < self.pinv_eps))
        Dn = min(rank, self.Dmax)
        if Dn < 1:
            raise FloatingPointError("all singular values fell below pinv_eps: the state collapsed")
        return Dn, float(np.sqrt(np.sum(colmax[Dn:] ** 2)))

    def _apply_and_bp(self, D: int, Dn: int, xtime: float, ztime: float) -> None:
        """Truncated simple update D -> Dn of every class + Rz/Rx layers + symmetric gauge, then BP to convergence."""
        st = self._stream()
        msgs_out = self._msgs[0]
        multi = self._use_multiclass()
        if multi:
            rows, n = self._class_table(out_side=True)
            self.lib.apply_update_classes(self.prec, n, C.byref(rows), D, Dn, self._canon.data_ptr(), self._lmbds.data_ptr(),
                                          msgs_out.data_ptr(), float(ztime), float(xtime), self._ws.data_ptr(),
                                          self._ws.numel(), st)
        for c in self.classes:
            if multi:
                c.cur = 1 - c.cur
                continue
            self.lib.apply_update(self.prec, c.degree, D, Dn, c.B, c.T[c.cur].data_ptr(), c.T[1 - c.cur].data_ptr(),
                                  self._canon.data_ptr(), self._lmbds.data_ptr(), msgs_out.data_ptr(),
                                  c.in_pos.data_ptr(), c.out_pos.data_ptr(), c.lmbd_pos.data_ptr(),
                                  c.node_ampls.data_ptr(), c.edge_ampls.data_ptr(), float(ztime), float(xtime),
                                  self._ws.data_ptr(), self._ws.numel(), st)
            c.cur = 1 - c.cur
        self._msgs_cur = 0
        self.D = Dn
        log.debug(f"Layer of interaction gates with truncation has been applied, current bond dimension is {Dn}")
        self._after_update()
        self.run_bp()

    def _exchange_ext(self) -> None:
        """Hook for the partitioned engine (extended messages of cut edges); no-op on one GPU."""

    def _after_update(self) -> None:
        """Hook for the partitioned engine (halo slots of the re-initialised messages); no-op on one GPU."""

    # ------------------------------------------------------------------------------------------
    # marginals  (state.py:77-94, utils.py:23-27)
    # ------------------------------------------------------------------------------------------
    def _compute_bloch(self) -> None:
        D = self.D
        self._ensure_ws(D, D)
        st = self._stream()
        cur = self.msgs_buffer
        for c in self.classes:
            self.lib.density(self.prec, c.degree, D, c.B, c.T[c.cur].data_ptr(), cur.data_ptr(),
                             c.in_pos.data_ptr(), c.node_ids.data_ptr(), self._bloch.data_ptr(),
                             self._ws.data_ptr(), self._ws.numel(), st)

    @_on_device
    def bloch_vectors(self) -> np.ndarray:
        """(N, 3) array of (x, y, z) per qubit in node-id order."""
        self._compute_bloch()
        log.info("Density matrices have been computed")
        return self._to_host(self._bloch).reshape(self.N, 4)[:, :3].astype(np.float64)

    def density_matrices(self) -> np.ndarray:
        """(N, 2, 2) trace-normalised single-qubit density matrices (state.py:77-94)."""
        b = self.bloch_vectors()
        rho = np.empty((b.shape[0], 2, 2), np.complex128)       # global node count on a partitioned engine
        rho[:, 0, 0] = 0.5 * (1 + b[:, 2])
        rho[:, 1, 1] = 0.5 * (1 - b[:, 2])
        rho[:, 0, 1] = 0.5 * (b[:, 0] - 1j * b[:, 1])
        rho[:, 1, 0] = 0.5 * (b[:, 0] + 1j * b[:, 1])
        return rho

    # ------------------------------------------------------------------------------------------
    # sampling  (state.py:250-312)
    # ------------------------------------------------------------------------------------------
    @_on_device
    def measure(self) -> list:
        st = self._stream()
        outcomes = torch.zeros(self.N, dtype=torch.int32, device=self.dev)
        log.debug("Measurement outcomes sampling started")
        while True:
            self._compute_bloch()
            self.lib.argmax_unmeasured(self.prec, self.N, self._bloch.data_ptr(), outcomes.data_ptr(),
                                       self._argmax_i.data_ptr(), self._argmax_p.data_ptr(), st)
            cand = self._to_host(self._argmax)
            node, left = (int(v) for v in cand[:8].view(np.int32))
            if left == 0:
                break
            p0 = float(cand[8:8 + self._argmax_p.element_size()].view(self.np_rdtype)[0])
            u = self.rng.uniform(0.0, 1.0)                  # one draw per pass, same stream as the reference
            bit = 0 if p0 > u else 1
            log.debug(f"Node {node} the most determined (spin-up probability {p0} and spin-down probability {1 - p0}) and has been measured")
            c = self.classes[int(self._node_class[node])]
            self.lib.project_node(self.prec, c.degree, self.D, c.T[c.cur].data_ptr(), int(self._node_slot[node]), bit, st)
            outcomes[node] = 1 - 2 * bit
            self.run_bp()
            self._compute_bloch()
            self._nproj.zero_()
            for c in self.classes:
                self.lib.threshold_project(self.prec, c.degree, self.D, c.B, c.T[c.cur].data_ptr(), c.node_ids.data_ptr(),
                                           self._bloch.data_ptr(), outcomes.data_ptr(), self.threshold,
                                           self._nproj.data_ptr(), st)
            self.run_bp()
        log.info("Measurement outcomes sampling completed")
        return [int(v) for v in self._to_host(outcomes)]
