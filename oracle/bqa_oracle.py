"""CPU oracle for the bqa belief-propagation annealing path.  TEST INFRASTRUCTURE ONLY.

A from-scratch numpy restatement of the reference algorithm (LuchnikovI/bqa v0.1.6, numpy backend).
It is imported only by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``; the product (``bqa_b200``) never imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the unmodified reference (imported from
``/root/reference/src``) in the build container and committed its outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this file against those vectors, against the reference's own golden
layout vectors (reference ``tests/test_config_to_context.py:68-119``), against the tree-exactness
properties of reference ``tests/test_core_subroutines.py:191-256`` and against an independent
state-vector simulation (reference ``src/bqa/exact_sim.py:14-70`` restated in numpy).

Each function cites the reference file:line it follows.  Arithmetic follows the reference operation by
operation (same LAPACK ``numpy.linalg.svd`` calls, same masks and thresholds, no sharing of leg
contractions), so that (i) results agree to rounding with the numpy backend and (ii) its running time
is representative of the reference CPU path.  dtype: complex128 (``BQA_PRECISION=double`` default) or
complex64 via ``dtype=`` (reference ``src/bqa/utils.py:9-30``).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from math import isclose

import numpy as np

log = logging.getLogger(__name__)

SQRT_NEG_1J = np.sqrt(2.0) / 2.0 - 1j * np.sqrt(2.0) / 2.0      # reference backends.py:20
SQRT_1J = np.sqrt(2.0) / 2.0 + 1j * np.sqrt(2.0) / 2.0          # reference backends.py:22


# ---------------------------------------------------------------------------------------------
# compile step (loop form, like the reference; the product uses a vectorised equivalent)
# ---------------------------------------------------------------------------------------------
@dataclass
class OLayout:
    node_ids: np.ndarray
    in_pos: list            # d arrays (B,)
    out_pos: list
    lmbd_pos: list
    node_ampls: np.ndarray  # (B,) complex dtype like the reference (backends.py:583)
    edge_ampls: list        # d arrays (B,)


@dataclass
class OContext:
    nodes_number: int
    edges_number: int           # directed count 2L
    max_bond_dim: int
    max_bp_iters: int
    bp_eps: float
    pinv_eps: float
    threshold: float
    damping: float
    seed: int
    layouts: dict               # degree -> OLayout
    path: dict                  # node id -> (degree, position)
    instructions: list
    graph: list
    edge_to_msg_pos: dict
    edge_to_lmbd_pos: dict
    node_to_ampl: dict
    edge_to_ampl: dict          # directed
    dtype: type = np.complex128

    @property
    def lmbds_number(self):
        return self.edges_number // 2


def _expand_schedule(schedule: dict) -> list:
    """reference schedule_syntax.py:90-169 (mixing bookkeeping) + schedule_canonicalization.py:6-33."""
    total_time = float(schedule.get("total_time", 10.0))
    mixing = float(schedule.get("starting_mixing", 1.0))
    actions = schedule.get("actions", [{"weight": 1.0, "steps_number": 100, "final_mixing": 0.0},
                                       "get_bloch_vectors"])
    out = []
    wsum = 0.0
    for a in actions:
        if isinstance(a, str):
            out.append(a)
            continue
        steps = a.get("steps_number", 100)
        p0 = mixing
        p1 = a.get("final_mixing", mixing)
        mixing = p1
        wsum += a["weight"]
        dt = total_time * float(a["weight"]) / steps
        delta = (p1 - p0) / steps
        for n in range(steps):
            p = p0 + n * delta
            out.append({"type": a.get("type", "real_time_evolution"), "xtime": p * dt, "ztime": (1.0 - p) * dt})
    assert isclose(wsum, 1.0), "weights must sum to 1 (reference schedule_syntax.py:85-88)"
    return out


def compile_config(config: dict, dtype=np.complex128) -> OContext:
    """reference config_syntax.py:118-150 (edge ordering: forward then backward),
    config_canonicalization.py:62-131 (degree classes) and :171-250 (context)."""
    edges_in = config["edges"]
    items = list(edges_in.items()) if isinstance(edges_in, dict) else [tuple(e) for e in edges_in]
    fwd = {(int(l), int(r)): float(a) for (l, r), a in items}
    bwd = {(r, l): a for (l, r), a in fwd.items()}
    directed = fwd | bwd                                   # insertion order = message slot order
    nodes_in = config.get("nodes") or {}
    nodes_items = nodes_in.items() if isinstance(nodes_in, dict) else nodes_in
    node_fields = {int(k): float(v) for k, v in nodes_items}
    n_nodes = 1 + max(max(node_fields.keys(), default=-1), max(max(l, r) for l, r in directed))
    default_field = float(config.get("default_field") or 0.0)
    node_to_ampl = {n: node_fields.get(n, default_field) for n in range(n_nodes)}
    graph = [[] for _ in range(n_nodes)]
    for l, r in directed:
        graph[l].append(r)
    L = len(directed) // 2
    msg_pos = {e: p for p, e in enumerate(directed)}
    lmbd_pos = {e: p % L for p, e in enumerate(directed)}
    raw: dict = {}
    for n, nbrs in enumerate(graph):
        d = len(nbrs)
        cls = raw.setdefault(d, {"ids": [], "in": [[] for _ in range(d)], "out": [[] for _ in range(d)],
                                 "lm": [[] for _ in range(d)], "na": [], "ea": [[] for _ in range(d)]})
        cls["ids"].append(n)
        cls["na"].append(node_to_ampl[n])
        for j, m in enumerate(nbrs):
            cls["in"][j].append(msg_pos[(m, n)])
            cls["out"][j].append(msg_pos[(n, m)])
            cls["lm"][j].append(lmbd_pos[(n, m)])
            cls["ea"][j].append(directed[(n, m)])
    layouts = {}
    path = {}
    for d, c in raw.items():
        layouts[d] = OLayout(
            np.asarray(c["ids"], np.intp),
            [np.asarray(x, np.intp) for x in c["in"]],
            [np.asarray(x, np.intp) for x in c["out"]],
            [np.asarray(x, np.intp) for x in c["lm"]],
            np.asarray(c["na"], dtype),
            [np.asarray(x, dtype) for x in c["ea"]])
        for pos, n in enumerate(c["ids"]):
            path[n] = (d, pos)
    schedule = config.get("schedule") or {}
    return OContext(
        nodes_number=n_nodes, edges_number=len(directed),
        max_bond_dim=int(config.get("max_bond_dim") or 4),
        max_bp_iters=int(config["max_bp_iter_number"]) if config.get("max_bp_iter_number") is not None else 75,
        bp_eps=float(config["bp_eps"]) if config.get("bp_eps") is not None else 1e-6,
        pinv_eps=float(config["pinv_eps"]) if config.get("pinv_eps") is not None else 1e-6,
        threshold=float(config["measurement_threshold"]) if config.get("measurement_threshold") is not None else 0.95,
        damping=float(config["damping"]) if config.get("damping") is not None else 0.0,
        seed=int(config["seed"]) if config.get("seed") is not None else 42,
        layouts=layouts, path=path, instructions=_expand_schedule(schedule), graph=graph,
        edge_to_msg_pos=msg_pos, edge_to_lmbd_pos=lmbd_pos, node_to_ampl=node_to_ampl,
        edge_to_ampl=directed, dtype=dtype)


# ---------------------------------------------------------------------------------------------
# state
# ---------------------------------------------------------------------------------------------
@dataclass
class OState:
    rng: np.random.Generator
    tensors: dict               # degree -> (B, 2, D, ..., D)
    msgs: np.ndarray            # (2L, D, D)
    lmbds: np.ndarray           # (L, D) stored in the complex dtype
    stats: dict = field(default_factory=lambda: {"bp_sweeps": [], "bp_dist": [], "bond_dims": [], "trunc_err": []})

    @property
    def bond_dim(self) -> int:
        return next(iter(self.tensors.values())).shape[-1]


def msgs_from_lmbds(lmbds: np.ndarray, ctx: OContext) -> np.ndarray:
    """reference state.py:56-57: msgs[p] = diag(lmbd[p mod L]) / trace."""
    L = ctx.lmbds_number
    lm = lmbds[np.arange(ctx.edges_number) % L]
    m = lm[..., None] * np.eye(lm.shape[-1])
    return m / np.trace(m, axis1=-2, axis2=-1)[:, None, None]


def init_state(ctx: OContext) -> OState:
    """reference state.py:21, :41-74: every qubit in |-> = (1, -1)/sqrt 2, bond dimension 1."""
    minus = np.array([np.sqrt(0.5), -np.sqrt(0.5)], ctx.dtype)
    tensors = {d: np.ascontiguousarray(np.broadcast_to(minus.reshape((2,) + (1,) * d),
                                                       (lay.node_ids.shape[0], 2) + (1,) * d))
               for d, lay in ctx.layouts.items()}
    lmbds = np.ones((ctx.lmbds_number, 1), ctx.dtype)
    return OState(np.random.default_rng(ctx.seed), tensors, msgs_from_lmbds(lmbds, ctx), lmbds)


# ---------------------------------------------------------------------------------------------
# tensor algebra (reference backends.py:329-408)
# ---------------------------------------------------------------------------------------------
def _contract_first_bond(t: np.ndarray, m: np.ndarray) -> np.ndarray:
    """t: (B, 2, D1, rest...), m: (B, X, D1).  Contracts the FIRST bond leg of t with the second index
    of m and appends the new index last (reference ``batch_tensordot(msg, [[1], [1]])``, backends.py:386,
    via the transpose+reshape+matmul route of :329-361)."""
    B = t.shape[0]
    moved = np.moveaxis(t, 2, -1)                               # (B, 2, rest..., D1)
    keep = moved.shape[1:-1]
    flat = np.ascontiguousarray(moved).reshape(B, -1, moved.shape[-1])
    res = flat @ np.ascontiguousarray(np.swapaxes(m, 1, 2))     # (B, prod(keep), X)
    return res.reshape((B,) + keep + (m.shape[1],))


def _rotate_first_bond_last(t: np.ndarray) -> np.ndarray:
    """reference backends.py:376-379."""
    return np.moveaxis(t, 2, -1)


def _extend_leg(t: np.ndarray, axis: int, theta: np.ndarray, conj: bool) -> np.ndarray:
    """ZZ half gate on one bond leg (reference backends.py:519-526): the leg of size D becomes 2D,
    upper block sqrt(cos th) * t, lower block e^{-i pi/4} sqrt(sin th) * Z t (complex principal roots;
    the bra side uses the conjugate factors)."""
    shp = (-1,) + (1,) * (t.ndim - 1)
    zt = t.copy()
    zt[:, 1] *= -1.0
    c = np.sqrt(np.cos(theta))
    s = np.sqrt(np.sin(theta))
    if conj:
        up = t * c.conj().reshape(shp)
        down = zt * (SQRT_1J * s.conj()).reshape(shp)
    else:
        up = t * c.reshape(shp)
        down = zt * (SQRT_NEG_1J * s).reshape(shp)
    return np.concatenate([up, down], axis)


def pass_msgs(t: np.ndarray, msgs_in: list, thetas: list | None = None) -> list:
    """All d outgoing messages of a degree class (reference backends.py:381-408).

    out_k[x, y] = sum conj(T[p, a_{!=k}, x]) prod_{j != k} m_j[a_j, b_j] T[p, b_{!=k}, y], trace-normalised;
    with ``thetas`` the open leg is first extended D -> 2D on ket and bra (state.py:127-139)."""
    d = len(msgs_in)
    tc = t.conj()
    outs = []
    for k in range(d):
        tm = t
        for j, m in enumerate(msgs_in):
            tm = _contract_first_bond(tm, m) if j != k else _rotate_first_bond_last(tm)
        bra = tc
        if thetas is not None:
            tm = _extend_leg(tm, k + 2, thetas[k], conj=False)
            bra = _extend_leg(tc, k + 2, thetas[k], conj=True)
        B = t.shape[0]
        bra_m = np.ascontiguousarray(np.moveaxis(bra, k + 2, 1)).reshape(B, bra.shape[k + 2], -1)
        ket_m = np.ascontiguousarray(np.moveaxis(tm, k + 2, -1)).reshape(B, -1, tm.shape[k + 2])
        out = bra_m @ ket_m
        outs.append(out / np.trace(out, axis1=-2, axis2=-1)[:, None, None])
    return outs


def density_of_class(t: np.ndarray, msgs_in: list) -> np.ndarray:
    """reference backends.py:440-448: rho[p, q] = sum T[p, b] prod m_j[a_j, b_j] conj(T[q, a]), / trace."""
    B = t.shape[0]
    tm = t
    for m in msgs_in:
        tm = _contract_first_bond(tm, m)
    rho = tm.reshape(B, 2, -1) @ np.swapaxes(t.conj().reshape(B, 2, -1), 1, 2)
    return rho / np.trace(rho, axis1=-2, axis2=-1)[:, None, None]


# ---------------------------------------------------------------------------------------------
# BP (reference state.py:97-124)
# ---------------------------------------------------------------------------------------------
def _max_abs(a: np.ndarray) -> float:
    return np.abs(a).max()                                       # reference backends.py:621-623


def run_bp(ctx: OContext, st: OState) -> int:
    """Damped BP fixed point with the reference's exact termination semantics: on convergence the
    *previous* iterate is kept; on hitting the cap the last (undamped) sweep output becomes the state."""
    D = st.bond_dim
    new = np.empty((ctx.edges_number, D, D), ctx.dtype)
    dist = np.inf
    for it in range(ctx.max_bp_iters):
        for d, t in st.tensors.items():
            lay = ctx.layouts[d]
            outs = pass_msgs(t, [st.msgs[p] for p in lay.in_pos])
            for o, p in zip(outs, lay.out_pos):
                new[p] = o
        dist = float(_max_abs(new - st.msgs) / _max_abs(new + st.msgs))
        if dist < ctx.bp_eps:
            st.stats["bp_sweeps"].append(it + 1)
            st.stats["bp_dist"].append(dist)
            return it + 1
        st.msgs *= ctx.damping                                   # reference backends.py:761-764
        st.msgs += (1.0 - ctx.damping) * new
    assert ctx.max_bp_iters > 0
    st.msgs = new
    st.stats["bp_sweeps"].append(ctx.max_bp_iters)
    st.stats["bp_dist"].append(dist)
    log.warning("BP exceeded the iteration cap %d, last dist %g", ctx.max_bp_iters, dist)
    return ctx.max_bp_iters


# ---------------------------------------------------------------------------------------------
# simple update (reference state.py:127-139, :171-200, :230-247)
# ---------------------------------------------------------------------------------------------
def extended_msgs(ctx: OContext, st: OState, ztime: float) -> np.ndarray:
    D = st.bond_dim
    ext = np.empty((ctx.edges_number, 2 * D, 2 * D), ctx.dtype)
    for d, t in st.tensors.items():
        lay = ctx.layouts[d]
        thetas = [a * ztime for a in lay.edge_ampls]
        outs = pass_msgs(t, [st.msgs[p] for p in lay.in_pos], thetas)
        for o, p in zip(outs, lay.out_pos):
            ext[p] = o
    return ext


def masked_svd(a: np.ndarray, eps: float, dtype):
    """reference backends.py:709-717."""
    u, s, vh = np.linalg.svd(a, full_matrices=False)
    mask = s > eps
    s = (s * mask).astype(dtype)
    return u * mask[..., None, :], s, vh * mask[..., None]


def _pinv(a: np.ndarray, dtype) -> np.ndarray:
    """reference backends.py:719-727 (cut at machine eps of the working dtype)."""
    out = np.zeros_like(a)
    np.divide(1.0, a, out=out, where=a.real > np.finfo(dtype).eps)
    return out


def canonicalizers(ext: np.ndarray, pinv_eps: float, dtype) -> tuple[np.ndarray, np.ndarray]:
    """reference state.py:171-200 + backends.py:483-490.  Returns (lmbds (L, 2D), canon (2L, 2D, 2D));
    slot p < L holds the backward canonicalizer, slot p >= L the forward one (state.py:182-183)."""
    L = ext.shape[0] // 2
    u, lam, uh = masked_svd(ext, pinv_eps, dtype)
    root = np.sqrt(lam)
    lu = root[..., :, None] * uh
    ul = u * _pinv(root, dtype)[..., None, :]
    ker = lu[:L] @ np.swapaxes(lu[L:], 1, 2)                    # batch_tensordot(bwd, [[1],[1]])
    us, s, vhs = masked_svd(ker, pinv_eps, dtype)
    vs = np.swapaxes(vhs, 1, 2)
    fwd_c = ul[:L] @ us
    bwd_c = ul[L:] @ vs
    nrm = np.linalg.norm(s, axis=1)
    return s / nrm[:, None], np.concatenate([bwd_c, fwd_c], 0)


def truncate_lmbds(lmbds: np.ndarray, max_dim: int, eps: float) -> tuple[np.ndarray, int, float]:
    """reference backends.py:297-303: one global bond dimension from the column-wise maxima."""
    colmax = np.abs(lmbds.max(0))
    rank = lmbds.shape[1] - int(np.sum(colmax < eps))
    dim = min(rank, max_dim)
    err = float(np.sqrt(np.sum(colmax[dim:] ** 2)))
    return lmbds[:, :dim], dim, err


def apply_canonicalizers_ext(t: np.ndarray, canons: list, thetas: list) -> np.ndarray:
    """reference backends.py:416-432: extend leg, contract with its (2D, dim) canonicalizer, next leg."""
    for c, th in zip(canons, thetas):
        t = _extend_leg(t, 2, th, conj=False)
        B = t.shape[0]
        moved = np.moveaxis(t, 2, -1)
        keep = moved.shape[1:-1]
        res = np.ascontiguousarray(moved).reshape(B, -1, moved.shape[-1]) @ c
        t = res.reshape((B,) + keep + (c.shape[2],))
    return t


def simple_update(ctx: OContext, st: OState, ztime: float) -> None:
    ext = extended_msgs(ctx, st, ztime)
    lmbds, canon = canonicalizers(ext, ctx.pinv_eps, ctx.dtype)
    st.lmbds, dim, err = truncate_lmbds(lmbds, ctx.max_bond_dim, ctx.pinv_eps)
    canon = canon[:, :, :dim]
    st.stats["bond_dims"].append(dim)
    st.stats["trunc_err"].append(err)
    for d, lay in ctx.layouts.items():
        st.tensors[d] = apply_canonicalizers_ext(
            st.tensors[d], [canon[p] for p in lay.in_pos], [a * ztime for a in lay.edge_ampls])


def z_layer(ctx: OContext, st: OState, ztime: float) -> None:
    """reference state.py:142-150, backends.py:509-510."""
    for d, t in st.tensors.items():
        phi = (ztime * ctx.layouts[d].node_ampls).reshape((-1,) + (1,) * (t.ndim - 1))
        zt = t.copy()
        zt[:, 1] *= -1.0
        st.tensors[d] = t * np.cos(phi) - 1j * zt * np.sin(phi)


def x_layer(st: OState, xtime: float) -> None:
    """reference state.py:153-156, backends.py:506-507."""
    for d, t in st.tensors.items():
        st.tensors[d] = np.cos(xtime) * t - 1j * np.sin(xtime) * t[:, ::-1]


def symmetric_gauge(ctx: OContext, st: OState) -> None:
    """reference state.py:219-227, backends.py:450-462."""
    st.msgs = msgs_from_lmbds(st.lmbds, ctx)
    for d, lay in ctx.layouts.items():
        t = st.tensors[d]
        for j, lp in enumerate(lay.lmbd_pos):
            root = np.sqrt(st.lmbds[lp])                         # (B, D)
            shp = [t.shape[0], 1] + [1] * d
            shp[2 + j] = root.shape[1]
            t = t * root.reshape(shp)
        nrm = np.linalg.norm(t.reshape(t.shape[0], -1), axis=1)
        st.tensors[d] = t / nrm.reshape((-1,) + (1,) * (t.ndim - 1))


def run_layer(ctx: OContext, st: OState, xtime: float, ztime: float) -> None:
    """One annealing (Trotter) step, reference state.py:315-321."""
    simple_update(ctx, st, ztime)
    z_layer(ctx, st, ztime)
    x_layer(st, xtime)
    symmetric_gauge(ctx, st)
    run_bp(ctx, st)


# ---------------------------------------------------------------------------------------------
# marginals and sampling (reference state.py:77-94, :250-312; utils.py:23-27)
# ---------------------------------------------------------------------------------------------
def density_matrices(ctx: OContext, st: OState) -> np.ndarray:
    rho = np.empty((ctx.nodes_number, 2, 2), ctx.dtype)
    for d, lay in ctx.layouts.items():
        rho[lay.node_ids] = density_of_class(st.tensors[d], [st.msgs[p] for p in lay.in_pos])
    return rho


def bloch_vectors(rho: np.ndarray) -> np.ndarray:
    x = (rho[:, 0, 1] + rho[:, 1, 0]).real
    y = (rho[:, 1, 0] - rho[:, 0, 1]).imag
    z = (rho[:, 0, 0] - rho[:, 1, 1]).real
    return np.stack([x, y, z], 1)


def measure(ctx: OContext, st: OState) -> list:
    """Sequential decimation sampler, reference state.py:250-312 and backends.py:729-734."""
    outcomes: dict = {}
    thr = ctx.threshold

    def ground_probs():
        rho = density_matrices(ctx, st)
        return {n: float(rho[n, 0, 0].real) for n in range(ctx.nodes_number) if n not in outcomes}

    def project(n: int, bit: int):
        outcomes[n] = 1 - 2 * bit
        d, pos = ctx.path[n]
        t = st.tensors[d]
        t[pos, 1 - bit] = 0.0
        t /= np.linalg.norm(t)

    while len(outcomes) < ctx.nodes_number:
        probs = ground_probs()
        n, p = max(probs.items(), key=lambda kv: abs(2 * kv[1] - 1))
        u = st.rng.uniform(0.0, 1.0)
        project(n, 0 if p > u else 1)
        run_bp(ctx, st)
        probs = ground_probs()
        for m, q in probs.items():
            if q > thr:
                project(m, 0)
        for m, q in probs.items():
            if q < 1.0 - thr:
                project(m, 1)
        run_bp(ctx, st)
    return [outcomes[n] for n in range(ctx.nodes_number)]


def run_qa(config: dict, dtype=np.complex128, return_state: bool = False):
    """reference core.py:13-35."""
    ctx = compile_config(config, dtype)
    st = init_state(ctx)
    results = []
    for ins in ctx.instructions:
        if isinstance(ins, dict):
            run_layer(ctx, st, ins["xtime"], ins["ztime"])
        elif ins == "measure":
            results.append(["measurement_outcomes", measure(ctx, st)])
        elif ins == "get_bloch_vectors":
            results.append(["bloch_vectors", bloch_vectors(density_matrices(ctx, st)).tolist()])
        else:
            raise ValueError(f"Unknown instruction {ins}")
    return (results, ctx, st) if return_state else results


def ising_energy(config_edges, config_nodes, spins) -> float:
    """E(s) = sum_ij J_ij s_i s_j + sum_i h_i s_i with s = +1 for bit 0 (the reference's own
    energy function lives in the absent mqlib_wrap; this definition is consistent with
    reference exact_sim.py:36-63 and state.py:276)."""
    e = 0.0
    items = config_edges.items() if isinstance(config_edges, dict) else config_edges
    for (l, r), j in items:
        e += j * spins[l] * spins[r]
    nitems = config_nodes.items() if isinstance(config_nodes, dict) else config_nodes
    for n, h in nitems:
        e += h * spins[n]
    return float(e)


# ---------------------------------------------------------------------------------------------
# independent state-vector simulation (reference exact_sim.py:14-70 restated; qem is not available)
# ---------------------------------------------------------------------------------------------
def run_exact_statevector(config: dict) -> np.ndarray:
    """Bloch vectors after the schedule for <= ~20 qubits.  Order of gates inside a layer follows
    reference exact_sim.py:51-63: Rz on every node, RZZ on every undirected edge, Rx on every node."""
    ctx = compile_config(config)
    n = ctx.nodes_number
    psi = np.ones(2 ** n, np.complex128)
    # |-> on every qubit: amplitude sign (-1)^{popcount}
    idx = np.arange(2 ** n)
    bits = ((idx[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1)      # qubit 0 = most significant
    psi = psi * (1 - 2 * (bits.sum(1) % 2)) / np.sqrt(2.0 ** n)
    z = 1 - 2 * bits                                                    # (2^n, n) eigenvalues of Z
    und = [(e, a) for e, a in ctx.edge_to_ampl.items() if ctx.edge_to_msg_pos[e] < ctx.lmbds_number]
    for ins in ctx.instructions:
        if not isinstance(ins, dict):
            continue
        zt, xt = ins["ztime"], ins["xtime"]
        phase = np.zeros(2 ** n)
        for q, h in ctx.node_to_ampl.items():
            phase += zt * h * z[:, q]
        for (l, r), a in und:
            phase += zt * a * z[:, l] * z[:, r]
        psi = psi * np.exp(-1j * phase)
        t = psi.reshape((2,) * n)
        c, s = np.cos(xt), np.sin(xt)
        for q in range(n):
            t = c * t - 1j * s * np.flip(t, axis=q)
        psi = t.reshape(-1)
    t = psi.reshape((2,) * n)
    out = np.zeros((n, 3))
    for q in range(n):
        m = np.moveaxis(t, q, 0).reshape(2, -1)
        rho = m @ m.conj().T
        rho = rho / np.trace(rho)
        out[q] = [(rho[0, 1] + rho[1, 0]).real, (rho[1, 0] - rho[0, 1]).imag, (rho[0, 0] - rho[1, 1]).real]
    return out
