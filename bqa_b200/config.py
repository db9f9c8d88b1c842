"""Config -> Context compiler (host side, numpy only).

This mirrors the *behaviour* of the reference compilation step so that a config dict that
is valid for ``bqa.run_qa`` is valid here and compiles to the same degree-classified layout:

* syntax / defaults / errors ........ reference ``src/bqa/config/config_syntax.py:63-192``,
                                       ``src/bqa/config/schedule_syntax.py:83-169``,
                                       ``src/bqa/config/utils.py:3-86``
* degree classes and message slots ... ``src/bqa/config/config_canonicalization.py:62-250``
* schedule expansion ................. ``src/bqa/config/schedule_canonicalization.py:6-33``

Conventions that the kernels rely on:

* the directed edge ``lhs -> rhs`` of the e-th config edge sits at message slot ``e``, its reverse
  ``rhs -> lhs`` at slot ``e + L`` (L = number of undirected edges);
* the legs of a node tensor are ordered like its directed out-edges sorted by slot;
* ``nodes_number = 1 + max id``; ids that never appear in an edge become degree-0 nodes.

Unlike the reference (python loops over nodes, ~3 s at 100k qubits) the layout is built with
vectorised numpy (stable sort by source node), which gives the identical arrays.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from math import isclose

import numpy as np

log = logging.getLogger(__name__)


class ConfigSyntaxError(ValueError):
    """Same role as ``bqa.config.utils.ConfigSyntaxError`` (reference config/utils.py:3-4)."""


# --- defaults (reference config_syntax.py:41-61, schedule_syntax.py:48-81) -------------------------
DEFAULTS = {
    "nodes": {},
    "default_field": 0.0,
    "max_bond_dim": 4,
    "max_bp_iter_number": 75,
    "seed": 42,
    "bp_eps": 1e-6,
    "pinv_eps": 1e-6,
    "measurement_threshold": 0.95,
    "damping": 0.0,
    "backend": "b200",
}
EVOLUTION_TYPES = ("real_time_evolution", "imag_time_evolution")
SIMPLE_ACTIONS = ("measure", "get_bloch_vectors")
DEFAULT_SCHEDULE = {
    "total_time": 10.0,
    "starting_mixing": 1.0,
    "actions": [
        {"type": "real_time_evolution", "weight": 1.0, "steps_number": 100, "final_mixing": 0.0},
        "get_bloch_vectors",
    ],
}

# Backend names accepted by the syntax check.  The reference validates against its registry dict
# (config_syntax.py:63-70) and defaults to "numpy"; this package provides exactly one backend, "b200",
# which is therefore also the default.  The reference's own names "numpy" / "cupy" are accepted so that its examples
# and benchmark configs (every one of them sets "backend" explicitly) run unmodified -- with a warning, because
# execution is on the b200 kernels all the same: there is no CPU backend here.
KNOWN_BACKENDS = {"b200"}
REFERENCE_BACKENDS = {"numpy", "cupy"}


def register_backend_name(name: str) -> None:
    KNOWN_BACKENDS.add(name)


# --- atomic validators ------------------------------------------------------------------------
def _is_num(x) -> bool:
    return isinstance(x, (float, int))


def _number(x) -> float:
    if not _is_num(x):
        raise ConfigSyntaxError(f"Invalid value {x}, must be a `float` or `int` number")
    return float(x)


def _ranged(x, lo, hi, what) -> float:
    if not (_is_num(x) and lo <= x <= hi):
        raise ConfigSyntaxError(f"Invalid value {x}, must be a `float` or `int` number from {what}")
    return float(x)


def _non_neg_number(x) -> float:
    if not (_is_num(x) and x >= 0.0):
        raise ConfigSyntaxError(f"Invalid value {x}, must be a non-negative `float` or `int` number")
    return float(x)


def _non_neg_int(x) -> int:
    if not (isinstance(x, int) and x >= 0):
        raise ConfigSyntaxError(f"Invalid value {x}, must be a non-negative `int`")
    return x


def _positive_int(x) -> int:
    if not (isinstance(x, int) and x > 0):
        raise ConfigSyntaxError(f"Invalid value {x}, must be a positive `int`")
    return x


def _get(dct, key, default, where):
    val = dct.get(key)
    if val is None:
        log.warning(f"`{key}` field is missing in {where}, set to default {default}")
        return default
    return val


# --- nodes / edges ------------------------------------------------------------------------------
def _analyse_nodes(nodes) -> dict[int, float]:
    if not isinstance(nodes, (tuple, list, dict)):
        raise ConfigSyntaxError(f"Invalid nodes {nodes}")
    out: dict[int, float] = {}
    items = nodes.items() if isinstance(nodes, dict) else nodes
    try:
        for node in items:
            if not (isinstance(node, (tuple, list)) and len(node) == 2):
                raise ConfigSyntaxError(f"Invalid node {node}")
            try:
                node_id = _non_neg_int(node[0])
                ampl = _number(node[1])
            except ConfigSyntaxError as e:
                raise ConfigSyntaxError(f"Invalid node {node}") from e
            if node_id in out:
                raise ConfigSyntaxError(f"Invalid node {node}") from ConfigSyntaxError(f"Duplicated node ID {node_id}")
            out[node_id] = ampl
    except ConfigSyntaxError as e:
        raise ConfigSyntaxError(f"Invalid nodes {_abbrev(nodes)}") from e
    return out


def _abbrev(obj, limit: int = 200) -> str:
    s = repr(obj)
    return s if len(s) <= limit else s[:limit] + "...}"


def _analyse_edges(edges) -> tuple[np.ndarray, np.ndarray]:
    """Returns (E, J): E int64 (L, 2) in config order, J float64 (L,).

    The reference materialises ``forward | backward`` dicts (config_syntax.py:118-150); the slot
    convention (forward e, backward e + L) is all that survives, so only E and J are kept."""
    if not isinstance(edges, (tuple, list, dict)):
        raise ConfigSyntaxError(f"Invalid edges {edges}")
    items = edges.items() if isinstance(edges, dict) else edges
    lhs_l, rhs_l, amp_l = [], [], []
    seen = set()
    try:
        for edge in items:
            if not (isinstance(edge, (tuple, list)) and len(edge) == 2):
                raise ConfigSyntaxError(f"Invalid edge {edge}")
            eid = edge[0]
            try:
                if not (isinstance(eid, (tuple, list)) and len(eid) == 2):
                    raise ConfigSyntaxError(f"Invalid edge ID {eid}")
                try:
                    lhs, rhs = _non_neg_int(eid[0]), _non_neg_int(eid[1])
                    if lhs == rhs:
                        raise ConfigSyntaxError(
                            f"LHS and RHS of the edge ID must not be equal, got edge ID {eid}")
                except ConfigSyntaxError as e:
                    raise ConfigSyntaxError(f"Invalid edge ID {eid}") from e
                coupling = _number(edge[1])
                if (lhs, rhs) in seen or (rhs, lhs) in seen:
                    raise ConfigSyntaxError(f"Duplicated edge ID {(lhs, rhs)}, {(rhs, lhs)}")
            except ConfigSyntaxError as e:
                raise ConfigSyntaxError(f"Invalid edge {edge}") from e
            seen.add((lhs, rhs))
            lhs_l.append(lhs)
            rhs_l.append(rhs)
            amp_l.append(coupling)
    except ConfigSyntaxError as e:
        raise ConfigSyntaxError(f"Invalid edges {_abbrev(edges)}") from e
    E = np.stack([np.asarray(lhs_l, np.int64), np.asarray(rhs_l, np.int64)], axis=1) if lhs_l \
        else np.zeros((0, 2), np.int64)
    return E, np.asarray(amp_l, np.float64)


# --- schedule -----------------------------------------------------------------------------------
def _analyse_schedule(schedule) -> dict:
    if not isinstance(schedule, dict):
        raise ConfigSyntaxError(f"Schedule must be a dict, got {schedule} of type {type(schedule)}")
    try:
        total_time = _non_neg_number(_get(schedule, "total_time", DEFAULT_SCHEDULE["total_time"], "schedule"))
        mixing = _ranged(_get(schedule, "starting_mixing", DEFAULT_SCHEDULE["starting_mixing"], "schedule"),
                         0.0, 1.0, "[0, 1]")
        actions = _get(schedule, "actions", DEFAULT_SCHEDULE["actions"], "schedule")
        if not isinstance(actions, (list, tuple)):
            raise ConfigSyntaxError(
                f"Actions must be either a list or a tuple, got {actions} of type {type(actions)}")
        analysed = []
        try:
            for action in actions:
                if isinstance(action, str):
                    # a simple action is a zero-weight pseudo action (schedule_syntax.py:93-100)
                    desug = {"type": action, "weight": 0, "initial_mixing": mixing,
                             "final_mixing": mixing, "steps_number": 1}
                elif isinstance(action, dict):
                    desug = dict(action)
                    if "initial_mixing" in desug:
                        raise ConfigSyntaxError(
                            f"`initial_mixing` must not be present in the action {desug}, it is infered automatically")
                    desug["initial_mixing"] = mixing
                    if "final_mixing" not in desug:
                        desug["final_mixing"] = mixing
                    else:
                        mixing = desug["final_mixing"]
                else:
                    raise TypeError(f"Invalid type {type(action)} of the action {action}")
                a_type = _get(desug, "type", "real_time_evolution", "action")
                if a_type not in EVOLUTION_TYPES + SIMPLE_ACTIONS:
                    raise ConfigSyntaxError(f"Invalid action type {a_type} in action {desug}")
                if desug.get("weight") is None:
                    raise ConfigSyntaxError("`weight` field is missing in action")
                analysed.append({
                    "type": a_type,
                    "weight": _ranged(desug["weight"], 0.0, 1.0, "[0, 1]"),
                    "steps_number": _positive_int(_get(desug, "steps_number", 100, "action")),
                    "initial_mixing": _ranged(desug["initial_mixing"], 0.0, 1.0, "[0, 1]"),
                    "final_mixing": _ranged(desug["final_mixing"], 0.0, 1.0, "[0, 1]"),
                })
            wsum = sum(a["weight"] for a in analysed)
            if not isclose(wsum, 1.0):
                raise ConfigSyntaxError(
                    f"`weight` fields must sum into 1. in actions {analysed}, but now it sums into {wsum}")
        except ConfigSyntaxError as e:
            raise ConfigSyntaxError(f"Invalid actions {actions}") from e
        return {"total_time": total_time, "actions": analysed}
    except ConfigSyntaxError as e:
        raise ConfigSyntaxError(f"Invalid schedule {schedule}") from e


def expand_schedule(schedule: dict) -> list:
    """Flat instruction list (reference schedule_canonicalization.py:6-33): every evolution action
    becomes ``steps_number`` dicts {type, xtime, ztime}; the mixing p is interpolated linearly from
    the initial value, *excluding* the end point."""
    total_time = schedule["total_time"]
    out: list = []
    for action in schedule["actions"]:
        if action["type"] in EVOLUTION_TYPES:
            steps = action["steps_number"]
            dt = total_time * action["weight"] / steps
            p0, p1 = action["initial_mixing"], action["final_mixing"]
            delta = (p1 - p0) / steps
            for n in range(steps):
                p = p0 + n * delta
                out.append({"type": action["type"], "xtime": p * dt, "ztime": (1.0 - p) * dt})
        else:
            out.append(action["type"])
    return out


# --- compiled objects ---------------------------------------------------------------------------
@dataclass
class Layout:
    """One degree class (reference ``Layout``, config_canonicalization.py:25-32) as plain arrays.

    node_ids (B,), input/output_msgs_position (d, B), lmbds_position (d, B): int64;
    node_ampls (B,), edge_ampls (d, B): float64."""
    node_ids: np.ndarray
    input_msgs_position: np.ndarray
    output_msgs_position: np.ndarray
    lmbds_position: np.ndarray
    node_ampls: np.ndarray
    edge_ampls: np.ndarray
    # partitioned runs over peer memory only (bqa_b200/partitioned.py): where the owner of the receiving node keeps
    # each outgoing message, (peer << 27 | slot on that peer), -1 if the receiver is owned by this rank
    remote_msgs_position: np.ndarray = None

    @property
    def degree(self) -> int:
        return int(self.input_msgs_position.shape[0])

    @property
    def batch_size(self) -> int:
        return int(self.node_ids.shape[0])


@dataclass
class Context:
    """Compiled problem (reference ``Context``, config_canonicalization.py:144-168).

    ``edges_number`` is the *directed* count 2L as in the reference."""
    backend: str
    bp_eps: float
    pinv_eps: float
    measurement_threshold: float
    nodes_number: int
    edges_number: int
    max_bond_dim: int
    max_bp_iters_number: int
    seed: int
    damping: float
    degree_to_layout: dict[int, Layout]
    instructions: list
    edges: np.ndarray          # (L, 2) config-ordered undirected edges
    couplings: np.ndarray      # (L,)
    fields: np.ndarray         # (N,) node amplitude incl. default_field
    node_degree: np.ndarray = field(repr=False, default=None)   # (N,)
    node_slot: np.ndarray = field(repr=False, default=None)     # (N,) position inside its degree class

    @property
    def lmbds_number(self) -> int:
        return self.edges_number // 2

    @property
    def msg_pos_to_lmbd_pos(self) -> np.ndarray:
        L = self.lmbds_number
        return np.arange(self.edges_number, dtype=np.int64) % max(L, 1)

    @property
    def path_to_tensors(self) -> dict[int, tuple[int, int]]:
        return {int(n): (int(self.node_degree[n]), int(self.node_slot[n])) for n in range(self.nodes_number)}

    @property
    def graph(self) -> list[list[int]]:
        """Adjacency in leg order (reference ``_make_graph``, config_canonicalization.py:182-186)."""
        g: list[list[int]] = [[] for _ in range(self.nodes_number)]
        for lhs, rhs in self.edges.tolist():
            g[lhs].append(rhs)
        for lhs, rhs in self.edges.tolist():
            g[rhs].append(lhs)
        return g


def analyse_config(config) -> dict:
    """Syntax check + defaults (reference ``_analyse_config``, config_syntax.py:157-192).
    Unknown keys are ignored silently, like in the reference."""
    if not isinstance(config, dict):
        raise ConfigSyntaxError(f"Config must be a dict, but got type {type(config)}")
    try:
        nodes = _get(config, "nodes", DEFAULTS["nodes"], "config")
        edges = config.get("edges")
        if edges is None:
            raise ConfigSyntaxError("`edges` field is missing in config")
        schedule = _get(config, "schedule", DEFAULT_SCHEDULE, "config")
        vals = {k: _get(config, k, DEFAULTS[k], "config") for k in (
            "max_bond_dim", "max_bp_iter_number", "bp_eps", "pinv_eps", "backend", "default_field",
            "measurement_threshold", "seed", "damping")}
        backend = vals["backend"]
        if not isinstance(backend, str):
            raise ConfigSyntaxError(f"Invalid backend \"{backend}\"")
        if backend in REFERENCE_BACKENDS:
            log.warning(f"Backend \"{backend}\" is a backend of the reference package; bqa_b200 executes the config on "
                        "its \"b200\" backend (CUDA kernels, complex64 unless BQA_PRECISION=double)")
            backend = "b200"
        elif backend not in KNOWN_BACKENDS:
            raise ConfigSyntaxError(f"Unknown backend \"{backend}\", available backends "
                                    f"{sorted(KNOWN_BACKENDS | REFERENCE_BACKENDS)} (all executed on b200)")
        E, J = _analyse_edges(edges)
        return {
            "nodes": _analyse_nodes(nodes),
            "edges": (E, J),
            "default_field": _number(vals["default_field"]),
            "schedule": _analyse_schedule(schedule),
            "max_bond_dim": _positive_int(vals["max_bond_dim"]),
            "max_bp_iter_number": _non_neg_int(vals["max_bp_iter_number"]),
            "seed": _non_neg_int(vals["seed"]),
            "bp_eps": _non_neg_number(vals["bp_eps"]),
            "pinv_eps": _ranged(vals["pinv_eps"], 0.0, 1.0, "[0, 1]"),
            "measurement_threshold": _ranged(vals["measurement_threshold"], 0.5, 1.0, "[0.5, 1]"),
            "damping": _ranged(vals["damping"], 0.0, 1.0, "[0, 1]"),
            "backend": backend,
        }
    except ConfigSyntaxError as e:
        raise ConfigSyntaxError("Invalid config") from e


def build_layouts(E: np.ndarray, J: np.ndarray, fields: np.ndarray) -> tuple[dict[int, Layout], np.ndarray, np.ndarray]:
    """Degree-classified layouts from the undirected edge list.

    Equivalent to reference ``_make_degree_to_layout`` (config_canonicalization.py:62-131): a node's
    legs are its out-edges in slot order, i.e. forward edges (slot e, node is lhs) before backward ones
    (slot e + L, node is rhs).  A stable sort of the slot-ordered directed edge list by source node
    produces exactly that order."""
    N = fields.shape[0]
    L = E.shape[0]
    src = np.concatenate([E[:, 0], E[:, 1]])
    slot = np.arange(2 * L, dtype=np.int64)
    order = np.argsort(src, kind="stable")            # directed edges grouped by source, slot-ordered inside
    deg = np.bincount(src, minlength=N).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(deg)])[:-1]
    out_sorted = slot[order]                           # out slot of the j-th leg of each node, concatenated
    layouts: dict[int, Layout] = {}
    node_slot = np.zeros(N, np.int64)
    # classes are keyed in order of first appearance by node id (dict insertion order in the reference)
    _, first_idx = np.unique(deg, return_index=True)
    for d in deg[np.sort(first_idx)].tolist():
        ids = np.nonzero(deg == d)[0].astype(np.int64)
        node_slot[ids] = np.arange(ids.shape[0])
        if d > 0 and L > 0:
            legs = start[ids][None, :] + np.arange(d, dtype=np.int64)[:, None]     # (d, B) index into out_sorted
            out_pos = out_sorted[legs]
            in_pos = (out_pos + L) % (2 * L)
            lm_pos = out_pos % L
            e_ampl = J[lm_pos]
        else:
            out_pos = in_pos = lm_pos = np.zeros((0, ids.shape[0]), np.int64)
            e_ampl = np.zeros((0, ids.shape[0]), np.float64)
        layouts[int(d)] = Layout(ids, in_pos, out_pos, lm_pos, fields[ids].astype(np.float64), e_ampl)
    return layouts, deg, node_slot


def config_to_context(config) -> Context:
    """Drop-in for ``bqa.config.core.config_to_context`` (reference config/core.py:8-11)."""
    c = analyse_config(config)
    E, J = c["edges"]
    if E.shape[0] == 0:
        raise ConfigSyntaxError("Invalid config") from ConfigSyntaxError("`edges` must not be empty")
    N = 1 + max(max(c["nodes"].keys(), default=-1), int(E.max()))
    fields = np.full(N, c["default_field"], np.float64)
    for nid, ampl in c["nodes"].items():
        fields[nid] = ampl
    layouts, deg, node_slot = build_layouts(E, J, fields)
    ctx = Context(
        backend=c["backend"], bp_eps=c["bp_eps"], pinv_eps=c["pinv_eps"],
        measurement_threshold=c["measurement_threshold"], nodes_number=N, edges_number=2 * E.shape[0],
        max_bond_dim=c["max_bond_dim"], max_bp_iters_number=c["max_bp_iter_number"], seed=c["seed"],
        damping=c["damping"], degree_to_layout=layouts, instructions=expand_schedule(c["schedule"]),
        edges=E, couplings=J, fields=fields, node_degree=deg, node_slot=node_slot)
    log.info("Context is built")
    return ctx
