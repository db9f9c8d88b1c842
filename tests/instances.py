"""Deterministic test instances shared by the golden generator and the tests.

Nothing here depends on the reference or on networkx: graphs are built explicitly and amplitudes come
from ``random.Random(seed)`` so the GPU box rebuilds bit-identical inputs."""
from __future__ import annotations

from random import Random

import numpy as np


def ring_with_chords(n: int, seed: int, zero_field: bool = False, unit_coupling: bool = False):
    """3-regular graph: ring i -- i+1 plus chords i -- i + n/2 (n even)."""
    assert n % 2 == 0
    rng = Random(seed)
    nodes = {i: (0.0 if zero_field else rng.uniform(-1.0, 1.0)) for i in range(n)}
    edges = {}
    for i in range(n):
        edges[(i, (i + 1) % n)] = 1.0 if unit_coupling else rng.uniform(-1.0, 1.0)
    for i in range(n // 2):
        # alternate orientation so that nodes are lhs and rhs in mixed order
        e = (i, i + n // 2) if i % 2 == 0 else (i + n // 2, i)
        edges[e] = 1.0 if unit_coupling else rng.uniform(-1.0, 1.0)
    return nodes, edges


def grid(m: int, n: int, seed: int):
    rng = Random(seed)
    nodes = {i * n + j: rng.uniform(-1.0, 1.0) for i in range(m) for j in range(n)}
    edges = {}
    for i in range(m):
        for j in range(n):
            if j + 1 < n:
                edges[(i * n + j, i * n + j + 1)] = rng.uniform(-1.0, 1.0)
            if i + 1 < m:
                edges[(i * n + j, (i + 1) * n + j)] = rng.uniform(-1.0, 1.0)
    return nodes, edges


def comb(n_spine: int, seed: int):
    """Heavy-hex flavoured tree: a spine with a pendant qubit on every other spine site (degrees 1, 2, 3),
    amplitudes +-1 like reference examples/full_size_ibm_heavy_hex.py:12-13."""
    rng = Random(seed)
    pm = lambda: float(2 * rng.randint(0, 1) - 1)
    n_pend = (n_spine + 1) // 2
    nodes = {i: pm() for i in range(n_spine + n_pend)}
    edges = {}
    for i in range(n_spine - 1):
        edges[(i, i + 1)] = pm()
    for k in range(n_pend):
        edges[(n_spine + k, 2 * k)] = pm()
    return nodes, edges


def _anneal(total_time, steps, tail):
    return {"total_time": total_time, "starting_mixing": 1.0,
            "actions": [{"weight": 1.0, "steps_number": steps, "final_mixing": 0.0}, *tail]}


def cfg_small6():
    # reference tests/test_small_circuit_final_density.py:9-15 (default schedule: 100 steps, T=10, Bloch vectors)
    return {
        "nodes": {0: 1., 1: -1., 2: 0.5, 3: -0.5, 4: 1.1, 5: 0.4},
        "edges": {(0, 1): 1., (2, 1): -1., (1, 3): 1., (4, 3): -1., (5, 3): 1.},
        "pinv_eps": 1e-9, "bp_eps": 1e-9, "max_bond_dim": 8,
    }


def cfg_ring24():
    nodes, edges = ring_with_chords(24, seed=7)
    return {"nodes": nodes, "edges": edges, "max_bond_dim": 4,
            "schedule": _anneal(4.0, 20, ["get_bloch_vectors", "measure"])}


def cfg_ring24_capped():
    nodes, edges = ring_with_chords(24, seed=11)
    return {"nodes": nodes, "edges": edges, "max_bond_dim": 3, "max_bp_iter_number": 3, "damping": 0.2,
            "schedule": _anneal(3.0, 12, ["get_bloch_vectors"])}


def cfg_grid4():
    nodes, edges = grid(4, 4, seed=42)
    return {"nodes": nodes, "edges": edges, "max_bond_dim": 4, "damping": 0.3,
            "schedule": _anneal(4.0, 16, ["get_bloch_vectors", "measure"])}


def cfg_comb():
    nodes, edges = comb(9, seed=42)
    return {"nodes": nodes, "edges": edges, "max_bond_dim": 8, "measurement_threshold": 0.9,
            "schedule": _anneal(10.0, 10, ["get_bloch_vectors", "measure"])}


GOLDEN_CONFIGS = {
    "small6": cfg_small6,
    "ring24": cfg_ring24,
    "ring24_capped": cfg_ring24_capped,
    "grid4": cfg_grid4,
    "comb": cfg_comb,
}


def random_psd_msgs(rng: np.random.Generator, B: int, D: int) -> np.ndarray:
    a = rng.normal(size=(B, D, D)) + 1j * rng.normal(size=(B, D, D))
    m = a @ np.swapaxes(a.conj(), 1, 2)
    return m / np.trace(m, axis1=1, axis2=2)[:, None, None]


def random_node_batch(B: int, d: int, D: int, seed: int):
    """(T (B, 2, D^d) complex128 unit-norm per node, d Hermitian PSD trace-1 messages (B, D, D),
    d coupling angles (B,) float64 in (-1, 1))."""
    rng = np.random.default_rng(seed)
    shape = (B, 2) + (D,) * d
    t = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    t /= np.linalg.norm(t.reshape(B, -1), axis=1).reshape((B,) + (1,) * (d + 1))
    msgs = [random_psd_msgs(rng, B, D) for _ in range(d)]
    thetas = [rng.uniform(-1.0, 1.0, size=B) for _ in range(d)]
    return t, msgs, thetas


def cfg_isolated_qubit():
    """A path 0-1-2 plus the isolated qubit 3 (degree 0).  The reference compiles such a config
    (tests/test_config_to_context.py:68-75) but asserts in get_density_matrices; here the
    isolated qubit simply follows its single-qubit Rz/Rx evolution."""
    return {"nodes": {0: 0.3, 1: -0.2, 2: 0.5, 3: 0.7}, "edges": {(0, 1): 0.8, (1, 2): -0.4}, "max_bond_dim": 4,
            "schedule": _anneal(1.0, 5, ["get_bloch_vectors", "measure"])}


def isolated_qubit_bloch(cfg, node=3):
    """|-> evolved by Rx(xt) Rz(h zt) per step (reference gate conventions, exact_sim.py:36-63)."""
    import math
    from bqa_b200.config import config_to_context
    h = cfg["nodes"][node]
    psi = np.array([1.0, -1.0], complex) / np.sqrt(2.0)
    for ins in config_to_context(cfg).instructions:
        if isinstance(ins, dict):
            phi, xt = h * ins["ztime"], ins["xtime"]
            rz = np.diag([np.exp(-1j * phi), np.exp(1j * phi)])
            rx = np.array([[math.cos(xt), -1j * math.sin(xt)], [-1j * math.sin(xt), math.cos(xt)]])
            psi = rx @ rz @ psi
    rho = np.outer(psi, psi.conj())
    return np.array([(rho[0, 1] + rho[1, 0]).real, (rho[1, 0] - rho[0, 1]).imag, (rho[0, 0] - rho[1, 1]).real])


def instance_fingerprint(nodes: dict, edges: dict) -> str:
    """sha256 over node ids / fields / edges / couplings in dict order: the golden generators store it, the GPU tests
    compare it before trusting a fixture (the random graph comes from networkx and must be rebuilt identically)."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.asarray(list(nodes.keys()), np.int64).tobytes())
    h.update(np.asarray(list(nodes.values()), np.float64).tobytes())
    h.update(np.asarray(list(edges.keys()), np.int64).tobytes())
    h.update(np.asarray(list(edges.values()), np.float64).tobytes())
    return h.hexdigest()


def bench_config(n: int, total_steps: int = 100, dt: float = 0.2) -> dict:
    """bench.py's workload (BASELINE.json configs[3] at n = 100 000): random 3-regular QUBO, max_bond_dim 4, defaults of
    reference benchmarks_against_mqlib/random_3_regular_qubo_100000.py:10-34, schedule total_time = dt * S, S steps."""
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    nodes, edges = generate_qubo_on_random_regular_graph(n, 3, seed=42)
    return {"nodes": nodes, "edges": edges, "max_bond_dim": 4, "bp_eps": 1e-6, "pinv_eps": 1e-6, "damping": 0.0,
            "max_bp_iter_number": 75, "seed": 42, "default_field": 0.0, "measurement_threshold": 0.95,
            "schedule": {"total_time": dt * total_steps, "starting_mixing": 1.0,
                         "actions": [{"type": "real_time_evolution", "weight": 1.0, "steps_number": total_steps,
                                      "final_mixing": 0.0}]}}
