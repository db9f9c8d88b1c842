"""Synthetic instance generators with the same call signature and the same random streams as the
reference (``src/bqa/benchmarking.py:31-72``): graph from networkx driven by ``random.Random(seed)``,
labels converted to integers, then node amplitudes followed by edge amplitudes drawn ~U(-1, 1) from the
same generator.  Identical (seed, size) therefore give identical instances with the same networkx."""
from __future__ import annotations

from random import Random
from typing import Callable

Edge = tuple[int, int]


def _uniform_ampl(rng: Random, _) -> float:
    return rng.uniform(-1.0, 1.0)


def _draw_ampls(nxgraph, rng: Random, node_ampl_func, edge_ampl_func):
    import networkx as nx
    g = nx.convert_node_labels_to_integers(nxgraph)
    nodes = {n: node_ampl_func(rng, n) for n in g.nodes}
    edges = {e: edge_ampl_func(rng, e) for e in g.edges}
    return nodes, edges


def generate_qubo_on_random_regular_graph(
        nodes_number: int, degree: int = 3, seed: int = 42,
        node_ampl_func: Callable[[Random, int], float] = _uniform_ampl,
        edge_ampl_func: Callable[[Random, Edge], float] = _uniform_ampl):
    import networkx as nx
    rng = Random(seed)
    return _draw_ampls(nx.random_regular_graph(degree, nodes_number, rng), rng, node_ampl_func, edge_ampl_func)


def generate_qubo_on_2d_grid(
        m: int, n: int, seed: int = 42,
        node_ampl_func: Callable[[Random, int], float] = _uniform_ampl,
        edge_ampl_func: Callable[[Random, Edge], float] = _uniform_ampl):
    import networkx as nx
    rng = Random(seed)
    return _draw_ampls(nx.grid_2d_graph(m, n), rng, node_ampl_func, edge_ampl_func)


def heavy_hex_127(seed: int = 42):
    """The 127-qubit heavy-hex lattice of reference examples/full_size_ibm_heavy_hex.py:15-78 with +-1
    amplitudes: node amplitudes are drawn first (ids 0..126), then edge amplitudes in the listed order."""
    rng = Random(seed)
    pm = lambda: 2 * rng.randint(0, 1) - 1
    nodes = {i: pm() for i in range(127)}
    rows = [range(13), range(18, 32), range(37, 51), range(56, 70), range(75, 89), range(94, 108), range(113, 126)]
    edges = {}
    for r in rows:
        for i in r:
            edges[(i, i + 1)] = pm()
    # bridge qubits between rows: (upper, bridge, lower)
    bridges = [(0, 14, 18), (4, 15, 22), (8, 16, 26), (12, 17, 30),
               (20, 33, 39), (24, 34, 43), (28, 35, 47), (32, 36, 51),
               (37, 52, 56), (41, 53, 60), (45, 54, 64), (49, 55, 68),
               (58, 71, 77), (62, 72, 81), (66, 73, 85), (70, 74, 89),
               (75, 90, 94), (79, 91, 98), (83, 92, 102), (87, 93, 106),
               (96, 109, 114), (100, 110, 118), (104, 111, 122), (108, 112, 126)]
    for up, br, lo in bridges:
        edges[(up, br)] = pm()
        edges[(br, lo)] = pm()
    return nodes, edges


def ising_energy(edges, nodes, spins) -> float:
    """E(s) = sum J_ij s_i s_j + sum h_i s_i, s_i = +1 for bit 0 / Bloch z > 0."""
    e = 0.0
    for (l, r), j in (edges.items() if isinstance(edges, dict) else edges):
        e += j * spins[l] * spins[r]
    for n, h in (nodes.items() if isinstance(nodes, dict) else nodes):
        e += h * spins[n]
    return float(e)
