"""Pins the CPU oracle (oracle/bqa_oracle.py) to the reference:

* committed outputs of the unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py);
* the reference's own known-answer properties: BP is exact on a tree and symmetric-gauge invariance
  (reference tests/test_core_subroutines.py:191-256), gate identities
  (reference tests/test_gpt_generated_gates_application.py:26-89);
* an independent state-vector simulation (reference tests/test_small_circuit_final_density.py:17-20
  with exact_sim.py restated in numpy because qem is not installable here)."""
import os

import numpy as np
import pytest

import instances
from oracle import bqa_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, f"{name}.npz"))


@pytest.mark.parametrize("name", list(instances.GOLDEN_CONFIGS))
def test_oracle_matches_reference_run(golden_dir, name):
    g = _load(golden_dir, name)
    res, ctx, st = O.run_qa(instances.GOLDEN_CONFIGS[name](), return_state=True)
    res = dict(res)
    assert np.abs(np.array(res["bloch_vectors"]) - g["bloch"]).max() < 1e-10
    if "outcomes" in g:
        assert res["measurement_outcomes"] == g["outcomes"].tolist()
    n_layers = len(g["bond_dims"])
    assert st.stats["bond_dims"] == g["bond_dims"].tolist()
    assert st.stats["bp_sweeps"][:n_layers] == g["bp_sweeps"].tolist()


@pytest.mark.parametrize("name", ["ring24", "comb"])
def test_oracle_lmbd_spectra(golden_dir, name):
    g = _load(golden_dir, name)
    cfg = instances.GOLDEN_CONFIGS[name]()
    cfg["schedule"]["actions"] = cfg["schedule"]["actions"][:1]
    _, ctx, st = O.run_qa(cfg, return_state=True)
    lm = np.sort(st.lmbds.real, axis=1)[:, ::-1]
    assert np.abs(lm - g["lmbds_last"]).max() < 1e-10


@pytest.mark.parametrize("d,D", [(1, 3), (2, 4), (3, 4), (3, 2), (4, 3)])
def test_oracle_kernel_level(golden_dir, d, D):
    g = _load(golden_dir, "kernel_level")
    t, msgs, thetas = instances.random_node_batch(5, d, D, seed=100 + 10 * d + D)
    plain = np.stack(O.pass_msgs(t, msgs))
    ext = np.stack(O.pass_msgs(t, msgs, [x.astype(np.complex128) for x in thetas]))
    rho = O.density_of_class(t, msgs)
    assert np.abs(plain - g[f"pass_d{d}_D{D}"]).max() < 1e-13
    assert np.abs(ext - g[f"ext_d{d}_D{D}"]).max() < 1e-13
    assert np.abs(rho - g[f"rho_d{d}_D{D}"]).max() < 1e-13


def test_oracle_vs_statevector_small6():
    # reference tests/test_small_circuit_final_density.py:17-20, tolerance 1e-5
    cfg = instances.cfg_small6()
    bloch = np.array(O.run_qa(cfg)[0][1])
    exact = O.run_exact_statevector(cfg)
    assert np.abs(bloch - exact).max() < 1e-5


def test_survey_golden_bloch_small6():
    # Bloch vectors printed from the numpy backend (double) for this config
    ref = np.array([[0.0181461526, -0.0126274979, -0.9926433340], [-0.0136529332, 0.0122463286, 0.9705519477],
                    [0.0920639996, 0.1934531474, 0.9381713597], [0.0394042819, 0.0044847983, -0.8980958243],
                    [-0.0363240975, -0.0205751776, -0.9640037751], [0.1626526209, -0.3086531788, 0.8279086032]])
    bloch = np.array(O.run_qa(instances.cfg_small6())[0][1])
    assert np.abs(bloch - ref).max() < 1e-9


TREE_CONFIG = {   # reference tests/test_core_subroutines.py:176-189
    "nodes": {1: 0.3, 3: -0.7, 5: 1., 6: -1., 7: 0.25},
    "edges": {(2, 0): 1., (1, 2): -1., (2, 4): 0.5, (4, 3): -0.5, (4, 5): 0.75, (4, 6): -0.75, (6, 7): 0.3, (8, 6): 0.6},
    "bp_eps": 1e-10, "pinv_eps": 1e-7, "default_field": 0.6,
}


def _randomize(st, rng):
    for d, t in st.tensors.items():
        new = rng.normal(size=t.shape) + 1j * rng.normal(size=t.shape)
        st.tensors[d] = (new / np.linalg.norm(new)).astype(t.dtype)


def _exact_tree_density(ctx, st):
    """Contracts the whole 9-qubit tree state and returns exact single-qubit marginals."""
    n = ctx.nodes_number
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    bond = {}
    nxt = n
    subs, ops = [], []
    for node in range(n):
        d, pos = ctx.path[node]
        s = letters[node]
        for m in ctx.graph[node]:
            key = (min(node, m), max(node, m))
            if key not in bond:
                bond[key] = letters[nxt]
                nxt += 1
            s += bond[key]
        subs.append(s)
        ops.append(st.tensors[d][pos])
    psi = np.einsum(",".join(subs) + "->" + letters[:n], *ops)
    psi = psi / np.linalg.norm(psi)
    rho = np.empty((n, 2, 2), complex)
    for q in range(n):
        m = np.moveaxis(psi, q, 0).reshape(2, -1)
        rho[q] = m @ m.conj().T
    return psi, rho


def test_bp_exact_on_tree():
    # reference tests/test_core_subroutines.py:207-220
    ctx = O.compile_config(TREE_CONFIG)
    st = O.init_state(ctx)
    # random tensors with bond dimension 2 on every leg
    rng = np.random.default_rng(42)
    for d, t in list(st.tensors.items()):
        st.tensors[d] = np.zeros((t.shape[0], 2) + (2,) * d, complex)
    _randomize(st, rng)
    st.lmbds = np.ones((ctx.lmbds_number, 2), complex)
    st.msgs = O.msgs_from_lmbds(st.lmbds, ctx)
    O.run_bp(ctx, st)
    _, rho_exact = _exact_tree_density(ctx, st)
    rho = O.density_matrices(ctx, st)
    assert np.abs(rho - rho_exact).max() < 1e-8


def test_gate_identities():
    # reference tests/test_gpt_generated_gates_application.py:26-64
    ctx = O.compile_config(TREE_CONFIG)
    st = O.init_state(ctx)
    before = {d: t.copy() for d, t in st.tensors.items()}
    O.x_layer(st, np.pi / 2)                               # exp(-i pi/2 X) = -i X ; X|-> = -|->
    for d in before:
        assert np.allclose(st.tensors[d], 1j * before[d])
    O.x_layer(st, np.pi / 2)
    for d in before:
        assert np.allclose(st.tensors[d], -before[d])
    O.z_layer(ctx, st, 0.0)
    for d in before:
        assert np.allclose(st.tensors[d], -before[d])


def test_first_step_keeps_bond_dim_one():
    # ztime = 0 in the first step => second block is exactly zero
    cfg = instances.cfg_ring24()
    _, ctx, st = O.run_qa({**cfg, "schedule": {"total_time": 4.0, "actions": [
        {"weight": 1.0, "steps_number": 20, "final_mixing": 0.0}]}}, return_state=True)
    assert st.stats["bond_dims"][0] == 1 and st.stats["bond_dims"][1] == 2
