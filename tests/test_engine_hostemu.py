"""CPU checks of the kernel math (bqa_b200/csrc/bqa_core.cuh) and of the whole Python engine, run through a
TEST-ONLY host emulation of the C ABI (tests/hostemu) and compared with the oracle and the reference goldens.
The GPU parity tests proper are in tests/test_gpu_parity.py (-m gpu)."""
import os

import numpy as np
import pytest

import instances
from bqa_b200 import _lib
from bqa_b200.config import config_to_context
from bqa_b200.core import run_context
from bqa_b200.engine import Engine
from hostemu.build import build as build_hostemu
from oracle import bqa_oracle as O


@pytest.fixture(scope="module")
def emu():
    return _lib.bind(build_hostemu())


def _run(cfg, emu, precision="double"):
    ctx = config_to_context(cfg)
    eng_holder = {}

    class Probe(Engine):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            eng_holder["e"] = self

    res = dict(run_context(ctx, precision=precision, engine_cls=Probe, _testing_lib=emu))
    return res, eng_holder["e"]


@pytest.mark.parametrize("name", list(instances.GOLDEN_CONFIGS))
def test_engine_matches_reference_goldens_double(golden_dir, emu, name):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    res, eng = _run(instances.GOLDEN_CONFIGS[name](), emu, "double")
    assert np.abs(np.array(res["bloch_vectors"]) - g["bloch"]).max() < 1e-8
    n = len(g["bond_dims"])
    assert eng.stats["bond_dims"] == g["bond_dims"].tolist()
    assert eng.stats["bp_sweeps"][:n] == g["bp_sweeps"].tolist()
    if "outcomes" in g:
        assert res["measurement_outcomes"] == g["outcomes"].tolist()


@pytest.mark.parametrize("name", ["ring24", "grid4"])
def test_engine_single_precision_tolerance(golden_dir, emu, name):
    # fp32 tolerance from BASELINE.md section 4: <= 5e-3 max-abs, <= 1e-4 mean-abs on Bloch components
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    res, eng = _run(instances.GOLDEN_CONFIGS[name](), emu, "single")
    diff = np.abs(np.array(res["bloch_vectors"]) - g["bloch"])
    assert diff.max() < 5e-3 and diff.mean() < 1e-4
    assert eng.stats["bond_dims"] == g["bond_dims"].tolist()
    assert np.abs(np.array(eng.stats["bp_sweeps"][:len(g["bp_sweeps"])]) - g["bp_sweeps"]).max() <= 1


def test_lmbd_spectra_match_reference(golden_dir, emu):
    g = np.load(os.path.join(golden_dir, "ring24.npz"))
    cfg = instances.cfg_ring24()
    cfg["schedule"]["actions"] = cfg["schedule"]["actions"][:1]
    _, eng = _run(cfg, emu, "double")
    lm = np.sort(eng.lmbds_numpy(), axis=1)[:, ::-1]
    assert np.abs(lm - g["lmbds_last"]).max() < 1e-8


@pytest.mark.parametrize("d,D", [(1, 3), (2, 4), (3, 4), (3, 2), (4, 3)])
def test_single_sweep_against_reference_kernel_goldens(golden_dir, emu, d, D):
    """One BP sweep / ext-message pass / marginal of a degree class on seeded inputs: elementwise parity
    (these stages involve no SVD, hence no gauge freedom)."""
    g = np.load(os.path.join(golden_dir, "kernel_level.npz"))
    B = 5
    t, msgs, thetas = instances.random_node_batch(B, d, D, seed=100 + 10 * d + D)
    import torch
    prec = _lib.C128
    T = torch.from_numpy(np.ascontiguousarray(t).reshape(-1))
    cur = torch.from_numpy(np.ascontiguousarray(np.concatenate(msgs, 0)).reshape(-1))     # slot j*B + i = message j of node i
    in_pos = torch.from_numpy(np.arange(d * B, dtype=np.int32).reshape(d, B))
    out_pos = in_pos.clone()
    nxt = torch.zeros_like(cur)
    resid = torch.zeros(2, dtype=torch.float64)
    status = torch.zeros(4, dtype=torch.int32)
    emu.bp_sweep(prec, d, D, B, T.data_ptr(), cur.data_ptr(), nxt.data_ptr(), in_pos.data_ptr(), out_pos.data_ptr(),
                 0.0, 0, 1e-6, 0, resid.data_ptr(), status.data_ptr(), 0, 0, 0)
    got = nxt.numpy().reshape(d, B, D, D)
    assert np.abs(got - g[f"pass_d{d}_D{D}"]).max() < 1e-12
    new, old = g[f"pass_d{d}_D{D}"], np.stack(msgs)
    r_ = resid.numpy()
    assert np.isclose(np.sqrt(r_[0] / r_[1]), np.abs(new - old).max() / np.abs(new + old).max(), rtol=1e-10)
    ext = torch.zeros(d * B * 4 * D * D, dtype=torch.complex128)
    ea = torch.from_numpy(np.ascontiguousarray(np.stack(thetas)))
    emu.ext_msgs(prec, d, D, B, T.data_ptr(), cur.data_ptr(), ext.data_ptr(), in_pos.data_ptr(), out_pos.data_ptr(),
                 ea.data_ptr(), 1.0, 0, 0, 0)
    assert np.abs(ext.numpy().reshape(d, B, 2 * D, 2 * D) - g[f"ext_d{d}_D{D}"]).max() < 1e-12
    bloch = torch.zeros(B * 4, dtype=torch.float64)
    ids = torch.arange(B, dtype=torch.int32)
    emu.density(prec, d, D, B, T.data_ptr(), cur.data_ptr(), in_pos.data_ptr(), ids.data_ptr(), bloch.data_ptr(), 0, 0, 0)
    assert np.abs(bloch.numpy().reshape(B, 4)[:, :3] - O.bloch_vectors(g[f"rho_d{d}_D{D}"])).max() < 1e-12


def test_jacobi_canonicalizers_are_gauge_equivalent_to_lapack(emu):
    """The canonicalizers depend on SVD conventions, their gauge-invariant content does not: lambdas and the
    products C_f^T-contracted projector.  Checked on random PSD extended messages (n = 8)."""
    import torch
    rng = np.random.default_rng(5)
    L, D = 7, 4
    n = 2 * D
    ext = instances.random_psd_msgs(rng, 2 * L, n)
    lm_ref, canon_ref = O.canonicalizers(ext, 1e-9, np.complex128)
    e = torch.from_numpy(np.ascontiguousarray(ext).reshape(-1))
    canon = torch.zeros_like(e)
    lm = torch.zeros(L * n, dtype=torch.float64)
    colmax = torch.zeros(n, dtype=torch.float64)
    emu.canonicalize(_lib.C128, D, L, e.data_ptr(), canon.data_ptr(), lm.data_ptr(), colmax.data_ptr(), 1e-9, n, 0)
    lm = lm.numpy().reshape(L, n)
    assert np.abs(lm - lm_ref.real).max() < 1e-10
    assert np.abs(colmax.numpy() - lm_ref.real.max(0)).max() < 1e-10
    c = canon.numpy().reshape(2 * L, n, n)
    # gauge-invariant combination: sum_k C_b[a, k] lambda_k C_f[b, k]  (column phases cancel between the two)
    inv = np.einsum("eak,ek,ebk->eab", c[:L], lm, c[L:])
    inv_ref = np.einsum("eak,ek,ebk->eab", canon_ref[:L], lm_ref.real, canon_ref[L:])
    assert np.abs(inv - inv_ref).max() < 1e-8


def test_checkpoint_roundtrip_and_oracle_state_injection(emu):
    """load_state(oracle state) followed by one layer equals the oracle's next layer (gauge-invariant check)."""
    cfg = instances.cfg_ring24()
    cfg["schedule"]["actions"] = cfg["schedule"]["actions"][:1]
    octx = O.compile_config(cfg)
    ost = O.init_state(octx)
    layers = [i for i in octx.instructions if isinstance(i, dict)]
    for ins in layers[:8]:
        O.run_layer(octx, ost, ins["xtime"], ins["ztime"])
    eng = Engine(config_to_context(cfg), precision="double", _testing_lib=emu)
    eng.load_state({"D": ost.bond_dim, "tensors": ost.tensors, "msgs": ost.msgs, "lmbds": ost.lmbds})
    snap = eng.state_to_host()
    assert np.abs(snap["msgs"] - ost.msgs).max() == 0 and snap["D"] == ost.bond_dim
    assert np.abs(eng.bloch_vectors() - O.bloch_vectors(O.density_matrices(octx, ost))).max() < 1e-12
    ins = layers[8]
    O.run_layer(octx, ost, ins["xtime"], ins["ztime"])
    eng.run_layer(ins["xtime"], ins["ztime"])
    assert eng.D == ost.bond_dim
    assert np.abs(eng.bloch_vectors() - O.bloch_vectors(O.density_matrices(octx, ost))).max() < 1e-9


def test_product_loader_refuses_host_emulation(emu, tmp_path, monkeypatch):
    """The package's own loader must never hand out the emulation: it is test infrastructure."""
    import shutil
    fake = tmp_path / "libbqa_b200.so"
    shutil.copy(emu.path, fake)
    monkeypatch.setattr(_lib, "LIB_PATH", str(fake))
    monkeypatch.setattr(_lib, "_cached", None)
    with pytest.raises(RuntimeError, match="not the CUDA build"):
        _lib.load_library()


def test_rank_deficient_ker_keeps_relative_accuracy(emu):
    """Extended messages whose masked ranks differ (4 vs 5 of 8) give a rank-deficient ker with more non-zero
    columns than its rank; the kept lambdas must still come out to high relative accuracy.  (The Jacobi SVD
    leaves numerically-zero columns alone; test_random_regular_400_lockstep_with_oracle is the regression test
    for the case where rotating such a column into the denormal range rescaled the smallest kept lambda.)"""
    import torch
    rng = np.random.default_rng(11)
    n, L = 8, 40
    spec_f = np.array([0.97, 2.4e-2, 5e-4, 1.3e-5, 7.9e-7, 3.5e-8, 9.5e-9, 3.2e-10])
    spec_b = np.array([0.97, 2.3e-2, 5e-4, 2.2e-5, 8.3e-6, 2.9e-7, 2.5e-8, 6.4e-10])

    def psd(spec):
        q, _ = np.linalg.qr(rng.normal(size=(L, n, n)) + 1j * rng.normal(size=(L, n, n)))
        # graded eigenvectors: close to the identity, like the extended messages of a weakly entangled bond
        q = np.linalg.qr(np.eye(n) + 0.05 * q)[0]
        m = (q * (spec / spec.sum())) @ np.swapaxes(q.conj(), 1, 2)
        return 0.5 * (m + np.swapaxes(m.conj(), 1, 2))
    ext = np.concatenate([psd(spec_f), psd(spec_b)], 0)
    lm_ref, _ = O.canonicalizers(ext, 1e-6, np.complex128)
    e = torch.from_numpy(np.ascontiguousarray(ext).reshape(-1))
    canon = torch.zeros_like(e)
    lm = torch.zeros(L * n, dtype=torch.float64)
    colmax = torch.zeros(n, dtype=torch.float64)
    emu.canonicalize(_lib.C128, 4, L, e.data_ptr(), canon.data_ptr(), lm.data_ptr(), colmax.data_ptr(), 1e-6, n, 0)
    lm = lm.numpy().reshape(L, n)
    assert (lm_ref.real[:, 3] > 0).all() and (lm_ref.real[:, 4] == 0).all()
    rel = np.abs(lm[:, :4] - lm_ref.real[:, :4]) / lm_ref.real[:, :4]
    assert rel.max() < 1e-9


def test_random_regular_400_lockstep_with_oracle(emu):
    """30 steps of the dt = 0.2 schedule on a seeded random 3-regular QUBO, engine and oracle in lock step:
    identical bond dimensions and sweep counts, Bloch vectors and lambda spectra to 1e-10 after every step."""
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    nodes, edges = generate_qubo_on_random_regular_graph(400, 3, seed=42)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 4,
           "schedule": {"total_time": 6.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 30, "final_mixing": 0.0}]}}
    ctx = config_to_context(cfg)
    eng = Engine(ctx, precision="double", _testing_lib=emu)
    octx = O.compile_config(cfg)
    ost = O.init_state(octx)
    for ins in [i for i in ctx.instructions if isinstance(i, dict)]:
        eng.run_layer(ins["xtime"], ins["ztime"])
        O.run_layer(octx, ost, ins["xtime"], ins["ztime"])
        assert eng.D == ost.bond_dim
        assert eng.stats["bp_sweeps"][-1] == ost.stats["bp_sweeps"][-1]
        lm = np.sort(eng.lmbds_numpy(), axis=1)[:, ::-1]
        olm = np.sort(np.real(ost.lmbds), axis=1)[:, ::-1]
        assert np.abs(lm - olm).max() < 1e-10
        assert np.abs(eng.bloch_vectors() - O.bloch_vectors(O.density_matrices(octx, ost))).max() < 1e-10


def test_isolated_qubit_follows_single_qubit_evolution(emu):
    """Degree-0 class (edge case the reference compiles but cannot run): exact single-qubit result, and the sampler
    terminates with an outcome for every qubit."""
    cfg = instances.cfg_isolated_qubit()
    res, _ = _run(cfg, emu, "double")
    assert np.abs(np.array(res["bloch_vectors"])[3] - instances.isolated_qubit_bloch(cfg)).max() < 1e-12
    assert sorted(set(res["measurement_outcomes"])) <= [-1, 1] and len(res["measurement_outcomes"]) == 4


def test_checkpoint_resume_continues_bit_for_bit(tmp_path, emu):
    """An interrupted run resumed from its checkpoint returns exactly what the uninterrupted run returns: Bloch
    vectors, sampled bitstring (host RNG state is part of the checkpoint), bond dimensions (core.save_checkpoint)."""
    cfg = instances.cfg_grid4()                             # 16 steps, then get_bloch_vectors, measure
    ctx = config_to_context(cfg)
    want = run_context(ctx, precision="double", _testing_lib=emu)
    path = str(tmp_path / "anneal.npz")

    class Crash(RuntimeError):
        pass

    class Flaky(Engine):
        def run_layer(self, xtime, ztime, next_ztime=None):
            if len(self.stats["bond_dims"]) == 11:
                raise Crash("power cut")
            return super().run_layer(xtime, ztime, next_ztime=next_ztime)

    with pytest.raises(Crash):
        run_context(ctx, precision="double", engine_cls=Flaky, checkpoint=path, checkpoint_every=4, _testing_lib=emu)
    assert os.path.exists(path)
    got = run_context(ctx, precision="double", checkpoint=path, checkpoint_every=4, resume=True, _testing_lib=emu)
    assert got == want                                      # lists of Python floats / ints: exact equality
    # a checkpoint of another schedule or precision is refused
    with pytest.raises(ValueError):
        run_context(ctx, precision="single", checkpoint=path, resume=True, _testing_lib=emu)
    short = config_to_context(instances.cfg_ring24())
    with pytest.raises(ValueError):
        run_context(short, precision="double", checkpoint=path, resume=True, _testing_lib=emu)
