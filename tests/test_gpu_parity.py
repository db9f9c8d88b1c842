"""GPU parity tests (-m gpu): the CUDA path, reached through the C ABI, against the reference goldens,
against the oracle on seeded instances, and through size-independent properties at full size."""
import os

import numpy as np
import pytest

import instances

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from bqa_b200 import _lib
    from bqa_b200.build import build
    build()
    lib = _lib.load_library()
    assert lib.version() > 0, "the CUDA build must be the one that is loaded"
    return lib


def _run(cfg, precision):
    from bqa_b200.config import config_to_context
    from bqa_b200.core import run_context
    from bqa_b200.engine import Engine
    holder = {}

    class Probe(Engine):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            holder["e"] = self

    res = dict(run_context(config_to_context(cfg), precision=precision, engine_cls=Probe))
    return res, holder["e"]


@pytest.mark.parametrize("name", list(instances.GOLDEN_CONFIGS))
def test_goldens_double(golden_dir, lib, name):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    before = lib.launch_count()
    res, eng = _run(instances.GOLDEN_CONFIGS[name](), "double")
    assert lib.launch_count() > before
    assert np.abs(np.array(res["bloch_vectors"]) - g["bloch"]).max() < 1e-8
    n = len(g["bond_dims"])
    assert eng.stats["bond_dims"] == g["bond_dims"].tolist()
    assert eng.stats["bp_sweeps"][:n] == g["bp_sweeps"].tolist()
    if "outcomes" in g:
        assert res["measurement_outcomes"] == g["outcomes"].tolist()     # same host RNG stream => same bitstring


@pytest.mark.parametrize("name", ["ring24", "grid4", "comb"])
def test_goldens_single(golden_dir, lib, name):
    # fp32 tolerance: <= 5e-3 max-abs and <= 5e-4 mean-abs on Bloch components.
    # (small6 is left out: its pinv_eps = 1e-9 rank cut is below complex64 resolution.)
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    res, eng = _run(instances.GOLDEN_CONFIGS[name](), "single")
    diff = np.abs(np.array(res["bloch_vectors"]) - g["bloch"])
    assert diff.max() < 5e-3 and diff.mean() < 5e-4
    assert eng.stats["bond_dims"] == g["bond_dims"].tolist()


@pytest.mark.parametrize("prec_name", ["double", "single"])
@pytest.mark.parametrize("d,D", [(1, 3), (2, 4), (3, 4), (3, 2), (4, 3)])
def test_kernel_level_goldens(golden_dir, lib, d, D, prec_name):
    import torch
    from bqa_b200 import _lib
    from oracle import bqa_oracle as O
    g = np.load(os.path.join(golden_dir, "kernel_level.npz"))
    dbl = prec_name == "double"
    prec = _lib.C128 if dbl else _lib.C64
    cdt, rdt = (np.complex128, np.float64) if dbl else (np.complex64, np.float32)
    tol = 1e-12 if dbl else 2e-5
    B = 5
    t, msgs, thetas = instances.random_node_batch(B, d, D, seed=100 + 10 * d + D)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    T = up(t.astype(cdt).reshape(-1))
    cur = up(np.concatenate(msgs, 0).astype(cdt).reshape(-1))
    in_pos = up(np.arange(d * B, dtype=np.int32).reshape(d, B))
    nxt = torch.zeros_like(cur)
    resid = torch.zeros(2, dtype=torch.float64 if dbl else torch.float32, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = torch.zeros(lib.workspace_bytes(prec, d, D, D), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lib.bp_sweep(prec, d, D, B, T.data_ptr(), cur.data_ptr(), nxt.data_ptr(), in_pos.data_ptr(), in_pos.data_ptr(),
                 0.0, 0, 1e-6, 0, resid.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), st)
    got = nxt.cpu().numpy().reshape(d, B, D, D)
    assert np.abs(got - g[f"pass_d{d}_D{D}"]).max() < tol
    new, old = g[f"pass_d{d}_D{D}"], np.stack(msgs)
    r = resid.cpu().numpy().astype(np.float64)
    assert np.isclose(np.sqrt(r[0] / r[1]), np.abs(new - old).max() / np.abs(new + old).max(), rtol=1e-9 if dbl else 1e-4)
    ext = torch.zeros(d * B * 4 * D * D, dtype=T.dtype, device=dev)
    ea = up(np.stack(thetas).astype(rdt))
    lib.ext_msgs(prec, d, D, B, T.data_ptr(), cur.data_ptr(), ext.data_ptr(), in_pos.data_ptr(), in_pos.data_ptr(),
                 ea.data_ptr(), 1.0, ws.data_ptr(), ws.numel(), st)
    assert np.abs(ext.cpu().numpy().reshape(d, B, 2 * D, 2 * D) - g[f"ext_d{d}_D{D}"]).max() < tol
    bloch = torch.zeros(B * 4, dtype=resid.dtype, device=dev)
    ids = torch.arange(B, dtype=torch.int32, device=dev)
    lib.density(prec, d, D, B, T.data_ptr(), cur.data_ptr(), in_pos.data_ptr(), ids.data_ptr(), bloch.data_ptr(),
                ws.data_ptr(), ws.numel(), st)
    assert np.abs(bloch.cpu().numpy().reshape(B, 4)[:, :3] - O.bloch_vectors(g[f"rho_d{d}_D{D}"])).max() < tol


def _rr_config(n, steps, total_time, tail, seed=42, **extra):
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    nodes, edges = generate_qubo_on_random_regular_graph(n, 3, seed=seed)
    return {"nodes": nodes, "edges": edges, "max_bond_dim": 4,
            "schedule": {"total_time": total_time, "starting_mixing": 1.0,
                         "actions": [{"weight": 1.0, "steps_number": steps, "final_mixing": 0.0}, *tail]}, **extra}


def test_random_regular_2000_vs_oracle(lib):
    """Seeded 3-regular QUBO, dt = 0.2 like the 100k benchmark script, against the oracle (complex128)."""
    from bqa_b200.benchmarking import ising_energy
    from oracle import bqa_oracle as O
    cfg = _rr_config(2000, 30, 6.0, ["get_bloch_vectors"])
    want, octx, ost = O.run_qa(cfg, return_state=True)
    want = np.array(want[0][1])
    res, eng = _run(cfg, "double")
    got = np.array(res["bloch_vectors"])
    assert np.abs(got - want).max() < 1e-7
    assert eng.stats["bond_dims"] == ost.stats["bond_dims"]
    assert eng.stats["bp_sweeps"] == ost.stats["bp_sweeps"][:len(eng.stats["bp_sweeps"])]
    res32, eng32 = _run(cfg, "single")
    got32 = np.array(res32["bloch_vectors"])
    diff = np.abs(got32 - want)
    assert diff.max() < 5e-3 and diff.mean() < 1e-4                     # stated fp32 tolerance
    assert eng32.stats["bond_dims"] == ost.stats["bond_dims"]
    assert np.abs(np.array(eng32.stats["bp_sweeps"]) - np.array(ost.stats["bp_sweeps"])).max() <= 1
    s_ref = np.where(want[:, 2] > 0, 1, -1)
    s_32 = np.where(got32[:, 2] > 0, 1, -1)
    e_ref = ising_energy(cfg["edges"], cfg["nodes"], s_ref)
    e_32 = ising_energy(cfg["edges"], cfg["nodes"], s_32)
    assert abs(e_32 - e_ref) <= 1e-4 * abs(e_ref)                         # 1e-4 relative on the energy


def test_full_size_properties_100k(lib):
    """100k-qubit 3-regular QUBO (BASELINE config 4), fp32: invariants that need no oracle run."""
    import torch
    cfg = _rr_config(100_000, 30, 6.0, [])
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    eng = Engine(config_to_context(cfg), precision="single")
    layers = [i for i in eng.ctx.instructions if isinstance(i, dict)]
    for ins in layers:
        eng.run_layer(ins["xtime"], ins["ztime"])
    assert eng.D == 4
    D = eng.D
    m = eng.msgs_buffer[: eng.E2 * D * D].view(eng.E2, D, D)
    tr = torch.diagonal(m, dim1=1, dim2=2).sum(1)
    assert float((tr - 1).abs().max()) < 1e-5                             # messages are trace-normalised
    assert float((m - m.transpose(1, 2).conj()).abs().max()) < 1e-5       # ... and Hermitian
    c = eng.classes[0]
    t = c.T[c.cur][: c.B * 2 * D ** 3].view(c.B, -1)
    assert float((torch.linalg.vector_norm(t, dim=1) - 1).abs().max()) < 1e-5   # node tensors are L2-normalised
    lm = eng.lmbds_numpy()
    assert np.all(lm >= 0) and np.all(np.diff(lm, axis=1) <= 1e-6)              # lambdas sorted descending
    # BP fixed point: one more run converges immediately and leaves the marginals unchanged
    b0 = eng.bloch_vectors()
    sweeps = eng.run_bp()
    assert sweeps <= 2
    assert np.abs(eng.bloch_vectors() - b0).max() < 1e-4
    assert np.all(np.abs(b0) <= 1 + 1e-5) and np.all(np.linalg.norm(b0, axis=1) <= 1 + 1e-4)
    # checkpoint round trip is exact
    snap = eng.state_to_host()
    eng2 = Engine(eng.ctx, precision="single")
    eng2.load_state(snap)
    assert np.abs(eng2.bloch_vectors() - eng.bloch_vectors()).max() == 0.0


@pytest.mark.parametrize("B", [1, 3, 4, 5, 7, 1003, 40_001])     # < 4: generic fallback; not a multiple of 4: overlapping last group
def test_fast_d3D4_kernels_match_generic_and_oracle(lib, B):
    """Specialised degree-3 / D=4 / complex64 kernels (bqa_fast_d3D4.cu) against the generic kernels and
    against the oracle's pass_msgs (reference backends.py:406-408) on scattered message slots, a ragged
    batch (B not a multiple of the 4 nodes a warp handles) and non-zero damping."""
    import torch
    from bqa_b200 import _lib
    from oracle import bqa_oracle as O
    d, D = 3, 4
    rng = np.random.default_rng(B)
    t, msgs, thetas = instances.random_node_batch(B, d, D, seed=7 + B)
    slots = rng.permutation(d * B + 5)                      # a few unused slots
    in_pos = slots[: d * B].reshape(d, B).astype(np.int32)
    out_pos = rng.permutation(d * B + 5)[: d * B].reshape(d, B).astype(np.int32)
    cur = np.zeros((d * B + 5, D, D), np.complex64)
    cur[:] = instances.random_psd_msgs(rng, d * B + 5, D)
    for j in range(d):
        cur[in_pos[j]] = msgs[j]
    msgs32 = [cur[in_pos[j]].astype(np.complex128) for j in range(d)]
    t32 = t.astype(np.complex64)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    T, C, ip, op = up(t32.reshape(-1)), up(cur.reshape(-1)), up(in_pos), up(out_pos)
    ea = up(np.stack(thetas).astype(np.float32))
    ws = torch.zeros(lib.workspace_bytes(_lib.C64, d, D, D), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for mode in (1, 0):
        lib.set_kernel_mode(mode)
        try:
            nxt = C.clone()
            resid = torch.zeros(2, dtype=torch.float32, device=dev)
            status = torch.zeros(4, dtype=torch.int32, device=dev)
            lib.bp_sweep(_lib.C64, d, D, B, T.data_ptr(), C.data_ptr(), nxt.data_ptr(), ip.data_ptr(), op.data_ptr(),
                         0.25, 0, 1e-6, 0, resid.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), st)
            ext = torch.zeros((d * B + 5) * 4 * D * D, dtype=torch.complex64, device=dev)
            lib.ext_msgs(_lib.C64, d, D, B, T.data_ptr(), C.data_ptr(), ext.data_ptr(), ip.data_ptr(), op.data_ptr(),
                         ea.data_ptr(), 0.7, ws.data_ptr(), ws.numel(), st)
            out[mode] = (nxt.cpu().numpy().reshape(-1, D, D), resid.cpu().numpy().astype(np.float64),
                         ext.cpu().numpy().reshape(-1, 2 * D, 2 * D))
        finally:
            lib.set_kernel_mode(0)
    gen, fast = out[1], out[0]
    assert np.abs(fast[0] - gen[0]).max() < 2e-6
    assert np.abs(fast[2] - gen[2]).max() < 2e-6
    assert np.allclose(np.sqrt(fast[1][0] / fast[1][1]), np.sqrt(gen[1][0] / gen[1][1]), rtol=1e-4)
    # untouched slots keep their content
    untouched = np.setdiff1d(np.arange(d * B + 5), out_pos.reshape(-1))
    assert np.array_equal(fast[0][untouched], cur[untouched])
    # oracle (complex128) on the same complex64 inputs
    want = O.pass_msgs(t32.astype(np.complex128), msgs32)
    want_ext = O.pass_msgs(t32.astype(np.complex128), msgs32, [(0.7 * th.astype(np.float32).astype(np.float64)).astype(np.complex128) for th in thetas])   # complex like the reference's edge_ampls (principal roots)
    for j in range(d):
        damped = 0.25 * cur[out_pos[j]] + 0.75 * want[j]
        assert np.abs(fast[0][out_pos[j]] - damped).max() < 2e-6
        assert np.abs(fast[2][out_pos[j]] - want_ext[j]).max() < 2e-6


@pytest.mark.parametrize("B", [1, 5, 130, 1003])
def test_fast_gram_d3D8_matches_generic_and_oracle(lib, B):
    """Degree 3, D = 8, complex64: the shared-memory node contraction (bqa_fast_gram_d3D8.cuh, taken by the table-driven
    kernels) against the generic contraction (kernel mode 1) and against the oracle's pass_msgs (reference
    backends.py:381-408) -- BP messages of a whole run of 2 sweeps incl. residuals, and the ZZ-extended messages -- on a
    table of two classes (degree 2 rides along on the generic path) with scattered message slots."""
    import ctypes as C
    import torch
    from bqa_b200 import _lib
    from oracle import bqa_oracle as O
    D = 8
    rng = np.random.default_rng(100 + B)
    B2 = 3
    nslots = 3 * B + 2 * B2 + 5
    perm_in, perm_out = rng.permutation(nslots), rng.permutation(nslots)
    in3, out3 = perm_in[: 3 * B].reshape(3, B).astype(np.int32), perm_out[: 3 * B].reshape(3, B).astype(np.int32)
    in2 = perm_in[3 * B: 3 * B + 2 * B2].reshape(2, B2).astype(np.int32)
    out2 = perm_out[3 * B: 3 * B + 2 * B2].reshape(2, B2).astype(np.int32)
    t3, _, th3 = instances.random_node_batch(B, 3, D, seed=11 + B)
    t2, _, th2 = instances.random_node_batch(B2, 2, D, seed=12 + B)
    cur = instances.random_psd_msgs(rng, nslots, D).astype(np.complex64)
    t3, t2 = t3.astype(np.complex64), t2.astype(np.complex64)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    T3, T2, Cm = up(t3.reshape(-1)), up(t2.reshape(-1)), up(cur.reshape(-1))
    i3, o3, i2, o2 = up(in3), up(out3), up(in2), up(out2)
    e3, e2 = up(np.stack(th3).astype(np.float32)), up(np.stack(th2).astype(np.float32))
    rows = (_lib.ClassDesc * 2)()
    for r, (deg, b, T, ip, op, ea) in zip(rows, ((2, B2, T2, i2, o2, e2), (3, B, T3, i3, o3, e3))):
        r.degree, r.B, r.T_in, r.T_out = deg, b, T.data_ptr(), None
        r.in_pos, r.out_pos, r.lmbd_pos, r.node_ampls, r.edge_ampls = ip.data_ptr(), op.data_ptr(), None, None, ea.data_ptr()
    ws = torch.zeros(max(lib.workspace_bytes(_lib.C64, 3, D, D), lib.workspace_bytes(_lib.C64, 2, D, D)),
                     dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for mode in (1, 0):
        lib.set_kernel_mode(mode)
        try:
            m0, m1 = Cm.clone(), Cm.clone()
            resid = torch.zeros(2 * 4, dtype=torch.float32, device=dev)
            status = torch.zeros(4, dtype=torch.int32, device=dev)
            lib.bp_run_classes(_lib.C64, 2, C.byref(rows), D, m0.data_ptr(), m1.data_ptr(), 0, 0.25, 1e-30, 2,
                               resid.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), st)
            ext = torch.zeros(nslots * 4 * D * D, dtype=torch.complex64, device=dev)
            lib.ext_msgs_classes(_lib.C64, 2, C.byref(rows), D, Cm.data_ptr(), ext.data_ptr(), 0.7, ws.data_ptr(), ws.numel(), st)
            torch.cuda.synchronize()
            out[mode] = (m1.cpu().numpy().reshape(-1, D, D), m0.cpu().numpy().reshape(-1, D, D),
                         resid.cpu().numpy().astype(np.float64), ext.cpu().numpy().reshape(-1, 2 * D, 2 * D),
                         status.cpu().numpy())
        finally:
            lib.set_kernel_mode(0)
    gen, fast = out[1], out[0]
    assert fast[4][1] == 2 and gen[4][1] == 2 and fast[4][0] == 0                  # two sweeps, cap reached
    for k in (0, 1, 3):
        assert np.abs(fast[k] - gen[k]).max() < 3e-6
    assert np.allclose(fast[2][:4], gen[2][:4], rtol=1e-3)
    # oracle (complex128) on the same complex64 inputs: first sweep (damped, into buffer 1) and extended messages
    for t, ip, op, th in ((t3, in3, out3, th3), (t2, in2, out2, th2)):
        msgs = [cur[ip[j]].astype(np.complex128) for j in range(ip.shape[0])]
        want = O.pass_msgs(t.astype(np.complex128), msgs)
        want_ext = O.pass_msgs(t.astype(np.complex128), msgs,
                               [(0.7 * x.astype(np.float32).astype(np.float64)).astype(np.complex128) for x in th])
        for j in range(ip.shape[0]):
            assert np.abs(fast[0][op[j]] - (0.25 * cur[op[j]] + 0.75 * want[j])).max() < 3e-6
            assert np.abs(fast[3][op[j]] - want_ext[j]).max() < 3e-6


def _graded_ext_msgs(rng, L, n=8):
    """Hermitian PSD trace-1 matrices with the graded spectra extended messages have in the symmetric gauge.
    The spectrum stays clear of the pinv_eps = 1e-6 mask (kept values >= 3e-6, masked ones <= 1e-7): an
    eigenvalue within complex64 rounding of the cut would be kept by one precision and dropped by the other."""
    spec = np.array([0.82, 0.18, 1.4e-3, 3e-4, 4e-5, 9e-6, 2.5e-8, 3e-9])

    def psd():
        q = np.linalg.qr(rng.normal(size=(L, n, n)) + 1j * rng.normal(size=(L, n, n)))[0]
        q = np.linalg.qr(np.eye(n) + 0.05 * q)[0]
        s = spec * np.exp(rng.normal(scale=0.3, size=(L, n)))
        m = (q * (s / s.sum(1, keepdims=True))[:, None, :]) @ np.swapaxes(q.conj(), 1, 2)
        return 0.5 * (m + np.swapaxes(m.conj(), 1, 2))
    return np.concatenate([psd(), psd()], 0)


@pytest.mark.parametrize("L", [1, 5, 16, 17, 3001])
def test_fast_canon8_matches_oracle(lib, L):
    """Specialised n = 8 complex64 canonicalizer kernels (bqa_fast_canon8v2.cu, the current one: mode 0; bqa_fast_canon8.cu,
    the first design: mode 2) and the generic kernel (mode 1) against the oracle's
    _get_canonicalizers restatement (reference state.py:171-200) in complex128.  Gauge-invariant checks: the lambdas,
    their column maxima, and the defining property of the canonicalizers -- with the oracle's square-root factors
    lu_f, lu_b of the masked messages (backends.py:483-490) and ker = lu_f lu_b^T (state.py:186-187),
    (lu_f C_f)^H ker conj(lu_b C_b) = diag(S) on the kept columns, i.e. lu_f C_f and lu_b C_b are the singular
    vectors of ker (well conditioned: the 1/sqrt(eigenvalue) factors inside C cancel against lu)."""
    import torch
    from bqa_b200 import _lib
    from oracle import bqa_oracle as O
    rng = np.random.default_rng(100 + L)
    n, D = 8, 4
    ext = _graded_ext_msgs(rng, L).astype(np.complex64)
    ext64 = ext.astype(np.complex128)
    lm_ref, _ = O.canonicalizers(ext64, 1e-6, np.complex128)
    lm_ref = lm_ref.real
    u, lam, uh = O.masked_svd(ext64, 1e-6, np.complex128)
    lu = np.sqrt(lam)[..., :, None] * uh
    ker = lu[:L] @ np.swapaxes(lu[L:], 1, 2)
    s_ref = np.linalg.svd(ker, compute_uv=False)[:, :4]
    dev = torch.device("cuda:0")
    e = torch.from_numpy(ext.reshape(-1)).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    for mode in (1, 2, 0):
        lib.set_kernel_mode(mode)
        try:
            canon = torch.zeros_like(e)
            lm = torch.zeros(L * n, dtype=torch.float32, device=dev)
            colmax = torch.zeros(n, dtype=torch.float32, device=dev)
            lib.canonicalize(_lib.C64, D, L, e.data_ptr(), canon.data_ptr(), lm.data_ptr(), colmax.data_ptr(), 1e-6, 4, st)
        finally:
            lib.set_kernel_mode(0)
        lmh = lm.cpu().numpy().reshape(L, n).astype(np.float64)
        c = canon.cpu().numpy().reshape(2 * L, n, n).astype(np.complex128)
        af, ab = lu[:L] @ c[L:, :, :4], lu[L:] @ c[:L, :, :4]
        G = np.swapaxes(af.conj(), 1, 2) @ ker @ ab.conj()
        assert np.abs(colmax.cpu().numpy() - lmh.max(0)).max() == 0.0      # exactly the maxima of what was written
        assert np.abs(lmh[:, :4] - lm_ref[:, :4]).max() < 5e-6                # lambdas (L2-normalised, <= 1)
        # (mode 0 takes the eigenvectors of a message from the rotated columns of its Cholesky factor instead of
        # accumulating the rotations: on these synthetic spectra, six decades deep, that costs a factor 30 on this
        # particular measure; on extended messages of real anneals the end-to-end error is what
        # test_rr100k_window_vs_reference_golden and bench.py's parity block bound)
        assert np.abs(np.abs(G) - s_ref[:, :, None] * np.eye(4)).max() < (6e-5 if mode == 0 else 2e-5), mode


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("D,L", [(8, 1), (8, 37), (16, 9)])
def test_generic_canonicalizer_round_robin_matches_serial_and_oracle(lib, D, L, precision):
    """n = 16 / 32 canonicalizer: the round-robin Jacobi over the lanes of a warp (mode 0, jacobi_svd_round_robin) and the
    serial cyclic routine (mode 1) against the oracle's _get_canonicalizers restatement (reference state.py:171-200) in
    complex128 -- lambdas, column maxima and the defining property of the canonicalizers (see
    test_fast_canon8_matches_oracle) on the D kept columns."""
    import torch
    from bqa_b200 import _lib
    from oracle import bqa_oracle as O
    n = 2 * D
    rng = np.random.default_rng(1000 * D + L)
    kept = np.array([0.5, 0.25, 0.12, 0.06, 0.03, 0.015, 7e-3, 3e-3, 1e-3, 3e-4, 1.2e-4, 5e-5, 2e-5, 9e-6][: min(n - 2, 14)])
    spec = np.concatenate([kept, np.full(n - kept.size, 2e-9)])

    def psd():
        q = np.linalg.qr(rng.normal(size=(L, n, n)) + 1j * rng.normal(size=(L, n, n)))[0]
        q = np.linalg.qr(np.eye(n) + 0.05 * q)[0]
        sp = spec * np.exp(rng.normal(scale=0.2, size=(L, n)))
        m = (q * (sp / sp.sum(1, keepdims=True))[:, None, :]) @ np.swapaxes(q.conj(), 1, 2)
        return 0.5 * (m + np.swapaxes(m.conj(), 1, 2))
    cdt, rdt, prec = (np.complex64, np.float32, _lib.C64) if precision == "single" else (np.complex128, np.float64, _lib.C128)
    ext = np.concatenate([psd(), psd()], 0).astype(cdt)
    ext128 = ext.astype(np.complex128)
    lm_ref, _ = O.canonicalizers(ext128, 1e-6, np.complex128)
    lm_ref = lm_ref.real
    u, lam, uh = O.masked_svd(ext128, 1e-6, np.complex128)
    lu = np.sqrt(lam)[..., :, None] * uh
    ker = lu[:L] @ np.swapaxes(lu[L:], 1, 2)
    s_ref = np.linalg.svd(ker, compute_uv=False)[:, :D]
    dev = torch.device("cuda:0")
    e = torch.from_numpy(ext.reshape(-1)).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    # complex64 at n = 32 with a spectrum five decades deep: the columns of the smallest kept singular values carry 1e-4
    tol_l, tol_g = (5e-6, 6e-5 if D == 8 else 3e-4) if precision == "single" else (1e-12, 1e-10)
    for mode in (1, 0):
        lib.set_kernel_mode(mode)
        try:
            canon = torch.zeros_like(e)
            lm = torch.zeros(L * n, dtype=torch.from_numpy(np.zeros(1, rdt)).dtype, device=dev)
            colmax = torch.zeros(n, dtype=lm.dtype, device=dev)
            lib.canonicalize(prec, D, L, e.data_ptr(), canon.data_ptr(), lm.data_ptr(), colmax.data_ptr(), 1e-6, D, st)
        finally:
            lib.set_kernel_mode(0)
        lmh = lm.cpu().numpy().reshape(L, n).astype(np.float64)
        c = canon.cpu().numpy().reshape(2 * L, n, n).astype(np.complex128)
        af, ab = lu[:L] @ c[L:, :, :D], lu[L:] @ c[:L, :, :D]
        G = np.swapaxes(af.conj(), 1, 2) @ ker @ ab.conj()
        assert np.abs(colmax.cpu().numpy() - lmh.max(0)).max() == 0.0
        assert np.abs(lmh[:, :D] - lm_ref[:, :D]).max() < tol_l, mode
        assert np.abs(np.abs(G) - s_ref[:, :, None] * np.eye(D)).max() < tol_g, mode


def test_partitioned_nccl(lib):
    """Node-partitioned engine over NCCL on 2 GPUs == single-GPU engine (skipped on a 1-GPU box; the gloo
    world_size-2/3 tests in tests/test_partitioned.py cover the same logic on CPU)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(here, "multigpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "multigpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("B", [1, 3, 5, 6, 2003])
def test_fast_apply_update_matches_generic(lib, B):
    """Specialised degree-3 / D = 4 -> 4 / complex64 simple-update application (bqa_fast_apply.cu) against the
    generic kernel (itself checked against the oracle end to end) on random canonicalizers, lambdas and scattered
    slots, incl. a ragged batch; the re-initialised messages must be identical."""
    import torch
    from bqa_b200 import _lib
    d, D = 3, 4
    rng = np.random.default_rng(50 + B)
    L = (3 * B + 1) // 2 + 3
    t, _, thetas = instances.random_node_batch(B, d, D, seed=B)
    canon = (rng.normal(size=(2 * L, 8, 8)) + 1j * rng.normal(size=(2 * L, 8, 8))).astype(np.complex64)
    lm = np.sort(rng.uniform(0.01, 1.0, size=(L, 8)), axis=1)[:, ::-1].astype(np.float32).copy()
    in_pos = rng.permutation(2 * L)[: d * B].reshape(d, B).astype(np.int32)
    out_pos = ((in_pos + L) % (2 * L)).astype(np.int32)
    lpos = (in_pos % L).astype(np.int32)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    T, Cn, Lm = up(t.astype(np.complex64).reshape(-1)), up(canon.reshape(-1)), up(lm.reshape(-1))
    ip, op, lp = up(in_pos), up(out_pos), up(lpos)
    na = up(rng.uniform(-1, 1, size=B).astype(np.float32))
    ea = up(np.stack(thetas).astype(np.float32))
    ws = torch.zeros(lib.workspace_bytes(_lib.C64, d, D, D), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for mode in (1, 0):
        lib.set_kernel_mode(mode)
        try:
            Tout = torch.zeros_like(T)
            msgs = torch.zeros(2 * L * D * D, dtype=torch.complex64, device=dev)
            lib.apply_update(_lib.C64, d, D, D, B, T.data_ptr(), Tout.data_ptr(), Cn.data_ptr(), Lm.data_ptr(),
                             msgs.data_ptr(), ip.data_ptr(), op.data_ptr(), lp.data_ptr(), na.data_ptr(), ea.data_ptr(),
                             0.13, 0.07, ws.data_ptr(), ws.numel(), st)
            out[mode] = (Tout.cpu().numpy().reshape(B, -1), msgs.cpu().numpy())
        finally:
            lib.set_kernel_mode(0)
    assert np.abs(np.linalg.norm(out[0][0], axis=1) - 1).max() < 1e-5
    assert np.abs(out[0][0] - out[1][0]).max() < 2e-6
    assert np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("B", [1, 3, 5, 6, 2003])
def test_apply_update_kernels_match_oracle(lib, B):
    """Kernel-level oracle vector for the simple-update application: both the specialised
    degree-3 / D = 4 kernel (bqa_fast_apply.cu) and the generic one against the oracle's restatement of
    apply_canonicalizers_with_extensions (reference backends.py:416-432) -> Rz layer (state.py:142-150) -> Rx layer
    (:153-156) -> symmetric gauge sqrt(lambda) per leg + L2 normalisation (:219-227, backends.py:450-462) in
    complex128 on the same complex64 inputs; the re-initialised messages are diag(lambda) / trace (state.py:56-57)."""
    import torch
    from bqa_b200 import _lib
    from oracle import bqa_oracle as O
    d, D = 3, 4
    rng = np.random.default_rng(150 + B)
    L = (3 * B + 1) // 2 + 3
    t, _, thetas = instances.random_node_batch(B, d, D, seed=11 + B)
    t32 = t.astype(np.complex64)
    canon = (rng.normal(size=(2 * L, 8, 8)) + 1j * rng.normal(size=(2 * L, 8, 8))).astype(np.complex64)
    lm = np.sort(rng.uniform(0.01, 1.0, size=(L, 8)), axis=1)[:, ::-1].astype(np.float32).copy()
    in_pos = rng.permutation(2 * L)[: d * B].reshape(d, B).astype(np.int32)
    out_pos = ((in_pos + L) % (2 * L)).astype(np.int32)
    lpos = (in_pos % L).astype(np.int32)
    h = rng.uniform(-1, 1, size=B).astype(np.float32)
    J = np.stack(thetas).astype(np.float32)
    zt, xt = 0.13, 0.07
    # oracle: complex couplings like the reference's edge_ampls (principal roots of negative sines)
    c128 = np.complex128
    want = O.apply_canonicalizers_ext(t32.astype(c128), [canon[in_pos[j]][:, :, :D].astype(c128) for j in range(d)],
                                      [(np.float64(np.float32(zt)) * J[j].astype(np.float64)).astype(c128) for j in range(d)])
    phi = (zt * h.astype(np.float64)).reshape(-1, 1, 1, 1, 1)
    zz = want.copy()
    zz[:, 1] *= -1.0
    want = want * np.cos(phi) - 1j * zz * np.sin(phi)
    want = np.cos(xt) * want - 1j * np.sin(xt) * want[:, ::-1]
    for j in range(d):
        shp = [B, 1, 1, 1, 1]
        shp[2 + j] = D
        want = want * np.sqrt(lm[lpos[j]][:, :D].astype(np.float64)).reshape(shp)
    want = want / np.linalg.norm(want.reshape(B, -1), axis=1).reshape(-1, 1, 1, 1, 1)
    lam = lm[lpos.reshape(-1)][:, :D].astype(np.float64)
    want_msgs = (lam / lam.sum(1, keepdims=True))[:, :, None] * np.eye(D)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    T, Cn, Lm = up(t32.reshape(-1)), up(canon.reshape(-1)), up(lm.reshape(-1))
    ip, op, lp, na, ea = up(in_pos), up(out_pos), up(lpos), up(h), up(J)
    ws = torch.zeros(lib.workspace_bytes(_lib.C64, d, D, D), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for mode in (1, 0):
        lib.set_kernel_mode(mode)
        try:
            Tout = torch.zeros_like(T)
            msgs = torch.zeros(2 * L * D * D, dtype=torch.complex64, device=dev)
            lib.apply_update(_lib.C64, d, D, D, B, T.data_ptr(), Tout.data_ptr(), Cn.data_ptr(), Lm.data_ptr(),
                             msgs.data_ptr(), ip.data_ptr(), op.data_ptr(), lp.data_ptr(), na.data_ptr(), ea.data_ptr(),
                             zt, xt, ws.data_ptr(), ws.numel(), st)
        finally:
            lib.set_kernel_mode(0)
        got = Tout.cpu().numpy().reshape(want.shape)
        assert np.abs(got - want).max() < 2e-6, mode
        got_msgs = msgs.cpu().numpy().reshape(2 * L, D, D)[out_pos.reshape(-1)]
        assert np.abs(got_msgs - want_msgs).max() < 1e-6, mode


def test_rr100k_window_vs_reference_golden(golden_dir, lib):
    """The benchmarked instance at full size against the UNMODIFIED reference (tests/golden/make_golden_100k.py: numpy
    backend, complex128): the 30 ramp steps (bond dimension 1 -> 4) plus 3 steady-state steps of bench.py's schedule.
    complex64 (the benchmarked precision) within the stated fp32 tolerance; complex128 to rounding."""
    from bqa_b200.benchmarking import ising_energy
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, "rr100k_window.npz"))
    cfg = instances.bench_config(100_000)
    assert instances.instance_fingerprint(cfg["nodes"], cfg["edges"]) == str(g["fingerprint"]), \
        "this box rebuilt a different random graph than the one the golden was computed on"
    ctx = config_to_context(cfg)
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    n_ramp, n_all = 30, len(g["bp_sweeps"])
    stride = 8
    for precision in ("single", "double"):
        eng = Engine(ctx, precision=precision)
        for ins in layers[:n_ramp]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        b_ramp = eng.bloch_vectors()
        for ins in layers[n_ramp:n_all]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        b = eng.bloch_vectors()
        lm = np.sort(eng.lmbds_numpy(), axis=1)[:, ::-1]
        assert eng.stats["bond_dims"] == g["bond_dims"].tolist()
        sw = np.abs(np.array(eng.stats["bp_sweeps"]) - g["bp_sweeps"])
        e_ref = float(g["energy"])
        e = ising_energy(cfg["edges"], cfg["nodes"], np.where(b[:, 2] > 0, 1.0, -1.0))
        if precision == "single":
            for got, want in ((b_ramp, g["bloch_ramp"]), (b, g["bloch"])):
                diff = np.abs(got - want)
                assert diff.max() < 5e-3 and diff.mean() < 1e-4                 # stated fp32 tolerance
            assert sw.max() <= 1
            dl = np.abs(lm[::stride] - g["lmbds_sorted_strided"])
            assert dl.max() < 2e-3 and dl.mean() < 5e-5                         # lambdas are L2-normalised (<= 1)
            assert abs(e - e_ref) <= 1e-4 * abs(e_ref)
        else:
            assert np.abs(b_ramp - g["bloch_ramp"]).max() < 1e-7 and np.abs(b - g["bloch"]).max() < 1e-7
            assert sw.max() == 0
            assert np.abs(lm[::stride] - g["lmbds_sorted_strided"]).max() < 1e-9
            assert np.abs(lm.max(0) - g["lmbds_colmax"]).max() < 1e-9
            assert abs(e - e_ref) <= 1e-12 * abs(e_ref)                         # same spins, another summation order
        del eng


# ---- BASELINE.json configs[0..2] as parity cases ------------------------------------
def _oracle_vs_gpu(cfg, lib, bloch_tol, check_outcomes=True):
    from oracle import bqa_oracle as O
    want, octx, ost = O.run_qa(cfg, return_state=True)
    want = dict(want)
    res, eng = _run(cfg, "double")
    assert eng.stats["bond_dims"] == ost.stats["bond_dims"]
    assert np.abs(np.array(res["bloch_vectors"]) - np.array(want["bloch_vectors"])).max() < bloch_tol
    if check_outcomes and "measurement_outcomes" in want:
        assert res["measurement_outcomes"] == want["measurement_outcomes"]
    return res, eng, ost


def test_config2_grid_20x20_double_vs_oracle(lib):
    """BASELINE configs[1]: 2D square-grid Ising anneal in the shape of reference examples/2d_greed.py:11-30
    (400 qubits, degrees 2/3/4, max_bond_dim 4, total_time 10; 40 of its 100 steps to bound the oracle's run time),
    validated against the numpy-backend restatement in double precision."""
    from bqa_b200.benchmarking import generate_qubo_on_2d_grid
    nodes, edges = generate_qubo_on_2d_grid(20, 20, seed=42)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 4,
           "schedule": {"total_time": 4.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 40, "final_mixing": 0.6}, "get_bloch_vectors"]}}
    _oracle_vs_gpu(cfg, lib, 1e-7)     # (the O(N^2) python sampler of the oracle is exercised on the smaller configs)


def test_config3_heavy_hex_127_double_vs_oracle(lib):
    """BASELINE configs[2]: the full-size IBM heavy-hex anneal of reference examples/full_size_ibm_heavy_hex.py:12-100
    (127 qubits, +-1 amplitudes, max_bond_dim 8, total_time 10, 10 steps, then measure), double precision.
    +-1 amplitudes give exactly degenerate spectra: this is the case the no-FMA complex128 build exists for."""
    from bqa_b200.benchmarking import heavy_hex_127
    nodes, edges = heavy_hex_127(seed=42)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 8,
           "schedule": {"total_time": 10.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 10, "final_mixing": 0.0}, "get_bloch_vectors", "measure"]}}
    _oracle_vs_gpu(cfg, lib, 1e-6)


def test_config1_maxcut_shape_double_vs_oracle(lib):
    """BASELINE configs[0] shape (benchmarks_against_mqlib/random_3_regular_maxcut_1000.py:10-44: zero fields, unit
    couplings, max_bond_dim 16, damping 0.5, eps 1e-5, 250 BP iterations, dt 0.2) on 200 qubits for the first 120
    steps -- the bond dimension passes through 1, 2, 3 on the way; degenerate spectra again."""
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    nodes, edges = generate_qubo_on_random_regular_graph(200, 3, seed=42, node_ampl_func=lambda *_: 0.0,
                                                         edge_ampl_func=lambda *_: 1.0)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 16, "measurement_threshold": 0.99, "damping": 0.5,
           "bp_eps": 1e-5, "pinv_eps": 1e-5, "max_bp_iter_number": 250,
           "schedule": {"total_time": 24.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 120, "final_mixing": 0.88}, "get_bloch_vectors"]}}
    _oracle_vs_gpu(cfg, lib, 1e-6)


def test_config1_maxcut_shape_single_bond_dimension_growth(lib):
    """The same shape in complex64 through the growth of the bond dimension past 4 (130 steps: 1, 2, 3, 4, 5, 6).  How far
    D may grow is decided by singular values around pinv_eps = 1e-5 of degenerate spectra: a canonicalizer that resolves the
    small end of the spectrum only to 3e-5 (the Cholesky-factor n = 8 kernel, meant for the truncation at D = max_bond_dim)
    lets D leave 4 six steps early and sends the anneal onto a trajectory with four times the BP sweeps.  The bond
    dimensions must appear within two steps of the complex128 oracle's and the sweep totals must agree."""
    from bqa_b200.benchmarking import generate_qubo_on_random_regular_graph
    from oracle import bqa_oracle as O
    nodes, edges = generate_qubo_on_random_regular_graph(200, 3, seed=42, node_ampl_func=lambda *_: 0.0,
                                                         edge_ampl_func=lambda *_: 1.0)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 16, "measurement_threshold": 0.99, "damping": 0.5,
           "bp_eps": 1e-5, "pinv_eps": 1e-5, "max_bp_iter_number": 250,
           "schedule": {"total_time": 26.0, "starting_mixing": 1.0,
                        "actions": [{"weight": 1.0, "steps_number": 130, "final_mixing": 0.87}, "get_bloch_vectors"]}}
    _, octx, ost = O.run_qa(cfg, return_state=True)
    _, eng = _run(cfg, "single")
    first = lambda dims: {d: dims.index(d) for d in sorted(set(dims))}
    want, got = first(ost.stats["bond_dims"]), first(eng.stats["bond_dims"])
    assert max(want) >= 5, want                                   # the window does reach the growth past 4
    assert set(got) == set(want), (got, want)
    assert max(abs(got[d] - want[d]) for d in want) <= 2, (got, want)
    assert abs(sum(eng.stats["bp_sweeps"]) - sum(ost.stats["bp_sweeps"])) <= 0.1 * sum(ost.stats["bp_sweeps"])


@pytest.mark.parametrize("precision,tol", [("double", 1e-12), ("single", 1e-5)])
def test_isolated_qubit_gpu(lib, precision, tol):
    """Degree-0 class on the GPU: the isolated qubit follows its exact single-qubit evolution; the other qubits
    equal the oracle run on the connected part."""
    from oracle import bqa_oracle as O
    cfg = instances.cfg_isolated_qubit()
    res, _ = _run(cfg, precision)
    b = np.array(res["bloch_vectors"])
    assert np.abs(b[3] - instances.isolated_qubit_bloch(cfg)).max() < tol
    connected = {**cfg, "nodes": {k: v for k, v in cfg["nodes"].items() if k != 3},
                 "schedule": {**cfg["schedule"], "actions": cfg["schedule"]["actions"][:2]}}
    want = np.array(dict(O.run_qa(connected))["bloch_vectors"])
    assert np.abs(b[:3] - want).max() < (1e-9 if precision == "double" else 5e-4)
    assert len(res["measurement_outcomes"]) == 4


def test_ext_msgs_enqueued_behind_the_bp_run(lib):
    """run_layers / run_layer(next_ztime=...): the next step's extended messages are enqueued behind the BP run and pick
    the message buffer on the device (bqa_b200_ext_msgs_after_run).  Same kernels on the same inputs: results identical
    to the step-by-step path, also when runs hit the iteration cap (the other buffer rule) and after a state upload."""
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    for cfg in (_rr_config(3000, 24, 4.8, []), _rr_config(1000, 24, 4.8, [], max_bp_iter_number=3, damping=0.2),
                _rr_config(1000, 24, 4.8, [], max_bp_iter_number=4)):
        ctx = config_to_context(cfg)
        layers = [i for i in ctx.instructions if isinstance(i, dict)]
        out, used = {}, 0
        for ahead in (False, True):
            eng = Engine(ctx, precision="single")
            eng.ext_ahead = ahead
            orig = eng.lib.ext_msgs_after_run
            calls = []
            eng.lib.ext_msgs_after_run = lambda *a: (calls.append(1), orig(*a))[1]
            try:
                eng.run_layers(layers[:-4])
                snap = eng.state_to_host()
                eng.run_layer(layers[-4]["xtime"], layers[-4]["ztime"], next_ztime=layers[-3]["ztime"])
                eng.load_state(snap)                      # drops the extended messages computed ahead
                eng.run_layers(layers[-4:])
            finally:
                eng.lib.ext_msgs_after_run = orig
            used += len(calls) if ahead else 0
            assert ahead or not calls
            out[ahead] = (eng.bloch_vectors(), eng.stats["bp_sweeps"], eng.stats["bond_dims"])
        assert used > 0 and out[True][2][-1] == 4
        assert out[True][1] == out[False][1] and out[True][2] == out[False][2]
        assert np.array_equal(out[True][0], out[False][0])


def test_single_launch_bp_run_equals_per_sweep_launches(lib):
    """bqa_b200_bp_run (the whole BP loop in one cooperative launch with in-kernel grid barriers) against the
    per-sweep launches of bqa_b200_bp_sweep: same arithmetic, so identical sweep counts, residuals and marginals."""
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    cfg = _rr_config(3000, 24, 4.8, [])
    ctx = config_to_context(cfg)
    out = {}
    for single in (False, True):
        eng = Engine(ctx, precision="single")
        eng._single_launch_ok = single
        for ins in [i for i in ctx.instructions if isinstance(i, dict)]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        out[single] = (eng.bloch_vectors(), eng.stats["bp_sweeps"], eng.stats["bp_dist"], eng.stats["bond_dims"])
    assert out[True][3] == out[False][3] and out[True][3][-1] == 4
    assert out[True][1] == out[False][1]
    assert out[True][2] == out[False][2]
    assert np.array_equal(out[True][0], out[False][0])
    # a capped run (3 iterations, damping): the undamped last sweep is kept (state.py:122-123) on both paths
    cfg = _rr_config(1000, 12, 2.4, [], max_bp_iter_number=3, damping=0.2)
    ctx = config_to_context(cfg)
    res = {}
    for single in (False, True):
        eng = Engine(ctx, precision="single")
        eng._single_launch_ok = single
        for ins in [i for i in ctx.instructions if isinstance(i, dict)]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        res[single] = (eng.bloch_vectors(), eng.stats["bp_sweeps"])
    assert res[True][1] == res[False][1] and np.array_equal(res[True][0], res[False][0])


@pytest.mark.parametrize("precision", ["double", "single"])
@pytest.mark.parametrize("name", ["grid4", "comb", "isolated"])
def test_multiclass_launches_equal_per_class_launches(lib, name, precision):
    """All degree classes in one launch (bqa_multiclass.cuh: ext_msgs_classes / apply_update_classes / bp_run_classes, the
    BP loop of state.py:97-124 on the device) against one launch per class and sweep: same per-node code, so identical
    sweep counts, residuals, bond dimensions, marginals and bitstrings -- with at most 4 launches per annealing step."""
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    cfg = instances.cfg_isolated_qubit() if name == "isolated" else instances.GOLDEN_CONFIGS[name]()
    ctx = config_to_context(cfg)
    out = {}
    for multi in (False, True):
        eng = Engine(ctx, precision=precision)
        assert eng._multiclass
        eng._multiclass = multi
        before = lib.launch_count()
        layers = [i for i in ctx.instructions if isinstance(i, dict)]
        # (generic kernels on both sides: per-class launches would give a degree-3 class at D = 4 in complex64 to the
        # specialised kernels, whose arithmetic is ordered differently)
        lib.set_kernel_mode(1)
        try:
            for ins in layers:
                eng.run_layer(ins["xtime"], ins["ztime"])
            launches = lib.launch_count() - before
            bloch = eng.bloch_vectors()
            outcomes = eng.measure()
        finally:
            lib.set_kernel_mode(0)
        out[multi] = (bloch, outcomes, eng.stats["bp_sweeps"], eng.stats["bp_dist"], eng.stats["bond_dims"], launches / len(layers))
    assert np.array_equal(out[True][0], out[False][0])
    assert out[True][1:5] == out[False][1:5]
    assert out[True][5] <= 4.0 + 1e-9 and out[False][5] > out[True][5]


def test_engine_on_a_device_that_is_not_current(lib):
    """Engine(device="cuda:1") while cuda:0 is the current device: the raw launches behind the C ABI (cooperative launch,
    function attributes, SM count) act on cudaGetDevice(), so the engine makes its device current around every call
    (skipped on a 1-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    torch.cuda.set_device(0)
    cfg = _rr_config(3000, 12, 2.4, [])
    ctx = config_to_context(cfg)
    out = []
    for dev in ("cuda:0", "cuda:1"):
        eng = Engine(ctx, precision="single", device=dev)
        for ins in [i for i in ctx.instructions if isinstance(i, dict)]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        assert torch.cuda.current_device() == 0
        out.append(eng.bloch_vectors())
    assert np.array_equal(out[0], out[1])


def test_nan_state_fails_loudly(lib):
    """A non-finite node tensor must not let BP "converge" on the finite entries: the residual becomes non-finite and the
    engine raises (the reference fails its `assert best_msgs is not None`, state.py:113-123)."""
    import torch
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    for precision, cfg in (("single", _rr_config(400, 12, 2.4, [])), ("double", instances.cfg_grid4())):
        ctx = config_to_context(cfg)
        eng = Engine(ctx, precision=precision)
        layers = [i for i in ctx.instructions if isinstance(i, dict)]
        for ins in layers[:10]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        c = eng.classes[-1]
        c.T[c.cur][3] = float("nan")
        with pytest.raises(FloatingPointError):
            eng.run_bp()
