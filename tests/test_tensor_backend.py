"""The backend interface itself: ``B200Backend`` (bqa_b200/tensor_backend.py) implements bqa's ``Tensor`` ABC (reference
src/bqa/backends.py:28-252) on device arrays.  The cases mirror the reference's own backend tests --
tests/test_gpt_generated_npbackend.py (raw ops and composites against numpy), tests/test_gpt_generated_gates_application.py
(Rx / Rz / ZZ half gate identities) and tests/test_small_circuit_final_density.py / test_core_subroutines.py (the
unmodified engine src/bqa/state.py driven through the backend) -- parametrised over the reference's numpy backend (CPU:
checks the cases themselves) and "b200" (GPU).  They need the reference package, installed unmodified into baseline/_ref
by baseline/install_ref.py (build container) and shipped to the GPU box with the snapshot."""
import os
import sys
from math import pi

import numpy as np
import pytest

import instances

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import install_ref  # noqa: E402

pytestmark = pytest.mark.skipif(not install_ref.installed(), reason="baseline/_ref (the reference install) is absent")

BACKENDS = ["numpy", pytest.param("b200", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", params=BACKENDS)
def B(request):
    install_ref.add_to_path()
    import bqa.backends as backends
    if request.param == "numpy":
        return backends.NumPyBackend
    from bqa_b200 import register_with_bqa
    from bqa_b200.build import build
    build()
    cls = register_with_bqa()
    assert backends.BACKEND_STR_TO_BACKEND["b200"] is cls and issubclass(cls, backends.Tensor)
    return cls


@pytest.fixture(scope="module")
def dtype():
    install_ref.add_to_path()
    from bqa.utils import NP_DTYPE
    return NP_DTYPE


def rand(rng, *shape, cplx=True):
    a = rng.normal(size=shape)
    return a + 1j * rng.normal(size=shape) if cplx else a


def close(t, want, tol=1e-10):
    got = t.numpy if hasattr(t, "numpy") and not isinstance(t, np.ndarray) else t
    assert got.shape == np.asarray(want).shape, (got.shape, np.asarray(want).shape)
    assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max())


# ---- constructors, shapes (reference test_gpt_generated_npbackend.py:25-89) -------------------------------------------
def test_constructors_and_dtypes(B, dtype):
    t = B.make_from_list([3, 1, 2])
    assert t.numpy.dtype == np.intp and t.numpy.tolist() == [3, 1, 2] and t.batch_size == 3 and t.batch_shape == ()
    t = B.make_from_list([1.0, -2.0])
    assert t.numpy.dtype == dtype and np.allclose(t.numpy, [1.0, -2.0])
    arr = np.arange(6).reshape(2, 3)
    assert np.array_equal(B.make_from_numpy(arr).numpy, arr)               # round trip (numpy: shared memory; b200: H2D, D2H)
    assert B.make_from_iter(iter([4, 5])).numpy.tolist() == [4, 5]
    e = B.make_empty(5, (2, 3))
    assert e.batch_size == 5 and e.batch_shape == (2, 3) and e.batch_rank == 2 and e.numpy.dtype == dtype


def test_sqrt_inv_pinv_sin_cos_conj(B):
    rng = np.random.default_rng(1)
    x = rand(rng, 3, 4, cplx=False) ** 2
    close(B.make_from_numpy(x).sqrt(), np.sqrt(x))
    z = rand(rng, 2, 3, 4)
    close(B.make_from_numpy(z).sqrt(), np.sqrt(z))                         # principal branch incl. negative real parts
    close(B.make_from_numpy(-x.astype(complex)).sqrt(), np.sqrt(-x.astype(complex)))
    close(B.make_from_numpy(z + 3.0).inv(), 1 / (z + 3.0))
    close(B.make_from_numpy(z).sin(), np.sin(z))
    close(B.make_from_numpy(z).cos(), np.cos(z))
    close(B.make_from_numpy(z).conj(), z.conj())
    lam = np.array([[0.5, 1e-3, 1e-20, 0.0]]).astype(complex)
    close(B.make_from_numpy(lam).pinv(), np.array([[2.0, 1e3, 0.0, 0.0]]))   # cut at machine epsilon (backends.py:719-727)


def test_reshape_transpose_slices(B):
    rng = np.random.default_rng(2)
    x = rand(rng, 4, 3, 2)
    t = B.make_from_numpy(x)
    assert (t.batch_size, t.batch_shape, t.batch_rank) == (4, (3, 2), 2)
    close(t.batch_reshape((6,)), x.reshape(4, 6))
    y = rand(rng, 3, 2, 4, 5)
    close(B.make_from_numpy(y).batch_transpose((2, 0, 1)), np.transpose(y, (0, 3, 1, 2)))
    w = rand(rng, 10, 3, 3)
    close(B.make_from_numpy(w).get_batch_slice(range(2, 7)), w[2:7])
    close(B.make_from_numpy(w).batch_slice(B.make_from_numpy(np.array([0, 2, 9, 2]))), w[[0, 2, 9, 2]])
    close(B.make_from_numpy(w).batch_truncate_all_but(2, [0]), w[:, :, :2])
    close(B.make_from_numpy(w).batch_concat(B.make_from_numpy(w[:3])), np.concatenate([w, w[:3]], 0))


def test_assign_at_batch_indices_is_in_place(B):
    rng = np.random.default_rng(3)
    dst, src = rand(rng, 6, 2, 2), rand(rng, 3, 2, 2)
    want = dst.copy()
    want[[5, 0, 2]] = src
    t = B.make_from_numpy(dst.copy())
    r = t.assign_at_batch_indices(B.make_from_numpy(src), B.make_from_numpy(np.array([5, 0, 2])))
    close(r, want)
    close(t, want)                                                        # the destination itself was written (state.py:111-112)


# ---- algebra (reference test_gpt_generated_npbackend.py:97-233) ---------------------------------------------------------
def test_matmul_tensordot_diag(B):
    rng = np.random.default_rng(4)
    a, b = rand(rng, 3, 4, 2), rand(rng, 3, 2, 5)
    close(B.make_from_numpy(a).batch_matmul(B.make_from_numpy(b)), a @ b)
    close(B.make_from_numpy(a).batch_tensordot(B.make_from_numpy(b), axes=1), a @ b)
    p, q = rand(rng, 2, 3, 4, 5), rand(rng, 2, 5, 7, 3)
    want = np.stack([np.tensordot(l, r, axes=((2, 0), (0, 2))) for l, r in zip(p, q)])
    close(B.make_from_numpy(p).batch_tensordot(B.make_from_numpy(q), axes=[[2, 0], [0, 2]]), want)
    x = rand(rng, 4, 6)
    close(B.make_from_numpy(x).batched_diag(), np.stack([np.diag(r) for r in x]))


def test_broadcast_arithmetic_and_constants(B):
    rng = np.random.default_rng(5)
    a, b = rand(rng, 5, 3, 2), rand(rng, 5, 1, 2)
    ta, tb = B.make_from_numpy(a), B.make_from_numpy(b)
    close(ta * tb, a * b)
    close(ta + tb, a + b)
    close(ta - tb, a - b)
    close(ta / tb, a / b)
    close(0.25 * ta, 0.25 * a)
    close(ta * (1 - 2j), a * (1 - 2j))
    c = np.array([2.0, 0.5, -1.0, 1.0, 3.0])
    close(ta._mul_by_constants(B.make_from_numpy(c)), c[:, None, None] * a)


def test_norms_traces_distances(B):
    rng = np.random.default_rng(6)
    m = rand(rng, 7, 3, 3)
    t = B.make_from_numpy(m)
    close(t.batch_trace_normalize(), m / np.trace(m, axis1=1, axis2=2)[:, None, None])
    close(t.batch_normalize(), m / np.linalg.norm(m.reshape(7, -1), axis=1)[:, None, None])
    other = rand(rng, 7, 3, 3)
    want = np.abs(m - other).max() / np.abs(m + other).max()
    assert abs(float(t.get_dist(B.make_from_numpy(other)).numpy) - want) < 1e-12          # get_dist (backends.py:492-495)
    lam = np.sort(rng.uniform(0, 1, size=(9, 4)), axis=1)[:, ::-1].astype(complex)
    lam[:, 3] = 1e-9
    new, dim, err = B.make_from_numpy(lam).truncate_lmbds(3, 1e-6)                          # backends.py:297-303
    assert dim == 3 and abs(err - 1e-9) < 1e-15
    close(new, lam[:, :3])
    d = B.make_from_numpy(m.copy())
    d.make_inplace_damping_update(B.make_from_numpy(other), 0.3)                            # backends.py:761-764
    close(d, 0.3 * m + 0.7 * other)


def test_masked_svd(B):
    rng = np.random.default_rng(7)
    x = rand(rng, 4, 5, 5)
    x[3] = np.outer(x[3, :, 0], x[3, 0])                                    # rank one: four masked singular values
    u, s, vh = B.make_from_numpy(x).get_batch_svd(1e-6)
    U, S, VH = u.numpy, s.numpy, vh.numpy
    assert np.abs(U @ (S[:, :, None] * VH) - x).max() < 1e-9
    assert np.all(np.diff(S.real, axis=1) <= 1e-12) and np.abs(S.imag).max() == 0
    assert np.abs(S[3, 1:]).max() == 0 and np.abs(U[3, :, 1:]).max() == 0 and np.abs(VH[3, 1:]).max() == 0
    assert np.abs(np.swapaxes(U[0].conj(), 0, 1) @ U[0] - np.eye(5)).max() < 1e-9


# ---- gates (reference test_gpt_generated_gates_application.py) --------------------------------------------------------
def qubit(B, dtype, state):
    return B.make_from_numpy(np.asarray(state, dtype=dtype)[None])


def test_x_and_z_gates(B, dtype):
    psi = qubit(B, dtype, [1.0, 0.0])
    close(psi._apply_x_to_phys_dim()._apply_x_to_phys_dim(), psi.numpy)
    close(psi.apply_x_gates(pi), -psi.numpy, 1e-9)
    minus = qubit(B, dtype, [np.sqrt(0.5), -np.sqrt(0.5)])
    r = minus.apply_x_gates(0.37).numpy[0]
    assert np.allclose(np.outer(r.conj(), r), np.outer(minus.numpy[0].conj(), minus.numpy[0]), atol=1e-12)
    phi = qubit(B, dtype, [0.4, 0.7])
    close(phi._apply_z_to_phys_dim()._apply_z_to_phys_dim(), phi.numpy)
    sv = qubit(B, dtype, [0.6, 0.8])
    close(sv.apply_z_gates(B.make_from_numpy(np.array([pi], dtype))), -sv.numpy, 1e-9)
    for angle in (0.1, 0.7, 1.3):
        out = sv.apply_x_gates(angle).apply_z_gates(B.make_from_numpy(np.array([angle], dtype))).numpy[0]
        rho = np.outer(out.conj(), out)
        assert abs(np.trace(rho) - 1) < 1e-12 and np.allclose(rho, rho.conj().T)


def test_zz_half_gate(B, dtype):
    psi = qubit(B, dtype, [[[1.0]], [[0.0]]])
    c = B.make_from_numpy(np.array([0.5], dtype))
    assert psi.apply_conditional_z_gates([c, c]).batch_shape == (2, 2, 2)     # every bond doubles (backends.py:519-536)
    phi = qubit(B, dtype, [[0.5], [0.8]]).batch_normalize()
    same = phi.apply_conditional_z_gates([B.make_from_numpy(np.array([0.0], dtype))])
    close(same.numpy[:, :, 0], phi.numpy[:, :, 0])
    neg = phi.apply_conditional_z_gates([B.make_from_numpy(np.array([-0.7], dtype))])     # negative coupling: imaginary root
    up, down = np.sqrt(np.cos(-0.7 + 0j)), (np.sqrt(0.5) - 1j * np.sqrt(0.5)) * np.sqrt(np.sin(-0.7 + 0j))
    want = np.stack([phi.numpy[:, :, 0] * up, phi.numpy[:, :, 0] * np.array([1, -1]) * down], -1)
    close(neg, want / np.linalg.norm(want), 1e-9)


def test_measure_projects_and_renormalises_the_batch(B, dtype):
    rng = np.random.default_rng(8)
    t = rand(rng, 3, 2, 2, 2).astype(dtype)
    want = t.copy()
    want[1, 0] = 0
    want /= np.linalg.norm(want)
    x = B.make_from_numpy(t.copy())
    x.measure(1, 1)                                                         # keep outcome 1 of node 1 (backends.py:729-734)
    close(x, want)


# ---- fused composites against the reference's own definitions -----------------------------------------------------------
@pytest.mark.parametrize("d,D", [(1, 3), (2, 4), (3, 4), (3, 2), (4, 3)])
def test_pass_msgs_and_densities_equal_reference_goldens(B, golden_dir, d, D):
    """Same inputs as tests/golden/make_golden.py::kernel_level_vectors (outputs of the unmodified numpy backend)."""
    g = np.load(os.path.join(golden_dir, "kernel_level.npz"))
    t, msgs, thetas = instances.random_node_batch(5, d, D, seed=100 + 10 * d + D)
    T = B.make_from_numpy(t)
    ms = tuple(B.make_from_numpy(m) for m in msgs)
    th = tuple(B.make_from_numpy(x.astype(np.complex128)) for x in thetas)
    close(np.stack([p.numpy for p in T.pass_msgs(ms)]), g[f"pass_d{d}_D{D}"], 1e-12)
    close(np.stack([p.numpy for p in T.pass_msgs(ms, th)]), g[f"ext_d{d}_D{D}"], 1e-12)
    close(T.get_density_matrices(ms), g[f"rho_d{d}_D{D}"], 1e-12)


# ---- the unmodified engine through the backend (reference tests/test_small_circuit_final_density.py:9-20) -----------------
def run_reference_engine(backend_name, cfg):
    """bqa's own run_qa loop (src/bqa/core.py:13-35, src/bqa/state.py) with the backend's Tensor methods"""
    install_ref.add_to_path()
    import bqa
    if backend_name == "b200":
        from bqa_b200 import register_with_bqa
        register_with_bqa()
        return dict(bqa.run_qa({**cfg, "backend": "b200"}, fused=False))
    return dict(bqa.run_qa({**cfg, "backend": "numpy"}))


@pytest.mark.parametrize("backend_name", BACKENDS)
def test_unmodified_engine_small_circuit_matches_exact_state_vector(backend_name):
    from oracle import bqa_oracle as O
    cfg = instances.cfg_small6()
    got = np.array(run_reference_engine(backend_name, cfg)["bloch_vectors"])
    assert np.abs(got - O.run_exact_statevector(cfg)).max() < 1e-5


@pytest.mark.parametrize("backend_name", BACKENDS)
@pytest.mark.parametrize("name", ["ring24", "grid4", "comb"])
def test_unmodified_engine_equals_reference_goldens(golden_dir, backend_name, name):
    """state.run_layer / _run_bp / measure / get_density_matrices of the reference, op by op on the backend: Bloch vectors
    and sampled bitstrings of the goldens (degree classes 1..4, damping, truncation, measurement)."""
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    res = run_reference_engine(backend_name, instances.GOLDEN_CONFIGS[name]())
    assert np.abs(np.array(res["bloch_vectors"]) - g["bloch"]).max() < 1e-8
    assert res["measurement_outcomes"] == g["outcomes"].tolist()


@pytest.mark.gpu
def test_fused_dispatch_is_the_default_and_agrees():
    install_ref.add_to_path()
    import bqa
    from bqa_b200 import _lib, register_with_bqa
    register_with_bqa()
    cfg = {**instances.cfg_ring24(), "backend": "b200"}
    lib = _lib.load_library()
    before = lib.launch_count()
    fused = dict(bqa.run_qa(cfg, precision="double"))
    n_fused = lib.launch_count() - before
    unfused = dict(bqa.run_qa(cfg, fused=False))
    n_unfused = lib.launch_count() - before - n_fused
    assert np.abs(np.array(fused["bloch_vectors"]) - np.array(unfused["bloch_vectors"])).max() < 1e-9
    assert fused["measurement_outcomes"] == unfused["measurement_outcomes"]
    assert n_fused * 5 < n_unfused                                          # op-by-op launches vs the fused engine


def test_backend_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    install_ref.add_to_path()
    from bqa_b200 import register_with_bqa
    cls = register_with_bqa()
    with pytest.raises(RuntimeError, match="CUDA|libbqa_b200"):
        cls.make_from_list([1, 2, 3])
