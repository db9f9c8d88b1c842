"""Generates tests/golden/*.npz by running the UNMODIFIED reference (LuchnikovI/bqa v0.1.6, numpy backend,
complex128) in the build container.  The reference is pure Python and cannot travel to the GPU box, so its
outputs are committed as fixtures.

Run (build container only):

    mkdir -p /tmp/bqa_shim/bqa-0.1.6.dist-info
    printf 'Metadata-Version: 2.1\\nName: bqa\\nVersion: 0.1.6\\n' > /tmp/bqa_shim/bqa-0.1.6.dist-info/METADATA
    PYTHONPATH=/root/reference/src:/tmp/bqa_shim python tests/golden/make_golden.py

(the dist-info stub is needed because reference src/bqa/__init__.py:1-3 asks importlib.metadata for its
version).  Instances are produced by tests/instances.py (shared with the tests, no reference code) so the
tests rebuild the identical inputs.
"""
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))          # tests/
logging.disable(logging.WARNING)

import bqa                                          # noqa: E402  (the reference)
from bqa.backends import NumPyBackend               # noqa: E402
from bqa.config.core import config_to_context       # noqa: E402
from bqa import state as rstate                     # noqa: E402
from bqa.utils import convert_density_matrix_to_bloch_vector   # noqa: E402

import instances                                    # noqa: E402


def run_instrumented(config):
    """Drives the reference engine instruction by instruction and records gauge-invariant observables."""
    ctx = config_to_context(dict(config))
    st = rstate._initialize_state(ctx)
    sweeps_per_bp = []
    orig_get_dist = NumPyBackend.get_dist
    counter = {"n": 0}

    def counting_get_dist(self, other):
        counter["n"] += 1
        return orig_get_dist(self, other)

    NumPyBackend.get_dist = counting_get_dist
    bond_dims, lmbd_snap, bloch_trace = [], [], []
    results = {}
    try:
        for ins in ctx.instructions:
            if isinstance(ins, dict):
                counter["n"] = 0
                rstate.run_layer(ctx, ins["xtime"], ins["ztime"], st)
                sweeps_per_bp.append(counter["n"])
                bond_dims.append(st.bond_dim)
                lmbd_snap.append(np.sort(st.lmbds.numpy.real, axis=1)[:, ::-1].copy())
            elif ins == "get_bloch_vectors":
                rho = rstate.get_density_matrices(ctx, st)
                results["bloch"] = np.array([convert_density_matrix_to_bloch_vector(r) for r in rho])
            elif ins == "measure":
                results["outcomes"] = np.array(rstate.measure(ctx, st), np.int64)
    finally:
        NumPyBackend.get_dist = orig_get_dist
    results["bp_sweeps"] = np.array(sweeps_per_bp, np.int64)
    results["bond_dims"] = np.array(bond_dims, np.int64)
    results["lmbds_last"] = lmbd_snap[-1]
    results["lmbds_mid"] = lmbd_snap[len(lmbd_snap) // 2]
    return results


def kernel_level_vectors():
    """Backend-composite goldens on seeded random inputs (reference backends.py:381-448, :416-432)."""
    out = {}
    for d, D in [(1, 3), (2, 4), (3, 4), (3, 2), (4, 3)]:
        B = 5
        t, msgs, thetas = instances.random_node_batch(B, d, D, seed=100 + 10 * d + D)
        T = NumPyBackend(t)
        ms = tuple(NumPyBackend(m) for m in msgs)
        th = tuple(NumPyBackend(x.astype(np.complex128)) for x in thetas)
        plain = T.pass_msgs(ms)
        ext = T.pass_msgs(ms, th)
        rho = T.get_density_matrices(ms)
        out[f"pass_d{d}_D{D}"] = np.stack([p.numpy for p in plain])
        out[f"ext_d{d}_D{D}"] = np.stack([p.numpy for p in ext])
        out[f"rho_d{d}_D{D}"] = rho.numpy
    return out


def main():
    for name, cfg in instances.GOLDEN_CONFIGS.items():
        res = run_instrumented(cfg())
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **res)
        print(name, {k: v.shape for k, v in res.items()})
    kv = kernel_level_vectors()
    np.savez_compressed(os.path.join(HERE, "kernel_level.npz"), **kv)
    print("kernel_level", len(kv))
    # the public entry point must agree with the instrumented drive
    cfg = instances.GOLDEN_CONFIGS["small6"]()
    direct = np.array(bqa.run_qa(cfg)[0][1])
    assert np.abs(direct - np.load(os.path.join(HERE, "small6.npz"))["bloch"]).max() < 1e-13


if __name__ == "__main__":
    main()
