"""Generates tests/golden/rr100k_window.npz: the UNMODIFIED reference (LuchnikovI/bqa v0.1.6, numpy backend,
complex128) on the benchmarked instance -- BASELINE.json configs[3], the 100 000-qubit random 3-regular QUBO of
bench.py -- for the truncated window of BASELINE.md section 4 step 4: the 30 ramp steps that take the bond dimension
1 -> 4 plus 3 steady-state steps of the bench schedule (S = 100, dt = 0.2).

Run (build container only; about 10 minutes on 8 cores, most of it in the D = 4 steps):

    mkdir -p /tmp/bqa_shim/bqa-0.1.6.dist-info
    printf 'Metadata-Version: 2.1\\nName: bqa\\nVersion: 0.1.6\\n' > /tmp/bqa_shim/bqa-0.1.6.dist-info/METADATA
    PYTHONPATH=/root/reference/src:/tmp/bqa_shim python tests/golden/make_golden_100k.py

Stored (gauge-invariant quantities only): Bloch vectors after the ramp and after the window,
sorted lambda spectra of every 8th edge after the window, column maxima of the lambdas, BP sweep counts, bond
dimensions, the energy of sign(z), and a fingerprint of the instance so that the GPU box can prove it rebuilt the
same graph and amplitudes.
"""
import hashlib
import logging
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
logging.disable(logging.WARNING)

import bqa                                          # noqa: E402,F401  (the reference)
from bqa.backends import NumPyBackend               # noqa: E402
from bqa.benchmarking import generate_qubo_on_random_regular_graph   # noqa: E402
from bqa.config.core import config_to_context       # noqa: E402
from bqa import state as rstate                     # noqa: E402

N_QUBITS = 100_000
SCHEDULE_STEPS = 100
RAMP = 30
WINDOW = 3
DT = 0.2
LMBD_STRIDE = 8


def instance_fingerprint(nodes: dict, edges: dict) -> str:
    h = hashlib.sha256()
    h.update(np.asarray(list(nodes.keys()), np.int64).tobytes())
    h.update(np.asarray(list(nodes.values()), np.float64).tobytes())
    h.update(np.asarray(list(edges.keys()), np.int64).tobytes())
    h.update(np.asarray(list(edges.values()), np.float64).tobytes())
    return h.hexdigest()


def bloch_of(ctx, st) -> np.ndarray:
    rho = rstate.get_density_matrices(ctx, st)          # (N, 2, 2), reference utils.py:23-27 vectorised below
    return np.stack([(rho[:, 0, 1] + rho[:, 1, 0]).real, (rho[:, 1, 0] - rho[:, 0, 1]).imag,
                     (rho[:, 0, 0] - rho[:, 1, 1]).real], axis=1)


def main():
    n = int(os.environ.get("BQA_GOLDEN_QUBITS", N_QUBITS))
    nodes, edges = generate_qubo_on_random_regular_graph(n, 3, seed=42)
    cfg = {"nodes": nodes, "edges": edges, "max_bond_dim": 4, "bp_eps": 1e-6, "pinv_eps": 1e-6,
           "damping": 0.0, "max_bp_iter_number": 75, "seed": 42, "default_field": 0.0,
           "measurement_threshold": 0.95, "backend": "numpy",
           "schedule": {"total_time": DT * SCHEDULE_STEPS, "starting_mixing": 1.0,
                        "actions": [{"type": "real_time_evolution", "weight": 1.0, "steps_number": SCHEDULE_STEPS,
                                     "final_mixing": 0.0}]}}
    ctx = config_to_context(cfg)
    st = rstate._initialize_state(ctx)
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    counter = {"n": 0}
    orig = NumPyBackend.get_dist

    def counting(self, other):
        counter["n"] += 1
        return orig(self, other)
    NumPyBackend.get_dist = counting
    sweeps, dims, secs = [], [], []
    out = {}
    for k, ins in enumerate(layers[:RAMP + WINDOW]):
        counter["n"] = 0
        t0 = time.perf_counter()
        rstate.run_layer(ctx, ins["xtime"], ins["ztime"], st)
        secs.append(time.perf_counter() - t0)
        sweeps.append(counter["n"])
        dims.append(st.bond_dim)
        print(f"step {k}: D = {st.bond_dim}, {sweeps[-1]} sweeps, {secs[-1]:.1f} s", flush=True)
        if k + 1 == RAMP:
            out["bloch_ramp"] = bloch_of(ctx, st)
    NumPyBackend.get_dist = orig
    bloch = bloch_of(ctx, st)
    lm = np.sort(st.lmbds.numpy.real, axis=1)[:, ::-1]
    spins = np.where(bloch[:, 2] > 0, 1.0, -1.0)
    e = sum(j * spins[a] * spins[b] for (a, b), j in edges.items()) + sum(h * spins[i] for i, h in nodes.items())
    out.update(bloch=bloch, lmbds_sorted_strided=lm[::LMBD_STRIDE].copy(), lmbds_colmax=lm.max(axis=0),
               lmbds_mean=lm.mean(axis=0), bp_sweeps=np.array(sweeps, np.int64), bond_dims=np.array(dims, np.int64),
               energy=np.float64(e), seconds_per_step=np.array(secs), cores=np.int64(os.cpu_count()),
               fingerprint=np.array(instance_fingerprint(nodes, edges)),
               meta=np.array(f"reference bqa v0.1.6 numpy complex128; n={n}; S={SCHEDULE_STEPS}; ramp={RAMP}; "
                             f"window={WINDOW}; dt={DT}; lmbd_stride={LMBD_STRIDE}"))
    name = "rr100k_window.npz" if n == N_QUBITS else f"rr{n}_window.npz"
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
