"""Accuracy of the complex64 engine against the reference golden window of the benchmarked 100k-qubit instance
(tests/golden/rr100k_window.npz: unmodified reference, complex128, 33 steps) per kernel mode:
0 = current kernels, 2 = first-design n = 8 canonicalizer.  One JSON line per mode."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import instances
    from bqa_b200 import _lib
    from bqa_b200.benchmarking import ising_energy
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    lib = _lib.load_library()
    g = np.load(os.path.join(ROOT, "tests", "golden", "rr100k_window.npz"))
    cfg = instances.bench_config(100_000)
    ctx = config_to_context(cfg)
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    for mode in [int(m) for m in (sys.argv[1:] or ["2", "0"])]:
        lib.set_kernel_mode(mode)
        eng = Engine(ctx, precision="single")
        for ins in layers[:len(g["bp_sweeps"])]:
            eng.run_layer(ins["xtime"], ins["ztime"])
        b = eng.bloch_vectors()
        lm = np.sort(eng.lmbds_numpy(), axis=1)[:, ::-1]
        d = np.abs(b - g["bloch"])
        dl = np.abs(lm[::8] - g["lmbds_sorted_strided"])
        e = ising_energy(cfg["edges"], cfg["nodes"], np.where(b[:, 2] > 0, 1.0, -1.0))
        print(json.dumps({"mode": mode, "bloch_max": float(d.max()), "bloch_mean": float(d.mean()),
                          "bloch_p999": float(np.quantile(d, 0.999)), "lmbd_max": float(dl.max()), "lmbd_mean": float(dl.mean()),
                          "lmbd_max_per_col": dl.max(0).tolist(), "energy_rel": abs(e - float(g["energy"])) / abs(float(g["energy"])),
                          "sweeps_diff_max": int(np.abs(np.array(eng.stats["bp_sweeps"]) - g["bp_sweeps"]).max()),
                          "dims_equal": eng.stats["bond_dims"] == g["bond_dims"].tolist()}), flush=True)
        del eng
    lib.set_kernel_mode(0)


if __name__ == "__main__":
    main()
