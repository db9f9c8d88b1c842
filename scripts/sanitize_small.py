"""Short anneals that touch every kernel family, for compute-sanitizer (memcheck / racecheck):
  * 24-qubit 3-regular instance, complex64, D 1 -> 4: generic kernels below D = 4, then the specialised BP run,
    extended messages (incl. the launch enqueued behind the BP run), n = 8 canonicalizer, apply, gauge rows; marginals
    and sampling;
  * 4 x 4 grid (three degree classes), complex64 and complex128: the table-driven multi-class kernels and the generic
    canonicalizer.
Usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import instances
    from bqa_b200 import run_qa
    cfg = instances.cfg_ring24()
    cfg["schedule"]["actions"][0]["steps_number"] = 8
    out = dict(run_qa(cfg, precision="single", device="cuda:0"))
    print("ring24 single ok", len(out["bloch_vectors"]))
    for precision in ("single", "double"):
        cfg = copy.deepcopy(instances.cfg_grid4())
        cfg["schedule"]["actions"][0]["steps_number"] = 6
        out = dict(run_qa(cfg, precision=precision, device="cuda:0"))
        print("grid4", precision, "ok", len(out["bloch_vectors"]))


if __name__ == "__main__":
    main()
