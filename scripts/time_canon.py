"""Times bqa_b200_canonicalize on the extended messages of the benchmarked instance (100k-qubit 3-regular QUBO after
`--steps` annealing steps) for the current n = 8 kernel (mode 0) and the first design (mode 2); prints one JSON line.

    python scripts/time_canon.py [--qubits 100000] [--steps 40] [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import instances
    from bqa_b200 import _lib
    from bqa_b200.config import config_to_context
    from bqa_b200.engine import Engine
    if os.environ.get("BQA_B200_LIB_EXPERIMENT"):            # side-by-side builds of the library (scratch/, not the product)
        _lib._cached = _lib.bind(os.environ["BQA_B200_LIB_EXPERIMENT"])
    lib = _lib.load_library()
    ctx = config_to_context(instances.bench_config(args.qubits))
    eng = Engine(ctx, precision="single")
    layers = [i for i in ctx.instructions if isinstance(i, dict)]
    for ins in layers[:args.steps]:
        eng.run_layer(ins["xtime"], ins["ztime"])
    # the extended messages of the next step, as the canonicalizer would see them
    ins = layers[args.steps]
    st = torch.cuda.current_stream().cuda_stream
    c = eng.classes[0]
    lib.ext_msgs(eng.prec, c.degree, eng.D, c.B, c.T[c.cur].data_ptr(), eng.msgs_buffer.data_ptr(), eng._ext.data_ptr(),
                 c.in_pos.data_ptr(), c.out_pos.data_ptr(), c.edge_ampls.data_ptr(), float(ins["ztime"]),
                 eng._ws.data_ptr(), eng._ws.numel(), st)
    out = {"qubits": args.qubits, "edges": eng.L, "D": eng.D, "after_steps": args.steps}
    res = {}
    for mode in (2, 0):
        lib.set_kernel_mode(mode)
        canon = torch.zeros_like(eng._canon)
        lm = torch.zeros_like(eng._lmbds)
        colmax = torch.zeros(8, dtype=torch.float32, device=eng.dev)
        call = lambda: lib.canonicalize(eng.prec, eng.D, eng.L, eng._ext.data_ptr(), canon.data_ptr(), lm.data_ptr(),
                                        colmax.data_ptr(), eng.pinv_eps, 4, st)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        s0 = lib.canon_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        s1 = lib.canon_stats()
        det = lib.canon_stats_detail() if mode == 0 else None
        ms = e0.elapsed_time(e1) / args.reps
        # every launch on its own: does the duration drift under back-to-back load (clocks)?
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for a, b in evs:
            a.record()
            call()
            b.record()
        torch.cuda.synchronize()
        each = [a.elapsed_time(b) for a, b in evs]
        res[mode] = (canon.cpu().numpy().reshape(2 * eng.L, 8, 8), lm.cpu().numpy().reshape(eng.L, 8), colmax.cpu().numpy())
        out[f"mode{mode}"] = {"ms": ms, "ms_first": each[0], "ms_min": min(each), "ms_last": each[-1], "edges_per_s": eng.L / (ms * 1e-3),
                              "sweeps_per_warp_run": (s1[1] - s0[1]) / max(s1[0] - s0[0], 1),
                              "ker_share_of_sweeps": (s1[2] - s0[2]) / max(s1[1] - s0[1], 1)}
        if det:
            out["mode0"]["mean_sweeps_per_matrix"] = {"eigen": det[3] / max(det[5], 1), "svd": det[4] / max(det[6], 1)}
    # one launch after an idle second vs back to back (boost clocks?), with nvidia-smi samples during a 2 s loop
    import subprocess, time as _t
    lib.set_kernel_mode(0)
    iso = []
    for _ in range(4):
        torch.cuda.synchronize(); _t.sleep(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize()
        iso.append(a.elapsed_time(b))
    out["mode0_single_launch_after_idle_ms"] = iso
    lib.canon_span()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); call(); b.record(); torch.cuda.synchronize()
    t0, t1 = lib.canon_span()
    out["mode0_one_launch"] = {"event_ms": a.elapsed_time(b), "first_cta_start_to_last_cta_end_ms": (t1 - t0) / 1e6}
    try:
        smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader",
                                "-lms", "100"], stdout=subprocess.PIPE, text=True)
        t0 = _t.time()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nloop = 0
        while _t.time() - t0 < 2.0:
            for _ in range(50):
                call()
            nloop += 50
            torch.cuda.synchronize()
        b.record(); torch.cuda.synchronize()
        smi.terminate()
        lines = smi.stdout.read().strip().splitlines()
        out["mode0_sustained_ms"] = a.elapsed_time(b) / nloop
        out["smi_during_loop"] = lines[2:: max(len(lines) // 6, 1)][:8]
    except Exception as e:
        out["smi_error"] = repr(e)
    out["lambda_max_abs_diff"] = float(np.abs(res[0][1] - res[2][1]).max())
    out["colmax"] = [res[0][2].tolist(), res[2][2].tolist()]
    import subprocess
    try:
        out["clocks_after"] = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                                             capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception:
        pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
