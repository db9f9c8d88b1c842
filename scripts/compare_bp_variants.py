"""Development tool: times the BP sweep kernels (degree 3, D = 4, complex64) of several builds of the library on the
same synthetic 100k-node layout and checks every build against the first one.

    python scripts/compare_bp_variants.py lib_base.so lib_a.so lib_b.so ...   [B=100000]

Per library: the per-sweep launch (bqa_b200_bp_sweep, median of 40) and the single-launch run (bqa_b200_bp_run with
bp_eps = 0, 30 sweeps, time / 30)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bqa_b200 import _lib  # noqa: E402


def main():
    libs = [a for a in sys.argv[1:] if not a.startswith("B=")]
    B = next((int(a[2:]) for a in sys.argv[1:] if a.startswith("B=")), 100_000)
    d, D = 3, 4
    rng = np.random.default_rng(1)
    dev = torch.device("cuda:0")
    nslots = d * B
    in_pos = rng.permutation(nslots).reshape(d, B).astype(np.int32)
    out_pos = rng.permutation(nslots).reshape(d, B).astype(np.int32)
    t = (rng.normal(size=(B, 2 * D ** 3)) + 1j * rng.normal(size=(B, 2 * D ** 3))).astype(np.complex64)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    a = (rng.normal(size=(nslots, D, D)) + 1j * rng.normal(size=(nslots, D, D))).astype(np.complex64)
    m = a @ np.swapaxes(a.conj(), 1, 2)
    m /= np.trace(m, axis1=1, axis2=2)[:, None, None]
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    T, M0, ip, op = up(t.reshape(-1)), up(m.astype(np.complex64).reshape(-1)), up(in_pos), up(out_pos)
    st = torch.cuda.current_stream().cuda_stream
    ref = None
    rows = []
    for path in libs:
        lib = _lib.bind(os.path.abspath(path))
        ws = torch.zeros(lib.workspace_bytes(_lib.C64, d, D, D), dtype=torch.uint8, device=dev)
        nxt = M0.clone()
        resid = torch.zeros(2, dtype=torch.float32, device=dev)
        status = torch.zeros(4, dtype=torch.int32, device=dev)

        def launch():
            lib.bp_sweep(_lib.C64, d, D, B, T.data_ptr(), M0.data_ptr(), nxt.data_ptr(), ip.data_ptr(), op.data_ptr(),
                         0.0, 0, 1e-6, 0, resid.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), st)
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
        ev[0].record()
        for i in range(40):
            launch()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(40))
        sweep_us = ts[len(ts) // 2] * 1e3
        out1 = nxt.cpu().numpy()
        res1 = resid.cpu().numpy().copy()
        # single-launch run: 30 sweeps that never converge (bp_eps = 0), damping 0.3 exercises the damped store
        iters = 30
        run_us, out2 = float("nan"), None
        bufs = [M0.clone(), torch.zeros_like(M0)]
        ctrl = torch.zeros(2 * iters + 8, dtype=torch.float32, device=dev)
        stat = torch.zeros(4, dtype=torch.int32, device=dev)
        null8 = (C.c_void_p * 8)()

        def run():
            ctrl.zero_(); stat.zero_()
            return lib.bp_run(_lib.C64, d, D, B, T.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), 0, ip.data_ptr(),
                              op.data_ptr(), 0.3, 0.0, iters, ctrl.data_ptr(), stat.data_ptr(), None, null8, null8, 0, 1,
                              null8, null8, 0, st)
        ok = run()
        if ok:
            torch.cuda.synchronize()
            tt = []
            for _ in range(5):
                bufs[0].copy_(M0)
                ctrl.zero_(); stat.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.bp_run(_lib.C64, d, D, B, T.data_ptr(), bufs[0].data_ptr(), bufs[1].data_ptr(), 0, ip.data_ptr(),
                           op.data_ptr(), 0.3, 0.0, iters, ctrl.data_ptr(), stat.data_ptr(), None, null8, null8, 0, 1,
                           null8, null8, 0, st)
                e1.record()
                torch.cuda.synchronize()
                tt.append(e0.elapsed_time(e1) * 1e3 / iters)
            run_us = sorted(tt)[len(tt) // 2]
            out2 = (bufs[iters & 1].cpu().numpy(), ctrl.cpu().numpy().copy(), stat.cpu().numpy().copy())
        row = {"lib": os.path.basename(path), "sweep_us": round(sweep_us, 2), "run_us_per_sweep": round(run_us, 2),
               "sweep_GBs": round(2176 * B / sweep_us * 1e-3), "run_GBs": round(2176 * B / run_us * 1e-3)}
        if ref is None:
            ref = (out1, res1, out2)
        else:
            row["sweep_maxdiff"] = float(np.abs(out1 - ref[0]).max())
            row["sweep_biteq"] = bool(np.array_equal(out1.view(np.uint32), ref[0].view(np.uint32)))
            row["resid_rel"] = float(np.abs(res1 - ref[1]).max() / np.abs(ref[1]).max())
            if out2 is not None and ref[2] is not None:
                row["run_maxdiff"] = float(np.abs(out2[0] - ref[2][0]).max())
                row["run_biteq"] = bool(np.array_equal(out2[0].view(np.uint32), ref[2][0].view(np.uint32)))
                row["run_resid_rel"] = float(np.abs(out2[1] - ref[2][1]).max() / np.abs(ref[2][1]).max())
                row["run_status"] = out2[2].tolist()
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bp_variants.jsonl", "a") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
