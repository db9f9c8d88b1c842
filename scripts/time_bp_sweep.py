"""Times one BP sweep launch (degree 3, D = 4, complex64) on a 100k-node random 3-regular layout and checks it against
the generic kernel.  Development tool: `python scripts/time_bp_sweep.py [B]`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bqa_b200 import _lib  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    d, D = 3, 4
    lib = _lib.bind(os.environ["BQA_LIB"]) if os.environ.get("BQA_LIB") else _lib.load_library()
    rng = np.random.default_rng(1)
    dev = torch.device("cuda:0")
    nslots = d * B
    # random 3-regular-like wiring: every slot is the in-slot of one (node, leg) and the out-slot of another
    in_pos = rng.permutation(nslots).reshape(d, B).astype(np.int32)
    out_pos = rng.permutation(nslots).reshape(d, B).astype(np.int32)
    t = (rng.normal(size=(B, 2 * D ** 3)) + 1j * rng.normal(size=(B, 2 * D ** 3))).astype(np.complex64)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    a = (rng.normal(size=(nslots, D, D)) + 1j * rng.normal(size=(nslots, D, D))).astype(np.complex64)
    m = a @ np.swapaxes(a.conj(), 1, 2)
    m /= np.trace(m, axis1=1, axis2=2)[:, None, None]
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    T, C, ip, op = up(t.reshape(-1)), up(m.astype(np.complex64).reshape(-1)), up(in_pos), up(out_pos)
    ws = torch.zeros(lib.workspace_bytes(_lib.C64, d, D, D), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    out = {}
    for mode in (1, 0):
        lib.set_kernel_mode(mode)
        nxt = C.clone()
        resid = torch.zeros(2, dtype=torch.float32, device=dev)
        status = torch.zeros(4, dtype=torch.int32, device=dev)

        def launch():
            lib.bp_sweep(_lib.C64, d, D, B, T.data_ptr(), C.data_ptr(), nxt.data_ptr(), ip.data_ptr(), op.data_ptr(),
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
        out[mode] = (nxt.cpu().numpy(), resid.cpu().numpy(), ts[len(ts) // 2] * 1e3)
    lib.set_kernel_mode(0)
    err = np.abs(out[0][0] - out[1][0]).max()
    us = out[0][2]
    print(f"B={B} fast {us:.1f} us/launch ({2176 * B / us * 1e-3:.0f} GB/s algorithmic), generic {out[1][2]:.1f} us, "
          f"max |fast - generic| = {err:.2e}, resid fast {out[0][1]} generic {out[1][1]}")
    assert err < 2e-6


if __name__ == "__main__":
    main()
