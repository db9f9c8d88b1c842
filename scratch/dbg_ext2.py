import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, instances
from bqa_b200 import _lib
lib = _lib.load_library()
d, D = 3, 4
for B in (1, 4):
    rng = np.random.default_rng(B)
    t, msgs, thetas = instances.random_node_batch(B, d, D, seed=7 + B)
    slots = rng.permutation(d * B + 5)
    in_pos = slots[: d * B].reshape(d, B).astype(np.int32)
    out_pos = rng.permutation(d * B + 5)[: d * B].reshape(d, B).astype(np.int32)
    cur = np.zeros((d * B + 5, D, D), np.complex64)
    cur[:] = instances.random_psd_msgs(rng, d * B + 5, D)
    for j in range(d):
        cur[in_pos[j]] = msgs[j]
    dev = torch.device('cuda:0')
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    T, C, ip, op = up(t.astype(np.complex64).reshape(-1)), up(cur.reshape(-1)), up(in_pos), up(out_pos)
    ea = up(np.stack(thetas).astype(np.float32))
    ws = torch.zeros(lib.workspace_bytes(_lib.C64, d, D, D), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for mode in (1, 0):
        lib.set_kernel_mode(mode)
        ext = torch.zeros((d * B + 5) * 4 * D * D, dtype=torch.complex64, device=dev)
        lib.ext_msgs(_lib.C64, d, D, B, T.data_ptr(), C.data_ptr(), ext.data_ptr(), ip.data_ptr(), op.data_ptr(), ea.data_ptr(), 0.7, ws.data_ptr(), ws.numel(), st)
        e = ext.cpu().numpy().reshape(-1, 8, 8)
        nanslots = np.where(np.isnan(e).any(axis=(1, 2)))[0]
        print('B', B, 'mode', mode, 'nan slots', nanslots, 'out_pos', out_pos.reshape(-1), 'in_pos', in_pos.reshape(-1))
    lib.set_kernel_mode(0)
