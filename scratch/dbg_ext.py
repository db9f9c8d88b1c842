import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, instances
from bqa_b200 import _lib
lib = _lib.load_library()
d, D, B = 3, 4, 4
t, msgs, thetas = instances.random_node_batch(B, d, D, seed=3)
dev = torch.device('cuda:0')
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
T = up(t.astype(np.complex64).reshape(-1))
cur = up(np.concatenate(msgs, 0).astype(np.complex64).reshape(-1))
ip = up(np.arange(d * B, dtype=np.int32).reshape(d, B))
ea = up(np.stack(thetas).astype(np.float32))
ws = torch.zeros(lib.workspace_bytes(_lib.C64, d, D, D), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
res = {}
for mode in (1, 0):
    lib.set_kernel_mode(mode)
    ext = torch.zeros(d * B * 4 * D * D, dtype=torch.complex64, device=dev)
    lib.ext_msgs(_lib.C64, d, D, B, T.data_ptr(), cur.data_ptr(), ext.data_ptr(), ip.data_ptr(), ip.data_ptr(), ea.data_ptr(), 0.7, ws.data_ptr(), ws.numel(), st)
    res[mode] = ext.cpu().numpy().reshape(d * B, 8, 8)
    print('mode', mode, 'nan count', np.isnan(res[mode]).sum(), 'of', res[mode].size)
np.set_printoptions(linewidth=250, precision=4, suppress=True)
print(np.isnan(res[0][0]).astype(int))
print('thetas', np.stack(thetas)[:, 0] * 0.7)
ok = ~np.isnan(res[0])
print('max diff where finite', np.abs(res[0][ok] - res[1][ok]).max())
print(np.abs(res[0][0]-res[1][0]))
