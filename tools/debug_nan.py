import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, _lib
z = np.load("tests/golden/ref_maps.npz"); q = float(z["qstep"])
x = torch.from_numpy(z["feat1_q"].astype(np.float32) * np.float32(q))[None].cuda()
y = torch.from_numpy(z["feat2_q"].astype(np.float32) * np.float32(q))[None].cuda()
v = torch.from_numpy(z["xyz2"])[None].cuda()
alpha = 10.0
o = ops.softmap_fwd(x, y, v, alpha=alpha, prec="f16", want_stats=True)
torch.cuda.synchronize()
ws = [b for k, b in _lib.workspace.buf.items() if k[2] == "softmap"][0]
rows, P4, KC, P = 1500, 4, 16, 2
def up(o): return (o + 255) // 256 * 256
off = 0
def take(n, size):
    global off
    off = up(off); s = off; off += n * size; return s
take(rows*P4*KC, 4); take(rows*P4*KC, 4); take(rows*P4, 4); take(rows*P4, 4); take(4, 4)
k_off = take(rows*P*KC, 4); i_off = take(rows*P*KC, 4); l_off = take(rows*P, 4); r_off = take(rows*P, 4)
raw = ws.cpu().numpy()
key = raw[k_off:k_off+rows*P*KC*4].view(np.float32).reshape(rows, P, KC)
idx = raw[i_off:i_off+rows*P*KC*4].view(np.int32).reshape(rows, P, KC)
l = raw[l_off:l_off+rows*P*4].view(np.float32).reshape(rows, P)
r = raw[r_off:r_off+rows*P*4].view(np.float32).reshape(rows, P)
bad = np.isnan(o.row_sum[0].cpu().numpy())
print("nan rows", bad.sum(), "l nan", np.isnan(l).sum(), "l inf", np.isinf(l).sum(), "r inf", np.isinf(r).sum(), "r nan", np.isnan(r).sum())
for i in np.nonzero(bad)[0][:3]:
    print("row", i, "l", l[i], "r", r[i], "keys0", key[i,0,:4], key[i,0,-2:], "keys1", key[i,1,:4], key[i,1,-2:], "idx0", idx[i,0,:3], "idx1", idx[i,1,:3])
good = np.nonzero(~bad)[0][:2]
for i in good:
    print("good row", i, "l", l[i], "r", r[i], "keys0", key[i,0,:3], key[i,0,-2:], "keys1", key[i,1,:3], key[i,1,-2:])
