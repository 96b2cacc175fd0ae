import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import _lib, ops, synthetic
n = 50000
d = synthetic.make_batch(1, n, n)
x = torch.cat([d["feat1"], d["feat2"]]).cuda(); y = torch.cat([d["feat2"], d["feat1"]]).cuda()
L = _lib.load()
for soft in (True, False):
    for _ in range(2):
        ops.softmap_fwd(x, y, None, alpha=100.0, topk=10 if soft else 1, soft=soft, prec="f16")
    torch.cuda.synchronize()
    L.dvm_profile_enable(1)
    for _ in range(5):
        ops.softmap_fwd(x, y, None, alpha=100.0, topk=10 if soft else 1, soft=soft, prec="f16")
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
    L.dvm_profile_read(ctypes.byref(tot), ctypes.byref(cnt)); L.dvm_profile_enable(0)
    print("split", os.environ.get("DVM_TC_SPLIT"), "soft", soft, "cand_ms", round(tot.value / cnt.value, 3), flush=True)
