import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import _lib, ops, synthetic
lib = ctypes.CDLL(_lib.LIB_PATH)
n = int(sys.argv[1]); soft = sys.argv[2] == "soft"
d = synthetic.make_batch(1, n, n)
x = torch.cat([d["feat1"], d["feat2"]]).cuda(); y = torch.cat([d["feat2"], d["feat1"]]).cuda()
out = (ctypes.c_ulonglong * 8)()
ops.softmap_fwd(x, y, None, alpha=100.0, topk=10 if soft else 1, soft=soft, prec="f16"); torch.cuda.synchronize()
lib.dvm_debug_tc_counters(out, 1)
ops.softmap_fwd(x, y, None, alpha=100.0, topk=10 if soft else 1, soft=soft, prec="f16"); torch.cuda.synchronize()
lib.dvm_debug_tc_counters(out, 1)
rows = 2 * n
names = ["entries", "rounds", "trips", "x", "consumer_idle_polls", "scanner_flow_spins"]
print({k: out[i] for i, k in enumerate(names)}, "per row:", {k: round(out[i] / rows, 2) for i, k in enumerate(names)})
