"""Experiment matrix for the sweep (DVM_TC_DEBUG switches): hard + soft candidate pass at 4 x 50k."""
import ctypes, json, os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dv_matcher_b200 import _lib, ops, synthetic
    lib = _lib.load()
    d = synthetic.make_batch(2, 50000, 50000)
    x = torch.cat([d["feat1"], d["feat2"]]).cuda(); y = torch.cat([d["feat2"], d["feat1"]]).cuda()
    out = {}
    for soft in (False, True):
        for _ in range(2):
            ops.softmap_fwd(x, y, None, alpha=100.0, topk=10 if soft else 1, soft=soft, prec="f16")
        torch.cuda.synchronize()
        lib.dvm_profile_enable(1)
        for _ in range(5):
            ops.softmap_fwd(x, y, None, alpha=100.0, topk=10 if soft else 1, soft=soft, prec="f16")
        torch.cuda.synchronize()
        tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
        lib.dvm_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
        lib.dvm_profile_enable(0)
        ms = tot.value / cnt.value
        out["soft" if soft else "hard"] = dict(cand_ms=round(ms, 3), tflops=round(2 * 50000 * 50000 * 128 * 4 / ms / 1e9, 1))
    print(json.dumps(dict(debug=os.environ.get("DVM_TC_DEBUG", "0"), **out)), flush=True)
else:
    for dbg in sys.argv[1:] or ["0", "4", "1", "5", "3", "7", "8", "12"]:
        env = dict(os.environ, DVM_TC_DEBUG=dbg)
        subprocess.run([sys.executable, __file__, "child"], env=env)
