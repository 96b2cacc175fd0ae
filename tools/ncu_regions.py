#!/usr/bin/env python
"""Per-region breakdown of executed instructions / stall samples from an .ncu-rep source page.
    python tools/ncu_regions.py rep [step] [lo hi]   (lo hi: dump individual SASS lines in that index range)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; body = rows[2:]
isrc = h.index('Source'); ismp = h.index('# Samples'); iex = h.index('Instructions Executed'); ith = h.index('Thread Instructions Executed'); iav = h.index('Avg. Threads Executed')
tot = sum(int(b[iex]) for b in body); ts = sum(int(b[ismp]) for b in body)
print('kernel', rows[0][1][:80]); print('total inst', tot, 'samples', ts, 'sass lines', len(body))
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for i in range(lo, min(hi, len(body))):
        b = body[i]
        print(f"{i:5d} ex={100*int(b[iex])/tot:5.2f}% smp={100*int(b[ismp])/ts:5.2f}% thr={b[iav]:>4s} {b[isrc].strip()[:110]}")
    sys.exit(0)
for s in range(0, len(body), step):
    seg = body[s:s + step]
    e = sum(int(b[iex]) for b in seg); sm = sum(int(b[ismp]) for b in seg); th = sum(int(b[ith]) for b in seg)
    ops = {}
    for b in seg:
        t = b[isrc].strip().split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op] = ops.get(op, 0) + int(b[iex])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    if e > 0.004 * tot or sm > 0.004 * ts:
        print(f"{s:5d} ex={100*e/tot:5.2f}% smp={100*sm/ts:5.2f}% thr={th/max(e,1):5.1f} {[(k, round(100*v/tot, 2)) for k, v in top]}")
