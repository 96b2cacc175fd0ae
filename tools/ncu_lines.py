#!/usr/bin/env python
"""Executed-instruction and stall-sample share per CUDA source line of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py rep kernel_regex [top] [skip]     (skip = number of matching launches to skip)"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx, "-s", skip, "-c", "1"], capture_output=True, text=True).stdout
cur = None; agg = collections.Counter(); smp = collections.Counter(); text = {}; tot = ts = 0
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': iex = r.index('Instructions Executed'); ism = r.index('# Samples'); continue
    try: ln = int(r[0]); ex = int(r[iex]); sm = int(r[ism])
    except (ValueError, IndexError): continue
    agg[(cur, ln)] += ex; smp[(cur, ln)] += sm; text[(cur, ln)] = r[1].strip()[:100]; tot += ex; ts += sm
print("total warp instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100*v/tot:5.1f}% ex {100*smp[k]/max(ts,1):5.1f}% smp  {k[0]}:{k[1]}  {text[k]}")
