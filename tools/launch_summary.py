#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals of ONE warm step.
A step = the launches from one `tc_prep_kernel<.., false/0>` (operand prep of X, the first kernel of a step) up to the next.
    python tools/launch_summary.py launches.csv [step_index_from_end=1] [min_rows=...]"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = list(csv.reader(lines))
hdr = r[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for row in r[1:]:
    if len(row) <= vi: continue
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    rows.append((row[ki], ms))
# a step starts at the first knn/linear kernel?  use the pair of tc_prep launches as the anchor inside the step, and cut at the
# kernel that starts match_deform: the first launch after the previous step's last kernel.  Simpler: cut at every 2nd tc_prep.
anchors = [i for i, (k, _) in enumerate(rows) if "tc_prep_kernel" in k]
starts = anchors[::2]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
# step = [starts[-which-1], starts[-which])
a, b = starts[-which - 1], starts[-which]
step = rows[a:b]
tot = sum(ms for _, ms in step)
agg = collections.OrderedDict()
for k, ms in step:
    e = agg.setdefault(k, [0, 0.0]); e[0] += 1; e[1] += ms
print(f"launches {a}..{b} of {len(rows)}: {len(step)} launches, {tot:.3f} ms\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:90]}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
