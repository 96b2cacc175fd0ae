#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key metrics table + top stall reasons by source line.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--source N]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
want = re.compile(r'^(gpu__time_duration.sum|dram__bytes_(read|write).sum|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic)|sm__cycles_elapsed.max|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum|sm__inst_executed_pipe_(xu|alu|fma|fmaheavy|lsu|tmem|uniform).avg.pct_of_peak_sustained_active|sm__pipe_tensor.*cycles_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum|lts__t_sector_hit_rate.pct|l1tex__t_bytes.sum|smsp__cycles_active.avg|sm__cycles_active.avg|lts__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__throughput.avg.pct_of_peak_sustained_elapsed|smsp__thread_inst_executed_per_inst_executed.ratio)$')
for row in r[2:]:
    print("## " + row[hdr.index("Kernel Name")][:100])
    print("| metric | value | unit |\n|---|---:|---|")
    for h, u, v in zip(hdr, units, row):
        if want.search(h):
            print(f"| {h} | {v} | {u} |")
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    for i, row in enumerate(rows):
        if "Source" in row and any("Samples" in c for c in row):
            h = row; body = rows[i + 1:]; break
    else:
        print("no source page"); sys.exit(0)
    si = h.index("Source")
    smp = [i for i, c in enumerate(h) if c.strip() in ("# Samples", "Warp Stall Sampling (All Samples)", "Sampling Data (All)")]
    ins = [i for i, c in enumerate(h) if c.strip() in ("Instructions Executed", "# Instructions Executed")]
    print("\ncolumns:", h[:12])
    key = smp[0] if smp else None
    if key is not None:
        body = [b for b in body if len(b) > key and b[key].replace(',', '').isdigit()]
        body.sort(key=lambda b: -int(b[key].replace(',', '')))
        tot = sum(int(b[key].replace(',', '')) for b in body)
        print(f"\ntop {n} lines by stall samples (total {tot}):")
        for b in body[:n]:
            ie = b[ins[0]] if ins else ""
            print(f"{int(b[key].replace(',', '')) * 100.0 / tot:5.1f}%  inst={ie:>10s}  {b[si][:140]}")
