#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + top stall lines from the source page."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
for k in keys:
    for i, h in enumerate(hdr):
        if h == k:
            print("%-75s %-10s %s" % (k, units[i], " | ".join(v[i][:60] for v in vals)))
for i, h in enumerate(hdr):
    if "tensor" in h and h not in keys:
        print("%-75s %-10s %s" % (h, units[i], " | ".join(v[i][:40] for v in vals)))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if "Source" in r or "# Samples" in " ".join(r))
    h = rows[hi]
    print(h)
    si = next((i for i, x in enumerate(h) if x.startswith("# Samples") or x == "Warp Stall Sampling (All Samples)"), None)
    body = [r for r in rows[hi + 1:] if len(r) == len(h)]
    def num(x):
        try: return float(x)
        except: return 0.0
    body.sort(key=lambda r: -num(r[si]))
    tot = sum(num(r[si]) for r in body)
    for r in body[:int(sys.argv[2])]:
        print("%6.1f%%  %s" % (100 * num(r[si]) / max(tot, 1), " | ".join(x[:110] for x in r[:3])))
