"""Key metrics + hottest source lines of an .ncu-rep (read here, no GPU): python tools/ncu_show.py gpurun_out/x.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
M = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
     "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print("==", row[hdr.index("Kernel Name")][:90])
    for m in M:
        if m in hdr:
            print(f"   {m:84s} {row[hdr.index(m)]} {units[hdr.index(m)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
# the source page prints one table per kernel instance; keep the first
blocks = src.split("\n\n")
rows = list(csv.reader(io.StringIO(src)))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "#"]
for bi, h in enumerate(hdrs[:2]):
    end = hdrs[bi + 1] if bi + 1 < len(hdrs) else len(rows)
    head = rows[h]
    try:
        ci = head.index("Source"); cs = head.index("# Samples") if "# Samples" in head else head.index("Sampling Data (All)")
    except ValueError:
        print(head); continue
    body = [r for r in rows[h + 1:end] if len(r) > cs and r[cs].replace(".", "").isdigit()]
    tot = sum(float(r[cs]) for r in body) or 1
    body.sort(key=lambda r: -float(r[cs]))
    print(f"-- hottest source lines of instance {bi} (samples, % of {tot:.0f})")
    for r in body[:top]:
        print(f"   {float(r[cs]):7.0f} {100 * float(r[cs]) / tot:5.1f}%  {r[ci].strip()[:150]}")
