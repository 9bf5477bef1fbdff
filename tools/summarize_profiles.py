#!/usr/bin/env python
"""Turns the ncu captures under gpurun_out/ into the committed summaries under profiles/."""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = {"Kernel Name": row[hdr.index("Kernel Name")]}
        for m in METRICS:
            if m in hdr:
                d[m] = (row[hdr.index(m)], units[hdr.index(m)])
        res.append(d)
    return res


def to_bytes(val, unit):
    v = float(val)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


traffic = {}
md = [f"# {R}: ncu --set full captures (clock-control none)\n"]
for tag, title in (("iter_cfg1", "the kernels of one LM iteration of the timed graph on BASELINE config 1 (bench.py workload: 10k pixel + 2k IMU factors)"),
                   ("iter_cfg4", "the kernels of one LM iteration on the 1M-factor window (833k pixel + 167k IMU, one GPU)"),
                   ("eval_cfg1", "factor kernels on BASELINE config 1 (bench.py workload: 10k pixel + 2k IMU factors)"),
                   ("eval_cfg4", "factor kernels on the 1M-factor window (833k pixel + 167k IMU)"),
                   ("band_cfg1", "band_solve_kernel on config 1 (dominant kernel of the step by time)"),
                   ("jtj_cfg1", "inertial J^T J and landmark Schur complement on config 1")):
    rep = os.path.join(G, f"{R}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    md.append(f"\n## {title}\n\nsource: `gpurun_out/{R}_{tag}.ncu-rep`\n")
    for d in raw(rep):
        name = d["Kernel Name"]
        md.append(f"\n### `{name}`\n\n| metric | value |\n|---|---|")
        for m in METRICS:
            if m in d:
                md.append(f"| {m} | {d[m][0]} {d[m][1]} |")
        rd = to_bytes(*d["dram__bytes_read.sum"]); wr = to_bytes(*d["dram__bytes_write.sum"])
        md.append(f"| **DRAM traffic (read + write)** | {rd + wr:.0f} B |")
        key = ("cfg1" if tag.endswith("cfg1") else "large_window")
        short = name.split("(")[0].replace("void ", "").replace("hb::", "")
        traffic.setdefault(key, {})[short] = dict(dram_bytes=rd + wr, dram_read=rd, dram_write=wr, duration_us=float(d["gpu__time_duration.sum"][0]))
with open(os.path.join(P, f"{R}_ncu_full.md"), "w") as f:
    f.write("\n".join(md) + "\n")

# launch list
lst = os.path.join(G, f"{R}_launches_bench.csv")
if os.path.exists(lst):
    lines = [l for l in open(lst) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        agg.setdefault(k, []).append(float(row["Metric Value"]) / 1e3)
    ours = {k: v for k, v in agg.items() if "hb::" in k or "hb200" in k}
    # one LM iteration = one launch of each of these (two for the eval kernels: <.,1> with Jacobians, <.,0> cost only)
    per_iter = {k: sum(v) / len(v) for k, v in ours.items() if not any(s in k for s in ("bind_", "calib_", "spin", "interpolate"))}
    tot = sum(per_iter.values())
    out = [f"# {R}: ncu launch list of `python bench.py` (first 600 launches; --metrics gpu__time_duration.sum --clock-control none)\n",
           "Per-launch device time under ncu is cold-cache and serialised: compare SHARES, not absolutes.\n",
           "| kernel | launches seen | mean us | share of one LM iteration |", "|---|---|---|---|"]
    for k, v in agg.items():
        share = f"{100 * per_iter[k] / tot:.1f} %" if k in per_iter else "-"
        out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {share} |")
    with open(os.path.join(P, f"{R}_launches_bench.md"), "w") as f:
        f.write("\n".join(out) + "\n")

tr = {}
if "cfg1" in traffic:
    # the roofline kernel of the bench line: the merged factor kernel WITH Jacobians (template <K, KB, WANT_J = 1, FUSE>)
    pe = [v for k, v in traffic["cfg1"].items() if k.startswith("factor_eval_kernel<4, 4, 1") or k.startswith("factor_eval_kernel<4, 4, true")]
    if pe:
        tr["factor_eval_kernel_dram_bytes"] = pe[0]["dram_bytes"]
    pe = [v for k, v in traffic["cfg1"].items() if k.startswith("pixel_eval_kernel<4, 1")]
    if pe:
        tr["pixel_eval_kernel_dram_bytes"] = pe[0]["dram_bytes"]
tr["cfg1"] = traffic.get("cfg1")
tr["large_window"] = traffic.get("large_window")
with open(os.path.join(P, "traffic.json"), "w") as f:
    json.dump(tr, f, indent=1)
for name in (f"{R}_bench.json", f"{R}_bench_reference.json", f"{R}_lscpu.txt", f"{R}_clocks_idle.csv"):
    src = os.path.join(G, name)
    if os.path.exists(src):
        with open(src) as a, open(os.path.join(P, name), "w") as b:
            b.write(a.read())
print(open(os.path.join(P, f"{R}_launches_bench.md")).read()[:3000])
print(json.dumps(tr, indent=1)[:1500])
