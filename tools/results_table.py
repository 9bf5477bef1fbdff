#!/usr/bin/env python
"""Markdown results table from the committed bench lines under profiles/."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"


def load(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def M(x):
    return f"{x / 1e6:.1f} M"


b1 = load(f"{R}_bench.json")
ref = load(f"{R}_bench_reference.json")
rows = []
hdr = "| Config (BASELINE.json index) | GPUs | ms / LM iteration | factor evals/s | e2e (host buffers) ms | dominant kernel | factor kernel: ms, fraction of HBM peak | source |"
sep = "|---|---|---|---|---|---|---|---|"
if b1:
    rf = b1["roofline"]
    rows.append(f"| 1: order 4, 50 knots, 10k pixel + 2k IMU | 1 | {b1['ms_per_step']:.4f} | {M(b1['value'])} | {b1['e2e']['ms_per_step']:.4f} ({M(b1['e2e']['value'])}) | "
                f"{rf['dominant_kernel_by_time']} {b1['kernel_ms'][rf['dominant_kernel_by_time']]*1e3:.1f} us | factor_eval_kernel (J, fused J^T J) {rf['launch_ms']*1e3:.1f} us, {rf['frac']:.3f} | `{R}_bench.json` |")
    for key in ("cfg2", "cfg4"):
        lw = rf.get("large_windows", {}).get(key)
        if not lw:
            continue
        km = lw.get("iteration_kernel_ms", {})
        dom = max(km, key=km.get) if km else "-"
        frs = ", ".join(f"{k} {lw[k]['launch_ms']:.3f} ms, {lw[k]['frac']:.3f}" for k in ("pixel_eval_kernel", "inertial_eval_kernel") if isinstance(lw.get(k), dict))
        idx = {"cfg2": "2: order 6, 200 knots, 100k pixel + 20k IMU", "cfg4": "4: 1 M factors (833k + 167k), 500 knots"}[key]
        rows.append(f"| {idx} | 1 | {lw['ms_per_iteration']:.4f} | {M(lw['factors'] / (lw['ms_per_iteration'] * 1e-3))} | - | {dom} {km.get(dom, 0):.3f} ms | {frs} | `{R}_bench.json` `roofline.large_windows.{key}` |")
for n in (2, 4, 8):
    bn = load(f"{R}_bench_n{n}.json")
    if not bn:
        continue
    cm = bn.get("comm", {})
    rows.append(f"| 1 per rank (weak scaling, {bn['config']['factors_per_step']} factors) | {n} | {bn['ms_per_step']:.4f} | {M(bn['value'])} | {bn['e2e']['ms_per_step']:.4f} ({M(bn['e2e']['value'])}) | "
                f"reduction: {cm.get('mode')} {cm.get('comm_ms', 0)*1e3:.1f} us, {cm.get('allreduce_bytes')} B | - | `{R}_bench_n{n}.json` |")
    for key, lw in bn["roofline"].get("large_windows", {}).items():
        if not isinstance(lw, dict) or "ms_per_iteration" not in lw:
            continue
        c2 = lw.get("comm", {}) or {}
        rows.append(f"| {lw.get('workload', key)} sharded | {n} | {lw['ms_per_iteration']:.4f} | {M(lw['factors'] / (lw['ms_per_iteration'] * 1e-3))} | - | reduction {c2.get('comm_ms', 0)*1e3:.1f} us, {c2.get('allreduce_bytes')} B, {c2.get('nvlink_busbw_gbs', 0):.0f} GB/s bus | - | `{R}_bench_n{n}.json` |")
print(hdr); print(sep); print("\n".join(rows))
if b1:
    cb = b1["cpu_baseline"]
    print(f"\nCPU restatement on the same box ({cb['cores']} threads for the full iteration): {cb['value']/1e6:.2f} M factor evals/s = {cb['gn_iters_per_s']:.1f} GN-iters/s; "
          f"Evaluate-only {cb['evaluate_only_all_cores']/1e6:.2f} M evals/s on {cb['evaluate_only_cores']} threads, {cb['evaluate_only_one_thread']/1e6:.2f} M/s on one.")
    print(f"GPU Evaluate sweep alone: {b1['evaluate_sweep']['ms']*1e3:.1f} us = {b1['evaluate_sweep']['evals_per_s']/1e6:.0f} M evals/s; "
          f"sliding-window frame loop {b1['e2e']['sliding_window']['ms_per_frame_median']:.2f} ms / frame (median); FP64 FMA peak measured {rf['fp64_peak_tflops_measured']:.1f} TFLOP/s; "
          f"cuSOLVER potrf+potrs on the dense n = {b1['dense_solver_bar']['n']} system {b1['dense_solver_bar']['cusolver_potrf_potrs_ms']:.3f} ms vs band_solve_kernel {b1['dense_solver_bar']['band_solve_kernel_ms']:.4f} ms.")
if ref:
    print(f"Reference arm (`--impl reference`, CPU restatement, {ref['cpu_baseline']['cores']} threads): {ref['value']/1e6:.2f} M factor evals/s.")
