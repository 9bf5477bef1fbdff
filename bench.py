#!/usr/bin/env python
"""bench.py -- factor evals/s & GN-iters/s of the B200 hot path (BASELINE.json metric).

A "step" is one Levenberg-Marquardt / Gauss-Newton iteration of the sliding-window problem: evaluate
every factor (residual + Jacobian) -> loss reweighting -> J^T J / J^T r -> landmark Schur complement
-> dense Cholesky -> back-substitution -> retraction -> cost at the trial point -> accept/reject.
`value` = factors per second with everything resident in HBM; `e2e` = the same through
hb200_optimize() with the variable blocks in pinned HOST memory (H2D + D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1] [--impl reference]
N > 1: launched by torch.distributed.run, one rank per GPU, weak scaling (each rank owns one
cfg-sized factor shard of an N-times larger window; one NCCL all-reduce of the reduced system per
iteration plus a 4-double all-reduce for step acceptance).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hyperslam_b200 import synthetic  # noqa: E402

METRIC = "factor_evals_per_s"
UNIT = "factor evals/s"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=300)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--config", type=int, default=1, help="index into BASELINE.json configs (default 1 = headline)")
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-large", action="store_true", help="skip the 1M-factor roofline section")
    return p.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_global_window(config, world):
    cfg = dict(synthetic.CONFIGS[config])
    cfg["num_landmarks"] *= world
    cfg["num_imu"] *= world
    return synthetic.make_window(seed=synthetic.SEED_BASE + config + 1, constant_knots=2, **cfg)


def algorithmic_bytes(win):
    """Per-launch algorithmic bytes of the two evaluate kernels (DESIGN.md 'HBM layout')."""
    k, kb = win.order, win.bias_order
    K, L, C = win.knots.shape[0], win.landmarks.shape[0], win.cameras.shape[0]
    pix_per = 8 + 16 + 16 + 16 + 2 * 6 * k * 8 + 48            # stamp, pixel, idx | r, Jp, Jl
    imu_per = 8 + 48 + 16 + 48 + 6 * 6 * k * 8 + 2 * kb * 8 + 96  # stamp, meas, idx | r, Jp, wg, wa, Jg
    shared = K * 224
    pix = win.v_stamp.size * pix_per + shared + L * 24 + C * 160
    imu = win.i_stamp.size * imu_per + shared + 48 * 8 + (win.gyro_bias.size + win.accel_bias.size) * 8
    return dict(pixel_eval_kernel=pix, inertial_eval_kernel=imu, pixel_per_factor=pix_per, inertial_per_factor=imu_per)


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, val in zip(names, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


def best_thread_count(ow, fn):
    """The CPU arm gets the thread count that serves it best (hyper-threads usually hurt)."""
    import psutil
    cands = sorted({c for c in (psutil.cpu_count(logical=False), os.cpu_count(), 32, 16, 8) if c and c <= (os.cpu_count() or 1)})
    best, best_t = cands[-1], float("inf")
    for c in cands:
        fn(c)
        t0 = time.perf_counter(); fn(c); fn(c); dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


def run_reference(args, rank, world):
    """CPU arm: the oracle (a restatement -- the reference cannot be built here) on the host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    win = make_global_window(args.config, 1)
    ow = ol.OracleWindow(win)
    nf = win.num_factors
    cores = best_thread_count(ow, lambda c: ow.iterate(apply=False, outputs=False, nthreads=c))
    for _ in range(max(args.warmup, 1)):
        ow.iterate(apply=False, outputs=False, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ow.iterate(apply=False, outputs=False, nthreads=cores)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = nf / dt
    ev_cores = best_thread_count(ow, lambda c: ow.evaluate(want_J=True, outputs=False, nthreads=c))
    t1 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t1 < 2.0:
        ow.evaluate(want_J=True, outputs=False, nthreads=ev_cores); reps += 1
    ev = nf * reps / (time.perf_counter() - t1)
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=synthetic.CONFIG_NAMES[args.config], factors_per_step=nf, note="CPU oracle (restatement of the reference path; Ceres/Eigen/HyperState are not installable here), OpenMP over factors, serial assembly + Schur + dense Cholesky"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=f"{args.steps} full LM iterations of {nf} factors each",
                                  evaluate_only_value=ev, evaluate_only_cores=ev_cores, gn_iters_per_s=1.0 / dt),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gn_iters_per_s=1.0 / dt,
                evaluate_sweep=dict(evals_per_s=ev))
    print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    import torch
    import torch.distributed as dist
    from hyperslam_b200 import runtime

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    gwin = make_global_window(args.config, world)
    win = gwin.shard(rank, world) if world > 1 else gwin
    nf_total = gwin.num_factors
    nf_local = win.num_factors

    ctx = runtime.Context(local_rank, use_graph=(world == 1))
    ctx.load_window(win)
    n_reduced = ctx.reduced_size()   # (the context is closed before the large-window section)
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    tensors = {}

    def allreduce(ptr, count, stream):
        key = (ptr, count)
        if key not in tensors:
            class _Arr:
                __cuda_array_interface__ = dict(shape=(count,), typestr="<f8", data=(ptr, False), version=2)
            tensors[key] = torch.as_tensor(_Arr(), device=torch.device("cuda", local_rank))
        with torch.cuda.stream(ext):
            dist.all_reduce(tensors[key])
        return 0

    if world > 1:
        ctx.set_allreduce(allreduce)

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # 256 MiB > 126 MB L2

    def l2_flush():
        with torch.cuda.stream(ext):
            flush.fill_(1.0)

    ctx.snapshot()

    # ---- device-resident steps ---------------------------------------------------------------
    def timed_steps(step_fn, steps, warmup):
        for _ in range(warmup):
            ctx.restore(); l2_flush(); step_fn()
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        launches0 = ctx.launch_count
        for i in range(steps):
            ctx.restore(); l2_flush()
            starts[i].record(ext)
            step_fn()
            stops[i].record(ext)
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count - launches0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, launches = timed_steps(lambda: ctx.iterate(1, records=False), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = nf_total / (ms_per_step * 1e-3)

    # evaluate-only sweep (prep + pixel + inertial kernels, residual + Jacobian)
    sweep_ms, _ = timed_steps(lambda: ctx.evaluate(jacobians=True), args.steps, 3)
    sweep_ms /= args.steps

    # ---- end to end through hb200_optimize with pinned host buffers ---------------------------
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[...] = a
        return t
    init = dict(knots=win.knots, gyro=win.gyro_bias, accel=win.accel_bias, gravity=win.gravity, landmarks=win.landmarks)
    src = {k_: pinned(v) for k_, v in init.items()}
    work = {k_: pinned(v) for k_, v in init.items()}
    h2d = sum(v.numel() * 8 for v in work.values())
    d2h = h2d + 56

    def e2e_step():
        for k_ in work:
            work[k_].copy_(src[k_])   # host-side reset of the in/out buffers (not device work)
        ctx.restore()                 # trust-region state back to the initial radius
        l2_flush()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.optimize(1, work["knots"].numpy(), work["gyro"].numpy(), work["accel"].numpy(), work["gravity"].numpy(),
                     work["landmarks"].numpy(), records=True)
        return time.perf_counter() - t0

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    if world > 1:
        dist.barrier()
    e2e_t = sum(e2e_step() for _ in range(args.steps)) / args.steps
    tt = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_t = float(tt.item())
    e2e_value = nf_total / e2e_t

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- attribution: per-kernel durations (CUDA events after every launch) -------------------
    kernel_ms, roofline = {}, None
    if world == 1:
        ctx.restore()
        prof = ctx.profile_iteration(reps=10)
        for name, ms in prof:
            kernel_ms[name] = kernel_ms.get(name, 0.0) + ms
        total_k = sum(kernel_ms.values())
        shares = {k_: round(v / total_k, 4) for k_, v in kernel_ms.items()}
        ab = algorithmic_bytes(win)
        peak, peak_src = load_peaks()
        # pixel_eval_kernel appears twice per iteration (Jacobian pass, cost-only pass): the first is the one with J.
        # the factor kernels as hb200_evaluate launches them (residual + materialised Jacobian, no fused J^T J)
        first = {}
        for name, ms in ctx.profile_iteration(reps=10, evaluate_only=True):
            first.setdefault(name, ms)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("pixel_eval_kernel_dram_bytes")
        dominant = max(kernel_ms, key=kernel_ms.get)
        ach = ab["pixel_eval_kernel"] / (first["pixel_eval_kernel"] * 1e-3) / 1e9
        roofline = dict(kernel="pixel_eval_kernel<4,true>" if win.order == 4 else "pixel_eval_kernel<6,true>", bound="hbm", achieved=ach, peak=peak,
                        unit="GB/s", frac=ach / peak, traffic=traffic, peak_source=peak_src,
                        algorithmic_bytes_per_launch=ab["pixel_eval_kernel"], launch_ms=first["pixel_eval_kernel"],
                        inertial_eval_kernel=dict(algorithmic_bytes_per_launch=ab["inertial_eval_kernel"], launch_ms=first.get("inertial_eval_kernel"),
                                                  achieved=(ab["inertial_eval_kernel"] / (first["inertial_eval_kernel"] * 1e-3) / 1e9) if "inertial_eval_kernel" in first else None),
                        dominant_kernel_by_time=dominant, kernel_share_of_step=shares,
                        kernel_ms_source="hb200_profile_iteration(evaluate sweep): CUDA event after every launch", note="cfg1 moves 7.7 MB per sweep (about 1.2 us at peak): launch/latency-bound by construction (SURVEY.md 8d); see profiles/ for larger windows")

    # ---- roofline of the factor kernels on a window large enough to be bandwidth-relevant (1 M factors) ----
    large = None
    if world == 1 and not args.no_large:
        ctx.close()
        big = synthetic.make_config(4, constant_knots=2)
        bctx = runtime.Context(local_rank)
        bctx.load_window(big)
        bext = torch.cuda.ExternalStream(bctx.stream, device=torch.device("cuda", local_rank))
        for _ in range(3):
            bctx.evaluate(jacobians=True)
        bctx.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        ev0.record(bext)
        for _ in range(reps):
            bctx.evaluate(jacobians=True)
        ev1.record(bext)
        bctx.synchronize()
        sweep = ev0.elapsed_time(ev1) / reps
        bprof = {}
        for name, ms in bctx.profile_iteration(reps=5, evaluate_only=True):
            bprof.setdefault(name, ms)
        biter = {}
        for name, ms in bctx.profile_iteration(reps=3):
            biter[name] = biter.get(name, 0.0) + ms
        bab = algorithmic_bytes(big)
        peak, peak_src = load_peaks()
        large = dict(workload=synthetic.CONFIG_NAMES[4], factors=big.num_factors, evaluate_sweep_ms=sweep,
                     evaluate_sweep_evals_per_s=big.num_factors / (sweep * 1e-3), peak=peak, unit="GB/s")
        for kname in ("pixel_eval_kernel", "inertial_eval_kernel"):
            ach = bab[kname] / (bprof[kname] * 1e-3) / 1e9
            large[kname] = dict(algorithmic_bytes_per_launch=bab[kname], launch_ms=bprof[kname], achieved=ach, frac=ach / peak)
        large["iteration_kernel_ms"] = {k_: round(v, 4) for k_, v in biter.items()}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                large["traffic"] = json.load(f).get("large_window")
        bctx.close()

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        ow = ol.OracleWindow(gwin)
        cores = best_thread_count(ow, lambda c: ow.iterate(apply=False, outputs=False, nthreads=c))
        ow.iterate(apply=False, outputs=False, nthreads=cores)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 6.0:
            ow.iterate(apply=False, outputs=False, nthreads=cores); reps += 1
        it_s = (time.perf_counter() - t0) / reps
        ev_cores = best_thread_count(ow, lambda c: ow.evaluate(want_J=True, outputs=False, nthreads=c))
        t0 = time.perf_counter(); r2 = 0
        while time.perf_counter() - t0 < 3.0:
            ow.evaluate(want_J=True, outputs=False, nthreads=ev_cores); r2 += 1
        ev_all = nf_total * r2 / (time.perf_counter() - t0)
        t0 = time.perf_counter(); r3 = 0
        while time.perf_counter() - t0 < 3.0:
            ow.evaluate(want_J=True, outputs=False, nthreads=1); r3 += 1
        ev_one = nf_total * r3 / (time.perf_counter() - t0)
        cpu = dict(value=nf_total / it_s, unit=UNIT, cores=cores, kind="port",
                   sample=f"{reps} full LM iterations + {r2} all-core and {r3} single-thread Evaluate sweeps of {nf_total} factors (oracle restatement; Ceres cannot be built here)",
                   gn_iters_per_s=1.0 / it_s, evaluate_only_all_cores=ev_all, evaluate_only_cores=ev_cores, evaluate_only_one_thread=ev_one)

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=synthetic.CONFIG_NAMES[args.config], factors_per_step=nf_total, factors_per_gpu=nf_local,
                            reduced_system_size=n_reduced, spline_order=win.order, knots=int(win.knots.shape[0]),
                            landmarks=int(gwin.landmarks.shape[0]), parallelism=f"factor-sharded x{world}" if world > 1 else "single GPU",
                            l2="flushed between timed steps (256 MiB device write)", step="one LM iteration (evaluate r+J, JtJ, Schur, Cholesky, retract, cost, accept) in one CUDA graph" if world == 1 else "one LM iteration, direct launches + 2 NCCL all-reduces"),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_t * 1e3,
                         api="hb200_optimize(1 iteration) with pinned host variable blocks"),
                gpu_launches=int(launches), clocks=clocks, gn_iters_per_s=1e3 / ms_per_step,
                evaluate_sweep=dict(ms=sweep_ms, evals_per_s=nf_total / (sweep_ms * 1e-3)), kernel_ms={k_: round(v, 5) for k_, v in kernel_ms.items()})
    if roofline:
        if large:
            roofline["large_window"] = large
        line["roofline"] = roofline
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        ctx.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
