"""Evaluate-only sweeps on a chosen BASELINE config (for ncu captures of the factor kernels)."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hyperslam_b200 import runtime, synthetic
p = argparse.ArgumentParser(); p.add_argument("--config", type=int, default=4); p.add_argument("--reps", type=int, default=5); p.add_argument("--scale", type=float, default=1.0)
a = p.parse_args()
win = synthetic.make_config(a.config, scale=a.scale, constant_knots=2)
ctx = runtime.Context(0); ctx.load_window(win)
for _ in range(a.reps):
    ctx.evaluate(jacobians=True)
ctx.synchronize()
import torch
ext = torch.cuda.ExternalStream(ctx.stream)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(ext)
for _ in range(a.reps): ctx.evaluate(jacobians=True)
e.record(ext); ctx.synchronize()
print(f"config {a.config}: {win.v_stamp.size} pixel + {win.i_stamp.size} inertial factors, sweep {s.elapsed_time(e)/a.reps:.4f} ms")
