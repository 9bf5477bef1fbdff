"""LM iterations on a chosen BASELINE config (for ncu captures of the in-iteration kernels at scale)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperslam_b200 import runtime, synthetic
import torch

p = argparse.ArgumentParser()
p.add_argument("--config", type=int, default=4)
p.add_argument("--iters", type=int, default=3)
p.add_argument("--profile", action="store_true", help="print the per-kernel event profile of one iteration")
a = p.parse_args()
win = synthetic.make_config(a.config, constant_knots=2)
ctx = runtime.Context(0)
ctx.load_window(win)
ctx.snapshot()
ext = torch.cuda.ExternalStream(ctx.stream)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(a.iters):
    ctx.restore(); s.record(ext); ctx.iterate(1, records=False); e.record(ext); ctx.synchronize(); tot += s.elapsed_time(e)
print(f"config {a.config}: {win.num_factors} factors, {tot / a.iters:.4f} ms / LM iteration (includes first-launch effects)")
if a.profile:
    agg = {}
    ctx.restore()
    for n, ms in ctx.profile_iteration(reps=3):
        agg[n] = agg.get(n, 0) + ms
    print({k: round(v, 4) for k, v in agg.items()})
