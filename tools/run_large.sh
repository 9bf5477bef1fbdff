# large-window iteration profile under env variants: run_large.sh VAR=val ...
for v in "$@"; do
  echo "== env $v"
  env $v timeout 300 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from hyperslam_b200 import runtime, synthetic
import torch
win = synthetic.make_config(4, constant_knots=2)
ctx = runtime.Context(0); ctx.load_window(win); ctx.snapshot()
for _ in range(3): ctx.restore(); ctx.iterate(1, records=False)
ctx.synchronize()
ext = torch.cuda.ExternalStream(ctx.stream)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0
for _ in range(5):
    ctx.restore(); s.record(ext); ctx.iterate(1, records=False); e.record(ext); ctx.synchronize(); tot += s.elapsed_time(e)
agg = {}
ctx.restore()
for n, ms in ctx.profile_iteration(reps=3): agg[n] = agg.get(n, 0) + ms
print("ms/iter", round(tot / 5, 4), {k: round(v, 4) for k, v in agg.items()})
PY
done
