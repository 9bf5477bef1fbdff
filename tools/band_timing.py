import os, sys, ctypes as C
os.environ["HB200_BAND_TIMING"]="1"
sys.path.insert(0,"/root/repo")
import numpy as np
from hyperslam_b200 import runtime, synthetic
win = synthetic.make_config(1, constant_knots=2)
ctx = runtime.Context(0); ctx.load_window(win)
ctx.iterate(3)
cyc = (C.c_longlong*72)()
ctx.lib.hb200_debug_band_timing(ctx.h, cyc)
names=["potf2 prologue","two-sided: panel phases","merge + separator columns","deferred corner update","corner Cholesky","corner back substitution","arrow pre-pass + outward back substitution","two-sided: update/look-ahead phases"]
tot=sum(cyc[0:8])
for n,c in zip(names,cyc): print(f"{n:40s} {c:9d} cycles {c/1.965e3:8.1f} us {100*c/max(tot,1):5.1f}%")

print("raw stamps, steps 4..11 of chain 0 (cycles relative to the step's panel start):")
print("step  panel_end  upd_start  upd_end  la_start  la_end   next_panel_start")
for i in range(8):
    r = cyc[8+8*i:8+8*i+6]
    nxt = cyc[8+8*(i+1)] if i < 7 else 0
    b = r[0]
    print(f"{i+4:4d} {r[1]-b:9d} {r[2]-b:9d} {r[3]-b:8d} {r[4]-b:9d} {r[5]-b:8d} {(nxt-b) if nxt else 0:9d}")
