import os, sys, ctypes as C
os.environ["HB200_BAND_TIMING"]="1"
sys.path.insert(0,"/root/repo")
import numpy as np
from hyperslam_b200 import runtime, synthetic
win = synthetic.make_config(1, constant_knots=2)
ctx = runtime.Context(0); ctx.load_window(win)
ctx.iterate(3)
cyc = (C.c_longlong*8)()
ctx.lib.hb200_debug_band_timing(ctx.h, cyc)
names=["warp0 tile0 (part of update)","potf2 prologue","trsm","update (incl. look-ahead potf2)","corner","backsub_corner","backsub_blocks","look-ahead chol6 (part of update)"]
tot=sum(cyc[1:7])
for n,c in zip(names,cyc): print(f"{n:40s} {c:9d} cycles {c/1.965e3:8.1f} us {100*c/max(tot,1):5.1f}%")
