"""Hottest CUDA source lines of one kernel in an .ncu-rep (warp-stall samples aggregated per source line, with the long-scoreboard /
local-memory share): python tools/ncu_lines.py rep.ncu-rep <kernel regex> [top]"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
func = None
fpath = ""
head = None
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if func is None:
            func = r[1]
        active = (r[1] == func)
        continue
    if r[0] == "Line No":
        head = r
        col = {n: i for i, n in reversed(list(enumerate(head)))}
        cs = col["# Samples"]; ins = col["Instructions Executed"]
        lsb = col.get("stall_long_sb"); lcl = col.get("L2 Theoretical Sectors Local"); ssb = col.get("stall_short_sb"); sw = col.get("stall_wait"); sbar = col.get("stall_barrier")
        cur = None
        continue
    if head is None or not active or len(r) <= cs:
        continue
    if r[0].strip():
        cur = (fpath, r[0], r[1].strip())
    if cur is None:
        continue
    def f(i):
        try:
            return float(r[i])
        except (ValueError, TypeError, IndexError):
            return 0.0
    a = agg.setdefault(cur, [0.0] * 7)
    a[0] += f(cs); a[1] += f(lsb); a[2] += f(lcl); a[3] += f(ins); a[4] += f(ssb); a[5] += f(sw); a[6] += f(sbar)
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[3] for a in agg.values()) or 1
print(f"{func[:110]}: {tot:.0f} samples, {toti:.0f} warp instructions")
print(" samples     %  long_sb short_sb  wait  barrier  local_sectors  instr%  line")
for (fp, ln, s), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]:8.0f} {100 * a[0] / tot:5.1f} {a[1]:8.0f} {a[4]:8.0f} {a[5]:6.0f} {a[6]:7.0f} {a[2]:12.0f}  {100 * a[3] / toti:5.1f}  {fp}:{ln} {s[:110]}")
