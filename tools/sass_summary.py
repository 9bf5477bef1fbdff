#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libhyperb200.so (cuobjdump -sass, read here without a GPU): tensor-core ops
(DMMA = mma.sync f64; tcgen05 has no FP64 kind), TMA bulk copies (UBLKCP), 256-bit global stores, local-memory
traffic (spills), atomics / reductions, peer-memory system-scope accesses.  Writes profiles/<round>_sass_summary.md."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "hyperslam_b200", "lib", "libhyperb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = dict(re.findall(r"(REG|STACK|SHARED):(\d+)", line))
        cur = None
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("hb::", "").replace("void ", "")
pats = collections.OrderedDict([("DMMA", r"\bDMMA"), ("UBLKCP (TMA 1-D bulk)", r"\bUBLKCP"), ("STG.256", r"\bSTG\.E(\.\w+)*\.256"), ("LDG.128", r"\bLDG\.E(\.\w+)*\.128"),
                                ("LDL", r"\bLDL"), ("STL", r"\bSTL"), ("REDG.ADD.F64", r"\bREDG\.E\.ADD\.F64"),
                                ("ATOMS", r"\bATOMS"), (".SYS accesses", r"\.SYS\b"), ("BAR", r"\bBAR\."), ("MUFU", r"\bMUFU"), ("total", r"^\s+/\*[0-9a-f]{4,}\*/")])
rows, cur, counts = [], None, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts))
        cur, counts = m.group(1), collections.Counter()
        continue
    if cur:
        for k_, p in pats.items():
            if re.search(p, line):
                counts[k_] += 1
if cur:
    rows.append((cur, counts))
arch = re.search(r"arch = (sm_\w+)", out)
md = [f"# {R}: SASS summary of `hyperslam_b200/lib/libhyperb200.so` ({arch.group(1) if arch else '?'})\n",
      "`cuobjdump -sass` mnemonic counts per kernel (static instruction counts, not executed counts) and `cuobjdump -res-usage`.",
      "tcgen05 has no FP64 kind: the tensor-core path of this FP64 problem is `mma.sync.m8n8k4.f64` = `DMMA.8x8x4`. `UBLKCP` = `cp.async.bulk` (1-D TMA).",
      "`.SYS` = system-scope loads / stores / releases on peer-mapped memory (the NVLink mailbox and the fused all-reduce).\n",
      "| kernel | regs | stack B | smem B | " + " | ".join(pats) + " |", "|---|---|---|---|" + "---|" * len(pats)]
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: -r[1]["total"]):
    u = usage.get(name, {})
    md.append(f"| `{demangle(name)[:70]}` | {u.get('REG', '')} | {u.get('STACK', '')} | {u.get('SHARED', '')} | " + " | ".join(str(c[k_]) for k_ in pats) + " |")
    tot.update(c)
md.append("| **all kernels** | | | | " + " | ".join(f"**{tot[k_]}**" for k_ in pats) + " |")
path = os.path.join(ROOT, "profiles", f"{R}_sass_summary.md")
open(path, "w").write("\n".join(md) + "\n")
print(path, dict(tot))
