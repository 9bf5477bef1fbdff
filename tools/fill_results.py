#!/usr/bin/env python
"""Regenerates the generated block of BASELINE.md section 4 from the committed bench lines (tools/results_table.py)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "results_table.py"), R], capture_output=True, text=True, check=True).stdout
p = os.path.join(ROOT, "BASELINE.md")
s = open(p).read()
a = s.index("<!-- results:begin -->") + len("<!-- results:begin -->")
b = s.index("<!-- results:end -->")
open(p, "w").write(s[:a] + "\n" + table + s[b:])
print(table)
