"""Pretty-print the parts of a bench.py JSON line that matter when iterating on kernels."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "gpu_launches")}, "clocks", d.get("clocks"))
print("e2e ms", d["e2e"]["ms_per_step"], "kernel_ms", d["kernel_ms"])
r = d["roofline"]
print("roofline", r["kernel"][:40], "launch_ms", r["launch_ms"], "frac", round(r["frac"], 4), "fp64 TF/s", r.get("fp64_peak_tflops_measured"))
for k, v in r.get("large_windows", {}).items():
    if isinstance(v, dict) and "ms_per_iteration" in v:
        print(k, "ms/iter", round(v["ms_per_iteration"], 4), "graph", v.get("graph"),
              {kk: (round(vv["launch_ms"], 4), round(vv["frac"], 3)) for kk, vv in v.items() if isinstance(vv, dict) and "frac" in vv})
        print("   ", v["iteration_kernel_ms"], v.get("comm"))
for key in ("comm", "parity", "dense_solver_bar"):
    if d.get(key):
        print(key, d[key])
if d["e2e"].get("sliding_window"):
    print("sliding", d["e2e"]["sliding_window"])
if d.get("cpu_baseline"):
    print("cpu", {k: d["cpu_baseline"][k] for k in ("value", "cores", "gn_iters_per_s")})
