# 1-GPU: GPU tests + bench line (default flags) + reference arm.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
    print('e2e',d['e2e']['ms_per_step'],'kernel_ms',d['kernel_ms'])
    r=d['roofline']; print('roofline',r['kernel'][:40],r['launch_ms'],r['frac'],'fp64',r.get('fp64_peak_tflops_measured'))
    for k,v in r.get('large_windows',{}).items():
        if isinstance(v,dict) and 'ms_per_iteration' in v: print(k,v['ms_per_iteration'],{kk:(vv['launch_ms'],round(vv['frac'],3)) for kk,vv in v.items() if isinstance(vv,dict) and 'frac' in vv}, v['iteration_kernel_ms'])
    print('dense',d.get('dense_solver_bar')); print('cpu',d.get('cpu_baseline'))
except Exception as e: print('parse failed',e)
PY
