# 1-GPU: GPU tests + bench line (default flags).  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench.err
python tools/show_bench.py gpurun_out/r02_bench.json
