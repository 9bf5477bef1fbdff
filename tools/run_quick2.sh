mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not large and not multi and not sliding and not termination and not calibration" 2>&1 | tail -3
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -3 gpurun_out/quick.err; python tools/show_bench.py gpurun_out/quick.json 2>/dev/null
