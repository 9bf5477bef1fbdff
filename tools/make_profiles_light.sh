#!/bin/bash
# Light refresh (no --set full captures): launch list under ncu + the real bench numbers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --no-cpu-baseline --no-large > gpurun_out/${R}_bench_under_ncu.json 2> gpurun_out/${R}_bench_under_ncu.err
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_clocks_idle.csv
ls -la gpurun_out | tail -8
