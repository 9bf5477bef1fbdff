#!/bin/bash
# Runs on the GPU box (under gpurun): ncu evidence for profiles/.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r01}
# (1) launch list of the default bench command (per-launch device time; cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py > gpurun_out/${R}_bench_under_ncu.json 2> gpurun_out/${R}_bench_under_ncu.err
# (2) full capture of the factor kernels at cfg1 (the bench workload; HB200_NO_MERGE=1 keeps the two factor families in
#     separate launches -- the production path runs the same bodies side by side in factor_eval_kernel) and on the 1M-factor window
HB200_NO_MERGE=1 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 8 -c 2 -o gpurun_out/${R}_eval_cfg1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large > /dev/null 2> gpurun_out/${R}_ncu_cfg1.err
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 2 -o gpurun_out/${R}_eval_cfg4 \
    python tools/eval_sweep.py --config 4 --reps 2 > /dev/null 2> gpurun_out/${R}_ncu_cfg4.err
# (3) full capture of the dominant kernel by time at cfg1
ncu --set full --clock-control none --import-source on -k regex:band_solve -s 3 -c 1 -o gpurun_out/${R}_band_cfg1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large > /dev/null 2> gpurun_out/${R}_ncu_band.err
# (3b) the J^T J / Schur kernels of the same workload (next-round targets)
ncu --set full --clock-control none --import-source on -k regex:'inertial_hessian|schur_kernel' -s 4 -c 2 -o gpurun_out/${R}_jtj_cfg1 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large > /dev/null 2> gpurun_out/${R}_ncu_jtj.err
# (4) the real numbers (never taken under a profiler)
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_clocks_idle.csv
lscpu | head -20 > gpurun_out/${R}_lscpu.txt
ls -la gpurun_out | tail -20
