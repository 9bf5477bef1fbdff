#!/bin/bash
# Runs on the GPU box (under gpurun): ncu evidence for profiles/.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r02}
BENCH1="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large"
# (1) launch list of the default bench command (per-launch device time; cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py > gpurun_out/${R}_bench_under_ncu.json 2> gpurun_out/${R}_bench_under_ncu.err
# (2) full capture of the kernels of the timed graph at cfg1 (the bench workload), as they run in the graph: the merged factor
#     kernel with the fused pixel J^T J, the inertial J^T J, the Schur complement, the band solver, the trial sweep, the accept
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'factor_eval|inertial_hessian|schur_kernel|band_solve|accept_kernel' \
    -s 18 -c 6 -f -o gpurun_out/${R}_iter_cfg1 $BENCH1 > /dev/null 2> gpurun_out/${R}_ncu_cfg1.err
# (3) the same iteration on the 1M-factor window (BASELINE config 5 on one GPU): factor kernels, J^T J / Schur, cyclic-reduction solver
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pixel_eval|inertial_eval|inertial_hessian|schur|bcr_solve|lm_backsub' \
    -s 16 -c 8 -f -o gpurun_out/${R}_iter_cfg4 python tools/large_iter.py --config 4 --iters 3 > /dev/null 2> gpurun_out/${R}_ncu_cfg4.err
# (4) the real numbers (never taken under a profiler)
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_clocks_idle.csv
lscpu | head -20 > gpurun_out/${R}_lscpu.txt
ls -la gpurun_out | tail -12
