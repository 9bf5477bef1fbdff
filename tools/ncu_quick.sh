#!/bin/bash
# Quick ncu --set full captures while iterating on kernels (1 GPU).  usage: ncu_quick.sh <tag> <kernel regex> [skip] [count] [-- cmd...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; RE=$2; SKIP=${3:-8}; CNT=${4:-2}
shift 4 2>/dev/null
CMD=${@:-python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o gpurun_out/$TAG $CMD > /dev/null 2> gpurun_out/$TAG.err
ls -la gpurun_out/$TAG.ncu-rep
