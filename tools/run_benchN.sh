# N-GPU bench line (torchrun, as the driver launches it) with a hard timeout.  usage: run_benchN.sh N [extra bench flags]
N=$1; shift
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 100 --warmup 5 "$@" > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
grep -v "^$\|OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/r02_bench_n$N.err | tail -c 1500
python tools/show_bench.py gpurun_out/r02_bench_n$N.json
