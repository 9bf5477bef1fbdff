mkdir -p gpurun_out
for mode in peer nccl+mailbox; do
  echo "=== $mode" >> gpurun_out/mg.log
  HB200_DEBUG=1 NCCL_DEBUG=WARN timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tests/multi_gpu_worker.py $mode >> gpurun_out/mg.log 2>&1
  echo "rc=$?" >> gpurun_out/mg.log
done
tail -60 gpurun_out/mg.log
