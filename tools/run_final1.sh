# final single-GPU evidence: the whole -m gpu suite, the default bench line, the reference arm, smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
tail -3 gpurun_out/r02_bench.err
