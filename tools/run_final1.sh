# final single-GPU evidence: the whole -m gpu suite, the ncu captures + launch list, the default bench line, the reference arm, smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
bash tools/make_profiles.sh r02 > gpurun_out/r02_make_profiles.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
tail -3 gpurun_out/r02_bench.err
ls -la gpurun_out | grep r02_ | tail -14
