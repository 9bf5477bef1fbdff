# quick check while iterating: parity subset + cfg1 bench line without the large sections / CPU arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not large and not multi" 2>&1 | tail -4
for v in "$@"; do
  echo "== env $v"
  env $v timeout 300 python bench.py --steps 200 --warmup 5 --no-large --no-cpu-baseline > gpurun_out/quick.json 2> gpurun_out/quick.err; python tools/show_bench.py gpurun_out/quick.json | head -3
done
