mkdir -p gpurun_out
export CB200_TILE=1
CB200_VERBOSE=1 timeout 900 python -m pytest tests/test_bulk_gpu.py -x -q -m gpu --timeout 300 --timeout-method thread -k "tile_path or parity_uniform or narrowphase or resizing" > gpurun_out/r2g_tests.log 2>&1; echo "bulk tests rc=$?"; tail -4 gpurun_out/r2g_tests.log; grep "capacity exceeded" gpurun_out/r2g_tests.log | sort | uniq -c
NOX="--no-cpu-baseline --no-collider-rows --no-parity --no-config4 --repeats 1"
for v in "X=1" "CB200_TILE_TARGET=96 CB200_TILE_CAP=160" "CB200_TILE_TARGET=24 CB200_TILE_CAP=64" "CB200_TILE=0"; do
  name=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 30 --warmup 5 $NOX > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err || { echo FAILED $v; tail -c 800 gpurun_out/r2g_$name.err; }
  echo "$v"; python profiles/bench_summary.py gpurun_out/r2g_$name.json | cut -c1-300
done
