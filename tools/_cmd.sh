mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bulk_gpu.py -x -q -m gpu --timeout 300 --timeout-method thread -k "window_drain or pipelined or tile_path or parity_uniform or chain_with or dense_cells" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2l_tests.log
NOX="--no-cpu-baseline --no-collider-rows --no-parity --no-config4 --repeats 1"
for v in "CB200_TILE=1 D=--drain" "CB200_TILE=0 D=--drain" "CB200_TILE=1 D="; do
  name=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --steps 30 --warmup 5 $NOX $(echo $v | sed 's/.*D=//') > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err || { echo FAILED $v; tail -c 1200 gpurun_out/r2l_$name.err; }
  echo "$v"; python profiles/bench_summary.py gpurun_out/r2l_$name.json | cut -c1-420
done
CB200_TILE=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tile|k_drain' --launch-skip 60 -c 3 \
    -f -o gpurun_out/prof_r2l python bench.py --steps 5 --warmup 5 $NOX --drain > gpurun_out/r2l_ncu_full.log 2>&1; echo "ncu full rc=$?"
