mkdir -p gpurun_out
bash tools/gpu_kickoff.sh r2p bench
bash tools/gpu_kickoff.sh r2p ncu
NOX="--no-cpu-baseline --no-collider-rows --no-parity --no-config4 --no-steady-state --repeats 1"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:'k_solve' --launch-skip 40 -c 4 --csv --log-file gpurun_out/r2p_dense_fp64.csv \
  python bench.py --steps 5 --warmup 5 $NOX --workload c3_dense > gpurun_out/r2p_dense_fp64.log 2>&1; echo "ncu fp64 rc=$?"
./tools/fp64_peak > gpurun_out/r2p_fp64_peak.json; cat gpurun_out/r2p_fp64_peak.json
timeout 300 python -m pytest tests/test_collider_gpu.py -x -q -m gpu -k "batched_queries" 2>&1 | tail -3
