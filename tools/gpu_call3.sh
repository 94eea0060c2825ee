#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/c3_bench_plain.json 2> gpurun_out/c3_bench_plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_stage_a|k_enum_pairs|k_scatter|k_scan_slots|k_step_zero' --launch-skip 60 -c 6 \
  -f -o gpurun_out/prof_r1k python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/c3_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1k.csv \
  python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/c3_ncu_launches.log 2>&1
ls -la gpurun_out | tail -8
