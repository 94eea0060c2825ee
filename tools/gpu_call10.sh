#!/bin/bash
# ncu: the two-phase dense kernel at config 4's density (full set, source), and the launch list of config 3 with the (empty) dense launch
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_solve' --launch-skip 50 -c 2 \
  -f -o gpurun_out/prof_r1m_dense python bench.py --steps 5 --warmup 5 --no-cpu-baseline --workload c3_dense > gpurun_out/c10_ncu_dense.log 2>&1
echo "ncu dense rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 300 --csv --log-file gpurun_out/launches_r1m.csv \
  python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/c10_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out | grep r1m
