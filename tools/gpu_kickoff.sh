#!/bin/bash
# One gpurun call that re-establishes the whole picture on a fresh box:
#   gpurun --timeout 1500 -- 'bash tools/gpu_kickoff.sh rNN [tests|bench|ncu|all]'
# GPU parity tests (full-size and multi-process ones included), the single-GPU bench line, the ncu launch list and one full
# capture of the step kernels.
tag=${1:-rXX}
what=${2:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,count --format=csv,noheader | head -1
if [[ $what == all || $what == tests ]]; then
  timeout 1200 python -m pytest tests -x -q -m gpu --timeout 900 --timeout-method thread --durations=15 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/${tag}_tests.log
fi
run() { # name, args...
  name=$1; shift
  timeout 600 python bench.py --steps 50 --warmup 5 "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err || { echo "$name FAILED"; tail -c 2500 gpurun_out/${tag}_$name.err; return; }
  python profiles/bench_summary.py gpurun_out/${tag}_$name.json
}
if [[ $what == all || $what == bench ]]; then
  run c3
  run dense --workload c3_dense --no-cpu-baseline --no-collider-rows --no-parity
  run c5 --workload c5 --no-cpu-baseline --no-collider-rows --no-parity
fi
if [[ $what == all || $what == ncu ]]; then
  NOX="--no-cpu-baseline --no-collider-rows --no-parity --no-config4 --repeats 1"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 300 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 5 --warmup 5 $NOX > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_stage_a|k_enum_pairs|k_scatter|k_scan_slots|k_step_zero|k_tile' --launch-skip 60 -c 8 \
    -f -o gpurun_out/prof_${tag} python bench.py --steps 5 --warmup 5 $NOX > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
# then, here:  python profiles/summarize.py launches gpurun_out/launches_${tag}.csv > profiles/${tag}_launches.txt
#              python profiles/summarize.py kernels gpurun_out/prof_${tag}.ncu-rep > profiles/${tag}_step_kernels.txt
