#!/bin/bash
# One gpurun call that re-establishes the whole picture on a fresh box (about 3.5 GPU-minutes):
#   gpurun --timeout 600 -- 'bash tools/gpu_kickoff.sh rNN'
# GPU parity tests, the three single-GPU workloads, the ncu launch list and one full capture of config 3 and of the dense scene.
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu --timeout 60 --timeout-method thread > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
run() { # name, args...
  name=$1; shift
  timeout 200 python bench.py --steps 50 --warmup 5 "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err || { echo "$name FAILED"; tail -c 1500 gpurun_out/${tag}_$name.err; return; }
  python profiles/bench_summary.py gpurun_out/${tag}_$name.json
}
run c3
run dense --workload c3_dense --no-cpu-baseline
run c5 --workload c5 --no-cpu-baseline
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 300 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_stage_a|k_enum_pairs|k_scatter|k_scan_slots|k_step_zero' --launch-skip 60 -c 6 \
  -f -o gpurun_out/prof_${tag} python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_solve' --launch-skip 50 -c 2 \
  -f -o gpurun_out/prof_${tag}_dense python bench.py --steps 5 --warmup 5 --no-cpu-baseline --workload c3_dense > gpurun_out/${tag}_ncu_dense.log 2>&1; echo "ncu dense rc=$?"
# then, here:  python profiles/summarize.py launches gpurun_out/launches_${tag}.csv > profiles/${tag}_launches.txt
#              python profiles/summarize.py kernels gpurun_out/prof_${tag}.ncu-rep > profiles/${tag}_step_kernels.txt
#              python profiles/hotlines.py gpurun_out/prof_${tag}_dense.ncu-rep k_solve_dense 30
