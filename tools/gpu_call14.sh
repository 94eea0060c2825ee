#!/bin/bash
# k_solve<WIDE, NET>: the sparse path back on the plain count; final numbers + ncu evidence of this round
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu --timeout 60 --timeout-method thread > gpurun_out/c14_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c14_tests.log
run() { # name, args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 200 python bench.py --steps 50 --warmup 5 "$@" > gpurun_out/c14_$name.json 2> gpurun_out/c14_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c14_$name.json"))
    print("$name", "wall %.1f dev %.1f e2e %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3, d["e2e"]["ms_per_step_median_rank0"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms"].items()}, d["per_step"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c14_$name.err").read()[-1500:])
PY
}
ENVV="X=1" run c3
ENVV="X=1" run dense --workload c3_dense --no-cpu-baseline
ENVV="X=1" run c5 --workload c5 --no-cpu-baseline
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 300 --csv --log-file gpurun_out/launches_r1n.csv \
  python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/c14_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_solve|k_stage_a|k_enum_pairs|k_scatter|k_scan_slots|k_step_zero' --launch-skip 60 -c 6 \
  -f -o gpurun_out/prof_r1n python bench.py --steps 5 --warmup 5 --no-cpu-baseline > gpurun_out/c14_ncu_full.log 2>&1
echo "ncu full rc=$?"
