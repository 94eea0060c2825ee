#!/bin/bash
# after the dense-buffer reservation fix (sharded direct exchange hung on a mid-run cudaMalloc) and the net solve counter
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu -k dense --timeout 40 --timeout-method thread > gpurun_out/c13_tests_sharded_dense.log 2>&1; echo "sharded dense rc=$?"; tail -2 gpurun_out/c13_tests_sharded_dense.log
timeout 240 python -m pytest tests -x -q -m gpu --timeout 60 --timeout-method thread > gpurun_out/c13_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c13_tests.log
run() { # name, args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 200 python bench.py --steps 30 --warmup 5 "$@" > gpurun_out/c13_$name.json 2> gpurun_out/c13_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c13_$name.json"))
    print("$name", "wall %.1f dev %.1f e2e %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3, d["e2e"]["ms_per_step_median_rank0"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms"].items()}, d["per_step"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c13_$name.err").read()[-1500:])
PY
}
L=collider-rs_b200/lib
ENVV="X=1" run c3
ENVV="CB200_LIB=$L/libcb_oldcount.so" run c3_oldcount --no-cpu-baseline
ENVV="X=1" run dense --workload c3_dense --no-cpu-baseline
ENVV="X=1" run c5 --workload c5 --no-cpu-baseline
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 300 --csv --log-file gpurun_out/launches_r1n_dense.csv \
  python bench.py --steps 5 --warmup 5 --no-cpu-baseline --workload c3_dense > gpurun_out/c13_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
