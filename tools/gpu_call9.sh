#!/bin/bash
# two-phase dense kernel (filter + solve): parity tests, then A/B on config 3 at config 4's density, config 5, config 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c9_tests.log
run() { # name, args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/c9_$name.json 2> gpurun_out/c9_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c9_$name.json"))
    print("$name", "wall %.1f dev %.1f e2e %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3, d["e2e"]["ms_per_step_median_rank0"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms"].items()}, d["per_step"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c9_$name.err").read()[-1500:])
PY
}
L=collider-rs_b200/lib
ENVV="X=1" run dense_d8 --workload c3_dense
ENVV="CB200_DENSE_RANK=4" run dense_d4 --workload c3_dense
ENVV="CB200_DENSE_RANK=2" run dense_d2 --workload c3_dense
ENVV="CB200_DENSE_RANK=1" run dense_d1 --workload c3_dense
ENVV="CB200_DENSE_RANK=16" run dense_d16 --workload c3_dense
ENVV="CB200_LIB=$L/libcb_m0.so" run dense_d8_m0 --workload c3_dense
ENVV="CB200_LIB=$L/libcb_l3.so CB200_PER_SM_D=3" run dense_d8_l3 --workload c3_dense
ENVV="CB200_LIB=$L/libcb_l6.so CB200_PER_SM_D=6" run dense_d8_l6 --workload c3_dense
ENVV="X=1" run c5_d8 --workload c5
ENVV="CB200_DENSE_RANK=2" run c5_d2 --workload c5
ENVV="X=1" run c3_d8
ENVV="CB200_DENSE_RANK=2" run c3_d2
