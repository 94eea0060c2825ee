#!/bin/bash
# dense kernel v3 after the k_end fix, predicated division, pair Bloom bits
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/c12_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c12_tests.log
run() { # name, args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/c12_$name.json 2> gpurun_out/c12_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c12_$name.json"))
    print("$name", "wall %.1f dev %.1f e2e %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3, d["e2e"]["ms_per_step_median_rank0"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms"].items()}, d["per_step"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c12_$name.err").read()[-1500:])
PY
}
L=collider-rs_b200/lib
ENVV="X=1" run c3_auto
ENVV="CB200_LIB=$L/libcb_twoarms.so" run c3_twoarms
ENVV="X=1" run dense_auto --workload c3_dense
ENVV="CB200_DENSE_RANK=4" run dense_d4 --workload c3_dense
ENVV="CB200_DENSE_RANK=2" run dense_d2 --workload c3_dense
ENVV="CB200_DENSE_RANK=4 CB200_LIB=$L/libcb_m0.so" run dense_d4_m0 --workload c3_dense
ENVV="CB200_DENSE_RANK=4 CB200_LIB=$L/libcb_twoarms.so" run dense_d4_twoarms --workload c3_dense
ENVV="X=1" run c5_auto --workload c5
ENVV="CB200_DENSE_RANK=4" run c5_d4 --workload c5
ENVV="CB200_DENSE_RANK=2" run c5_d2 --workload c5
