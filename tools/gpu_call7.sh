#!/bin/bash
mkdir -p gpurun_out
run() { # name, args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/c7_$name.json 2> gpurun_out/c7_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c7_$name.json"))
print("$name", "wall %.1f dev %.1f e2e %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3, d["e2e"]["ms_per_step_median_rank0"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms"].items()})
PY
}
L=collider-rs_b200/lib
ENVV="CB200_SOLVE=1" run base
ENVV="CB200_SOLVE=1 CB200_LIB=$L/libcb_lb3.so CB200_PER_SM_S=3" run lb3
ENVV="CB200_SOLVE=1 CB200_LIB=$L/libcb_lb2.so CB200_PER_SM_S=2" run lb2
ENVV="CB200_SOLVE=1 CB200_LIB=$L/libcb_lb5.so CB200_PER_SM_S=5" run lb5
ENVV="CB200_SOLVE=1 CB200_LIB=$L/libcb_lb6.so CB200_PER_SM_S=6" run lb6
ENVV="CB200_SOLVE=1 CB200_LIB=$L/libcb_lb3.so CB200_PER_SM_S=3" run lb3_dense --workload c3_dense
ENVV="CB200_SOLVE=1 CB200_LIB=$L/libcb_lb6.so CB200_PER_SM_S=6" run lb6_dense --workload c3_dense
