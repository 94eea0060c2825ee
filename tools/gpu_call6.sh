#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c6_tests.log 2>&1
tail -3 gpurun_out/c6_tests.log
run() { # name, args..., env via ENVV
  name=$1; shift
  env $ENVV timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/c6_$name.json 2> gpurun_out/c6_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c6_$name.json"))
print("$name", "wall %.1f dev %.1f e2e %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3, d["e2e"]["ms_per_step_median_rank0"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms"].items()})
PY
}
ENVV="CB200_SOLVE=2" run two
ENVV="CB200_SOLVE=1" run one
ENVV="CB200_SOLVE=2 CB200_PER_SM_S=8" run two_s8
ENVV="CB200_SOLVE=2" run two_c5 --workload c5
ENVV="CB200_SOLVE=1" run one_c5 --workload c5
ENVV="CB200_SOLVE=2" run two_dense --workload c3_dense
ENVV="CB200_SOLVE=1" run one_dense --workload c3_dense
