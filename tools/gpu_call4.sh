#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c4_tests.log 2>&1
tail -3 gpurun_out/c4_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/c4_$name.json 2> gpurun_out/c4_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c4_$name.json"))
print("$name", "wall %.1f dev %.1f e2e %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3), {k:round(v*1e3,1) for k,v in d["stage_ms"].items()})
PY
}
run base X=1
run s4 CB200_PER_SM_S=4
run s4a4 CB200_PER_SM_S=4 CB200_PER_SM_A=4
run s4a4p5 CB200_PER_SM_S=4 CB200_PER_SM_A=4 CB200_PER_SM_P=5
run s12 CB200_PER_SM_S=12
run s16a16 CB200_PER_SM_S=16 CB200_PER_SM_A=16 CB200_PER_SM_P=16
timeout 200 python profiles/e2e_breakdown.py > gpurun_out/c4_e2e_breakdown.txt 2>&1
cat gpurun_out/c4_e2e_breakdown.txt
