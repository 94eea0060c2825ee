#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_sharded_gpu.py -m gpu -q 2>&1 | tail -1; done
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/c5_$name.json 2> gpurun_out/c5_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/c5_$name.json"))
print("$name", "wall %.1f dev %.1f" % (d["ms_per_step"]*1e3, d["device_ms_per_step"]*1e3), d["e2e"])
PY
}
run base X=1
run s4 CB200_PER_SM_S=4
run base2 X=1
CB200_VERBOSE=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 >/dev/null | grep -c "captured step graph"
