"""Stage A timing with parts switched off (CB200_DBG_A bits: 1 no cell count, 2 no record store, 4 no state write-back).
Results are wrong in those modes; only the stage time is read.  Diagnostic."""
import importlib, sys
sys.path.insert(0, ".")
import numpy as np
cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")
import bench
scene = bench.make_workload(scenes, "c3", 1_000_000, 0, 1)
eng = cb.Engine(4.0, 0.01)
eng.upload_state(**scene)
T = 0.0
for s in range(12):
    if s == 4:
        eng.reset_timing()
    try:
        eng.bulk_update(T, end_time=T + 0.1)
    except Exception as e:
        pass
    T += 0.1
tot, st = eng.step_timing()
print("stage_a %.1f us  solve %.1f us  (step %.1f us)" % (st["stage_a"] * 1e3, st["solve"] * 1e3, tot * 1e3))
