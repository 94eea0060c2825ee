import importlib, sys, time
sys.path.insert(0, ".")
cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")
import bench
scene = bench.make_workload(scenes, "c3", 1_000_000, 0, 1)
eng = cb.Engine(4.0, 0.01)
eng.upload_state(**scene)
eng.set_stage_timing(False)
T = 0.0
for s in range(20):
    eng.reset_timing()
    t0 = time.perf_counter()
    res = eng.bulk_update(T, end_time=T + 0.1)
    wall = (time.perf_counter() - t0) * 1e3
    T += 0.1
    tot, st = eng.step_timing()
    print(s, "wall %.3f dev %.3f" % (wall, tot), "launches", eng.launch_count(), flush=True)
