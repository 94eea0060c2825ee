"""Per-step stage timings for one workload (diagnostic): python profiles/step_trace.py c3_dense 12"""
import importlib, sys, time, math
sys.path.insert(0, ".")
import numpy as np
cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")
import bench
w, steps = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
scene = bench.make_workload(scenes, w, n, 0, 1)
eng = cb.Engine(4.0, 0.01)
eng.upload_state(**scene)
res = eng.bulk_update(0.0, end_time=0.1)
ev = eng.fetch_events()
t = ev[(ev["kind"] == 1) & (ev["time"] == 0.0)]
eng.set_overlaps(t["a"], t["b"])
T = 0.1
for s in range(steps):
    eng.reset_timing()
    t0 = time.perf_counter()
    res = eng.bulk_update(T, end_time=T + 0.1)
    wall = (time.perf_counter() - t0) * 1e3
    T = T + 0.1
    tot, st = eng.step_timing()
    print(s, "wall %.3f dev %.3f" % (wall, tot), {k: round(v, 3) for k, v in st.items()}, "cand", res.n_collide_solves, "sep", res.n_separate_solves, "ev", res.n_events)
