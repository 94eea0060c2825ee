"""Where the end-to-end step time goes (diagnostic): python profiles/e2e_breakdown.py [workload] [n]"""
import importlib, sys, time
sys.path.insert(0, ".")
import numpy as np
cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
scene = bench.make_workload(scenes, w, n, 0, 1)
eng = cb.Engine(4.0, 0.01)
eng.upload_state(**scene)
res = eng.bulk_update(0.0, end_time=0.1)
ev = eng.fetch_events()
t = ev[(ev["kind"] == 1) & (ev["time"] == 0.0)]
eng.set_overlaps(t["a"], t["b"])
pvx, pvy = cb.PinnedArray(n, np.float64), cb.PinnedArray(n, np.float64)
pvx.array[:] = scene["vel_x"]; pvy.array[:] = scene["vel_y"]
T = 0.1
acc = {}
def tick(name, t0):
    acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
for s in range(24):
    t0 = time.perf_counter(); res = eng.bulk_update(T, end_time=T + 0.1); tick("step_device_resident", t0)
    T += 0.1
    t0 = time.perf_counter(); res = eng.bulk_update(T, end_time=T + 0.1, vel_x=pvx.array, vel_y=pvy.array); tick("step_with_h2d_vel", t0)
    T += 0.1
    t0 = time.perf_counter(); e1 = eng.fetch_events(pinned=True); tick("fetch_sorted_pinned(first: sort+d2h)", t0)
    t0 = time.perf_counter(); e2 = eng.fetch_events(pinned=True); tick("fetch_again_pinned(d2h only)", t0)
    t0 = time.perf_counter(); e3 = eng.fetch_events(); tick("fetch_again_pageable", t0)
for k, v in acc.items():
    v = v[4:]
    print(f"{k:45s} median {np.median(v):8.3f} ms   min {min(v):8.3f}")
print("events", len(e1), "resorts", eng.resort_count())
