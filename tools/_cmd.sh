timeout 100 python - <<'PY'
import importlib, time, sys, numpy as np
sys.path.insert(0, ".")
cb = importlib.import_module("collider-rs_b200"); scenes = importlib.import_module("collider-rs_b200.scenes")
s = scenes.c3_dense(1_000_000)
eng = cb.Engine(4.0, 0.01); eng.upload_state(**s); eng.set_window_drain(True); eng.set_stage_timing(False)
pin = cb.PinnedArray(4_000_000, cb.EVENT16_DTYPE)
T = 0.0
for k in range(24):
    t0 = time.perf_counter(); r = eng.bulk_update(T, end_time=T + 0.1); t1 = time.perf_counter()
    got = eng.fetch_events_begin(pin.array); t2 = time.perf_counter(); eng.fetch_events_end(); t3 = time.perf_counter()
    T += 0.1
    d = eng.last_drain()
    if k % 3 == 0 or k > 18:
        print(k, "step %.2f ms  begin %.2f  end %.2f" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), "events", r.n_events, "live", d["pairs_now"], "added", d["pairs_added"], "removed", d["pairs_removed"], "launches", eng.launch_count(), "fallbacks", eng.event_sort_fallbacks(), flush=True)
PY
