import importlib, sys
sys.path.insert(0, ".")
import numpy as np
cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")
from oracle import oracle_py as oracle
n = 20000
def seeded(scene):
    orc = oracle.Collider(4.0, 0.01)
    orc.add_hitboxes(scene, end_time=cb.INF)
    eng = cb.Engine(4.0, 0.01)
    eng.upload_state(**orc.dump_state())
    eng.set_overlaps(*orc.dump_overlaps())
    return orc, eng
scene = scenes.uniform_density(n, seed=99)
half = np.arange(n) % 2 == 0
vx = np.where(half, 2.0, -1.0); vy = np.where(half, 2.0, 0.5)
orc, eng = seeded(scene)
res = eng.bulk_update(0.0, end_time=cb.INF, vel_x=vx, vel_y=vy)
print("step1 ok", res.n_events, flush=True)
if "nofetch" not in sys.argv:
    ev = eng.fetch_events()
    print("fetch1 ok", len(ev), eng.event_sort_fallbacks(), flush=True)
if "del" in sys.argv:
    del eng
    import gc; gc.collect()
scene2 = scenes.uniform_density(n, seed=100)
orc2, eng2 = seeded(scene2)
print("seeded2 ok", flush=True)
v = np.full(n, 2.0)
res = eng2.bulk_update(0.0, end_time=cb.INF, vel_x=v, vel_y=v)
print("step2 ok", res.n_events, res.n_solitaire, flush=True)
ev = eng2.fetch_events()
print("fetch2 ok", len(ev), flush=True)
