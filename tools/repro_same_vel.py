import importlib, sys
sys.path.insert(0, ".")
import numpy as np
cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")
from oracle import oracle_py as oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
scene2 = scenes.uniform_density(n, seed=100)
orc = oracle.Collider(4.0, 0.01)
orc.add_hitboxes(scene2, end_time=cb.INF)
eng = cb.Engine(4.0, 0.01)
eng.upload_state(**orc.dump_state())
eng.set_overlaps(*orc.dump_overlaps())
v = np.full(n, 2.0)
res = eng.bulk_update(0.0, end_time=cb.INF, vel_x=v, vel_y=v)
print("step ok", res.n_events, res.n_solitaire)
ev = eng.fetch_events()
print("fetch ok", len(ev))
