"""Config 2 through the Collider API only (diagnostic timing): n hitboxes, bulk step every 0.1, every event processed
with next() in between, velocity halving on Collide (two set_hitbox_vel per Collide).
   python profiles/collider_flow.py [n] [steps]"""
import importlib, sys, time
sys.path.insert(0, ".")
import numpy as np
col = importlib.import_module("collider-rs_b200.collider")
scenes = importlib.import_module("collider-rs_b200.scenes")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
s = scenes.uniform_density(n, seed=scenes.SEED_BASE + 2)
c = col.Collider(scenes.CELL_WIDTH, scenes.PADDING)
t0 = time.perf_counter()
a, b = c.add_hitboxes(s, end_time=0.1)
print(f"add_hitboxes({n}): {(time.perf_counter() - t0) * 1e3:.1f} ms, {len(a)} immediate overlaps")
T = 0.0
ev_total = vel_total = 0
t_bulk = t_events = 0.0
for step in range(steps):
    T_next = T + 0.1
    t0 = time.perf_counter()
    if step:
        c.bulk_update(end_time=T_next)
    t_bulk += time.perf_counter() - t0
    t0 = time.perf_counter()
    while c.time() < T_next:
        c.set_time(min(c.next_time(), T_next))
        ev = c.next()
        while ev is not None:
            ev_total += 1
            if ev[0] == col.COLLIDE:
                for i in (ev[1], ev[2]):
                    v = c.get_hitbox(i).vel
                    v.vx *= 0.5
                    v.vy *= 0.5
                    c.set_hitbox_vel(i, v)
                    vel_total += 1
            ev = c.next()
    t_events += time.perf_counter() - t0
    T = T_next
print(f"{steps} steps: bulk {t_bulk / max(steps - 1, 1) * 1e3:.3f} ms/step; events {ev_total} ({ev_total / t_events:.0f}/s incl. {vel_total} set_hitbox_vel + get_hitbox), "
      f"{t_events / steps * 1e3:.2f} ms/step; prefetch launches {c.prefetch_count()}, kernel launches {c.launch_count()}, rebins {c.rebin_count()}")
