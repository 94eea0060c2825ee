"""GPU parity at the sizes the benchmark runs (BASELINE.json configs 2-5), through the C ABI.

  config 3   1,000,000 mixed shapes, uniform density: TWO chained bulk steps against ONE single-threaded oracle collider
             (the ground truth as it is: ascending set_hitbox_vel loop) — events in order, times and the whole rebased table
             bit-equal, candidate-pair set, cell membership.  The second step is taken after the oracle drained its events,
             with the overlap set re-uploaded, i.e. exactly the state a config-2/3 run is in.
  config 5   1,000,000 clustered shapes with fast small ones, long slot runs through k_solve_dense
  config 4   its density (1 shape per unit^2) at 1,000,000 hitboxes, through k_solve_dense
  f-4        config 3's density with resize velocities, four groups (one HbGroup None, asymmetric interact lists) and
             coordinates centred on the origin, 1,000,000 hitboxes
             Both against the TILED oracle (tests/tiled_oracle.py: 16 independent oracle colliders on the host threads,
             proven equal to the single oracle by tests/test_tiled_oracle.py) — the single oracle needs minutes for them.
  config 2   100,000 shapes through the `Collider` API only: batched add, bulk steps every 0.1 with the Collide events in
             between processed by next() and answered with set_hitbox_vel, device and oracle in lockstep.
"""
import importlib

import numpy as np
import pytest

from tiled_oracle import STATE_COLS, tiled_bulk_step

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")

CW, PAD = scenes.CELL_WIDTH, scenes.PADDING


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def pair_key(p, q):
    p, q = np.asarray(p).astype(np.uint64), np.asarray(q).astype(np.uint64)
    return np.sort(np.maximum(p, q) << np.uint64(32) | np.minimum(p, q))


def test_config3_one_million_two_chained_steps(oracle):
    from test_bulk_gpu import assert_step_parity, drain_until, seeded

    scene = scenes.c3(1_000_000)
    orc, eng = seeded(oracle, scene)
    T = 0.0
    for step in range(2):
        orc.bulk_update(end_time=T + 0.1, record_candidates=True)
        res = eng.bulk_update(T, end_time=T + 0.1)
        assert res.n_hitboxes == 1_000_000 and res.n_collide_solves > 400_000
        assert_step_parity(orc, eng, res)
        T += 0.1
        drain_until(orc, T)
        eng.set_overlaps(*orc.dump_overlaps())
    assert eng.event_sort_fallbacks() == 0


@pytest.mark.parametrize("name", ["config5_clustered", "config4_density", "config3_groups_resizing"])
def test_one_million_dense_scenes(oracle, name, monkeypatch):
    if name != "config3_groups_resizing":
        monkeypatch.setenv("CB200_DENSE_RANK", "4")  # the dense kernel from the first step (a run reaches it from its second step)
    scene = {"config5_clustered": scenes.c5, "config4_density": scenes.c3_dense, "config3_groups_resizing": scenes.c3_groups}[name](1_000_000)
    want = tiled_bulk_step(oracle, scene, CW, PAD, 0.1, tiles=4)
    eng = cb.Engine(CW, PAD)
    eng.upload_state(**want["state"])
    eng.set_overlaps(*want["overlaps"])
    res = eng.bulk_update(0.0, end_time=0.1)
    t, kind, a, b = want["events"]
    assert res.n_events == len(t) and res.n_collide_solves > {"config4_density": 20_000_000, "config5_clustered": 5_000_000, "config3_groups_resizing": 150_000}[name]
    ev = eng.fetch_events()
    np.testing.assert_array_equal(ev["kind"], kind)
    np.testing.assert_array_equal(ev["a"], a)
    np.testing.assert_array_equal(ev["b"], np.where(kind == 0, 0, b))
    np.testing.assert_array_equal(bits(ev["time"]), bits(t))
    assert bits([res.next_time])[0] == bits(t[:1])[0] and (res.next_event.kind, res.next_event.a) == (kind[0], a[0])
    dhi, dlo = eng.fetch_candidate_pairs()
    assert res.n_collide_solves == len(dhi) == len(want["candidates"][0])
    np.testing.assert_array_equal(pair_key(dhi, dlo), pair_key(*want["candidates"]))
    st = eng.download_state()
    for k in STATE_COLS:
        np.testing.assert_array_equal(bits(st[k]), bits(want["state_after"][k]), err_msg=k)
    gx, gy, gi = want["grid"]
    dx, dy, di = eng.fetch_cell_entries()
    assert res.n_entries == len(gx) == len(dx)
    o, d = np.lexsort((gy, gx, gi)), np.lexsort((dy, dx, di))
    np.testing.assert_array_equal(dx[d], gx[o])
    np.testing.assert_array_equal(dy[d], gy[o])
    np.testing.assert_array_equal(di[d], gi[o])


def test_config2_hundred_thousand_through_the_collider_api(oracle):
    """BASELINE config 2: 100k mixed circles/rects, bulk update every 0.1, events processed between the steps.  Oracle and
    device Collider in lockstep: every event tuple and every get_hitbox answer of the Collide handler compared."""
    from test_collider_gpu import Lockstep, readme_loop

    n = 100_000
    s = scenes.c2()
    c = Lockstep(oracle, CW, PAD)
    c.o.add_hitboxes(s, end_time=0.1)
    a, b = c.d.add_hitboxes(s, end_time=0.1)
    np.testing.assert_array_equal(pair_key(a, b), pair_key(*c.o.dump_overlaps()))
    T = 0.0
    for step in range(3):
        vx = scenes.uniform(300 + step, n, 1, -2.0, 2.0)
        vy = scenes.uniform(300 + step, n, 2, -2.0, 2.0)
        T_next = T + 0.1
        res = c.bulk_update(end_time=T_next, vx=vx, vy=vy)
        assert res.n_hitboxes == n

        def react(p, q):
            for i in (p, q):
                v = c.get_hitbox(i).vel
                v.vx *= 0.5
                v.vy *= 0.5
                c.set_hitbox_vel(i, v)
        readme_loop(c, T_next, react)
        T = T_next
    assert c.events > 1500
    for i in (0, 4711, n - 1):
        c.get_overlaps(i)
        c.get_hitbox(i)
