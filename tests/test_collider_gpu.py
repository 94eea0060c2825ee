"""GPU parity of the reference's `Collider` API on the device (cb200_collider_*, through the C ABI).

Part 1 transcribes the reference's own integration tests (src/tests.rs:67-282, the lib.rs doctest) onto the device
Collider.  Part 2 drives the device Collider and the CPU oracle in lockstep with identical call sequences — config 1
(README loop with velocity halving), adds/removes, groups, can_interact, bulk steps with events processed in between —
and requires every return value to be equal: event tuples in order, times and get_hitbox results bit-for-bit.
"""
import importlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cb = importlib.import_module("collider-rs_b200")
col = importlib.import_module("collider-rs_b200.collider")
scenes = importlib.import_module("collider-rs_b200.scenes")

Collider, Shape, HbVel, Hitbox, PlacedShape = col.Collider, col.Shape, col.HbVel, col.Hitbox, col.PlacedShape
COLLIDE, SEPARATE, INF = col.COLLIDE, col.SEPARATE, col.INF


# ---------------------------------------------------------------- src/tests.rs:39-60 helpers
def advance_to_event(c, time):
    advance_c(c, time)
    assert c.next_time() == c.time()


def advance_c(c, time):
    while c.time() < time:
        assert c.next() is None
        c.set_time(min(c.next_time(), time))
    assert c.time() == time


def advance_through_events(c, time):
    while c.time() < time:
        c.next()
        c.set_time(min(c.next_time(), time))
    assert c.time() == time


# ---------------------------------------------------------------- part 1: src/tests.rs:67-282 on the device
def test_smoke():  # tests.rs:67-92
    c = Collider(4.0, 0.25)
    hb = Shape.square(2.0).place(-10.0, 0.0).still()
    hb.vel.vx = 1.0
    assert c.add_hitbox(0, hb) == []
    hb = Shape.circle(2.0).place(10.0, 0.0).still()
    hb.vel.vx = -1.0
    assert c.add_hitbox(1, hb) == []
    advance_to_event(c, 9.0)
    assert c.next() == (COLLIDE, 0, 1)
    advance_to_event(c, 11.125)
    assert c.next() == (SEPARATE, 0, 1)
    advance_c(c, 23.0)
    assert c.launch_count() > 0


def test_hitbox_updates():  # tests.rs:94-177
    c = Collider(4.0, 0.25)
    hb = Shape.square(2.0).place(-10.0, 0.0).still()
    hb.vel.vx = 1.0
    assert c.add_hitbox(0, hb) == []
    hb = Shape.circle(2.0).place(10.0, 0.0).still()
    hb.vel.vx = 1.0
    assert c.add_hitbox(1, hb) == []
    advance_c(c, 11.0)

    hb = c.get_hitbox(0)
    assert hb.value == Shape.square(2.0).place(1.0, 0.0)
    assert (hb.vel.vx, hb.vel.vy, hb.vel.rw, hb.vel.rh, hb.vel.end_time) == (1.0, 0.0, 0.0, 0.0, INF)
    hb.value.x, hb.value.y = 0.0, 2.0
    hb.vel.vx, hb.vel.vy = 0.0, -1.0
    assert c.remove_hitbox(0) == []
    assert c.add_hitbox(0, hb) == []
    advance_c(c, 14.0)

    hb = c.get_hitbox(1)
    assert hb.value == Shape.circle(2.0).place(24.0, 0.0)
    assert (hb.vel.vx, hb.vel.vy, hb.vel.rw, hb.vel.rh, hb.vel.end_time) == (1.0, 0.0, 0.0, 0.0, INF)
    hb.value.x, hb.value.y = 0.0, -8.0
    hb.vel.vx, hb.vel.vy = 0.0, 0.0
    assert c.remove_hitbox(1) == []
    assert c.add_hitbox(1, hb) == []
    advance_to_event(c, 19.0)

    assert c.next() == (COLLIDE, 0, 1)
    hb = c.get_hitbox(0)
    assert hb.value == Shape.square(2.0).place(0.0, -6.0)
    assert (hb.vel.vx, hb.vel.vy, hb.vel.end_time) == (0.0, -1.0, INF)
    hb.vel.vx, hb.vel.vy = 0.0, 0.0
    c.set_hitbox_vel(0, hb.vel)

    hb = c.get_hitbox(1)
    assert hb.value == Shape.circle(2.0).place(0.0, -8.0)
    assert (hb.vel.vx, hb.vel.vy, hb.vel.end_time) == (0.0, 0.0, INF)
    hb.vel.vx, hb.vel.vy = 0.0, 2.0
    c.set_hitbox_vel(1, hb.vel)

    hb = Shape.rect(2.0, 20.0).place(0.0, 0.0).still()
    assert sorted(c.add_hitbox(2, hb)) == [0, 1]
    advance_to_event(c, 21.125)
    assert c.next() == (SEPARATE, 0, 1)
    advance_c(c, 26.125)
    assert c.remove_hitbox(1) == [2]
    advance_c(c, 37.125)


def test_get_overlaps():  # tests.rs:179-232
    c = Collider(4.0, 0.25)
    c.add_hitbox(0, Shape.square(2.0).place(-10.0, 0.0).moving(1.0, 0.0))
    c.add_hitbox(1, Shape.circle(2.0).place(10.0, 0.0).moving(-1.0, 0.0))
    c.add_hitbox(2, Shape.square(2.0).place(0.0, 0.0).still())
    for i in range(3):
        assert c.get_overlaps(i) == []
    assert not c.is_overlapping(0, 1) and not c.is_overlapping(0, 2)
    assert not c.is_overlapping(1, 2) and not c.is_overlapping(1, 0)

    advance_through_events(c, 10.0)
    assert sorted(c.get_overlaps(0)) == [1, 2]
    assert sorted(c.get_overlaps(1)) == [0, 2]
    assert sorted(c.get_overlaps(2)) == [0, 1]
    assert c.is_overlapping(0, 1) and c.is_overlapping(0, 2) and c.is_overlapping(1, 2) and c.is_overlapping(1, 0)

    c.set_hitbox_vel(1, HbVel(1.0, 0.0))
    advance_through_events(c, 20.0)
    assert c.get_overlaps(0) == [1]
    assert c.get_overlaps(1) == [0]
    assert c.get_overlaps(2) == []
    assert c.is_overlapping(0, 1) and not c.is_overlapping(0, 2) and not c.is_overlapping(1, 2)

    c.remove_hitbox(2)
    assert c.get_overlaps(0) == [1] and c.get_overlaps(1) == [0]
    assert c.is_overlapping(0, 1)
    c.remove_hitbox(1)
    assert c.get_overlaps(0) == []


def test_query_overlaps():  # tests.rs:234-261
    c = Collider(4.0, 0.25)
    c.add_hitbox(0, Shape.square(2.0).place(-5.0, 0.0).moving(1.0, 0.0))
    c.add_hitbox(1, Shape.circle(2.0).place(0.0, 0.0).still())
    c.add_hitbox(2, Shape.circle(2.0).place(10.0, 0.0).moving(-1.0, 0.0))
    shape = Shape.circle(2.0).place(-1.0, 0.5)
    assert c.query_overlaps(shape, 5) == [1]
    advance_c(c, 3.0)
    assert sorted(c.query_overlaps(shape, 5)) == [0, 1]


def test_separate_initial_overlap():  # tests.rs:263-282
    c = Collider(4.0, 0.25)
    assert c.add_hitbox(0, Shape.square(1.0).place(0.0, 0.0).moving(0.0, 1.0)) == []
    assert c.add_hitbox(1, Shape.square(1.0).place(0.0, 0.0).still()) == [0]
    advance_to_event(c, 1.25)
    assert c.next() == (SEPARATE, 0, 1)
    advance_c(c, 1.5)


def test_doctest_velocity_halving():  # src/lib.rs:55-85 == README.md:44-87; documented output 9 / 13.01
    c = Collider(4.0, 0.01)
    assert c.add_hitbox(0, Shape.square(2.0).place(-10.0, 0.0).moving(1.0, 0.0)) == []
    assert c.add_hitbox(1, Shape.square(2.0).place(10.0, 0.0).moving(-1.0, 0.0)) == []
    log = []
    while c.time() < 20.0:
        c.set_time(min(c.next_time(), 20.0))
        ev = c.next()
        if ev is not None:
            kind, a, b = ev
            log.append((kind, a, b, c.time()))
            if kind == COLLIDE:
                for i in (a, b):
                    v = c.get_hitbox(i).vel
                    v.vx *= 0.5
                    v.vy *= 0.5
                    c.set_hitbox_vel(i, v)
    assert log == [(COLLIDE, 0, 1, 9.0), (SEPARATE, 0, 1, 13.01)]


def test_contract_panics():  # collider.rs:60-61, 98-100, 235; core/mod.rs:157-161
    with pytest.raises(cb.ColliderError, match="cell_width > padding"):
        Collider(0.25, 0.25)
    with pytest.raises(cb.ColliderError, match="padding > 0.0"):
        Collider(4.0, 0.0)
    c = Collider(4.0, 0.25)
    c.add_hitbox(0, Shape.square(2.0).place(0.0, 0.0).moving(1.0, 0.0))
    with pytest.raises(cb.ColliderError, match="next_time"):
        c.set_time(100.0)
    c.set_time(1.0)
    with pytest.raises(cb.ColliderError, match="rewind"):
        c.set_time(0.5)
    with pytest.raises(cb.ColliderError, match="not found"):
        c.get_hitbox(7)
    with pytest.raises(cb.ColliderError, match="at least"):
        c.add_hitbox(3, Shape.square(0.1).place(0.0, 0.0).still())
    assert c.len() == 1  # the failed add left nothing behind
    with pytest.raises(cb.ColliderError, match="end time"):
        c.set_hitbox_vel(0, HbVel(1.0, 0.0, end_time=0.5))


# ---------------------------------------------------------------- part 2: lockstep with the oracle
def f64_bits(x):
    return np.float64(x).view(np.uint64)


class Lockstep:
    """Applies every call to the oracle and to the device Collider and compares what they return."""

    def __init__(self, oracle, cell_width, padding):
        self.o = oracle.Collider(cell_width, padding)
        self.d = Collider(cell_width, padding)
        self.events = 0
        self.filtered = False

    def set_can_interact(self, fn):
        self.filtered = fn is not None
        self.o.set_can_interact(fn)
        self.d.set_can_interact(fn)

    def time(self):
        assert f64_bits(self.o.time()) == f64_bits(self.d.time())
        return self.o.time()

    def next_time(self):
        a, b = self.o.next_time(), self.d.next_time()
        assert f64_bits(a) == f64_bits(b), (a, b)
        return a

    def set_time(self, t):
        self.o.set_time(t)
        self.d.set_time(t)

    def next(self):
        a, b = self.o.next(), self.d.next()
        assert a == b, (a, b, self.o.time())
        self.events += a is not None
        return a

    def add_hitbox(self, id, hb, group=0, interact_mask=1):
        a, b = self.o.add_hitbox(id, hb, group, interact_mask), self.d.add_hitbox(id, hb, group, interact_mask)
        assert a == b, (id, a, b)
        return a

    def set_hitbox_vel(self, id, vel):
        self.o.set_hitbox_vel(id, vel)
        self.d.set_hitbox_vel(id, vel)

    def remove_hitbox(self, id):
        a, b = self.o.remove_hitbox(id), self.d.remove_hitbox(id)
        assert a == b
        return a

    def get_hitbox(self, id):
        a, b = self.o.get_hitbox(id), self.d.get_hitbox(id)
        va = (a.value.x, a.value.y, a.value.shape.w, a.value.shape.h, a.vel.vx, a.vel.vy, a.vel.rw, a.vel.rh, a.vel.end_time)
        vb = (b.value.x, b.value.y, b.value.shape.w, b.value.shape.h, b.vel.vx, b.vel.vy, b.vel.rw, b.vel.rh, b.vel.end_time)
        assert [f64_bits(x) for x in va] == [f64_bits(x) for x in vb], (id, va, vb)
        assert a.value.shape.kind == b.value.shape.kind
        return a

    def get_overlaps(self, id):
        a, b = self.o.get_overlaps(id), self.d.get_overlaps(id)
        assert a == b
        return a

    def query_overlaps(self, shape, id, group=0, interact_mask=1):
        a, b = sorted(self.o.query_overlaps(shape, id, group, interact_mask)), sorted(self.d.query_overlaps(shape, id, group, interact_mask))
        assert a == b, (a, b)
        return a

    def bulk_update(self, end_time=INF, vx=None, vy=None):
        self.o.bulk_update(end_time=end_time, vx=vx, vy=vy)
        res = self.d.bulk_update(end_time=end_time, vx=vx, vy=vy)
        t, idx, kind, a, b = self.o.dump_events()
        assert self.filtered or res.n_events == len(t)  # (the device counts events before the host's can_interact filter)
        return res


def scene_hitboxes(oracle, s, end_time=INF):
    out = []
    for i in range(len(s["id"])):
        shape = oracle.Shape(int(s["kind"][i]), s["dim_x"][i], s["dim_y"][i])
        out.append((int(s["id"][i]), oracle.Hitbox(oracle.PlacedShape(s["pos_x"][i], s["pos_y"][i], shape),
                                                    oracle.HbVel(s["vel_x"][i], s["vel_y"][i], end_time=end_time))))
    return out


def readme_loop(c, t_end, on_collide):
    """README.md:66-80: advance to the next event, pop it, react."""
    while c.time() < t_end:
        c.set_time(min(c.next_time(), t_end))
        ev = c.next()
        if ev is not None and ev[0] == COLLIDE:
            on_collide(ev[1], ev[2])


def halve(c):
    def react(a, b):
        for i in (a, b):
            v = c.get_hitbox(i).vel
            v.vx *= 0.5
            v.vy *= 0.5
            c.set_hitbox_vel(i, v)
    return react


def test_config1_readme_loop_lockstep(oracle):
    """BASELINE config 1 at 1/4 size and time: mixed circles/squares, cell 4, padding 0.01, velocity halving on Collide."""
    s = scenes.make_scene(250, 100.0, seed=scenes.SEED_BASE + 1, squares=True)
    c = Lockstep(oracle, scenes.CELL_WIDTH, scenes.PADDING)
    for id, hb in scene_hitboxes(oracle, s):
        c.add_hitbox(id, hb)
    readme_loop(c, 25.0, halve(c))
    assert c.events > 50
    assert c.d.hitbox_cache_hits() > 50, "the Collide handler's get_hitbox calls must be served from the prefetch"
    for i in (0, 100, 249):
        c.get_hitbox(i)
        c.get_overlaps(i)


def test_config1_with_small_rebin_threshold(oracle):
    """Same flow with the grid rebuilt every few updates: static-grid lookups + dirty list must agree with the oracle's hash grid."""
    s = scenes.make_scene(300, 110.0, seed=scenes.SEED_BASE + 11, squares=True)
    c = Lockstep(oracle, scenes.CELL_WIDTH, scenes.PADDING)
    c.d.set_rebin_threshold(7)
    for id, hb in scene_hitboxes(oracle, s):
        c.add_hitbox(id, hb)
    readme_loop(c, 12.0, halve(c))
    assert c.d.rebin_count() > 10 and c.events > 20


def test_adds_removes_groups_and_queries(oracle):
    rng = np.random.default_rng(7)
    c = Lockstep(oracle, 4.0, 0.05)
    c.d.set_rebin_threshold(16)
    live = {}
    next_id = 0
    masks = [0b01, 0b10, 0b11]

    def new_hitbox():
        kind = int(rng.integers(0, 2))
        w = float(rng.uniform(0.5, 3.0))
        h = w if kind == 0 else float(rng.uniform(0.5, 3.0))
        shape = oracle.Shape(kind, w, h)
        hb = oracle.Hitbox(oracle.PlacedShape(rng.uniform(-30, 30), rng.uniform(-30, 30), shape),
                           oracle.HbVel(rng.uniform(-3, 3), rng.uniform(-3, 3)))
        if rng.integers(0, 4) == 0:
            r = float(rng.uniform(-0.02, 0.2))
            hb.vel.rw, hb.vel.rh = r, (r if kind == 0 else float(rng.uniform(-0.02, 0.2)))
        return hb

    for step in range(400):
        op = rng.integers(0, 10)
        if op < 4 or len(live) < 20:
            # ids are mostly ascending, sometimes inserted between existing ones (forces a relabel on the device side)
            id = next_id * 10 if rng.integers(0, 6) else next_id * 10 - int(rng.integers(1, 9))
            next_id += 1
            if id in live or id < 0:
                continue
            g = int(rng.integers(0, 3))
            group = None if g == 2 and rng.integers(0, 3) == 0 else g % 2
            live[id] = (group, masks[g])
            c.add_hitbox(id, new_hitbox(), group, masks[g])
        elif op < 6:
            id = int(rng.choice(list(live)))
            c.remove_hitbox(id)
            del live[id]
        elif op < 8:
            id = int(rng.choice(list(live)))
            v = c.get_hitbox(id).vel
            v.vx, v.vy = float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3))
            c.set_hitbox_vel(id, v)
        else:
            kind = int(rng.integers(0, 2))
            w = float(rng.uniform(1, 12))
            q = oracle.PlacedShape(rng.uniform(-30, 30), rng.uniform(-30, 30), oracle.Shape(kind, w, w if kind == 0 else float(rng.uniform(1, 12))))
            c.query_overlaps(q, 999999, 0, int(rng.integers(1, 4)))
        # let time pass, processing events
        t_stop = c.time() + float(rng.uniform(0.0, 0.4))
        while c.time() < t_stop:
            c.set_time(min(c.next_time(), t_stop))
            while c.next() is not None:
                pass
    assert c.events > 30
    for id in live:
        c.get_overlaps(id)


def test_batched_queries_and_arbitrary_group_values(oracle):
    """cb200_collider_query_overlaps_batch answers many queries in one launch (same answers as one call each and as the
    oracle), and cb200_collider_group_index lets a wrapper use ANY u32 HbGroup value (core/mod.rs:199): the device sees dense
    indices in order of first use, the oracle is driven with those indices."""
    rng = np.random.default_rng(11)
    c = Lockstep(oracle, 4.0, 0.05)
    values = [4_000_000_000, 7, 1_000_000_007]                      # arbitrary HbGroup values
    wants = {4_000_000_000: [7, 4_000_000_000], 7: [1_000_000_007], 1_000_000_007: [4_000_000_000, 7, 1_000_000_007]}
    index = {}
    for id in range(600):
        v = values[id % 3]
        g, m = c.d.profile_of(v, wants[v])
        index[v] = g
        kind = int(rng.integers(0, 2))
        w = float(rng.uniform(0.5, 3.0))
        hb = oracle.Hitbox(oracle.PlacedShape(rng.uniform(-40, 40), rng.uniform(-40, 40), oracle.Shape(kind, w, w if kind == 0 else float(rng.uniform(0.5, 3.0)))),
                           oracle.HbVel(rng.uniform(-2, 2), rng.uniform(-2, 2)))
        c.add_hitbox(id, hb, g, m)
    assert sorted(index.values()) == [0, 1, 2] and c.d.group_index(7) == index[7]
    readme_loop(c, 1.5, lambda a, b: None)
    shapes, ids, masks = [], [], []
    for k in range(200):
        kind = int(rng.integers(0, 2))
        w = float(rng.uniform(1, 15))
        shapes.append(oracle.PlacedShape(rng.uniform(-40, 40), rng.uniform(-40, 40), oracle.Shape(kind, w, w if kind == 0 else float(rng.uniform(1, 15)))))
        ids.append(10_000 + k)
        masks.append(int(rng.integers(1, 8)))
    got = c.d.query_overlaps_batch(shapes, ids, masks)
    assert sum(len(g) for g in got) > 100
    for k in range(200):
        assert got[k] == sorted(c.o.query_overlaps(shapes[k], ids[k], 0, masks[k])) == sorted(c.d.query_overlaps(shapes[k], ids[k], 0, masks[k]))
    with pytest.raises(col.ColliderError):
        for v in range(100, 140):
            c.d.group_index(v)  # the 33rd distinct value


def test_can_interact_filter(oracle):
    s = scenes.make_scene(200, 90.0, seed=scenes.SEED_BASE + 21, squares=True)
    c = Lockstep(oracle, scenes.CELL_WIDTH, scenes.PADDING)
    c.set_can_interact(lambda a, b: (a + b) % 3 != 0)
    for id, hb in scene_hitboxes(oracle, s):
        c.add_hitbox(id, hb)
    readme_loop(c, 10.0, halve(c))
    c.bulk_update(end_time=c.time() + 5.0)
    readme_loop(c, 14.0, halve(c))
    assert c.events > 10


def test_bulk_steps_with_events_between_lockstep(oracle):
    """Config 2 shape through the Collider API only: bulk step every 0.1 with new velocities, events in between processed
    by next() with velocity halving on Collide — the device keeps overlaps and the queue consistent by itself."""
    n = 3000
    s = scenes.uniform_density(n, seed=scenes.SEED_BASE + 2)
    c = Lockstep(oracle, scenes.CELL_WIDTH, scenes.PADDING)
    for id, hb in scene_hitboxes(oracle, s, end_time=0.1):
        c.add_hitbox(id, hb)
    T = 0.0
    for step in range(6):
        vx = scenes.uniform(300 + step, n, 1, -2.0, 2.0)
        vy = scenes.uniform(300 + step, n, 2, -2.0, 2.0)
        T_next = T + 0.1
        c.bulk_update(end_time=T_next, vx=vx, vy=vy)

        def react(a, b):
            for i in (a, b):
                v = c.get_hitbox(i).vel
                v.vx *= 0.5
                v.vy *= 0.5
                c.set_hitbox_vel(i, v)
        readme_loop(c, T_next, react)
        T = T_next
    assert c.events > 50


def test_config1_full_size_lockstep(oracle):
    """BASELINE config 1 as stated: 1,000 mixed circles/squares in a 200x200 box, cell 4, padding 0.01, random velocities,
    the README loop with velocity halving on Collide — oracle and device in lockstep, every event and time compared."""
    s = scenes.c1(1000)
    c = Lockstep(oracle, scenes.CELL_WIDTH, scenes.PADDING)
    for id, hb in scene_hitboxes(oracle, s):
        c.add_hitbox(id, hb)
    readme_loop(c, 100.0, halve(c))
    assert c.events > 300
    assert 0 < c.d.prefetch_count() < c.events  # the narrowphase solves behind next() are computed many events per launch
    for i in range(0, 1000, 97):
        c.get_hitbox(i)
        c.get_overlaps(i)


def test_cpp_wrapper_runs_the_reference_tests(tmp_path):
    """include/collider_b200.hpp (the C++ twin of the Rust `Collider<P>` wrapper) + tests/host_collider_check.cpp:
    test_smoke, test_separate_initial_overlap and the lib.rs doctest of the reference, compiled with g++ against the .so."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(cb.LIB_PATH)
    exe = str(tmp_path / "host_collider_check")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(root, "tests", "host_collider_check.cpp"), "-L" + libdir, "-lcollider_b200",
                    "-Wl,-rpath," + libdir], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "host_collider_check ok" in out.stdout, out.stdout + out.stderr


def test_batched_add_equals_sequential_adds(oracle):
    """cb200_collider_add_hitboxes on an empty collider (one device pass) against the oracle's add_hitbox loop: the same
    immediate overlaps, then the same event stream through the README loop.  Dense enough that many shapes overlap at t = 0."""
    n = 4000
    s = scenes.make_scene(n, 220.0, seed=scenes.SEED_BASE + 77, squares=True)
    c = Lockstep(oracle, scenes.CELL_WIDTH, scenes.PADDING)
    want = []
    for id, hb in scene_hitboxes(oracle, s):
        for other in c.o.add_hitbox(id, hb):
            want.append((id, other))
    a, b = c.d.add_hitboxes(s)
    assert sorted(zip(a.tolist(), b.tolist())) == sorted(want) and len(want) > 50
    assert c.d.len() == n and c.d.launch_count() < 200  # one pass (+ buffer growth re-runs, the queue sort), not 4000 calls
    for i in (0, 1234, n - 1):
        c.get_overlaps(i)
        c.get_hitbox(i)
    readme_loop(c, 3.0, halve(c))
    assert c.events > 200
    # and on a non-empty collider it falls back to single adds with the same semantics
    extra = scenes.make_scene(50, 220.0, seed=scenes.SEED_BASE + 78, squares=True)
    extra["id"] = extra["id"] + np.uint64(n)
    want = []
    t_now = c.time()
    for id, hb in scene_hitboxes(oracle, extra):
        for other in c.o.add_hitbox(id, hb):
            want.append((id, other))
    a, b = c.d.add_hitboxes(extra)
    assert sorted(zip(a.tolist(), b.tolist())) == sorted(want)
    readme_loop(c, t_now + 1.0, halve(c))


def test_batched_add_dense(oracle):
    """The batched add at config 4's density: the immediate overlaps found by k_solve_dense (add mode) are the oracle's."""
    n = 3000
    s = scenes.uniform_density(n, density=1.0, seed=scenes.SEED_BASE + 79)
    c = Lockstep(oracle, scenes.CELL_WIDTH, scenes.PADDING)
    want = []
    for id, hb in scene_hitboxes(oracle, s):
        for other in c.o.add_hitbox(id, hb):
            want.append((id, other))
    a, b = c.d.add_hitboxes(s)
    assert sorted(zip(a.tolist(), b.tolist())) == sorted(want) and len(want) > 2000
    for i in (0, 1234, n - 1):
        c.get_overlaps(i)
    readme_loop(c, 0.25, halve(c))
    assert c.events > 200
