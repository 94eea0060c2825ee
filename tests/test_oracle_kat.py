"""Pins the CPU oracle against every known-answer assertion the reference's own tests hold for the
hot path.  Each test names the reference test it restates (paths under
/root/reference).  All f64 comparisons that are `assert_eq!` in the reference are exact (`==`) here.
"""
import math

import pytest

from oracle.oracle_py import (CIRCLE, COLLIDE, RECT, SEPARATE, Collider, HbVel, OraclePanic, PlacedShape, Shape, advance,
                              collide_time, dur, edges, index_bounds, index_rect_contains, index_rect_iter,
                              quad_root_ascending, separate_time)

INF = math.inf
SQRT2 = math.sqrt(2.0)


# ---------------------------------------------------------------- src/core/dur_hitbox/mod.rs:116-280
def test_rect_rect_collision():  # dur_hitbox/mod.rs:117-128
    a = dur(-11.0, 0.0, 2.0, 2.0, vx=2.0, duration=100.0)
    b = dur(12.0, 2.0, 2.0, 4.0, vx=-0.5, rw=1.0, duration=100.0)
    assert collide_time(a, RECT, b, RECT) == 7.0
    assert collide_time(b, RECT, a, RECT) == 7.0
    assert separate_time(a, RECT, b, RECT, 0.1) == 0.0


def test_circle_circle_collision():  # :131-145
    a = dur(-0.1 * SQRT2, 0.0, 2.0, 2.0, vx=0.1, duration=100.0)
    d = 2.0 + SQRT2 * 0.1
    b = dur(3.0 * SQRT2, 0.0, d, d, vx=-2.0, vy=1.0, rw=-0.1, rh=-0.1, duration=100.0)
    assert abs(collide_time(a, CIRCLE, b, CIRCLE) - SQRT2) < 1e-7
    assert separate_time(a, CIRCLE, b, CIRCLE, 0.1) == 0.0


def test_rect_circle_collision():  # :148-158
    a = dur(-11.0, 0.0, 2.0, 2.0, vx=2.0, duration=100.0)
    b = dur(12.0, 2.0, 2.0, 4.0, vx=-1.0, duration=100.0)
    assert collide_time(a, CIRCLE, b, RECT) == 7.0
    assert collide_time(b, RECT, a, CIRCLE) == 7.0
    assert separate_time(a, CIRCLE, b, RECT, 0.1) == 0.0


def test_rect_circle_angled_collision():  # :161-170
    a = dur(0.0, 0.0, 2.0, 2.0, duration=100.0)
    b = dur(5.0, 5.0, 2.0, 2.0, vx=-1.0, vy=-1.0, duration=100.0)
    assert collide_time(a, RECT, b, CIRCLE) == 4.0 - 1.0 / math.sqrt(2.0)


def test_rect_rect_separation():  # :173-184
    a = dur(0.0, 0.0, 6.0, 4.0, vx=1.0, vy=1.0, duration=100.0)
    b = dur(1.0, 0.0, 4.0, 4.0, vx=0.5, duration=100.0)
    assert separate_time(a, RECT, b, RECT, 0.1) == 4.1
    assert separate_time(b, RECT, a, RECT, 0.1) == 4.1
    assert collide_time(a, RECT, b, RECT) == 0.0


def test_circle_circle_separation():  # :187-197
    a = dur(2.0, 5.0, 2.0, 2.0, duration=100.0)
    b = dur(3.0, 4.0, 1.8, 1.8, vx=-1.0, vy=1.0, duration=100.0)
    assert separate_time(a, CIRCLE, b, CIRCLE, 0.1) == 1.0 + SQRT2
    assert separate_time(b, CIRCLE, a, CIRCLE, 0.1) == 1.0 + SQRT2
    assert collide_time(a, CIRCLE, b, CIRCLE) == 0.0


def test_rect_circle_separation():  # :200-210
    a = dur(4.0, 2.0, 4.0, 6.0, duration=100.0)
    b = dur(3.0, 4.0, 3.8, 3.8, vx=-1.0, vy=1.0, duration=100.0)
    assert separate_time(a, RECT, b, CIRCLE, 0.1) == 1.0 + SQRT2
    assert separate_time(b, CIRCLE, a, RECT, 0.1) == 1.0 + SQRT2
    assert collide_time(a, RECT, b, CIRCLE) == 0.0


def test_rect_circle_angled_separation():  # :213-222
    a = dur(0.0, 0.0, 2.0, 2.0, duration=100.0)
    b = dur(-1.0, 1.0, 2.0, 2.0, vx=1.0, vy=-1.0, duration=100.0)
    assert separate_time(a, RECT, b, CIRCLE, 0.1) == 2.0 + 1.1 / math.sqrt(2.0)


def test_no_collision():  # :225-243
    a = dur(-11.0, 0.0, 2.0, 2.0, vx=2.0, duration=100.0)
    b = dur(12.0, 2.0, 2.0, 4.0, vx=-1.0, vy=1.0, duration=100.0)
    assert collide_time(a, RECT, b, RECT) == INF
    assert separate_time(a, RECT, b, RECT, 0.1) == 0.0
    b = dur(12.0, 2.0, 2.0, 2.0, vx=-1.0, vy=1.0, duration=100.0)
    assert collide_time(a, RECT, b, CIRCLE) == INF
    assert separate_time(a, RECT, b, CIRCLE, 0.1) == 0.0
    assert collide_time(a, CIRCLE, b, CIRCLE) == INF
    assert separate_time(a, CIRCLE, b, CIRCLE, 0.1) == 0.0


def test_no_separation():  # :246-265
    a = dur(5.0, 1.0, 2.0, 2.0, vx=2.0, vy=1.0, duration=100.0)
    b = dur(5.0, 1.0, 2.0, 4.0, vx=2.0, vy=1.0, duration=100.0)
    assert separate_time(a, RECT, b, RECT, 0.1) == INF
    assert collide_time(a, RECT, b, RECT) == 0.0
    b = dur(5.0, 1.0, 2.0, 2.0, vx=2.0, vy=1.0, duration=100.0)
    assert separate_time(a, RECT, b, CIRCLE, 0.1) == INF
    assert collide_time(a, RECT, b, CIRCLE) == 0.0
    assert separate_time(a, CIRCLE, b, CIRCLE, 0.1) == INF
    assert collide_time(a, CIRCLE, b, CIRCLE) == 0.0


def test_low_duration():  # :268-280
    d = 4.0 - SQRT2 + 0.01
    a = dur(0.0, 0.0, 2.0, 2.0, duration=d)
    b = dur(4.0, 4.0, 2.0, 2.0, vx=-1.0, vy=-1.0, duration=d)
    assert collide_time(a, CIRCLE, b, CIRCLE) == 4.0 - SQRT2
    a[8] -= 0.02
    assert collide_time(a, CIRCLE, b, CIRCLE) == INF
    b[8] -= 0.02
    assert collide_time(a, CIRCLE, b, CIRCLE) == INF


# ---------------------------------------------------------------- src/util.rs:149-156
def test_quad_root_ascending():
    assert abs(quad_root_ascending(1e-14, 2.0, -1.0) - 0.5) < 1e-7
    assert abs(quad_root_ascending(0.0, 2.0, -1.0) - 0.5) < 1e-7
    assert abs(quad_root_ascending(100.0, -1.0, -1e-16) - 0.01) < 1e-7
    assert math.isinf(quad_root_ascending(0.0, -2.0, 1.0))
    assert quad_root_ascending(-3.0, 0.0, -1.0) is None
    assert quad_root_ascending(1.0, 1.0, 1.0) is None


# ---------------------------------------------------------------- src/index_rect.rs:79-117
def test_index_rect_iterator():
    cells = index_rect_iter(2, 3, 5, 7)
    assert len(cells) == 12 and len(set(cells)) == 12
    assert all(2 <= x < 5 and 3 <= y < 7 for x, y in cells)
    # column-major walk: x outer, y inner (index_rect.rs:53-72)
    assert cells[:5] == [(2, 3), (2, 4), (2, 5), (2, 6), (3, 3)]


def test_index_rect_contains():
    r = (2, 3, 5, 7)
    assert index_rect_contains(r, 2, 3) and index_rect_contains(r, 4, 6)
    for x, y in [(1, 3), (2, 2), (5, 6), (4, 7)]:
        assert not index_rect_contains(r, x, y)


def test_index_rect_new():
    with pytest.raises(OraclePanic):
        index_rect_iter(4, -5, 4, -4)
    with pytest.raises(OraclePanic):
        index_rect_iter(4, -5, 5, -5)
    assert index_rect_iter(4, -5, 5, -4) == [(4, -5)]


# ---------------------------------------------------------------- src/geom/shape/tests.rs:17-49
def test_circle_advance():
    s = advance(Shape.circle(2.0).place(3.0, 5.0), 1.0, 2.0, -0.25, -0.25, 2.0)
    assert s == Shape.circle(1.5).place(5.0, 9.0)


def test_rect_advance():
    s = advance(Shape.rect(2.0, 5.0).place(3.0, 5.0), 1.0, 2.0, -0.25, 1.0, 2.0)
    assert s == Shape.rect(1.5, 7.0).place(5.0, 9.0)


def test_illegal_circle_advance():
    with pytest.raises(OraclePanic):
        advance(Shape.circle(2.0).place(3.0, 5.0), 1.0, 2.0, -0.25, -0.24, 2.0)


def test_edges():
    assert edges(Shape.rect(4.0, 6.0).place(3.0, 5.0)) == (1.0, 2.0, 5.0, 8.0)


# ---------------------------------------------------------------- Grid::index_bounds, grid.rs:101-113
# (no reference test exists — grid.rs:28 TODO; pinned by reading the code)
def test_index_bounds_edge_cases():
    assert index_bounds(4.0, 7.0, 1.0, 2.0, 2.0) == (1, 0, 2, 1)    # [6,8] -> column 1 only
    assert index_bounds(4.0, 9.0, 1.0, 2.0, 2.0) == (2, 0, 3, 1)    # [8,10] -> column 2 only
    assert index_bounds(4.0, 8.0, 8.0, 0.0, 0.0) == (2, 2, 3, 3)    # zero extent on a line -> one cell
    assert index_bounds(4.0, -1.0, -1.0, 1.0, 1.0) == (-1, -1, 0, 0)  # negative coords floor toward -inf
    assert index_bounds(4.0, -1e300, 0.5, 1.0, 0.5)[0::2] == (-2147483648, -2147483647)  # saturating `as i32`
    with pytest.raises(OraclePanic):  # +saturation: start == end == i32::MAX -> "IndexRect contains no elements"
        index_bounds(4.0, 1e300, 0.5, 1.0, 0.5)


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


# ---------------------------------------------------------------- src/tests.rs:67-282
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


# ---------------------------------------------------------------- contract panics (collider.rs:60-61, 98-100)
def test_contract_panics():
    with pytest.raises(OraclePanic, match="cell_width > padding"):
        Collider(0.25, 0.25)
    with pytest.raises(OraclePanic, match="padding > 0.0"):
        Collider(4.0, 0.0)
    c = Collider(4.0, 0.25)
    c.add_hitbox(0, Shape.square(2.0).place(0.0, 0.0).moving(1.0, 0.0))
    with pytest.raises(OraclePanic, match="next_time"):
        c.set_time(100.0)
    c.set_time(1.0)
    with pytest.raises(OraclePanic, match="rewind"):
        c.set_time(0.5)
    with pytest.raises(OraclePanic, match="not found"):
        c.get_hitbox(7)
    with pytest.raises(OraclePanic, match="at least"):
        c.add_hitbox(3, Shape.square(0.1).place(0.0, 0.0).still())
