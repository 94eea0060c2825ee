"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module.  It mirrors the reference's public API (`Collider`, src/core/collider.rs:59-318) so the
KAT tests read like the reference's own tests (src/tests.rs).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

CIRCLE, RECT = 0, 1
COLLIDE, SEPARATE = 1, 2
EV_REITERATE, EV_COLLIDE, EV_SEPARATE = 0, 1, 2
INF = math.inf


class OraclePanic(RuntimeError):
    """A Rust `panic!`/`assert!` of the reference, restated."""


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle_collider.hpp", "oracle_geom.hpp")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.c_double, C.c_double]
        L.orc_free.argtypes = [C.c_void_p]
        for f in ("orc_time", "orc_next_time"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_set_time.argtypes = [C.c_void_p, C.c_double]
        L.orc_next.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_add_hitbox.restype = C.c_int64
        L.orc_add_hitbox.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_set_hitbox_vel.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_remove_hitbox.restype = C.c_int64
        L.orc_remove_hitbox.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_get_hitbox.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_get_raw.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_get_overlaps.restype = C.c_int64
        L.orc_get_overlaps.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_is_overlapping.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_query_overlaps.restype = C.c_int64
        L.orc_query_overlaps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_void_p, C.c_uint64]
        L.orc_len.restype = C.c_uint64
        L.orc_len.argtypes = [C.c_void_p]
        L.orc_bulk_update.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_double, C.c_int]
        L.orc_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_halving_loop.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        L.orc_bulk_candidates.restype = C.c_int64
        L.orc_bulk_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_dump_events.restype = C.c_int64
        L.orc_dump_events.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_uint64]
        L.orc_dump_grid.restype = C.c_int64
        L.orc_dump_grid.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_uint64]
        L.orc_collide_time_batch.argtypes = [C.c_uint64] + [C.c_void_p] * 5
        L.orc_separate_time_batch.argtypes = [C.c_uint64] + [C.c_void_p] * 4 + [C.c_double, C.c_void_p]
        L.orc_quad_root_ascending.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)]
        L.orc_index_rect_iter.restype = C.c_int64
        L.orc_index_rect_iter.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_index_rect_contains.argtypes = [C.c_int32] * 6
        L.orc_index_bounds.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        L.orc_advance.argtypes = [C.c_int, C.c_void_p] + [C.c_double] * 5 + [C.c_void_p]
        L.orc_edges.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_shape_overlaps.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_swept_rect.argtypes = [C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_set_can_interact.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_add_hitboxes.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 5
        L.orc_dump_state.restype = C.c_int64
        L.orc_dump_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_uint64]
        L.orc_dump_overlaps.restype = C.c_int64
        L.orc_dump_overlaps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f64(*vals) -> np.ndarray:
    return np.array(vals, dtype=np.float64)


# ---------------------------------------------------------------- value types (mirror geom/core)
class Shape:
    """Shape (src/geom/shape/mod.rs:37-102)."""

    def __init__(self, kind: int, w: float, h: float):
        if not (w >= 0.0 and h >= 0.0):
            raise OraclePanic("dims must be non-negative")
        if kind == CIRCLE and w != h:
            raise OraclePanic("circle width must equal height")
        self.kind, self.w, self.h = kind, float(w), float(h)

    @staticmethod
    def circle(diam):
        return Shape(CIRCLE, diam, diam)

    @staticmethod
    def rect(w, h):
        return Shape(RECT, w, h)

    @staticmethod
    def square(w):
        return Shape(RECT, w, w)

    def place(self, x, y):
        return PlacedShape(x, y, self)

    def __eq__(self, o):
        return (self.kind, self.w, self.h) == (o.kind, o.w, o.h)

    def __repr__(self):
        return f"Shape({'circle' if self.kind == CIRCLE else 'rect'}, {self.w}, {self.h})"


class PlacedShape:
    """PlacedShape (src/geom/shape/mod.rs:105-266)."""

    def __init__(self, x, y, shape: Shape):
        self.x, self.y, self.shape = float(x), float(y), shape

    def moving(self, vx, vy):
        return Hitbox(self, HbVel(vx, vy))

    def moving_until(self, vx, vy, end_time):
        return Hitbox(self, HbVel(vx, vy, end_time=end_time))

    def still(self):
        return Hitbox(self, HbVel(0.0, 0.0))

    def still_until(self, end_time):
        return Hitbox(self, HbVel(0.0, 0.0, end_time=end_time))

    def arr(self):
        return _f64(self.x, self.y, self.shape.w, self.shape.h)

    def __eq__(self, o):
        return (self.x, self.y) == (o.x, o.y) and self.shape == o.shape

    def __repr__(self):
        return f"PlacedShape(({self.x}, {self.y}), {self.shape})"


class HbVel:
    """HbVel (src/core/mod.rs:34-107)."""

    def __init__(self, vx=0.0, vy=0.0, rw=0.0, rh=0.0, end_time=INF):
        self.vx, self.vy, self.rw, self.rh, self.end_time = float(vx), float(vy), float(rw), float(rh), float(end_time)

    def copy(self):
        return HbVel(self.vx, self.vy, self.rw, self.rh, self.end_time)

    def arr(self):
        return _f64(self.vx, self.vy, self.rw, self.rh, self.end_time)

    def __repr__(self):
        return f"HbVel(({self.vx}, {self.vy}), ({self.rw}, {self.rh}), {self.end_time})"


class Hitbox:
    """Hitbox (src/core/mod.rs:119-188)."""

    def __init__(self, value: PlacedShape, vel: HbVel):
        self.value, self.vel = value, vel

    def __repr__(self):
        return f"Hitbox({self.value}, {self.vel})"


class Collider:
    """Mirror of `Collider<P>` (src/core/collider.rs:30-318) over the C++ restatement.

    Profiles are (id, group, interact_mask); `can_interact` is an optional Python callback
    `(id_a, id_b) -> bool` (default: always true, like the reference tests' TestHbProfile).
    """

    def __init__(self, cell_width: float, padding: float):
        self._L = lib()
        self._h = self._L.orc_new(cell_width, padding)
        if not self._h:
            raise OraclePanic(self._L.orc_last_error(None).decode())
        self._cb = None

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_free(self._h)
            self._h = None

    def _check(self, rc):
        if rc < 0:
            raise OraclePanic(self._L.orc_last_error(self._h).decode())
        return rc

    def set_can_interact(self, fn):
        if fn is None:
            self._cb = None
            self._L.orc_set_can_interact(self._h, None)
        else:
            self._cb = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_uint64)(lambda a, b: 1 if fn(a, b) else 0)
            self._L.orc_set_can_interact(self._h, C.cast(self._cb, C.c_void_p))

    def time(self):
        return self._L.orc_time(self._h)

    def next_time(self):
        return self._L.orc_next_time(self._h)

    def set_time(self, t):
        self._check(self._L.orc_set_time(self._h, t))

    def next(self):
        k, a, b = C.c_uint32(), C.c_uint64(), C.c_uint64()
        rc = self._check(self._L.orc_next(self._h, C.byref(k), C.byref(a), C.byref(b)))
        return (k.value, a.value, b.value) if rc == 1 else None

    def add_hitbox(self, id, hitbox: Hitbox, group=0, interact_mask=1):
        hb = np.concatenate([hitbox.value.arr(), hitbox.vel.arr()])
        out = np.zeros(max(64, self.len() + 1), dtype=np.uint64)
        g = -1 if group is None else int(group)
        n = self._check(self._L.orc_add_hitbox(self._h, id, g, interact_mask, hitbox.value.shape.kind, _p(hb), _p(out), out.size))
        return [int(x) for x in out[:n]]

    def set_hitbox_vel(self, id, vel: HbVel):
        self._check(self._L.orc_set_hitbox_vel(self._h, id, _p(vel.arr())))

    def remove_hitbox(self, id):
        out = np.zeros(max(64, self.len() + 1), dtype=np.uint64)
        n = self._check(self._L.orc_remove_hitbox(self._h, id, _p(out), out.size))
        return [int(x) for x in out[:n]]

    def get_hitbox(self, id) -> Hitbox:
        out = np.zeros(9)
        kind = C.c_int()
        self._check(self._L.orc_get_hitbox(self._h, id, _p(out), C.byref(kind)))
        return Hitbox(PlacedShape(out[0], out[1], Shape(kind.value, out[2], out[3])), HbVel(*out[4:9]))

    def get_raw(self, id):
        out = np.zeros(11)
        kind = C.c_int()
        self._check(self._L.orc_get_raw(self._h, id, _p(out), C.byref(kind)))
        return out, kind.value

    def get_overlaps(self, id):
        out = np.zeros(max(64, self.len() + 1), dtype=np.uint64)
        n = self._check(self._L.orc_get_overlaps(self._h, id, _p(out), out.size))
        return [int(x) for x in out[:n]]

    def is_overlapping(self, a, b):
        return bool(self._L.orc_is_overlapping(self._h, a, b))

    def query_overlaps(self, shape: PlacedShape, id, group=0, interact_mask=1):
        out = np.zeros(max(64, self.len() + 1), dtype=np.uint64)
        g = -1 if group is None else int(group)
        n = self._check(self._L.orc_query_overlaps(self._h, shape.shape.kind, _p(shape.arr()), id, g, interact_mask, _p(out), out.size))
        return [int(x) for x in out[:n]]

    def len(self):
        return int(self._L.orc_len(self._h))

    # ---- bulk helpers for large scenes (one call instead of one per hitbox) ----
    def add_hitboxes(self, scene: dict, end_time=INF):
        """add_hitbox for every row of a scene dict (id, pos_x, pos_y, dim_x, dim_y, vel_x, vel_y, kind[, rsz_x, rsz_y, group, interact_mask])."""
        n = len(scene["id"])
        z = np.zeros(n)
        cols = np.ascontiguousarray(np.stack([
            scene["pos_x"], scene["pos_y"], scene["dim_x"], scene["dim_y"], scene["vel_x"], scene["vel_y"],
            scene.get("rsz_x", z), scene.get("rsz_y", z), np.broadcast_to(np.asarray(end_time, dtype=np.float64), (n,)),
        ]), dtype=np.float64)
        ids = np.ascontiguousarray(scene["id"], dtype=np.uint64)
        kind = np.ascontiguousarray(scene["kind"], dtype=np.uint8)
        group = np.ascontiguousarray(scene.get("group", np.zeros(n, np.uint32)), dtype=np.uint32)
        mask = np.ascontiguousarray(scene.get("interact_mask", np.ones(n, np.uint32)), dtype=np.uint32)
        self._check(self._L.orc_add_hitboxes(self._h, n, _p(ids), _p(cols), _p(kind), _p(group), _p(mask)))

    def dump_state(self) -> dict:
        """Raw internal table in ascending id order, as the columns cb200_upload_state takes."""
        n = self.len()
        ids = np.zeros(n, dtype=np.uint64)
        cols = np.zeros((11, n))
        kind = np.zeros(n, dtype=np.uint8)
        group = np.zeros(n, dtype=np.uint32)
        mask = np.zeros(n, dtype=np.uint32)
        self._L.orc_dump_state(self._h, _p(ids), _p(cols), _p(kind), _p(group), _p(mask), n)
        names = ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "start_time", "end_time", "pub_end_time")
        out = {k: cols[i].copy() for i, k in enumerate(names)}
        out.update(id=ids, kind=kind, group=group, interact_mask=mask)
        return out

    def dump_overlaps(self):
        n = self._L.orc_dump_overlaps(self._h, None, None, 0)
        a = np.zeros(n, dtype=np.uint64)
        b = np.zeros(n, dtype=np.uint64)
        self._L.orc_dump_overlaps(self._h, _p(a), _p(b), n)
        return a, b

    # ---- bulk step + introspection (not in the reference API) ----
    def bulk_update(self, end_time=INF, vx=None, vy=None, rw=None, rh=None, end_time_arr=None, record_candidates=False):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (vx, vy, rw, rh, end_time_arr)]
        ptrs = [None if a is None else _p(a) for a in arrs]
        self._check(self._L.orc_bulk_update(self._h, *ptrs, float(end_time), 1 if record_candidates else 0))

    def halving_loop(self, t_end: float):
        """README.md:66-80 loop up to t_end with velocity halving on Collide, run natively; (events, collides)."""
        out = np.zeros(2, dtype=np.uint64)
        self._check(self._L.orc_halving_loop(self._h, float(t_end), _p(out)))
        return int(out[0]), int(out[1])

    def stats(self, reset=False):
        out = np.zeros(4, dtype=np.uint64)
        self._L.orc_stats(self._h, _p(out), 1 if reset else 0)
        return dict(updates=int(out[0]), collide_solves=int(out[1]), separate_solves=int(out[2]), grid_cells=int(out[3]))

    def bulk_candidates(self):
        n = self._L.orc_bulk_candidates(self._h, None, None, 0)
        hi = np.zeros(n, dtype=np.uint64)
        lo = np.zeros(n, dtype=np.uint64)
        self._L.orc_bulk_candidates(self._h, _p(hi), _p(lo), n)
        return hi, lo

    def dump_events(self):
        n = self._L.orc_dump_events(self._h, None, None, None, None, None, 0)
        t = np.zeros(n)
        idx = np.zeros(n, dtype=np.uint64)
        kind = np.zeros(n, dtype=np.uint32)
        a = np.zeros(n, dtype=np.uint64)
        b = np.zeros(n, dtype=np.uint64)
        self._L.orc_dump_events(self._h, _p(t), _p(idx), _p(kind), _p(a), _p(b), n)
        return t, idx, kind, a, b

    def dump_grid(self):
        n = self._L.orc_dump_grid(self._h, None, None, None, None, 0)
        x = np.zeros(n, dtype=np.int32)
        y = np.zeros(n, dtype=np.int32)
        g = np.zeros(n, dtype=np.uint32)
        i = np.zeros(n, dtype=np.uint64)
        self._L.orc_dump_grid(self._h, _p(x), _p(y), _p(g), _p(i), n)
        return x, y, g, i


# ---------------------------------------------------------------- primitives
def dur(px, py, dw, dh, vx=0.0, vy=0.0, rw=0.0, rh=0.0, duration=INF):
    return _f64(px, py, dw, dh, vx, vy, rw, rh, duration)


def collide_time_batch(a, ka, b, kb):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 9)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 9)
    ka = np.ascontiguousarray(ka, dtype=np.uint8).reshape(-1)
    kb = np.ascontiguousarray(kb, dtype=np.uint8).reshape(-1)
    out = np.zeros(a.shape[0])
    lib().orc_collide_time_batch(a.shape[0], _p(a), _p(ka), _p(b), _p(kb), _p(out))
    return out


def separate_time_batch(a, ka, b, kb, padding):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 9)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 9)
    ka = np.ascontiguousarray(ka, dtype=np.uint8).reshape(-1)
    kb = np.ascontiguousarray(kb, dtype=np.uint8).reshape(-1)
    out = np.zeros(a.shape[0])
    lib().orc_separate_time_batch(a.shape[0], _p(a), _p(ka), _p(b), _p(kb), float(padding), _p(out))
    return out


def collide_time(a, ka, b, kb):
    return float(collide_time_batch(a, [ka], b, [kb])[0])


def separate_time(a, ka, b, kb, padding):
    return float(separate_time_batch(a, [ka], b, [kb], padding)[0])


def quad_root_ascending(a, b, c):
    r = C.c_double()
    return r.value if lib().orc_quad_root_ascending(a, b, c, C.byref(r)) else None


def index_rect_iter(sx, sy, ex, ey):
    n = lib().orc_index_rect_iter(sx, sy, ex, ey, None, None, 0)
    if n < 0:
        raise OraclePanic(lib().orc_last_error(None).decode())
    xs = np.zeros(n, dtype=np.int32)
    ys = np.zeros(n, dtype=np.int32)
    lib().orc_index_rect_iter(sx, sy, ex, ey, _p(xs), _p(ys), n)
    return list(zip(xs.tolist(), ys.tolist()))


def index_rect_contains(rect, x, y):
    return bool(lib().orc_index_rect_contains(*rect, x, y))


def index_bounds(cell_width, px, py, dw, dh):
    out = np.zeros(4, dtype=np.int32)
    if lib().orc_index_bounds(cell_width, _p(_f64(px, py, dw, dh)), _p(out)) != 0:
        raise OraclePanic(lib().orc_last_error(None).decode())
    return tuple(out.tolist())


def advance(shape: PlacedShape, vx, vy, rw, rh, elapsed) -> PlacedShape:
    out = np.zeros(4)
    if lib().orc_advance(shape.shape.kind, _p(shape.arr()), vx, vy, rw, rh, elapsed, _p(out)) != 0:
        raise OraclePanic(lib().orc_last_error(None).decode())
    s = Shape.__new__(Shape)
    s.kind, s.w, s.h = shape.shape.kind, float(out[2]), float(out[3])
    return PlacedShape(out[0], out[1], s)


def edges(shape: PlacedShape):
    out = np.zeros(4)
    lib().orc_edges(_p(shape.arr()), _p(out))
    return tuple(out.tolist())


def shape_overlaps(a: PlacedShape, b: PlacedShape) -> bool:
    return bool(lib().orc_shape_overlaps(a.shape.kind, _p(a.arr()), b.shape.kind, _p(b.arr())))


def swept_rect(cell_width, d, kind):
    bbox = np.zeros(4)
    rect = np.zeros(4, dtype=np.int32)
    if lib().orc_swept_rect(cell_width, _p(np.ascontiguousarray(d, dtype=np.float64)), kind, _p(bbox), _p(rect)) != 0:
        raise OraclePanic(lib().orc_last_error(None).decode())
    return bbox, tuple(rect.tolist())
