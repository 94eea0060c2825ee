"""`Collider` — the reference's public API (src/core/collider.rs:59-318) over the CUDA library.

Python mirror of what a Rust `Collider<P: HbProfile>` wrapper does with the `cb200_collider_*` C ABI
(include/collider_b200.h): every method forwards one call; geometry runs on the device.  The value types
(`Shape`, `PlacedShape`, `HbVel`, `Hitbox`) mirror src/geom/shape/mod.rs and src/core/mod.rs closely enough that
the reference's own tests (src/tests.rs, the lib.rs doctest) can be transcribed line by line.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import ColliderError, StepResult, VelView, load_library, _col, GROUP_NONE, KIND_CIRCLE, KIND_RECT, EV_COLLIDE, EV_SEPARATE

INF = math.inf
COLLIDE, SEPARATE = EV_COLLIDE, EV_SEPARATE  # HbEvent (collider.rs:500-511)


class HitboxC(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "end_time")] + [
        ("kind", C.c_uint32), ("_pad", C.c_uint32)]


class ProfileC(C.Structure):
    _fields_ = [("id", C.c_uint64), ("group", C.c_uint32), ("interact_mask", C.c_uint32)]


CAN_INTERACT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64)


class Shape:
    """Shape (src/geom/shape/mod.rs:37-102)."""

    def __init__(self, kind: int, w: float, h: float):
        self.kind, self.w, self.h = kind, float(w), float(h)

    @staticmethod
    def circle(diam):
        return Shape(KIND_CIRCLE, diam, diam)

    @staticmethod
    def rect(w, h):
        return Shape(KIND_RECT, w, h)

    @staticmethod
    def square(w):
        return Shape(KIND_RECT, w, w)

    def place(self, x, y):
        return PlacedShape(x, y, self)

    def __eq__(self, o):
        return (self.kind, self.w, self.h) == (o.kind, o.w, o.h)

    def __repr__(self):
        return f"Shape({'circle' if self.kind == KIND_CIRCLE else 'rect'}, {self.w}, {self.h})"


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

    def __repr__(self):
        return f"HbVel(({self.vx}, {self.vy}), ({self.rw}, {self.rh}), {self.end_time})"


class Hitbox:
    """Hitbox (src/core/mod.rs:119-188)."""

    def __init__(self, value: PlacedShape, vel: HbVel):
        self.value, self.vel = value, vel

    def __repr__(self):
        return f"Hitbox({self.value}, {self.vel})"


def _bind(L):
    if getattr(L, "_collider_bound", False):
        return
    vp, u64, dbl, i32, u32 = C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_uint32
    L.cb200_collider_new.argtypes = [dbl, dbl, i32, C.POINTER(vp)]
    L.cb200_collider_free.argtypes = [vp]
    L.cb200_collider_free.restype = None
    L.cb200_collider_last_error.argtypes = [vp]
    L.cb200_collider_last_error.restype = C.c_char_p
    L.cb200_collider_set_can_interact.argtypes = [vp, vp, vp]
    L.cb200_collider_time.argtypes = [vp]
    L.cb200_collider_time.restype = dbl
    L.cb200_collider_next_time.argtypes = [vp, C.POINTER(dbl)]
    L.cb200_collider_set_time.argtypes = [vp, dbl]
    L.cb200_collider_next.argtypes = [vp, C.POINTER(i32), C.POINTER(u32), C.POINTER(u64), C.POINTER(u64)]
    L.cb200_collider_get_hitbox.argtypes = [vp, u64, C.POINTER(HitboxC)]
    L.cb200_collider_add_hitbox.argtypes = [vp, C.POINTER(ProfileC), C.POINTER(HitboxC), vp, u64, C.POINTER(u64)]
    L.cb200_collider_add_hitboxes.argtypes = [vp, u64, C.POINTER(ProfileC), C.POINTER(HitboxC), vp, vp, u64, C.POINTER(u64)]
    L.cb200_collider_set_hitbox_vel.argtypes = [vp, u64, dbl, dbl, dbl, dbl, dbl]
    L.cb200_collider_remove_hitbox.argtypes = [vp, u64, vp, u64, C.POINTER(u64)]
    L.cb200_collider_get_overlaps.argtypes = [vp, u64, vp, u64, C.POINTER(u64)]
    L.cb200_collider_is_overlapping.argtypes = [vp, u64, u64, C.POINTER(i32)]
    L.cb200_collider_query_overlaps.argtypes = [vp, u32, dbl, dbl, dbl, dbl, C.POINTER(ProfileC), vp, u64, C.POINTER(u64)]
    L.cb200_collider_bulk_update.argtypes = [vp, C.POINTER(VelView), dbl, C.POINTER(StepResult)]
    L.cb200_collider_len.argtypes = [vp]
    L.cb200_collider_len.restype = u64
    L.cb200_collider_set_rebin_threshold.argtypes = [vp, u32]
    L.cb200_collider_rebin_count.argtypes = [vp]
    L.cb200_collider_rebin_count.restype = u64
    L.cb200_collider_prefetch_count.argtypes = [vp]
    L.cb200_collider_prefetch_count.restype = u64
    L.cb200_collider_hitbox_cache_hits.argtypes = [vp]
    L.cb200_collider_hitbox_cache_hits.restype = u64
    L.cb200_collider_run_halving_loop.argtypes = [vp, C.c_double, C.POINTER(u64), C.POINTER(u64)]
    L.cb200_collider_query_overlaps_batch.argtypes = [vp, u64, vp, vp, u64, vp]
    L.cb200_collider_group_index.argtypes = [vp, u32, C.POINTER(u32)]
    L.cb200_collider_engine.argtypes = [vp]
    L.cb200_collider_engine.restype = vp
    L._collider_bound = True


class Collider:
    """`Collider<P>` with P = (id, group, interact_mask); `can_interact` is an optional `(id_a, id_b) -> bool`."""

    def __init__(self, cell_width: float, padding: float, device: int = 0):
        self._L = load_library()
        _bind(self._L)
        h = C.c_void_p()
        rc = self._L.cb200_collider_new(cell_width, padding, device, C.byref(h))
        if rc != 0:
            raise ColliderError(rc, self._L.cb200_collider_last_error(None).decode())
        self._h = h
        self._cb = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.cb200_collider_free(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise ColliderError(rc, self._L.cb200_collider_last_error(self._h).decode())

    def _ids(self, call):
        cap = 64
        while True:
            out = np.zeros(cap, dtype=np.uint64)
            n = C.c_uint64()
            self._check(call(out.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
            if n.value <= cap:
                return [int(x) for x in out[: n.value]]
            cap = int(n.value)

    @staticmethod
    def _profile(id, group, interact_mask):
        return ProfileC(int(id), GROUP_NONE if group is None else int(group), int(interact_mask))

    def set_can_interact(self, fn):
        if fn is None:
            self._cb = None
            self._check(self._L.cb200_collider_set_can_interact(self._h, None, None))
        else:
            self._cb = CAN_INTERACT_FN(lambda user, a, b: 1 if fn(a, b) else 0)
            self._check(self._L.cb200_collider_set_can_interact(self._h, C.cast(self._cb, C.c_void_p), None))

    def time(self):
        return self._L.cb200_collider_time(self._h)

    def next_time(self):
        t = C.c_double()
        self._check(self._L.cb200_collider_next_time(self._h, C.byref(t)))
        return t.value

    def set_time(self, t):
        self._check(self._L.cb200_collider_set_time(self._h, t))

    def next(self):
        has, k, a, b = C.c_int(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._check(self._L.cb200_collider_next(self._h, C.byref(has), C.byref(k), C.byref(a), C.byref(b)))
        return (k.value, a.value, b.value) if has.value else None

    def add_hitbox(self, id, hitbox, group=0, interact_mask=1):
        v, s, w = hitbox.value, hitbox.value.shape, hitbox.vel
        hb = HitboxC(v.x, v.y, s.w, s.h, w.vx, w.vy, w.rw, w.rh, w.end_time, s.kind, 0)
        p = self._profile(id, group, interact_mask)
        # the add happens once; a too-small id buffer is re-read through get_overlaps
        out = np.zeros(64, dtype=np.uint64)
        n = C.c_uint64()
        self._check(self._L.cb200_collider_add_hitbox(self._h, C.byref(p), C.byref(hb), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        if n.value <= out.size:
            return [int(x) for x in out[: n.value]]
        return self.get_overlaps(id)

    def add_hitboxes(self, scene: dict, end_time=INF):
        """Batched add_hitbox over a scene dict (id, pos_x, pos_y, dim_x, dim_y, vel_x, vel_y, kind[, rsz_x, rsz_y, group,
        interact_mask]); returns the immediately overlapping pairs as two id arrays (later-added id, earlier id)."""
        n = len(scene["id"])
        z = np.zeros(n)
        # the two C arrays, filled column-wise through numpy views of the same layout
        pd = np.zeros(n, dtype=np.dtype([("id", "<u8"), ("group", "<u4"), ("interact_mask", "<u4")]))
        hd = np.zeros(n, dtype=np.dtype([(k, "<f8") for k in ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "end_time")]
                                        + [("kind", "<u4"), ("_pad", "<u4")]))
        assert pd.dtype.itemsize == C.sizeof(ProfileC) and hd.dtype.itemsize == C.sizeof(HitboxC)
        pd["id"] = scene["id"]
        pd["group"] = scene.get("group", np.zeros(n, np.uint32))
        pd["interact_mask"] = scene.get("interact_mask", np.ones(n, np.uint32))
        for k in ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y"):
            hd[k] = scene[k]
        hd["rsz_x"], hd["rsz_y"] = scene.get("rsz_x", z), scene.get("rsz_y", z)
        hd["end_time"] = np.broadcast_to(np.asarray(end_time, dtype=np.float64), (n,))
        hd["kind"] = scene["kind"]
        profs = C.cast(pd.ctypes.data, C.POINTER(ProfileC))
        hbs = C.cast(hd.ctypes.data, C.POINTER(HitboxC))
        cap = 1024
        while True:
            a = np.zeros(cap, dtype=np.uint64)
            b = np.zeros(cap, dtype=np.uint64)
            m = C.c_uint64()
            self._check(self._L.cb200_collider_add_hitboxes(self._h, n, profs, hbs, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), cap, C.byref(m)))
            if m.value <= cap:
                return a[: m.value], b[: m.value]
            # the hitboxes are in; only the pair list was cut short: read it back from the overlap sets
            pairs = [(i, j) for i in scene["id"] for j in self.get_overlaps(int(i)) if j < i]
            return np.array([p[0] for p in pairs], dtype=np.uint64), np.array([p[1] for p in pairs], dtype=np.uint64)

    def set_hitbox_vel(self, id, vel):
        self._check(self._L.cb200_collider_set_hitbox_vel(self._h, id, vel.vx, vel.vy, vel.rw, vel.rh, vel.end_time))

    def remove_hitbox(self, id):
        before = self.get_overlaps(id)
        n = C.c_uint64()
        self._check(self._L.cb200_collider_remove_hitbox(self._h, id, None, 0, C.byref(n)))
        return before

    def get_hitbox(self, id) -> Hitbox:
        hb = HitboxC()
        self._check(self._L.cb200_collider_get_hitbox(self._h, id, C.byref(hb)))
        return Hitbox(PlacedShape(hb.pos_x, hb.pos_y, Shape(hb.kind, hb.dim_x, hb.dim_y)), HbVel(hb.vel_x, hb.vel_y, hb.rsz_x, hb.rsz_y, hb.end_time))

    def get_overlaps(self, id):
        return self._ids(lambda out, cap, n: self._L.cb200_collider_get_overlaps(self._h, id, out, cap, n))

    def is_overlapping(self, a, b):
        r = C.c_int()
        self._check(self._L.cb200_collider_is_overlapping(self._h, a, b, C.byref(r)))
        return bool(r.value)

    def query_overlaps(self, shape, id, group=0, interact_mask=1):
        p = self._profile(id, group, interact_mask)
        return self._ids(lambda out, cap, n: self._L.cb200_collider_query_overlaps(self._h, shape.shape.kind, shape.x, shape.y, shape.shape.w, shape.shape.h,
                                                                                  C.byref(p), out, cap, n))

    def query_overlaps_batch(self, shapes, ids, interact_masks=None):
        """query_overlaps for many shapes in one device launch.  shapes: PlacedShape list; ids: the querying profiles' ids;
        returns a list of ascending id lists."""
        n = len(shapes)
        qd = np.zeros(n, dtype=np.dtype([("kind", "<u4"), ("_pad", "<u4"), ("pos_x", "<f8"), ("pos_y", "<f8"), ("dim_x", "<f8"), ("dim_y", "<f8"),
                                         ("id", "<u8"), ("group", "<u4"), ("interact_mask", "<u4")]))
        assert qd.dtype.itemsize == 56
        for k, sh in enumerate(shapes):
            qd[k] = (sh.shape.kind, 0, sh.x, sh.y, sh.shape.w, sh.shape.h, ids[k], 0, 1 if interact_masks is None else interact_masks[k])
        offsets = np.zeros(n + 1, dtype=np.uint64)
        cap = max(64, 8 * n)
        while True:
            out = np.zeros(cap, dtype=np.uint64)
            self._check(self._L.cb200_collider_query_overlaps_batch(self._h, n, qd.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), cap,
                                                                    offsets.ctypes.data_as(C.c_void_p)))
            if int(offsets[n]) <= cap:
                return [out[int(offsets[k]):int(offsets[k + 1])].tolist() for k in range(n)]
            cap = int(offsets[n])

    def group_index(self, group_value: int) -> int:
        """Device group index (0..31) of an arbitrary HbGroup value (cb200_collider_group_index)."""
        i = C.c_uint32()
        self._check(self._L.cb200_collider_group_index(self._h, group_value, C.byref(i)))
        return i.value

    def profile_of(self, group, interact_groups):
        """(group index or None, interact mask) for arbitrary HbGroup values — what add_hitbox / query_overlaps take."""
        g = None if group is None else self.group_index(group)
        m = 0
        for v in interact_groups:
            m |= 1 << self.group_index(v)
        return g, m

    def len(self):
        return int(self._L.cb200_collider_len(self._h))

    def bulk_update(self, end_time=INF, vx=None, vy=None, rw=None, rh=None, end_time_arr=None) -> StepResult:
        """The bulk advance/rebuild entry point: for id ascending, set_hitbox_vel(id, vel_id) at the current time, in one
        device pass.  Arrays are in ascending-id order."""
        n = self.len()
        arrs = [vx, vy, rw, rh, end_time_arr]
        res = StepResult()
        if all(a is None for a in arrs):
            rc = self._L.cb200_collider_bulk_update(self._h, None, end_time, C.byref(res))
        else:
            keep = [None if a is None else _col(a, np.float64, n) for a in arrs]
            v = VelView(*[None if a is None else a.ctypes.data for a in keep])
            rc = self._L.cb200_collider_bulk_update(self._h, C.byref(v), end_time, C.byref(res))
        self._check(rc)
        return res

    def halving_loop(self, t_end: float):
        """README.md:66-80 loop up to t_end with velocity halving on Collide, run natively in the library; (events, collides)."""
        ne, nc = C.c_uint64(), C.c_uint64()
        self._check(self._L.cb200_collider_run_halving_loop(self._h, t_end, C.byref(ne), C.byref(nc)))
        return ne.value, nc.value

    def stage_velocities(self, vx=None, vy=None):
        """Start the upload of the NEXT bulk_update's velocities on the engine's upload stream (cb200_stage_velocities on the
        collider's context); the next bulk_update() without velocity arguments consumes them.  Arrays: ascending-id order, pinned."""
        n = self.len()
        self._staged_keep = [None if a is None else _col(a, np.float64, n) for a in (vx, vy, None, None, None)]
        v = VelView(*[None if a is None else a.ctypes.data for a in self._staged_keep])
        rc = self._L.cb200_stage_velocities(self._L.cb200_collider_engine(self._h), C.byref(v))
        if rc != 0:
            raise ColliderError(rc, self._L.cb200_last_error(self._L.cb200_collider_engine(self._h)).decode())

    def set_rebin_threshold(self, rows: int):
        self._check(self._L.cb200_collider_set_rebin_threshold(self._h, rows))

    def rebin_count(self) -> int:
        return int(self._L.cb200_collider_rebin_count(self._h))

    def hitbox_cache_hits(self) -> int:
        return int(self._L.cb200_collider_hitbox_cache_hits(self._h))

    def prefetch_count(self) -> int:
        return int(self._L.cb200_collider_prefetch_count(self._h))

    def launch_count(self) -> int:
        return int(self._L.cb200_launch_count(self._L.cb200_collider_engine(self._h)))
