"""collider-rs_b200 — B200-native engine for collider-rs's data-parallel hot path.

This package is the Python-side loader/binding of ``lib/libcollider_b200.so`` (the C ABI declared in
``include/collider_b200.h``).  It holds no algorithm of its own: every call below forwards to the
CUDA library, and importing the binding fails loudly if the library has not been built
(``python collider-rs_b200/build.py``) — there is no CPU fallback.

Import with ``importlib.import_module("collider-rs_b200")`` (the directory name carries a hyphen).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CB200_LIB") or os.path.join(_HERE, "lib", "libcollider_b200.so")

KIND_CIRCLE, KIND_RECT = 0, 1
GROUP_NONE = 0xFFFFFFFF
EV_REITERATE, EV_COLLIDE, EV_SEPARATE = 0, 1, 2
E_ARG, E_CUDA, E_CONTRACT, E_STATE = -1, -2, -3, -4
N_STAGES = 8
STAGE_NAMES = ("pre", "stage_a", "exchange", "scan", "scatter", "candidates", "solve", "tail")
INF = math.inf


class ColliderError(RuntimeError):
    """A failed C-ABI call; ``code`` is the CB200_E_* value.  E_CONTRACT restates a reference panic."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


class StateView(C.Structure):
    _fields_ = [("n", C.c_uint64), ("id", C.c_void_p)] + [
        (k, C.c_void_p)
        for k in ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "start_time", "end_time", "pub_end_time",
                  "kind", "group", "interact_mask")
    ]


class StateOut(C.Structure):
    _fields_ = [("n", C.c_uint64)] + [
        (k, C.c_void_p) for k in ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "start_time", "end_time", "pub_end_time")
    ]


ROW_F64 = ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "start_time", "end_time", "pub_end_time")


class RowsOut(C.Structure):
    """cb200_rows_out"""
    _fields_ = [("cap", C.c_uint64), ("id", C.c_void_p)] + [(k, C.c_void_p) for k in ROW_F64] + [
        ("kind", C.c_void_p), ("group", C.c_void_p), ("interact_mask", C.c_void_p), ("dest", C.c_void_p)]


class VelView(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("vel_x", "vel_y", "rsz_x", "rsz_y", "end_time")]


class Event(C.Structure):
    _fields_ = [("time", C.c_double), ("a", C.c_uint64), ("b", C.c_uint64), ("kind", C.c_uint32), ("_pad", C.c_uint32)]


EVENT16_DTYPE = np.dtype([("time", "<f8"), ("a", "<u4"), ("b", "<u4")])
EVENT_DTYPE = np.dtype([("time", "<f8"), ("a", "<u8"), ("b", "<u8"), ("kind", "<u4"), ("_pad", "<u4")])


class StepResult(C.Structure):
    _fields_ = [
        ("next_time", C.c_double),
        ("next_event", Event),
        ("n_hitboxes", C.c_uint64),
        ("n_entries", C.c_uint64),
        ("n_cells", C.c_uint64),
        ("n_collide_solves", C.c_uint64),
        ("n_separate_solves", C.c_uint64),
        ("n_events", C.c_uint64),
        ("n_solitaire", C.c_uint64),
        ("error_flags", C.c_uint32),
        ("_pad", C.c_uint32),
        ("error_slot", C.c_uint64),
        ("n_work_items", C.c_uint64),
    ]


class P2PInfo(C.Structure):
    _fields_ = [("rec_handle", C.c_uint8 * 64), ("mail_handle", C.c_uint8 * 64), ("rec_ptr", C.c_uint64), ("mail_ptr", C.c_uint64),
                ("recv_offset_bytes", C.c_uint64), ("pid", C.c_uint64), ("device", C.c_int32), ("slab_records", C.c_uint32)]


EXPORTS = (
    "cb200_abi_version", "cb200_create", "cb200_destroy", "cb200_last_error", "cb200_set_grid", "cb200_get_grid",
    "cb200_host_alloc", "cb200_host_free", "cb200_upload_state", "cb200_download_state", "cb200_set_overlaps",
    "cb200_bulk_update", "cb200_fetch_events", "cb200_fetch_candidate_pairs", "cb200_fetch_cell_entries",
    "cb200_fetch_index_rects", "cb200_collide_time_batch", "cb200_separate_time_batch", "cb200_launch_count",
    "cb200_last_step_timing", "cb200_reset_timing", "cb200_shard_configure", "cb200_shard_set_slab", "cb200_set_stream", "cb200_shard_begin", "cb200_shard_finish",
    "cb200_shard_result", "cb200_set_resort_period", "cb200_resort_count", "cb200_event_sort_fallbacks", "cb200_set_stage_timing",
    "cb200_shard_p2p_export", "cb200_shard_p2p_import", "cb200_shard_step", "cb200_shard_global_next", "cb200_shard_local_key", "cb200_shard_abort", "cb200_shard_set_row_capacity", "cb200_shard_rows", "cb200_shard_take_movers", "cb200_shard_add_rows", "cb200_stage_velocities", "cb200_fetch_events_begin", "cb200_fetch_events_end", "cb200_set_window_drain", "cb200_fetch_overlaps", "cb200_last_drain", "cb200_last_step_path", "cb200_tile_fallbacks", "cb200_get_tile",
    "cb200_collider_new", "cb200_collider_free", "cb200_collider_last_error", "cb200_collider_set_can_interact", "cb200_collider_time",
    "cb200_collider_next_time", "cb200_collider_set_time", "cb200_collider_next", "cb200_collider_get_hitbox", "cb200_collider_add_hitbox",
    "cb200_collider_set_hitbox_vel", "cb200_collider_remove_hitbox", "cb200_collider_get_overlaps", "cb200_collider_is_overlapping",
    "cb200_collider_query_overlaps", "cb200_collider_bulk_update", "cb200_collider_len", "cb200_collider_set_rebin_threshold",
    "cb200_collider_rebin_count", "cb200_collider_engine", "cb200_collider_add_hitboxes", "cb200_collider_prefetch_count",
    "cb200_collider_hitbox_cache_hits", "cb200_collider_run_halving_loop", "cb200_collider_query_overlaps_batch", "cb200_collider_group_index",
)

_lib = None


def load_library() -> C.CDLL:
    """dlopen the CUDA library.  Raises if it is missing — the product has no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run `python collider-rs_b200/build.py` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, dbl, i32 = C.c_void_p, C.c_uint64, C.c_double, C.c_int
    L.cb200_abi_version.restype = i32
    L.cb200_create.argtypes = [dbl, dbl, i32, C.POINTER(vp)]
    L.cb200_destroy.argtypes = [vp]
    L.cb200_destroy.restype = None
    L.cb200_last_error.argtypes = [vp]
    L.cb200_last_error.restype = C.c_char_p
    L.cb200_set_grid.argtypes = [vp, i32, i32]
    L.cb200_get_grid.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.cb200_host_alloc.argtypes = [u64]
    L.cb200_host_alloc.restype = vp
    L.cb200_host_free.argtypes = [vp]
    L.cb200_host_free.restype = None
    L.cb200_upload_state.argtypes = [vp, C.POINTER(StateView)]
    L.cb200_download_state.argtypes = [vp, C.POINTER(StateOut)]
    L.cb200_set_overlaps.argtypes = [vp, vp, vp, u64]
    L.cb200_bulk_update.argtypes = [vp, dbl, C.POINTER(VelView), dbl, C.POINTER(StepResult)]
    L.cb200_fetch_events.argtypes = [vp, vp, u64, u64, C.POINTER(u64)]
    L.cb200_fetch_candidate_pairs.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    L.cb200_fetch_cell_entries.argtypes = [vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.cb200_fetch_index_rects.argtypes = [vp, vp, u64]
    L.cb200_collide_time_batch.argtypes = [vp, u64, vp, vp, vp, vp, vp]
    L.cb200_separate_time_batch.argtypes = [vp, u64, vp, vp, vp, vp, dbl, vp]
    L.cb200_launch_count.argtypes = [vp]
    L.cb200_launch_count.restype = u64
    L.cb200_last_step_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), i32]
    L.cb200_reset_timing.argtypes = [vp]
    L.cb200_shard_configure.argtypes = [vp, i32, i32, vp, u64]
    L.cb200_set_stream.argtypes = [vp, vp]
    L.cb200_shard_set_slab.argtypes = [vp, u64]
    L.cb200_shard_begin.argtypes = [vp, dbl, C.POINTER(VelView), dbl, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    L.cb200_shard_finish.argtypes = [vp, C.POINTER(vp)]
    L.cb200_shard_result.argtypes = [vp, C.POINTER(StepResult)]
    L.cb200_set_stage_timing.argtypes = [vp, i32]
    L.cb200_shard_p2p_export.argtypes = [vp, C.POINTER(P2PInfo)]
    L.cb200_shard_p2p_import.argtypes = [vp, C.POINTER(P2PInfo)]
    L.cb200_shard_step.argtypes = [vp, dbl, C.POINTER(VelView), dbl]
    L.cb200_shard_global_next.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]
    L.cb200_shard_local_key.argtypes = [vp]
    L.cb200_shard_local_key.restype = vp
    L.cb200_shard_abort.argtypes = [vp]
    L.cb200_shard_set_row_capacity.argtypes = [vp, u64]
    L.cb200_shard_rows.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.cb200_shard_take_movers.argtypes = [vp, i32, C.POINTER(RowsOut), C.POINTER(u64)]
    L.cb200_shard_add_rows.argtypes = [vp, C.POINTER(StateView)]
    L.cb200_last_step_path.argtypes = [vp]
    L.cb200_stage_velocities.argtypes = [vp, C.POINTER(VelView)]
    L.cb200_fetch_events_begin.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.cb200_fetch_events_end.argtypes = [vp]
    L.cb200_set_window_drain.argtypes = [vp, i32]
    L.cb200_fetch_overlaps.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    L.cb200_last_drain.argtypes = [vp] + [C.POINTER(u64)] * 5
    L.cb200_tile_fallbacks.argtypes = [vp]
    L.cb200_tile_fallbacks.restype = u64
    L.cb200_get_tile.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.cb200_set_resort_period.argtypes = [vp, i32]
    L.cb200_resort_count.argtypes = [vp]
    L.cb200_resort_count.restype = u64
    L.cb200_event_sort_fallbacks.argtypes = [vp]
    L.cb200_event_sort_fallbacks.restype = u64
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _col(a, dtype, n):
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.shape != (n,):
        raise ValueError(f"column has shape {a.shape}, expected ({n},)")
    return a


class PinnedArray:
    """numpy view over page-locked host memory from cb200_host_alloc."""

    def __init__(self, shape, dtype):
        self._L = load_library()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        self._p = self._L.cb200_host_alloc(max(1, n * dtype.itemsize))
        if not self._p:
            raise MemoryError("cb200_host_alloc failed")
        buf = (C.c_char * (n * dtype.itemsize)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        if getattr(self, "_p", None):
            self.array = None
            self._L.cb200_host_free(self._p)
            self._p = None


class Engine:
    """One device context: the hitbox table, the cell table and the event buffer of one `Collider`."""

    def __init__(self, cell_width: float, padding: float, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.cb200_create(cell_width, padding, device, C.byref(h))
        if rc != 0:
            raise ColliderError(rc, self._L.cb200_last_error(None).decode())
        self._h = h
        self.cell_width, self.padding, self.device = cell_width, padding, device
        self.n = 0
        self._keep = None
        self._ev_pin = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.cb200_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise ColliderError(rc, self._L.cb200_last_error(self._h).decode())

    def set_grid(self, log2_w: int, log2_h: int):
        self._check(self._L.cb200_set_grid(self._h, log2_w, log2_h))

    def get_grid(self):
        a, b = C.c_int(), C.c_int()
        self._check(self._L.cb200_get_grid(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def upload_state(self, *, id, pos_x, pos_y, dim_x, dim_y, vel_x, vel_y, kind, rsz_x=None, rsz_y=None, start_time=None, end_time=None,
                     pub_end_time=None, group=None, interact_mask=None):
        n = len(id)
        z = np.zeros(n)
        cols = dict(
            id=_col(id, np.uint64, n),
            pos_x=_col(pos_x, np.float64, n), pos_y=_col(pos_y, np.float64, n),
            dim_x=_col(dim_x, np.float64, n), dim_y=_col(dim_y, np.float64, n),
            vel_x=_col(vel_x, np.float64, n), vel_y=_col(vel_y, np.float64, n),
            rsz_x=_col(z if rsz_x is None else rsz_x, np.float64, n), rsz_y=_col(z if rsz_y is None else rsz_y, np.float64, n),
            start_time=_col(z if start_time is None else start_time, np.float64, n),
            end_time=_col(np.full(n, INF) if end_time is None else end_time, np.float64, n),
            pub_end_time=_col(np.full(n, INF) if pub_end_time is None else pub_end_time, np.float64, n),
            kind=_col(kind, np.uint8, n),
            group=_col(np.zeros(n, np.uint32) if group is None else group, np.uint32, n),
            interact_mask=_col(np.ones(n, np.uint32) if interact_mask is None else interact_mask, np.uint32, n),
        )
        v = StateView()
        v.n = n
        for k, a in cols.items():
            setattr(v, k, a.ctypes.data)
        self._check(self._L.cb200_upload_state(self._h, C.byref(v)))
        self.n = n

    def download_state(self):
        n = self.n
        names = ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "start_time", "end_time", "pub_end_time")
        out = {k: np.zeros(n) for k in names}
        o = StateOut()
        o.n = n
        for k in names:
            setattr(o, k, out[k].ctypes.data)
        self._check(self._L.cb200_download_state(self._h, C.byref(o)))
        return out

    def set_overlaps(self, id_a, id_b):
        a = np.ascontiguousarray(id_a, dtype=np.uint64)
        b = np.ascontiguousarray(id_b, dtype=np.uint64)
        if a.shape != b.shape:
            raise ValueError("id_a / id_b length mismatch")
        self._check(self._L.cb200_set_overlaps(self._h, _ptr(a), _ptr(b), a.size))

    def bulk_update(self, time: float, end_time: float = INF, vel_x=None, vel_y=None, rsz_x=None, rsz_y=None, end_time_arr=None) -> StepResult:
        """cb200_bulk_update.  Arrays are host arrays of length n (numpy, or PinnedArray.array)."""
        arrs = [vel_x, vel_y, rsz_x, rsz_y, end_time_arr]
        res = StepResult()
        if all(a is None for a in arrs):
            rc = self._L.cb200_bulk_update(self._h, time, None, end_time, C.byref(res))
        else:
            keep = [None if a is None else _col(a, np.float64, self.n) for a in arrs]
            v = VelView(*[None if a is None else a.ctypes.data for a in keep])
            rc = self._L.cb200_bulk_update(self._h, time, C.byref(v), end_time, C.byref(res))
        self.last_result = res
        self._check(rc)
        return res

    # ---- sharded step (one context per GPU; the exchange between begin and finish is the caller's, see sharding.py) ----
    def shard_configure(self, rank: int, world: int, col_bounds, id_space: int):
        b = np.ascontiguousarray(col_bounds, dtype=np.int32)
        if b.shape != (world + 1,):
            raise ValueError("col_bounds must have world + 1 entries")
        self._check(self._L.cb200_shard_configure(self._h, rank, world, _ptr(b), id_space))
        self.rank, self.world = rank, world

    def shard_set_slab(self, records: int):
        self._check(self._L.cb200_shard_set_slab(self._h, records))

    def shard_set_row_capacity(self, rows: int):
        self._check(self._L.cb200_shard_set_row_capacity(self._h, rows))

    def shard_rows(self) -> dict:
        v = [C.c_uint64(0) for _ in range(4)]
        self._check(self._L.cb200_shard_rows(self._h, *[C.byref(x) for x in v]))
        return dict(rows=v[0].value, capacity=v[1].value, migrated_out=v[2].value, migrated_in=v[3].value)

    def shard_take_movers(self, margin_cols: int = 1):
        """Owner migration, sending side: (rows, dest) — `rows` a dict of columns (id, the 11 state columns, kind, group,
        interact_mask) of the rows that left this rank's strip, dest[i] the rank that should take row i.  They are removed here."""
        need = C.c_uint64(0)
        rc = self._L.cb200_shard_take_movers(self._h, margin_cols, None, C.byref(need))
        m = need.value
        rows = {"id": np.zeros(m, np.uint64), **{k: np.zeros(m) for k in ROW_F64}, "kind": np.zeros(m, np.uint8),
                "group": np.zeros(m, np.uint32), "interact_mask": np.zeros(m, np.uint32)}
        dest = np.zeros(m, np.int32)
        if m == 0:
            self._check(rc)
            return rows, dest
        o = RowsOut()
        o.cap = m
        for k, a in rows.items():
            setattr(o, k, a.ctypes.data)
        o.dest = dest.ctypes.data
        self._check(self._L.cb200_shard_take_movers(self._h, margin_cols, C.byref(o), C.byref(need)))
        assert need.value == m
        self.n -= m
        return rows, dest

    def shard_add_rows(self, rows: dict):
        """Owner migration, receiving side: append rows (a dict like shard_take_movers returns; any order)."""
        m = len(rows["id"])
        if m == 0:
            return 0
        o = np.argsort(np.asarray(rows["id"], dtype=np.uint64), kind="stable")
        cols = {"id": _col(np.asarray(rows["id"])[o], np.uint64, m), **{k: _col(np.asarray(rows[k])[o], np.float64, m) for k in ROW_F64},
                "kind": _col(np.asarray(rows["kind"])[o], np.uint8, m), "group": _col(np.asarray(rows["group"])[o], np.uint32, m),
                "interact_mask": _col(np.asarray(rows["interact_mask"])[o], np.uint32, m)}
        v = StateView()
        v.n = m
        for k, a in cols.items():
            setattr(v, k, a.ctypes.data)
        self._check(self._L.cb200_shard_add_rows(self._h, C.byref(v)))
        self.n += m
        return m

    def set_stream(self, cuda_stream: int):
        self._check(self._L.cb200_set_stream(self._h, C.c_void_p(cuda_stream)))

    def shard_begin(self, time: float, end_time: float = INF, vel_x=None, vel_y=None, rsz_x=None, rsz_y=None, end_time_arr=None) -> dict:
        """Asynchronous front half.  Returns the exchange buffers: the slab at send_ptr is all-gathered into slot `rank` of every rank's recv_ptr."""
        arrs = [vel_x, vel_y, rsz_x, rsz_y, end_time_arr]
        per_peer = C.c_uint64()
        send, recv = C.c_void_p(), C.c_void_p()
        v = None
        if not all(a is None for a in arrs):
            self._keep = [None if a is None else _col(a, np.float64, self.n) for a in arrs]  # alive until the step completes
            v = C.byref(VelView(*[None if a is None else a.ctypes.data for a in self._keep]))
        self._check(self._L.cb200_shard_begin(self._h, time, v, end_time, C.byref(send), C.byref(recv), C.byref(per_peer)))
        return dict(send_ptr=send.value, recv_ptr=recv.value, slab_bytes=per_peer.value)

    # ---- direct exchange over peer memory (cb200_shard_p2p_*) ----
    def p2p_export(self) -> bytes:
        info = P2PInfo()
        self._check(self._L.cb200_shard_p2p_export(self._h, C.byref(info)))
        return bytes(info)

    def p2p_import(self, infos):
        """infos: one exported blob per rank, indexed by rank."""
        arr = (P2PInfo * len(infos))(*[P2PInfo.from_buffer_copy(b) for b in infos])
        self._check(self._L.cb200_shard_p2p_import(self._h, arr))

    def shard_step(self, time: float, end_time: float = INF, vel_x=None, vel_y=None, rsz_x=None, rsz_y=None, end_time_arr=None):
        """The whole sharded step with the direct exchange, asynchronous; shard_result() waits for it."""
        arrs = [vel_x, vel_y, rsz_x, rsz_y, end_time_arr]
        v = None
        if not all(a is None for a in arrs):
            self._keep = [None if a is None else _col(a, np.float64, self.n) for a in arrs]
            v = C.byref(VelView(*[None if a is None else a.ctypes.data for a in self._keep]))
        self._check(self._L.cb200_shard_step(self._h, time, v, end_time))

    def shard_global_next(self):
        """(time key, tie, complete) of the minimum over all ranks after a direct-exchange step."""
        t, tie, ok = C.c_uint64(), C.c_uint64(), C.c_int()
        self._check(self._L.cb200_shard_global_next(self._h, C.byref(t), C.byref(tie), C.byref(ok)))
        return t.value, tie.value, bool(ok.value)

    def shard_abort(self):
        self._L.cb200_shard_abort(self._h)

    def shard_local_key_ptr(self) -> int:
        return int(self._L.cb200_shard_local_key(self._h))

    def shard_finish(self) -> int:
        """Asynchronous back half; returns the device address of this rank's (time key, tie) minimum (2 x u64)."""
        key = C.c_void_p()
        self._check(self._L.cb200_shard_finish(self._h, C.byref(key)))
        return key.value

    def shard_result(self) -> StepResult:
        res = StepResult()
        rc = self._L.cb200_shard_result(self._h, C.byref(res))
        self.last_result = res
        self._check(rc)
        return res

    def fetch_events(self, offset: int = 0, cap: int | None = None, pinned: bool = False) -> np.ndarray:
        """Sorted events of the last step.  pinned=True copies into a reusable page-locked buffer (full PCIe speed)
        and returns a view that stays valid until the next pinned fetch."""
        if cap is None:
            cap = max(0, int(self.last_result.n_events) - offset)
        if pinned:
            if self._ev_pin is None or self._ev_pin.array.shape[0] < cap:
                self._ev_pin = PinnedArray(max(cap + cap // 4, 1024), EVENT_DTYPE)
            buf = self._ev_pin.array
        else:
            buf = np.zeros(cap, dtype=EVENT_DTYPE)
        got = C.c_uint64()
        self._check(self._L.cb200_fetch_events(self._h, _ptr(buf), cap, offset, C.byref(got)))
        return buf[: got.value]

    # ---- pipelined I/O (cb200_stage_velocities / cb200_fetch_events_begin / _end) ----
    def stage_velocities(self, vel_x=None, vel_y=None, rsz_x=None, rsz_y=None, end_time_arr=None):
        """Start the H2D copy of the NEXT step's velocities (pinned arrays: PinnedArray.array); the next bulk_update() without
        velocity arguments consumes them."""
        arrs = [vel_x, vel_y, rsz_x, rsz_y, end_time_arr]
        self._staged_keep = [None if a is None else _col(a, np.float64, self.n) for a in arrs]
        v = VelView(*[None if a is None else a.ctypes.data for a in self._staged_keep])
        self._check(self._L.cb200_stage_velocities(self._h, C.byref(v)))

    def fetch_events_begin(self, buf: np.ndarray) -> int:
        """Queue the sort + download of the last step's events as EVENT16_DTYPE records into `buf` (ideally pinned); returns the
        number of events that will be there once fetch_events_end() has returned."""
        got = C.c_uint64()
        self._check(self._L.cb200_fetch_events_begin(self._h, _ptr(buf), buf.shape[0], C.byref(got)))
        return int(got.value)

    def fetch_events_end(self):
        self._check(self._L.cb200_fetch_events_end(self._h))

    @staticmethod
    def decode_events16(ev16: np.ndarray, ids: np.ndarray) -> np.ndarray:
        """EVENT16_DTYPE records -> EVENT_DTYPE records (ids looked up from the labels; `ids` = the uploaded ascending id array)."""
        out = np.zeros(len(ev16), dtype=EVENT_DTYPE)
        reit = (ev16["a"] >> np.uint32(31)).astype(bool)
        collide = (ev16["b"] >> np.uint32(31)).astype(bool)
        la, lb = ev16["a"] & np.uint32(0x7FFFFFFF), ev16["b"] & np.uint32(0x7FFFFFFF)
        out["time"] = ev16["time"]
        out["a"] = ids[la]
        out["b"] = np.where(reit, 0, ids[np.where(reit, 0, lb)])
        out["kind"] = np.where(reit, EV_REITERATE, np.where(collide, EV_COLLIDE, EV_SEPARATE))
        return out

    def set_resort_period(self, steps: int):
        self._check(self._L.cb200_set_resort_period(self._h, steps))

    def resort_count(self) -> int:
        return int(self._L.cb200_resort_count(self._h))

    def event_sort_fallbacks(self) -> int:
        return int(self._L.cb200_event_sort_fallbacks(self._h))

    def fetch_candidate_pairs(self):
        n = C.c_uint64()
        self._check(self._L.cb200_fetch_candidate_pairs(self._h, None, None, 0, C.byref(n)))
        hi = np.zeros(n.value, dtype=np.uint64)
        lo = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self._L.cb200_fetch_candidate_pairs(self._h, _ptr(hi), _ptr(lo), n.value, C.byref(n)))
        return hi, lo

    def fetch_cell_entries(self):
        n = C.c_uint64()
        self._check(self._L.cb200_fetch_cell_entries(self._h, None, None, None, 0, C.byref(n)))
        x = np.zeros(n.value, dtype=np.int32)
        y = np.zeros(n.value, dtype=np.int32)
        i = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self._L.cb200_fetch_cell_entries(self._h, _ptr(x), _ptr(y), _ptr(i), n.value, C.byref(n)))
        return x, y, i

    def fetch_index_rects(self) -> np.ndarray:
        r = np.zeros((self.n, 4), dtype=np.int32)
        self._check(self._L.cb200_fetch_index_rects(self._h, _ptr(r), self.n))
        return r

    def _batch(self, fn, a, ka, b, kb, *extra):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 9)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 9)
        ka = np.ascontiguousarray(ka, dtype=np.uint8).reshape(-1)
        kb = np.ascontiguousarray(kb, dtype=np.uint8).reshape(-1)
        out = np.zeros(a.shape[0])
        self._check(fn(self._h, a.shape[0], _ptr(a), _ptr(ka), _ptr(b), _ptr(kb), *extra, _ptr(out)))
        return out

    def collide_time_batch(self, a, ka, b, kb):
        return self._batch(self._L.cb200_collide_time_batch, a, ka, b, kb)

    def separate_time_batch(self, a, ka, b, kb, padding):
        return self._batch(self._L.cb200_separate_time_batch, a, ka, b, kb, C.c_double(padding))

    def launch_count(self) -> int:
        return int(self._L.cb200_launch_count(self._h))

    def set_window_drain(self, on: bool = True):
        """Process the pair events inside every step's own window on the device and keep the overlap set there (cb200_set_window_drain)."""
        self._check(self._L.cb200_set_window_drain(self._h, 1 if on else 0))

    def fetch_overlaps(self):
        n = C.c_uint64()
        self._check(self._L.cb200_fetch_overlaps(self._h, None, None, 0, C.byref(n)))
        a = np.zeros(n.value, dtype=np.uint64)
        b = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self._L.cb200_fetch_overlaps(self._h, _ptr(a), _ptr(b), n.value, C.byref(n)))
        return a, b

    def last_drain(self) -> dict:
        v = [C.c_uint64() for _ in range(5)]
        self._check(self._L.cb200_last_drain(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("followup_events", "pairs_added", "pairs_removed", "pairs_now", "reiterates_in_windows"), [x.value for x in v]))

    def last_step_path(self) -> str:
        return {1: "tile", 0: "global"}.get(self._L.cb200_last_step_path(self._h), "none")

    def tile_fallbacks(self) -> int:
        return int(self._L.cb200_tile_fallbacks(self._h))

    def get_tile(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self._L.cb200_get_tile(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_stage_timing(self, on: bool):
        self._check(self._L.cb200_set_stage_timing(self._h, 1 if on else 0))

    def reset_timing(self):
        self._check(self._L.cb200_reset_timing(self._h))

    def step_timing(self):
        total = C.c_float()
        stages = (C.c_float * N_STAGES)()
        self._check(self._L.cb200_last_step_timing(self._h, C.byref(total), stages, N_STAGES))
        return total.value, dict(zip(STAGE_NAMES, [float(s) for s in stages]))
