"""Spatial sharding of the bulk step across the GPUs of one box: host-side plumbing.

One process per GPU, `torch.distributed` (backend "nccl" on the GPUs; the same code runs over "gloo" on CPU
tensors in tests/test_sharding_cpu.py).  This module owns only the plumbing — who sends how many 96-byte
records to whom, and the global minimum of the per-rank next events; every computation is in the CUDA library
(cb200_shard_begin / cb200_shard_finish).
"""
from __future__ import annotations

import math
import struct

import numpy as np

RECORD_BYTES = 96
INT32_MIN, INT32_MAX = -(2**31), 2**31 - 1


def strip_bounds(world: int, x_lo: float, x_hi: float, cell_width: float) -> np.ndarray:
    """Column bounds of `world` equal-width strips covering [x_lo, x_hi); the outer strips are open-ended."""
    c_lo, c_hi = math.floor(x_lo / cell_width), math.ceil(x_hi / cell_width)
    inner = [c_lo + round((c_hi - c_lo) * d / world) for d in range(1, world)]
    b = np.array([INT32_MIN] + inner + [INT32_MAX], dtype=np.int64)
    if world > 1 and not np.all(np.diff(b) > 0):
        raise ValueError("world too large for this extent: strips would be empty")
    return b.astype(np.int32)


def strips_touched(bounds: np.ndarray, sx: int, ex: int) -> list:
    """Ranks whose strip intersects the columns [sx, ex + 1) — the rule the device applies (kernels.cuh, stage A)."""
    return [d for d in range(len(bounds) - 1) if int(bounds[d + 1]) > sx and int(bounds[d]) < ex + 1]


def event_key(next_time: float, kind: int, a: int, b: int):
    """Total order of queued events across ranks: (time, class[solitaire < pair], hi, kind[Separate < Collide], lo)
    — events.rs:28-68 with the canonical tie-break of SURVEY.md 8c.  kind: 0 Reiterate, 1 Collide, 2 Separate."""
    if math.isinf(next_time) and next_time > 0:
        return (next_time, 2, 0, 0, 0)
    if kind == 0:
        return (next_time, 0, a, 0, 0)
    return (next_time, 1, a, 1 if kind == 1 else 0, b)


KEY_MAX = (1 << 64) - 1
PAIR_CLASS = 1 << 63
COLLIDE_BIT = 1 << 31


def key_to_time(tkey: int) -> float:
    """Inverse of the device's order-preserving u64 image of an f64 (common.cuh time_key)."""
    bits = (tkey & 0x7FFFFFFFFFFFFFFF) if (tkey >> 63) else (~tkey & KEY_MAX)
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def decode_key(tkey: int, tie: int):
    """(time, kind, a, b) of a device MinKey; kind None when the queue is empty."""
    if tkey == KEY_MAX and tie == KEY_MAX:
        return (math.inf, None, 0, 0)
    t = key_to_time(tkey)
    if tie & PAIR_CLASS:
        return (t, 1 if tie & COLLIDE_BIT else 2, (tie >> 32) & 0x7FFFFFFF, tie & 0x7FFFFFFF)
    return (t, 0, tie & 0xFFFFFFFF, 0)


def min_key(rows):
    """Lexicographic minimum of (time key, tie) rows given as signed int64 pairs (the all-gather output)."""
    return min(((int(a) & KEY_MAX), (int(b) & KEY_MAX)) for a, b in rows)


def rows_for(parts, rank: int) -> dict:
    """Of the (rows, dest) every rank took out of its table, the rows addressed to `rank`, as one dict of columns."""
    picks = [{k: v[dest == rank] for k, v in rows.items()} for rows, dest in parts]
    return {k: np.concatenate([p[k] for p in picks]) for k in picks[0]}


class DevMem:
    """A raw device allocation exposed through __cuda_array_interface__ so torch can wrap it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 3}


def wrap(ptr: int, nbytes: int, device):
    import torch

    return torch.as_tensor(DevMem(ptr, nbytes), device=device)[:nbytes]


class Exchange:
    """The two collectives of a sharded step over a torch.distributed process group (NCCL on GPUs, gloo in CPU tests):
    an all-to-all of the per-destination halo slabs and an all-gather of the 16-byte per-rank minimum keys."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_slabs(self, recv_buf, send_buf):
        """send_buf: uint8 tensor of `world` slabs, slab d addressed to rank d; recv_buf: world slabs.  Slot d of recv_buf
        receives the slab rank d addressed to this rank (the slot of the own rank is not meaningful)."""
        if send_buf.device.type == "cuda":
            self.dist.all_to_all_single(recv_buf, send_buf, group=self.group)
            return
        import torch

        per = send_buf.numel() // self.world
        rows = [torch.empty_like(send_buf) for _ in range(self.world)]
        self.dist.all_gather(rows, send_buf, group=self.group)  # (gloo: no all-to-all on every build; tests only)
        recv_buf.copy_(torch.cat([rows[d][self.rank * per:(self.rank + 1) * per] for d in range(self.world)]))

    def all_gather_keys(self, key_local):
        """key_local: int64 tensor [2] (time key, tie).  Returns an int64 tensor [world, 2]."""
        import torch

        out = torch.empty(self.world * 2, dtype=torch.int64, device=key_local.device)
        if key_local.device.type == "cuda":
            self.dist.all_gather_into_tensor(out, key_local, group=self.group)
        else:
            rows = [torch.empty_like(key_local) for _ in range(self.world)]
            self.dist.all_gather(rows, key_local, group=self.group)
            out = torch.cat(rows)
        return out.view(self.world, 2)


class ShardedEngine:
    """One rank of a sharded `Collider`: a device context on torch's current stream + the two collectives of its step.
    Per step the host enqueues: stage A + halo pack -> all-to-all of the halo slabs -> binning / candidates / solve / local min
    -> all-gather of the keys, and synchronises once at the end."""


    def __init__(self, cb, cell_width: float, padding: float, device_index: int, bounds, id_space: int, exchange: Exchange, direct: bool = True):
        import torch

        self.cb, self.ex = cb, exchange
        self.device = torch.device("cuda", device_index)
        self.eng = cb.Engine(cell_width, padding, device=device_index)
        self.eng.shard_configure(exchange.rank, exchange.world, bounds, id_space)
        self._bufs = None
        self.want_direct, self.direct = direct, False
        # The collective exchange is stream-ordered with the step on torch's current stream.  The direct exchange has no
        # collective in the step, so the engine keeps its own stream (which it can capture into a CUDA graph).
        if not direct:
            self.eng.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def upload_state(self, **cols):
        self.eng.upload_state(**cols)
        self._bufs = None
        self.direct = False
        if self.want_direct:
            # Direct exchange over peer memory: the ranks swap their cudaIpc handles once (host side, through the process
            # group) and from then on the step kernels write halo records and keys straight into the peers' memory.
            # If a peer cannot be opened (no peer access between two devices), every rank falls back to the collectives.
            infos = [None] * self.ex.world
            self.ex.dist.all_gather_object(infos, self.eng.p2p_export(), group=self.ex.group)
            ok = 1
            try:
                self.eng.p2p_import(infos)
            except self.cb.ColliderError as e:
                self.direct_error = str(e)
                ok = 0
            oks = [None] * self.ex.world
            self.ex.dist.all_gather_object(oks, ok, group=self.ex.group)
            self.direct = all(oks)
            if not self.direct:
                import torch

                self.eng.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def set_overlaps(self, a, b):
        self.eng.set_overlaps(a, b)

    def migrate(self, margin_cols: int = 1):
        """Owner migration, collective, between two steps: every rank hands the rows whose centre column left its
        strip (by more than margin_cols columns — hysteresis against ping-pong at an edge) to the rank whose strip holds them.
        Returns (rows sent, rows received).  The rows travel through the host (they are few: a row crosses an edge only after
        moving a whole strip margin); the step's halo exchange is untouched — the receive areas do not move."""
        rows, dest = self.eng.shard_take_movers(margin_cols)
        parts = [None] * self.ex.world
        self.ex.dist.all_gather_object(parts, (rows, dest), group=self.ex.group)
        return len(dest), self.eng.shard_add_rows(rows_for(parts, self.ex.rank))

    def bulk_update(self, time: float, end_time: float, **vel):
        import torch

        if self.direct:
            try:
                self.eng.shard_step(time, end_time=end_time, **vel)
                res = self.eng.shard_result()  # the one host synchronisation of the step
            except BaseException:
                self.eng.shard_abort()  # this rank stops stepping: its peers' waits end with CB200_F_PEER_LOST, not a hang
                raise
            tkey, tie, complete = self.eng.shard_global_next()
            if not complete:  # some rank re-ran the step with larger buffers: every rank sees this and repeats the key exchange
                keys = self.ex.all_gather_keys(wrap(self.eng.shard_local_key_ptr(), 16, self.device).view(torch.int64))
                tkey, tie = min_key(keys.tolist())
            self.local_result = res
            self.next_key = (tkey, tie)
            self.next_event = decode_key(tkey, tie)
            return res
        st = self.eng.shard_begin(time, end_time=end_time, **vel)
        key = (st["send_ptr"], st["recv_ptr"], st["slab_bytes"])
        if self._bufs is None or self._bufs[0] != key:
            self._bufs = (key, wrap(st["send_ptr"], st["slab_bytes"] * self.ex.world, self.device), wrap(st["recv_ptr"], st["slab_bytes"] * self.ex.world, self.device))
        _, send, recv = self._bufs
        self.ex.exchange_slabs(recv, send)
        key_ptr = self.eng.shard_finish()
        # shard_result may re-run the back half after growing a buffer that overflowed, and the re-run rewrites the local
        # key (the first run dropped candidate pairs): the keys are gathered only after it, from the final value
        res = self.eng.shard_result()
        keys = self.ex.all_gather_keys(wrap(key_ptr, 16, self.device).view(torch.int64))
        self.local_result = res
        self.next_key = min_key(keys.tolist())
        self.next_event = decode_key(*self.next_key)
        return res

    def halo_counts(self):
        """(records sent, records received) of the last step — reads the slab headers (D2H; diagnostics only)."""
        import torch

        if self._bufs is None:
            return 0, 0  # direct exchange: the slabs live in the engine
        _, send, recv = self._bufs
        per = send.numel()
        s = int(send[:4].view(torch.int32).item())
        r = sum(int(recv[d * per:d * per + 4].view(torch.int32).item()) for d in range(self.ex.world) if d != self.ex.rank)
        return s, r

    @property
    def next_time(self) -> float:
        return self.next_event[0]
