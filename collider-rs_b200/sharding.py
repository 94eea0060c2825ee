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


class DevMem:
    """A raw device allocation exposed through __cuda_array_interface__ so torch can wrap it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 3}


def wrap(ptr: int, nbytes: int, device):
    import torch

    return torch.as_tensor(DevMem(ptr, nbytes), device=device)[:nbytes]


class Exchange:
    """Record exchange + global min over a torch.distributed process group."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_counts(self, send_counts, device):
        import torch

        s = torch.tensor([int(c) for c in send_counts], dtype=torch.int64, device=device)
        r = torch.empty_like(s)
        self.dist.all_to_all_single(r, s, group=self.group) if device.type == "cuda" else self._a2a_cpu(r, s)
        return [int(x) for x in r.tolist()]

    def _a2a_cpu(self, r, s):
        # gloo has no all_to_all_single on every build: all_gather the count rows and pick our column
        import torch

        rows = [torch.empty_like(s) for _ in range(self.world)]
        self.dist.all_gather(rows, s, group=self.group)
        for d in range(self.world):
            r[d] = rows[d][self.rank]

    def exchange_records(self, send_buf, send_stride: int, send_counts, recv_buf, recv_counts) -> int:
        """send_buf: uint8 tensor [world * send_stride]; the first send_counts[d] records of slab d go to rank d.
        Received records are packed contiguously into recv_buf in rank order.  Returns the number received."""
        ops, off = [], 0
        for d in range(self.world):
            if d == self.rank:
                continue
            n_out = int(send_counts[d]) * RECORD_BYTES
            n_in = int(recv_counts[d]) * RECORD_BYTES
            if n_in:
                ops.append(self.dist.P2POp(self.dist.irecv, recv_buf[off:off + n_in], d, self.group))
                off += n_in
            if n_out:
                ops.append(self.dist.P2POp(self.dist.isend, send_buf[d * send_stride:d * send_stride + n_out], d, self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        return off // RECORD_BYTES

    def global_min(self, key, device):
        """key: tuple (time f64, class, hi, kind bit, lo).  Returns the minimum over all ranks (same on every rank)."""
        import torch

        bits = struct.unpack("<q", struct.pack("<d", float(key[0])))[0]
        mine = torch.tensor([bits, int(key[1]), int(key[2]), int(key[3]), int(key[4])], dtype=torch.int64, device=device)
        rows = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(rows, mine, group=self.group)
        best = None
        for r in rows:
            v = r.tolist()
            k = (struct.unpack("<d", struct.pack("<q", v[0]))[0], v[1], v[2], v[3], v[4])
            if best is None or k < best:
                best = k
        return best


class ShardedEngine:
    """One rank of a sharded `Collider`: a device context + the exchange around its two-phase step."""

    def __init__(self, cb, cell_width: float, padding: float, device_index: int, bounds, exchange: Exchange):
        import torch

        self.cb, self.ex = cb, exchange
        self.device = torch.device("cuda", device_index)
        self.eng = cb.Engine(cell_width, padding, device=device_index)
        self.eng.shard_configure(exchange.rank, exchange.world, bounds)
        self.h2d = self.d2h = 0

    def upload_state(self, **cols):
        self.eng.upload_state(**cols)

    def set_overlaps(self, a, b):
        self.eng.set_overlaps(a, b)

    def bulk_update(self, time: float, end_time: float, **vel):
        import torch

        st = self.eng.shard_begin(time, end_time=end_time, **vel)
        recv_counts = self.ex.exchange_counts(st["send_counts"], self.device)
        n_in = sum(recv_counts)
        if n_in > st["recv_cap"]:
            raise RuntimeError(f"rank {self.ex.rank}: {n_in} halo records exceed the visitor table ({st['recv_cap']})")
        send = wrap(st["send_ptr"], st["send_stride"] * self.ex.world, self.device)
        recv = wrap(st["recv_ptr"], max(n_in, 1) * RECORD_BYTES, self.device)
        got = self.ex.exchange_records(send, st["send_stride"], st["send_counts"], recv, recv_counts)
        torch.cuda.current_stream(self.device).synchronize()  # the engine runs on its own stream
        res = self.eng.shard_finish(got)
        ev = res.next_event
        self.local_result = res
        self.halo_sent = int(sum(int(c) for c in st["send_counts"]))
        self.halo_received = got
        self.next_key = self.ex.global_min(event_key(res.next_time, ev.kind, ev.a, ev.b), self.device)
        return res

    @property
    def next_time(self) -> float:
        return self.next_key[0]
