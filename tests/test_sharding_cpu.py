"""Host-side sharding logic over torch.distributed with the gloo backend, world_size 2, on CPU tensors:
strip partition, count/record exchange, global event minimum (the N>1 plumbing of SURVEY.md 8e)."""
import importlib
import os
import socket
import struct

import numpy as np
import pytest

sharding = importlib.import_module("collider-rs_b200.sharding")


def test_strip_bounds_and_destinations():
    b = sharding.strip_bounds(4, 0.0, 4096.0, 4.0)
    assert b[0] == sharding.INT32_MIN and b[-1] == sharding.INT32_MAX and list(b[1:-1]) == [256, 512, 768]
    assert sharding.strips_touched(b, 10, 12) == [0]
    assert sharding.strips_touched(b, 254, 256) == [0, 1]       # ex + 1 reaches column 256: one extra column of halo
    assert sharding.strips_touched(b, 255, 258) == [0, 1]
    assert sharding.strips_touched(b, 256, 257) == [1]           # records never travel left of their start column
    assert sharding.strips_touched(b, -5, -3) == [0] and sharding.strips_touched(b, 5000, 5001) == [3]
    with pytest.raises(ValueError):
        sharding.strip_bounds(64, 0.0, 8.0, 4.0)


def test_migration_routing():
    """ShardedEngine.migrate: of the (rows, dest) every rank took out of its table, rank r keeps the rows addressed to r."""
    def part(ids, dest):
        ids = np.asarray(ids, dtype=np.uint64)
        return ({"id": ids, "pos_x": ids.astype(np.float64) * 0.5, "kind": (ids % np.uint64(2)).astype(np.uint8)}, np.asarray(dest, dtype=np.int32))

    parts = [part([5, 9, 11], [1, 2, 1]), part([], []), part([3, 40], [0, 1])]
    got = sharding.rows_for(parts, 1)
    assert got["id"].tolist() == [5, 11, 40] and got["pos_x"].tolist() == [2.5, 5.5, 20.0] and got["kind"].dtype == np.uint8
    assert sharding.rows_for(parts, 0)["id"].tolist() == [3] and sharding.rows_for(parts, 2)["id"].tolist() == [9]
    assert len(sharding.rows_for([part([], []), part([], [])], 0)["id"]) == 0


def test_event_key_order():
    k = sharding.event_key
    assert k(1.0, 0, 7, 0) < k(1.0, 2, 3, 1) < k(1.0, 1, 3, 1) < k(1.0, 2, 4, 0) < k(1.5, 0, 0, 0) < k(float("inf"), 0, 0, 0)
    assert sharding.decode_key(sharding.KEY_MAX, sharding.KEY_MAX)[0] == float("inf")
    assert sharding.key_to_time((0x3FF0000000000000) | (1 << 63)) == 1.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = sharding.Exchange()
        per = 3 * sharding.RECORD_BYTES
        # rank r's halo slab for rank d is filled with the tag 16 * (r + 1) + d; after the exchange slot d of rank r must hold
        # what rank d addressed to r
        send = torch.cat([torch.full((per,), 16 * (rank + 1) + d, dtype=torch.uint8) for d in range(world)])
        recv = torch.zeros(world * per, dtype=torch.uint8)
        ex.exchange_slabs(recv, send)
        for d in range(world):
            assert bool((recv[d * per:(d + 1) * per] == 16 * (d + 1) + rank).all())
        # all-gather of the per-rank (time key, tie) minima + lexicographic min, in the device's u64 encoding
        def enc(t, tie):
            bits = struct.unpack("<Q", struct.pack("<d", t))[0]
            k = (~bits & sharding.KEY_MAX) if bits >> 63 else (bits | (1 << 63))
            to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v
            return torch.tensor([to_i64(k), to_i64(tie)], dtype=torch.int64)

        pair = lambda hi, collide, lo: sharding.PAIR_CLASS | (hi << 32) | (sharding.COLLIDE_BIT if collide else 0) | lo
        mine = enc(2.5, pair(10 + rank, True, 3)) if rank else enc(2.5, pair(11, False, 4))
        rows = ex.all_gather_keys(mine).tolist()
        assert sharding.decode_key(*sharding.min_key(rows)) == (2.5, 2, 11, 4)   # Separate(11, 4) < Collide(11, 3) at equal time
        mine = torch.tensor([-1, -1], dtype=torch.int64) if rank else enc(7.25, 5)  # rank 1: empty queue (all ones)
        assert sharding.decode_key(*sharding.min_key(ex.all_gather_keys(mine).tolist())) == (7.25, 0, 5, 0)
        mine = enc(-1.5, 9) if rank else enc(0.25, 1)                               # negative times order correctly
        assert sharding.decode_key(*sharding.min_key(ex.all_gather_keys(mine).tolist())) == (-1.5, 0, 9, 0)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_exchange_over_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
