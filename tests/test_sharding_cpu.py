"""Host-side sharding logic over torch.distributed with the gloo backend, world_size 2, on CPU tensors:
strip partition, count/record exchange, global event minimum (the N>1 plumbing of SURVEY.md 8e)."""
import importlib
import os
import socket

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


def test_event_key_order():
    k = sharding.event_key
    assert k(1.0, 0, 7, 0) < k(1.0, 2, 3, 1) < k(1.0, 1, 3, 1) < k(1.0, 2, 4, 0) < k(1.5, 0, 0, 0) < k(float("inf"), 0, 0, 0)


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
        dev = torch.device("cpu")
        RB = sharding.RECORD_BYTES
        stride = 8 * RB
        # rank r sends (r + 1) * (d + 1) records to rank d, each filled with a byte tag
        counts = [0 if d == rank else (rank + 1) * (d + 1) for d in range(world)]
        send = torch.zeros(world * stride, dtype=torch.uint8)
        for d in range(world):
            send[d * stride:d * stride + counts[d] * RB] = 16 * rank + d
        recv_counts = ex.exchange_counts(counts, dev)
        assert recv_counts == [0 if d == rank else (d + 1) * (rank + 1) for d in range(world)]
        recv = torch.zeros(sum(recv_counts) * RB, dtype=torch.uint8)
        got = ex.exchange_records(send, stride, counts, recv, recv_counts)
        assert got == sum(recv_counts)
        off = 0
        for d in range(world):
            nb = recv_counts[d] * RB
            assert bool((recv[off:off + nb] == 16 * d + rank).all())
            off += nb
        # global minimum of the per-rank next events; same answer on every rank, ties broken by the key
        mine = sharding.event_key(2.5, 1, 10 + rank, 3) if rank else sharding.event_key(2.5, 2, 11, 4)
        best = ex.global_min(mine, dev)
        assert best == (2.5, 1, 11, 0, 4)
        best = ex.global_min(sharding.event_key(float("inf"), 0, 0, 0) if rank else sharding.event_key(7.25, 0, 5, 0), dev)
        assert best == (7.25, 0, 5, 0, 0)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
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
