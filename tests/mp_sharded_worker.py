"""One rank of the multi-PROCESS sharded parity check (launched by tests/test_sharded_multiprocess_gpu.py through
torch.distributed.run).  Every rank is its own process with its own CUDA context: the direct exchange goes through
cudaIpcOpenMemHandle and the flag words in the peers' mailboxes — the path bench.py --gpus N runs — and the union of the
per-rank results is compared with the oracle's single collider (event sequence, times bit-equal, candidate-pair set, global
next event).  With fewer GPUs than ranks all ranks share cuda:0 (cudaIpc works between processes on one device; their
kernels are time-sliced, so a waiting kernel of one rank is preempted for the peer it waits for).

  --exchange direct   cb200_shard_step (peer memory); host plumbing over gloo
  --exchange nccl     cb200_shard_begin / finish with all_gather_into_tensor over NCCL (needs one GPU per rank)
  --migrate           owner migration (ShardedEngine.migrate: take_movers -> all_gather_object -> add_rows) after every step,
                      with new fast velocities each step: rows cross strip edges and change rank between the steps
  --scenario abort    rank 1 stops after the first step and raises its peers' abort words: rank 0's next step must FAIL with
                      CB200_F_PEER_LOST instead of hanging
  --scenario timeout  the same without the abort call: rank 0's wait ends by the timeout (CB200_PEER_TIMEOUT_MS)
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exchange", default="direct", choices=["direct", "nccl"])
    ap.add_argument("--scenario", default="parity", choices=["parity", "abort", "timeout"])
    ap.add_argument("--hitboxes", dest="n", type=int, default=6000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--scene", default="uniform", choices=["uniform", "dense", "clustered"])
    ap.add_argument("--migrate", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    device = local if torch.cuda.device_count() >= world else 0
    torch.cuda.set_device(device)
    if args.exchange == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo")

    cb = importlib.import_module("collider-rs_b200")
    scenes = importlib.import_module("collider-rs_b200.scenes")
    sharding = importlib.import_module("collider-rs_b200.sharding")
    from oracle import oracle_py as oracle
    from tiled_oracle import canonical_order

    CW, PAD = scenes.CELL_WIDTH, scenes.PADDING
    scene = {"uniform": lambda: scenes.uniform_density(args.n), "dense": lambda: scenes.uniform_density(args.n, density=1.0, seed=scenes.SEED_BASE + 41),
             "clustered": lambda: scenes.c5(args.n)}[args.scene]()
    n = len(scene["id"])
    orc = oracle.Collider(CW, PAD)
    orc.add_hitboxes(scene)
    st = orc.dump_state()
    bounds = sharding.strip_bounds(world, float(scene["pos_x"].min()), float(scene["pos_x"].max()) + 1e-9, CW)
    owner = np.clip(np.searchsorted(bounds[1:-1], np.floor(scene["pos_x"] / CW), side="right"), 0, world - 1)
    sh = sharding.ShardedEngine(cb, CW, PAD, device, bounds, n, sharding.Exchange(), direct=args.exchange == "direct")
    m = owner == rank
    if args.migrate:
        m = (st["id"] % np.uint64(world)) == np.uint64(rank)  # residents unrelated to position: the first migration re-homes most rows
        sh.eng.shard_set_row_capacity(n)
    sh.upload_state(**{k: v[m] for k, v in st.items()})
    my_ids = st["id"][m].astype(np.int64)
    moved = 0
    assert sh.direct == (args.exchange == "direct"), getattr(sh, "direct_error", "direct exchange was not established")

    T, dt = 0.0, 0.1
    for step in range(args.steps):
        ov_a, ov_b = orc.dump_overlaps()
        sh.set_overlaps(ov_a, ov_b)
        if args.scenario != "parity" and step == 1:
            if rank == 1:
                if args.scenario == "abort":
                    sh.eng.shard_abort()
                dist.barrier()
                break
            try:
                sh.bulk_update(T, end_time=T + dt)
            except cb.ColliderError as e:
                assert "peer rank did not deliver" in e.msg, e.msg
                print(f"MP_SHARDED_OK rank {rank}: step failed as it must ({e.msg})", flush=True)
                dist.barrier()
                break
            raise AssertionError("the step of a rank whose peer stopped must fail, not succeed")
        vel = {}
        if args.migrate:
            vx, vy = scenes.uniform(9000 + step, n, 1, -30.0, 30.0), scenes.uniform(9000 + step, n, 2, -30.0, 30.0)
            orc.bulk_update(end_time=T + dt, vx=vx, vy=vy, record_candidates=True)
            vel = dict(vel_x=vx[my_ids], vel_y=vy[my_ids])
        else:
            orc.bulk_update(end_time=T + dt, record_candidates=True)
        res = sh.bulk_update(T, end_time=T + dt, **vel)
        ev = sh.eng.fetch_events()
        hi, lo = sh.eng.fetch_candidate_pairs()
        parts = [None] * world
        dist.all_gather_object(parts, (ev, hi, lo, res.n_separate_solves, sh.next_event, sh.eng.event_sort_fallbacks()))
        t, idx, kind, a, b = orc.dump_events()
        # every rank holds the same global next event, reduced on the device from the peers' keys
        if len(t):
            assert sh.next_event == (t[0], int(kind[0]), int(a[0]), 0 if kind[0] == 0 else int(b[0])), (sh.next_event, t[0], kind[0], a[0], b[0])
        assert all(p[4] == sh.next_event for p in parts)
        if rank == 0:
            evs = np.concatenate([p[0] for p in parts])
            evs = evs[canonical_order(evs["time"], evs["kind"], evs["a"], evs["b"])]
            assert len(evs) == len(t), (len(evs), len(t))
            np.testing.assert_array_equal(evs["kind"], kind)
            np.testing.assert_array_equal(evs["a"], a)
            np.testing.assert_array_equal(evs["b"], np.where(kind == 0, 0, b))
            np.testing.assert_array_equal(bits(evs["time"]), bits(t))
            ohi, olo = orc.bulk_candidates()
            dkey = np.sort(np.concatenate([p[1] << np.uint64(32) | p[2] for p in parts]))
            np.testing.assert_array_equal(dkey, np.sort(ohi.astype(np.uint64) << np.uint64(32) | olo))  # disjoint union == the oracle's set
            assert sum(p[3] for p in parts) == len(ov_a)
            assert all(p[5] == 0 for p in parts), "the one-shot event sort must hold on every rank"
        if args.migrate:
            st_after = orc.dump_state()
            before = sh.eng.shard_rows()
            sent, got = sh.migrate(0)
            info = sh.eng.shard_rows()
            assert info["rows"] == before["rows"] - sent + got
            moved += sent
            col = np.floor(st_after["pos_x"] / CW)
            my_ids = np.nonzero((col >= float(bounds[rank])) & (col < float(bounds[rank + 1])))[0]  # residence follows space now
            assert info["rows"] == len(my_ids), (info, len(my_ids))
            mine = sh.eng.download_state()
            for k, colv in mine.items():
                np.testing.assert_array_equal(bits(colv), bits(st_after[k][my_ids]), err_msg=f"rank {rank} {k}")
        T += dt
        while True:  # drain the oracle's events up to the next step (no velocity changes)
            tt = min(orc.next_time(), T)
            orc.set_time(tt)
            while orc.next() is not None:
                pass
            if tt >= T:
                break
    else:
        assert not args.migrate or moved > 0
        print(f"MP_SHARDED_OK rank {rank}: {args.steps} steps, world {world}, exchange {args.exchange}, device {device}, rows migrated out {moved}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
