"""GPU parity of the spatially sharded step: R ranks emulated as R device contexts on ONE GPU, the
record exchange done with device-to-device copies (the real run uses NCCL via sharding.Exchange — same engine calls).
The union of the per-rank results must equal the oracle's single-collider result: event sequence, times (bit-equal),
candidate-pair set (each pair solved by exactly one rank), global next event."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")
sharding = importlib.import_module("collider-rs_b200.sharding")

CW, PAD = scenes.CELL_WIDTH, scenes.PADDING


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def exchange_on_one_gpu(engs, states):
    """The all-to-all of the per-destination halo slabs, done with device-to-device copies between the contexts of one GPU."""
    import torch

    dev = torch.device("cuda", 0)
    R = len(engs)
    torch.cuda.synchronize()  # every context runs on its own stream
    for r in range(R):
        per = states[r]["slab_bytes"]
        for d in range(R):
            src = sharding.wrap(states[d]["send_ptr"] + r * per, per, dev)  # the slab rank d addressed to rank r
            dst = sharding.wrap(states[r]["recv_ptr"] + d * per, per, dev)
            dst.copy_(src)
    torch.cuda.synchronize()


def key_of(ev):
    cls = np.where(ev["kind"] == 0, 0, 1)
    kbit = np.where(ev["kind"] == 1, 1, 0)
    return np.lexsort((ev["b"], kbit, ev["a"], cls, ev["time"]))


def run_sharded_case(oracle, scene, world, steps, dt, end_dt, resident_by="space", direct=False, migrate_margin=None, speed=None):
    """migrate_margin: after every step the rows that left their rank's strip (by more than that many columns) migrate to
    the rank that holds them (cb200_shard_take_movers / cb200_shard_add_rows), and every rank's table is compared with the
    oracle's.  speed: every step installs new random velocities in [-speed, speed]^2 (arrays in the ranks' ascending-id order)."""
    n = len(scene["id"])
    orc = oracle.Collider(CW, PAD)
    orc.add_hitboxes(scene)
    st = orc.dump_state()
    x_lo, x_hi = float(scene["pos_x"].min()), float(scene["pos_x"].max()) + 1e-9
    bounds = sharding.strip_bounds(world, x_lo, x_hi, CW)
    if resident_by == "space":
        owner = np.clip(np.searchsorted(bounds[1:-1], np.floor(scene["pos_x"] / CW), side="right"), 0, world - 1)
    else:  # residents unrelated to position: every record may have to travel
        owner = (scene["id"] % np.uint64(world)).astype(np.int64)
    engs, ids_of = [], []
    for r in range(world):
        e = cb.Engine(CW, PAD)
        e.shard_configure(r, world, bounds, n)
        if migrate_margin is not None:
            e.shard_set_row_capacity(n)
        m = owner == r
        e.upload_state(**{k: v[m] for k, v in st.items()})
        engs.append(e)
        ids_of.append(st["id"][m].astype(np.int64))
    moved = 0
    if direct:
        # direct exchange over peer memory: here the "peers" are contexts of this process, addressed by raw pointers
        infos = [e.p2p_export() for e in engs]
        for e in engs:
            e.p2p_import(infos)
    T = 0.0
    total_halo = 0
    for step in range(steps):
        ov_a, ov_b = orc.dump_overlaps()
        for e in engs:
            e.set_overlaps(ov_a, ov_b)
        end = np.inf if end_dt is None else T + end_dt
        vel = [{} for _ in engs]
        if speed is None:
            orc.bulk_update(end_time=end, record_candidates=True)
        else:  # (ids are 0 .. n-1: the oracle's ascending-id arrays are indexed by id)
            vx, vy = scenes.uniform(9000 + step, n, 1, -speed, speed), scenes.uniform(9000 + step, n, 2, -speed, speed)
            orc.bulk_update(end_time=end, vx=vx, vy=vy, record_candidates=True)
            vel = [dict(vel_x=vx[ids], vel_y=vy[ids]) for ids in ids_of]
        st_after = orc.dump_state()
        import torch

        dev = torch.device("cuda", 0)
        if direct:
            for e, v in zip(engs, vel):  # every rank's step is enqueued before any is waited for: the kernels wait for each other's flags
                e.shard_step(T, end_time=end, **v)
            results = [e.shard_result() for e in engs]
            gk = [e.shard_global_next() for e in engs]
            rows = [sharding.wrap(e.shard_local_key_ptr(), 16, dev).view(torch.int64).tolist() for e in engs]
            want = sharding.min_key(rows)
            for tkey, tie, complete in gk:
                assert not complete or (tkey, tie) == want  # every rank reduced the same global minimum on the device
            assert step == 0 or all(c for _, _, c in gk), "only the first step may have to grow buffers"
            global_event = sharding.decode_key(*want)
            total_halo += 1
        else:
            states = [e.shard_begin(T, end_time=end, **v) for e, v in zip(engs, vel)]
            exchange_on_one_gpu(engs, states)
            key_ptrs = [e.shard_finish() for e in engs]
            results = [e.shard_result() for e in engs]
            for st in states:  # records that actually travelled: the slab headers
                for d in range(world):
                    total_halo += int(sharding.wrap(st["send_ptr"] + d * st["slab_bytes"], 4, dev).view(torch.int32).item())
            # the device-side local keys, all-gathered and min-reduced like sharding.ShardedEngine does
            rows = [sharding.wrap(p, 16, dev).view(torch.int64).tolist() for p in key_ptrs]
            global_event = sharding.decode_key(*sharding.min_key(rows))
        # events: union over ranks, merged in the canonical order
        ev = np.concatenate([e.fetch_events() for e in engs])
        ev = ev[key_of(ev)]
        t, idx, kind, a, b = orc.dump_events()
        assert len(ev) == len(t)
        np.testing.assert_array_equal(ev["kind"], kind)
        np.testing.assert_array_equal(ev["a"], a)
        np.testing.assert_array_equal(ev["b"], np.where(kind == 0, 0, b))
        np.testing.assert_array_equal(bits(ev["time"]), bits(t))
        # global next event = min over ranks
        keys = [sharding.event_key(r.next_time, r.next_event.kind, r.next_event.a, r.next_event.b) for r in results]
        assert min(keys)[0] == orc.next_time() == global_event[0]
        if len(t):
            assert min(keys) == sharding.event_key(t[0], int(kind[0]), int(a[0]), int(b[0]))
            assert global_event == (t[0], int(kind[0]), int(a[0]), 0 if kind[0] == 0 else int(b[0]))
        # candidate pairs: disjoint union == oracle set
        ohi, olo = orc.bulk_candidates()
        pairs = [e.fetch_candidate_pairs() for e in engs]
        dkey = np.sort(np.concatenate([hi << np.uint64(32) | lo for hi, lo in pairs]))
        np.testing.assert_array_equal(dkey, np.sort(ohi.astype(np.uint64) << np.uint64(32) | olo))
        assert sum(r.n_separate_solves for r in results) == len(ov_a)
        if migrate_margin is not None:
            parts = [e.shard_take_movers(migrate_margin) for e in engs]
            for r, e in enumerate(engs):
                mine = sharding.rows_for(parts, r)
                e.shard_add_rows(mine)
                gone = parts[r][0]["id"].astype(np.int64)
                ids_of[r] = np.sort(np.concatenate([np.setdiff1d(ids_of[r], gone), mine["id"].astype(np.int64)]))
                moved += len(gone)
                info = e.shard_rows()
                assert info["rows"] == len(ids_of[r]) == e.n
            assert np.array_equal(np.sort(np.concatenate(ids_of)), np.arange(n)), "every hitbox is resident on exactly one rank"
            # residence follows space now: the centre column of every row lies in (or within the margin of) its rank's strip
            col = np.floor(st_after["pos_x"] / CW)
            for r in range(world):
                c = col[ids_of[r]]
                assert np.all(c >= float(bounds[r]) - migrate_margin) and np.all(c < float(bounds[r + 1]) + migrate_margin)
        if migrate_margin is not None or speed is not None:
            # every rank's table, in its ascending-id order, is the oracle's table of those hitboxes
            for r, e in enumerate(engs):
                got = e.download_state()
                for k, colv in got.items():
                    np.testing.assert_array_equal(bits(colv), bits(st_after[k][ids_of[r]]), err_msg=f"rank {r} {k}")
        T = T + dt
        # drain the oracle's events up to the next step (no velocity changes)
        while True:
            tt = min(orc.next_time(), T)
            orc.set_time(tt)
            while orc.next() is not None:
                pass
            if tt >= T:
                break
    return moved if migrate_margin is not None else total_halo


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_equals_single_collider(oracle, world):
    halo = run_sharded_case(oracle, scenes.uniform_density(6000), world, steps=3, dt=0.1, end_dt=0.1)
    assert halo > 0, "the scene must actually exercise the halo exchange"


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_direct_exchange(oracle, world):
    """The same parity bar with the halo records and keys pushed straight into the peers' memory by the step kernels
    (cb200_shard_step), no collective and no host copy in between."""
    run_sharded_case(oracle, scenes.uniform_density(6000), world, steps=4, dt=0.1, end_dt=0.1, direct=True)


def test_sharded_direct_exchange_clustered(oracle):
    run_sharded_case(oracle, scenes.c5(8000), 4, steps=1, dt=0.1, end_dt=0.1, direct=True)  # (one step: the harness does not replay Reiterates)


def test_sharded_full_cell_sweeps_and_reiterates(oracle):
    run_sharded_case(oracle, scenes.c1(1000), 3, steps=1, dt=0.5, end_dt=None)


def test_sharded_residents_unrelated_to_position(oracle):
    """Ownership is by id, work is by space: with residents scattered over all ranks everything still matches."""
    run_sharded_case(oracle, scenes.uniform_density(3000), 2, steps=2, dt=0.1, end_dt=0.1, resident_by="id")


def test_sharded_clustered(oracle):
    run_sharded_case(oracle, scenes.c5(8000), 4, steps=1, dt=0.1, end_dt=0.1)


@pytest.mark.parametrize("direct", [False, True])
def test_sharded_dense_cells(oracle, direct):
    """Config 4's density across 2 ranks: the long slot runs go through k_solve_dense, pairs straddling the strip edge are
    still solved by exactly one rank."""
    run_sharded_case(oracle, scenes.uniform_density(4000, density=1.0, seed=scenes.SEED_BASE + 41), 2, steps=2, dt=0.1, end_dt=0.1, direct=direct)


@pytest.mark.parametrize("direct", [False, True])
def test_sharded_owner_migration(oracle, direct):
    """SURVEY.md 8e owner migration: residents start out unrelated to position (id mod world), so the first migration moves
    about two thirds of all rows to the rank whose strip holds them; later steps install fast new velocities and rows keep
    crossing strip edges.  After every migration: each hitbox is resident exactly once, residence follows space, every rank's
    table equals the oracle's, and the following steps (new velocities in the ranks' new array order) still match the oracle
    event for event.  direct: the peers' mapped receive areas stay where they were."""
    moved = run_sharded_case(oracle, scenes.uniform_density(6000), 3, steps=5, dt=0.1, end_dt=0.1, resident_by="id", direct=direct,
                             migrate_margin=0, speed=30.0)
    assert moved > 4000 + 100, "the first migration re-homes ~2/3 of the rows, later ones move rows that crossed an edge"


def test_sharded_migration_hysteresis(oracle):
    """With a margin of one column, rows whose centre is in the column next to the strip stay where they are."""
    moved = run_sharded_case(oracle, scenes.uniform_density(6000), 2, steps=8, dt=0.1, end_dt=0.1, migrate_margin=1, speed=30.0)
    moved0 = run_sharded_case(oracle, scenes.uniform_density(6000), 2, steps=8, dt=0.1, end_dt=0.1, migrate_margin=0, speed=30.0)
    assert 0 < moved < moved0

