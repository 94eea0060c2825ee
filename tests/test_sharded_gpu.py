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
    """The all-gather of the halo slabs, done with device-to-device copies between the contexts of one GPU."""
    import torch

    dev = torch.device("cuda", 0)
    R = len(engs)
    torch.cuda.synchronize()  # every context runs on its own stream
    for r in range(R):
        per = states[r]["slab_bytes"]
        for d in range(R):
            src = sharding.wrap(states[d]["send_ptr"], per, dev)
            dst = sharding.wrap(states[r]["recv_ptr"] + d * per, per, dev)
            dst.copy_(src)
    torch.cuda.synchronize()


def key_of(ev):
    cls = np.where(ev["kind"] == 0, 0, 1)
    kbit = np.where(ev["kind"] == 1, 1, 0)
    return np.lexsort((ev["b"], kbit, ev["a"], cls, ev["time"]))


def run_sharded_case(oracle, scene, world, steps, dt, end_dt, resident_by="space", direct=False):
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
    engs = []
    for r in range(world):
        e = cb.Engine(CW, PAD)
        e.shard_configure(r, world, bounds, n)
        m = owner == r
        e.upload_state(**{k: v[m] for k, v in st.items()})
        engs.append(e)
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
        orc.bulk_update(end_time=end, record_candidates=True)
        import torch

        dev = torch.device("cuda", 0)
        if direct:
            for e in engs:  # every rank's step is enqueued before any is waited for: the kernels wait for each other's flags
                e.shard_step(T, end_time=end)
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
            states = [e.shard_begin(T, end_time=end) for e in engs]
            exchange_on_one_gpu(engs, states)
            key_ptrs = [e.shard_finish() for e in engs]
            results = [e.shard_result() for e in engs]
            for st in states:  # records that actually travelled: the slab headers
                total_halo += int(sharding.wrap(st["send_ptr"], 4, dev).view(torch.int32).item())
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
        T = T + dt
        # drain the oracle's events up to the next step (no velocity changes)
        while True:
            tt = min(orc.next_time(), T)
            orc.set_time(tt)
            while orc.next() is not None:
                pass
            if tt >= T:
                break
    return total_halo


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
