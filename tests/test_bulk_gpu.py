"""GPU parity: cb200_bulk_update (CUDA, through the C ABI) against the CPU oracle on the same seeded scenes.

Bar: identical candidate-pair set, identical cell membership, identical (kind, id, id) event
sequence under the canonical order, event times and rebased state bit-equal (the north_star allows 1e-9
relative for times; the f64 contract of DESIGN.md makes them bit-equal, and that is what is asserted).
"""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cb = importlib.import_module("collider-rs_b200")
scenes = importlib.import_module("collider-rs_b200.scenes")

CW, PAD = scenes.CELL_WIDTH, scenes.PADDING


@pytest.fixture(autouse=True, params=["tile", "global"])
def device_path(request, monkeypatch):
    """Every case runs on both device paths of the bulk step: the tile path (tile.cuh: binning / enumeration / solve per
    block of cells in shared memory; the default) and the global path (slot table, entry array, work lists)."""
    monkeypatch.setenv("CB200_TILE", "1" if request.param == "tile" else "0")
    return request.param


STATE_COLS = ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "start_time", "end_time", "pub_end_time")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def oracle_events(orc):
    t, idx, kind, a, b = orc.dump_events()
    return t, kind, a, b


def assert_step_parity(orc, eng, res, check_grid=True):
    """Compare everything the last bulk step produced on both sides."""
    # event queue, in order
    t, kind, a, b = oracle_events(orc)
    ev = eng.fetch_events()
    assert len(ev) == len(t) == res.n_events
    np.testing.assert_array_equal(ev["kind"], kind)
    np.testing.assert_array_equal(ev["a"], a)
    np.testing.assert_array_equal(ev["b"], np.where(kind == 0, 0, b))
    np.testing.assert_array_equal(bits(ev["time"]), bits(t))
    # min-reduction
    assert res.next_time == orc.next_time()
    if len(t):
        assert (res.next_event.kind, res.next_event.a) == (kind[0], a[0])
        assert bits([res.next_event.time])[0] == bits(t[:1])[0]
        if kind[0] != 0:
            assert res.next_event.b == b[0]
    # candidate-pair set
    ohi, olo = orc.bulk_candidates()
    dhi, dlo = eng.fetch_candidate_pairs()
    assert res.n_collide_solves == len(ohi) == len(dhi)
    okey = np.sort(ohi.astype(np.uint64) << np.uint64(32) | olo)
    dkey = np.sort(dhi << np.uint64(32) | dlo)
    np.testing.assert_array_equal(dkey, okey)
    assert len(np.unique(dkey)) == len(dkey), "a pair was produced twice"
    # rebased table
    st = eng.download_state()
    ost = orc.dump_state()
    for k in STATE_COLS:
        np.testing.assert_array_equal(bits(st[k]), bits(ost[k]), err_msg=k)
    if check_grid:
        gx, gy, gg, gi = orc.dump_grid()
        dx, dy, di = eng.fetch_cell_entries()
        assert res.n_entries == len(gx) == len(dx)
        o = np.lexsort((gy, gx, gi))
        d = np.lexsort((dy, dx, di))
        np.testing.assert_array_equal(dx[d], gx[o])
        np.testing.assert_array_equal(dy[d], gy[o])
        np.testing.assert_array_equal(di[d], gi[o])


def seeded(oracle, scene, end_time=cb.INF):
    orc = oracle.Collider(CW, PAD)
    orc.add_hitboxes(scene, end_time=end_time)
    eng = cb.Engine(CW, PAD)
    st = orc.dump_state()
    eng.upload_state(**st)
    a, b = orc.dump_overlaps()
    eng.set_overlaps(a, b)
    return orc, eng


def drain_until(orc, t_stop):
    """The user loop of README.md:66-80 without velocity changes: process every event up to t_stop."""
    n = 0
    while True:
        t = min(orc.next_time(), t_stop)
        orc.set_time(t)
        while orc.next() is not None:
            n += 1
        if t >= t_stop:
            return n


@pytest.mark.parametrize("solve", ["two_phase", "one_phase"])
@pytest.mark.parametrize("n", [1, 2, 1000, 20000])
def test_bulk_step_parity_uniform(oracle, n, solve, monkeypatch):
    monkeypatch.setenv("CB200_SOLVE", "1" if solve == "one_phase" else "2")  # (read when the engine is created)
    scene = scenes.uniform_density(n)
    orc, eng = seeded(oracle, scene)
    orc.bulk_update(end_time=0.1, record_candidates=True)
    res = eng.bulk_update(0.0, end_time=0.1)
    assert res.n_hitboxes == n
    assert_step_parity(orc, eng, res)


def test_bulk_step_parity_full_cell_sweeps(oracle):
    """end_time = +inf: every moving hitbox sweeps one cell period and queues a Reiterate (a-3)."""
    scene = scenes.c1()
    orc, eng = seeded(oracle, scene)
    orc.bulk_update(end_time=cb.INF, record_candidates=True)
    res = eng.bulk_update(0.0, end_time=cb.INF)
    assert res.n_solitaire == 1000
    assert_step_parity(orc, eng, res)


def test_bulk_chain_with_events_between_steps(oracle):
    """Config 2 shape: bulk step every 0.1, Collide/Separate events processed in between (they change the overlap
    sets, which the host re-uploads); device state is carried across steps without re-upload."""
    scene = scenes.uniform_density(5000, seed=scenes.SEED_BASE + 2)
    orc, eng = seeded(oracle, scene)
    T = 0.0
    for step in range(8):
        orc.bulk_update(end_time=T + 0.1, record_candidates=True)
        res = eng.bulk_update(T, end_time=T + 0.1)
        assert_step_parity(orc, eng, res, check_grid=(step % 3 == 0))
        T = T + 0.1
        drain_until(orc, T)
        a, b = orc.dump_overlaps()
        eng.set_overlaps(a, b)


def test_bulk_new_velocities(oracle):
    """The bulk entry point installs new velocities (reference-equivalent: set_hitbox_vel for every id ascending)."""
    n = 3000
    scene = scenes.uniform_density(n, seed=77)
    orc, eng = seeded(oracle, scene)
    T = 0.0
    for step in range(3):
        vx = scenes.uniform(1000 + step, n, 1, -3.0, 3.0)
        vy = scenes.uniform(1000 + step, n, 2, -3.0, 3.0)
        orc.bulk_update(end_time=T + 0.25, vx=vx, vy=vy, record_candidates=True)
        res = eng.bulk_update(T, end_time=T + 0.25, vel_x=vx, vel_y=vy)
        assert_step_parity(orc, eng, res)
        T += 0.25
        drain_until(orc, T)
        eng.set_overlaps(*orc.dump_overlaps())


def test_resizing_groups_and_negative_coordinates(oracle):
    """Rows the reference leaves untested: resize velocities, several groups with asymmetric
    interact_groups, group None, coordinates around the origin (floor toward -inf in index_bounds)."""
    n = 4000
    s = scenes.make_scene(n, 300.0, seed=4242, x0=-150.0)
    s["pos_y"] = s["pos_y"] - 150.0
    r = scenes.uniform(4242, n, 20, -0.05, 0.3)
    s["rsz_x"] = r
    s["rsz_y"] = np.where(s["kind"] == 0, r, scenes.uniform(4242, n, 21, -0.05, 0.3))
    g = (scenes.splitmix64(4242, n, 22) % np.uint64(4)).astype(np.uint32)
    s["group"] = np.where(g == 3, np.uint32(cb.GROUP_NONE), g)
    s["interact_mask"] = np.array([0b011, 0b101, 0b110, 0b111], dtype=np.uint32)[g]
    orc, eng = seeded(oracle, s)
    T = 0.0
    for step in range(3):
        orc.bulk_update(end_time=T + 0.2, record_candidates=True)
        res = eng.bulk_update(T, end_time=T + 0.2)
        assert_step_parity(orc, eng, res)
        T += 0.2
        drain_until(orc, T)
        eng.set_overlaps(*orc.dump_overlaps())


@pytest.mark.parametrize("solve", ["two_phase", "one_phase"])
def test_clustered_fast_small_shapes(oracle, solve, monkeypatch):
    """Config 5 at a size the oracle finishes quickly: dense cells and multi-cell sweeps."""
    monkeypatch.setenv("CB200_SOLVE", "1" if solve == "one_phase" else "2")
    s = scenes.c5(20000)
    orc, eng = seeded(oracle, s)
    orc.bulk_update(end_time=0.1, record_candidates=True)
    res = eng.bulk_update(0.0, end_time=0.1)
    assert_step_parity(orc, eng, res)


@pytest.mark.parametrize("dense_rank,kernel", [("1", "filter"), ("8", "filter"), ("0", "filter"), ("8", "lockstep"), ("auto", "filter")])
@pytest.mark.parametrize("density", [1.0, 6.0])
def test_dense_cells(oracle, density, dense_rank, kernel, monkeypatch):
    """Config 4's density (1 shape per unit^2, ~30 entries per slot run) and six times that (runs of ~200 entries: several
    descriptors per run, a warp's 32 ranks ending inside the run).  CB200_DENSE_RANK = first rank of a run that k_solve_dense
    takes: 1 = every pair through the dense kernel, 8 = the default split, 0 = none (work items only); CB200_DENSE_KERNEL=1
    = the first (lockstep) dense kernel, kept for A/B; "auto" = the default: chosen per step from the previous step's counters
    (work items on the first step, the dense kernel from the second)."""
    if dense_rank == "auto":
        monkeypatch.delenv("CB200_DENSE_RANK", raising=False)
    else:
        monkeypatch.setenv("CB200_DENSE_RANK", dense_rank)
    monkeypatch.setenv("CB200_DENSE_KERNEL", "1" if kernel == "lockstep" else "2")
    scene = scenes.uniform_density(5000, density=density, seed=scenes.SEED_BASE + 40)
    orc, eng = seeded(oracle, scene)
    T = 0.0
    for step in range(3 if dense_rank == "auto" else 2):
        orc.bulk_update(end_time=T + 0.1, record_candidates=True)
        res = eng.bulk_update(T, end_time=T + 0.1)
        assert res.n_collide_solves > 5000 * (10 if density < 2 else 60)
        assert_step_parity(orc, eng, res, check_grid=(step == 0))
        T += 0.1
        drain_until(orc, T)
        eng.set_overlaps(*orc.dump_overlaps())


@pytest.mark.parametrize("period", [0, 1, 3])
def test_table_order_never_changes_results(oracle, period):
    """The device keeps its rows in grid-slot order and re-sorts them every `period` steps (0: only after upload);
    with new velocities arriving in id order and the state read back in id order, nothing may depend on it."""
    n = 6000
    scene = scenes.uniform_density(n, seed=991)
    orc, eng = seeded(oracle, scene)
    eng.set_resort_period(period)
    T = 0.0
    for step in range(5):
        vx = scenes.uniform(500 + step, n, 1, -12.0, 12.0)  # fast enough to shuffle the cell order between steps
        vy = scenes.uniform(500 + step, n, 2, -12.0, 12.0)
        orc.bulk_update(end_time=T + 0.3, vx=vx, vy=vy, record_candidates=True)
        res = eng.bulk_update(T, end_time=T + 0.3, vel_x=vx, vel_y=vy)
        assert_step_parity(orc, eng, res, check_grid=(step == 4))
        T += 0.3
        drain_until(orc, T)
        eng.set_overlaps(*orc.dump_overlaps())
    assert eng.resort_count() == (1 if period == 0 else 1 + 4 // period)
    rects = eng.fetch_index_rects()
    gx, gy, gg, gi = orc.dump_grid()
    for i in (0, 17, n - 1):  # fetch_index_rects is in id order
        sel = gi == i
        assert (rects[i][0], rects[i][1], rects[i][2], rects[i][3]) == (gx[sel].min(), gy[sel].min(), gx[sel].max() + 1, gy[sel].max() + 1)


def canonical_order(ev):
    cls = np.where(ev["kind"] == 0, 0, 1)
    kbit = np.where(ev["kind"] == 1, 1, 0)
    return np.lexsort((ev["b"], kbit, ev["a"], cls, ev["time"]))


@pytest.mark.parametrize("sort_mode", ["oneshot", "slow"])
def test_event_queue_order_at_scale(oracle, sort_mode, monkeypatch):
    """The queue materialisation (one-shot two-region bucket sort; CB200_SORT=slow: bucket sort on the time key, recursing
    into heavy ties, radix fallback) on 120k hitboxes:
    step 1 starts at T = 0 (event times span every binade) and is compared with the oracle event by event; step 2 runs
    without processing the Collide events of step 1, so every pair already in contact reports time == T exactly (tens of
    thousands of ties ordered by the tie word) — checked for canonical order, uniqueness and count."""
    monkeypatch.setenv("CB200_SORT", sort_mode)
    n = 120_000
    scene = scenes.uniform_density(n, density=0.05, seed=4711)
    orc, eng = seeded(oracle, scene)
    orc.bulk_update(end_time=0.1)
    res = eng.bulk_update(0.0, end_time=0.1)
    t, kind, a, b = oracle_events(orc)
    ev = eng.fetch_events()
    assert len(ev) == len(t) == res.n_events > 2_000
    np.testing.assert_array_equal(ev["kind"], kind)
    np.testing.assert_array_equal(ev["a"], a)
    np.testing.assert_array_equal(bits(ev["time"]), bits(t))
    res = eng.bulk_update(0.1, end_time=0.2)
    ev = eng.fetch_events(pinned=True).copy()
    assert len(ev) == res.n_events
    assert (ev["time"] == 0.1).sum() > 200, "the case must contain a heavy tie"
    np.testing.assert_array_equal(canonical_order(ev), np.arange(len(ev)))
    key = (ev["a"].astype(np.uint64) << np.uint64(32) | ev["b"]) * np.uint64(4) + ev["kind"].astype(np.uint64)
    assert len(np.unique(key)) == len(ev)
    assert eng.event_sort_fallbacks() == 0, "ties at the step time must not leave the one-shot path"


def test_event_queue_heavy_ties_at_later_times(oracle):
    """Two velocity classes and end_time = +inf: half of the Reiterate events fall on exactly T + 2, the other half on
    exactly T + 4 (cell periods, grid.rs:61-72), with the pair events spread around them.  The tie at T + 4 is not at the
    first event time, so the one-shot sort reports an oversized bucket and the recursive path takes over; then a second
    step where every event time is equal (no pair events at all)."""
    n = 20_000
    scene = scenes.uniform_density(n, seed=99)
    half = np.arange(n) % 2 == 0
    vx = np.where(half, 2.0, -1.0)
    vy = np.where(half, 2.0, 0.5)
    orc, eng = seeded(oracle, scene)
    orc.bulk_update(end_time=cb.INF, vx=vx, vy=vy)
    res = eng.bulk_update(0.0, end_time=cb.INF, vel_x=vx, vel_y=vy)
    t, kind, a, b = oracle_events(orc)
    ev = eng.fetch_events()
    assert len(ev) == len(t) == res.n_events
    assert (t == 2.0).sum() >= n // 2 and (t == 4.0).sum() >= n // 2 - 1
    np.testing.assert_array_equal(ev["kind"], kind)
    np.testing.assert_array_equal(ev["a"], a)
    np.testing.assert_array_equal(ev["b"], np.where(kind == 0, 0, b))
    np.testing.assert_array_equal(bits(ev["time"]), bits(t))
    assert eng.event_sort_fallbacks() == 1
    # all hitboxes move alike: no pair can collide, every event is a Reiterate at exactly T + 2 -> region A alone
    scene2 = scenes.uniform_density(n, seed=100)
    orc2, eng2 = seeded(oracle, scene2)
    v = np.full(n, 2.0)
    orc2.bulk_update(end_time=cb.INF, vx=v, vy=v)
    res2 = eng2.bulk_update(0.0, end_time=cb.INF, vel_x=v, vel_y=v)
    t2, kind2, a2, b2 = oracle_events(orc2)
    ev2 = eng2.fetch_events()
    assert len(ev2) == len(t2) == res2.n_events
    sol = kind2 == 0
    assert sol.sum() == n and np.all(t2[sol] == 2.0)
    np.testing.assert_array_equal(ev2["kind"], kind2)
    np.testing.assert_array_equal(ev2["a"], a2)
    np.testing.assert_array_equal(bits(ev2["time"]), bits(t2))


def test_tile_path_is_taken_and_splits_or_falls_back_when_a_tile_does_not_fit(oracle, device_path, monkeypatch):
    """The tile path handles a tile whose records exceed the shared-memory capacity by splitting its cell range (tiny
    CB200_TILE_CAP), and a visitor slab that overflows by repeating the step on the global path (tiny CB200_TILE_VCAP) —
    results identical either way."""
    if device_path != "tile":
        pytest.skip("tile path only")
    scene = scenes.uniform_density(20000, density=0.5, seed=scenes.SEED_BASE + 50)
    for env, want_fallback in (({}, False), ({"CB200_TILE_CAP": "40"}, False), ({"CB200_TILE_VCAP": "2"}, True), ({"CB200_TILE_LOG2": "11", "CB200_TILE_CAP": "64", "CB200_TILE_VCAP": "400"}, False),
                               ({"CB200_TILE_LOG2": "6"}, False)):
        for k in ("CB200_TILE_CAP", "CB200_TILE_VCAP", "CB200_TILE_LOG2"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        orc, eng = seeded(oracle, scene)
        T = 0.0
        for step in range(2):
            orc.bulk_update(end_time=T + 0.1, record_candidates=True)
            res = eng.bulk_update(T, end_time=T + 0.1)
            assert_step_parity(orc, eng, res, check_grid=(step == 1))
            T += 0.1
            drain_until(orc, T)
            eng.set_overlaps(*orc.dump_overlaps())
        assert (eng.tile_fallbacks() > 0) == want_fallback, env
        assert eng.last_step_path() == ("global" if want_fallback else "tile"), env


def drain_recording(orc, t_stop):
    """drain_until, keeping what the reference delivers: (time, kind, id_1, id_2) of every user-visible event up to t_stop."""
    out = []
    while True:
        t = min(orc.next_time(), t_stop)
        orc.set_time(t)
        while True:
            ev = orc.next()
            if ev is None:
                break
            out.append((t, ev[0], ev[1], ev[2]))
        if t >= t_stop:
            return out


@pytest.mark.parametrize("scene_name", ["sparse", "dense", "dense_cold"])
def test_window_drain_keeps_the_overlap_set_on_the_device(oracle, scene_name):
    """cb200_set_window_drain: the pair events inside every step's window are processed on the device (chains included) and the
    overlap set stays there — no set_overlaps between the steps.  Per step: the events up to the end of the window equal what
    the oracle delivers through next() (times bit-equal), and the device overlap set equals the oracle's at the window's end.
    dense_cold: the device starts WITHOUT the overlap set (the shapes in contact at t = 0 come in as zero-delay Collide events
    of the first step — tens of thousands more pairs than the freshly sized table holds, so the set outgrows its table inside
    one step and is rebuilt); from the end of the first window on everything must agree again."""
    scene = {"sparse": lambda: scenes.uniform_density(30000, seed=scenes.SEED_BASE + 60),
             "dense": lambda: scenes.uniform_density(4000, density=1.0, seed=scenes.SEED_BASE + 61),
             "dense_cold": lambda: scenes.uniform_density(20000, density=1.0, seed=scenes.SEED_BASE + 62)}[scene_name]()
    orc, eng = seeded(oracle, scene)
    cold = scene_name == "dense_cold"
    if cold:
        eng.set_overlaps(np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    eng.set_window_drain(True)
    T, dt, followups = 0.0, 0.1, 0
    key = lambda p, q: np.sort(np.maximum(p, q).astype(np.uint64) << np.uint64(32) | np.minimum(p, q).astype(np.uint64))
    for step in range(4 if cold else 6):
        orc.bulk_update(end_time=T + dt)
        res = eng.bulk_update(T, end_time=T + dt)
        assert res.n_solitaire == 0  # (no Reiterate inside a window: the drain is exact)
        want = drain_recording(orc, T + dt)
        ev = eng.fetch_events()
        assert len(ev) == res.n_events
        if not (cold and step == 0):
            ev = ev[ev["time"] <= T + dt]
            got = sorted((bits([t])[0], int(k), int(min(a, b)), int(max(a, b))) for t, k, a, b in zip(ev["time"], ev["kind"], ev["a"], ev["b"]))
            assert got == sorted((bits([t])[0], k, a, b) for t, k, a, b in want), step
            assert np.all(np.diff(ev["time"]) >= 0)
        followups += eng.last_drain()["followup_events"]
        oa, ob = orc.dump_overlaps()
        da, db = eng.fetch_overlaps()
        np.testing.assert_array_equal(key(da, db), key(oa, ob))
        assert eng.last_drain()["pairs_now"] == len(oa)
        if cold and step == 0:
            assert len(oa) > 40000
        T += dt
    assert eng.last_drain()["reiterates_in_windows"] == 0
    if scene_name != "sparse":
        assert followups > 0, "the dense cases must contain chains (a pair that collides and separates inside one window)"


def test_pipelined_io_matches_plain_calls(oracle):
    """cb200_stage_velocities + cb200_fetch_events_begin / _end (transfers on their own streams, compact 16-byte events)
    give exactly what the plain calls give: same events in the same order, same state."""
    n = 8000
    scene = scenes.uniform_density(n, seed=scenes.SEED_BASE + 70)
    orc, eng = seeded(oracle, scene)
    ids = np.ascontiguousarray(orc.dump_state()["id"], dtype=np.uint64)
    pin = cb.PinnedArray(4 * n, cb.EVENT16_DTYPE)
    pv = [(cb.PinnedArray(n, np.float64), cb.PinnedArray(n, np.float64)) for _ in range(2)]
    T = 0.0
    for step in range(4):
        vx = scenes.uniform(900 + step, n, 1, -3.0, 3.0)
        vy = scenes.uniform(900 + step, n, 2, -3.0, 3.0)
        pv[step & 1][0].array[:] = vx
        pv[step & 1][1].array[:] = vy
        eng.stage_velocities(vel_x=pv[step & 1][0].array, vel_y=pv[step & 1][1].array)
        orc.bulk_update(end_time=T + 0.2, vx=vx, vy=vy, record_candidates=True)
        res = eng.bulk_update(T, end_time=T + 0.2)  # consumes the staged velocities
        got = eng.fetch_events_begin(pin.array)
        eng.fetch_events_end()
        assert got == res.n_events
        ev = cb.Engine.decode_events16(pin.array[:got], ids)
        t, kind, a, b = oracle_events(orc)
        np.testing.assert_array_equal(ev["kind"], kind)
        np.testing.assert_array_equal(ev["a"], a)
        np.testing.assert_array_equal(ev["b"], np.where(kind == 0, 0, b))
        np.testing.assert_array_equal(bits(ev["time"]), bits(t))
        assert_step_parity(orc, eng, res, check_grid=False)  # (the plain fetch after the pipelined one: the same list)
        T += 0.2
        drain_until(orc, T)
        eng.set_overlaps(*orc.dump_overlaps())


def test_aliased_grid_is_still_exact(oracle):
    """Force a tiny grid period so that many cells share a slot: results must not change."""
    scene = scenes.uniform_density(5000)
    orc, eng = seeded(oracle, scene)
    eng.set_grid(6, 6)
    orc.bulk_update(end_time=0.1, record_candidates=True)
    res = eng.bulk_update(0.0, end_time=0.1)
    assert_step_parity(orc, eng, res)


def test_contract_errors(oracle):
    eng = cb.Engine(CW, PAD)
    s = scenes.uniform_density(100)
    eng.upload_state(**s)
    with pytest.raises(cb.ColliderError) as e:
        eng.bulk_update(1.0, end_time=0.5)  # end time must exceed present time (core/mod.rs:147-150)
    assert e.value.code == cb.E_CONTRACT and "end time must exceed present time" in e.value.msg
    with pytest.raises(cb.ColliderError):
        cb.Engine(1.0, 2.0)  # requires cell_width > padding (collider.rs:60)
    with pytest.raises(cb.ColliderError):
        eng.upload_state(**{**s, "id": s["id"][::-1].copy()})


def test_narrowphase_batch_parity(oracle):
    """collide_time / separate_time kernels against the oracle on random and lattice (exactly-touching) pairs."""
    n = 200000
    rng = np.random.default_rng(12345)

    def hitboxes(lattice):
        kind = rng.integers(0, 2, n).astype(np.uint8)
        if lattice:
            px, py = rng.integers(-6, 7, n).astype(float), rng.integers(-6, 7, n).astype(float)
            w = rng.integers(1, 5, n).astype(float)
            h = np.where(kind == 0, w, rng.integers(1, 5, n).astype(float))
            vx, vy = rng.integers(-4, 5, n) * 0.5, rng.integers(-4, 5, n) * 0.5
        else:
            px, py = rng.uniform(-6, 6, n), rng.uniform(-6, 6, n)
            w = rng.uniform(0.05, 4, n)
            h = np.where(kind == 0, w, rng.uniform(0.05, 4, n))
            vx, vy = rng.uniform(-3, 3, n), rng.uniform(-3, 3, n)
        rw = np.where(rng.integers(0, 4, n) == 0, rng.uniform(-0.2, 0.5, n), 0.0)
        rh = np.where(kind == 0, rw, np.where(rw != 0, rng.uniform(-0.2, 0.5, n), 0.0))
        still = rng.integers(0, 7, n) == 0
        vx, vy, rw, rh = [np.where(still, 0.0, v) for v in (vx, vy, rw, rh)]
        dur = np.where(rng.integers(0, 4, n) == 0, np.inf, rng.uniform(0.01, 8, n))
        dur = np.where(~still & (dur == np.inf), 5.0, dur)
        return np.stack([px, py, w, h, vx, vy, rw, rh, dur], axis=1), kind

    eng = cb.Engine(CW, PAD)
    for lattice in (False, True):
        a, ka = hitboxes(lattice)
        b, kb = hitboxes(lattice)
        want = oracle.collide_time_batch(a, ka, b, kb)
        got = eng.collide_time_batch(a, ka, b, kb)
        np.testing.assert_array_equal(bits(got), bits(want))
        assert np.isfinite(want).sum() > n // 50
        want = oracle.separate_time_batch(a, ka, b, kb, PAD)
        got = eng.separate_time_batch(a, ka, b, kb, PAD)
        np.testing.assert_array_equal(bits(got), bits(want))
