"""The tiled form of the oracle (tests/tiled_oracle.py) equals the single oracle — on scenes small enough to run both.
CPU only; pins the harness the full-size GPU parity tests rely on."""
import importlib

import numpy as np
import pytest

from tiled_oracle import STATE_COLS, canonical_order, tiled_bulk_step

scenes = importlib.import_module("collider-rs_b200.scenes")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("name", ["sparse", "clustered", "dense", "groups_resizing"])
def test_tiled_equals_single(oracle, name):
    scene = {"sparse": lambda: scenes.uniform_density(20000), "clustered": lambda: scenes.c5(20000),
             "dense": lambda: scenes.uniform_density(6000, density=1.0, seed=scenes.SEED_BASE + 40),
             # (f-4: resize velocities, four groups with asymmetric interact lists, coordinates around the origin — at a density
             # that gives this small scene enough events)
             "groups_resizing": lambda: scenes.with_groups_and_resizing(scenes.uniform_density(20000, density=0.1, seed=scenes.SEED_BASE + 6), scenes.SEED_BASE + 6)}[name]()
    orc = oracle.Collider(scenes.CELL_WIDTH, scenes.PADDING)
    orc.add_hitboxes(scene)
    st0 = orc.dump_state()
    oa, ob = orc.dump_overlaps()
    orc.bulk_update(end_time=0.1, record_candidates=True)
    t, idx, kind, a, b = orc.dump_events()
    got = tiled_bulk_step(oracle, scene, scenes.CELL_WIDTH, scenes.PADDING, 0.1, tiles=3, threads=4)
    for k in STATE_COLS:
        np.testing.assert_array_equal(bits(got["state"][k]), bits(st0[k]), err_msg=k)
    pair_key = lambda p, q: np.sort(np.maximum(p, q) << np.uint64(32) | np.minimum(p, q))
    np.testing.assert_array_equal(pair_key(*got["overlaps"]), pair_key(oa, ob))
    gt, gk, ga, gb = got["events"]
    assert len(gt) == len(t) > 100
    np.testing.assert_array_equal(canonical_order(t, kind, a, b), np.arange(len(t)))  # the oracle's own order is the canonical one
    np.testing.assert_array_equal(gk, kind)
    np.testing.assert_array_equal(ga, a)
    np.testing.assert_array_equal(np.where(gk == 0, 0, gb), np.where(kind == 0, 0, b))
    np.testing.assert_array_equal(bits(gt), bits(t))
    np.testing.assert_array_equal(pair_key(*got["candidates"]), pair_key(*orc.bulk_candidates()))
    gx, gy, gg, gi = orc.dump_grid()
    tx, ty, ti = got["grid"]
    o, d = np.lexsort((gy, gx, gi)), np.lexsort((ty, tx, ti))
    np.testing.assert_array_equal(tx[d], gx[o])
    np.testing.assert_array_equal(ty[d], gy[o])
    np.testing.assert_array_equal(ti[d], gi[o])
    st1 = orc.dump_state()
    for k in STATE_COLS:
        np.testing.assert_array_equal(bits(got["state_after"][k]), bits(st1[k]), err_msg=k)
