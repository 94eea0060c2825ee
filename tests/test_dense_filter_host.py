"""Pre-GPU check of k_solve_dense's arithmetic on REAL candidate pairs (tests/host_narrowphase_check.cpp does it on random
ones): for every candidate pair of a dense and of a clustered seeded scene, as the step's records hold the two hitboxes,
 - a pair the swept-edge filter rejects has collide_time = +inf in the oracle (the filter is exact, never lossy);
 - collide_time_merged (one box stage + one disc stage for all kind pairs) equals the oracle's collide_time bit for bit;
 - the filter does reject a large share of the pairs (otherwise the dense kernel has no point).
The product library never runs this host instantiation."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scenes = importlib.import_module("collider-rs_b200.scenes")


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = tmp_path_factory.mktemp("filter") / "libfiltercheck.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "host_filter_check.cpp")], check=True)
    L = C.CDLL(str(so))
    L.filter_and_solve.argtypes = [C.c_uint64] + [C.c_void_p] * 7
    return L


def records_of(state):
    """n x 9 doubles as HbRec holds them after the step: shape at T, velocity, remaining duration."""
    return np.ascontiguousarray(np.stack([state["pos_x"], state["pos_y"], state["dim_x"], state["dim_y"], state["vel_x"], state["vel_y"],
                                          state["rsz_x"], state["rsz_y"], state["end_time"] - state["start_time"]], axis=1))


@pytest.mark.parametrize("name,scene_fn,min_rejected", [
    ("dense", lambda: scenes.uniform_density(6000, density=1.0, seed=scenes.SEED_BASE + 40), 0.4),
    ("very_dense", lambda: scenes.uniform_density(3000, density=6.0, seed=scenes.SEED_BASE + 41), 0.2),
    ("clustered_fast_small", lambda: scenes.c5(20000), 0.3),
])
def test_filter_is_exact_on_scene_pairs(oracle, hostlib, name, scene_fn, min_rejected):
    orc = oracle.Collider(scenes.CELL_WIDTH, scenes.PADDING)
    orc.add_hitboxes(scene_fn())
    T = 0.0
    for step in range(2):
        orc.bulk_update(end_time=T + 0.1, record_candidates=True)
        st = orc.dump_state()
        hi, lo = orc.bulk_candidates()
        assert len(hi) > 10000
        row = {int(i): k for k, i in enumerate(st["id"])}
        ih = np.array([row[int(i)] for i in hi])
        il = np.array([row[int(i)] for i in lo])
        rec = records_of(st)
        a, b = np.ascontiguousarray(rec[ih]), np.ascontiguousarray(rec[il])
        ka, kb = np.ascontiguousarray(st["kind"][ih], dtype=np.uint8), np.ascontiguousarray(st["kind"][il], dtype=np.uint8)
        n = len(hi)
        passed = np.zeros(n, dtype=np.uint8)
        merged = np.zeros(n)
        plain = np.zeros(n)
        hostlib.filter_and_solve(n, a.ctypes.data, ka.ctypes.data, b.ctypes.data, kb.ctypes.data, passed.ctypes.data, merged.ctypes.data, plain.ctypes.data)
        # the oracle's collide_time of (hi, lo rebased to T) — lo is at its start time already, the advance is by zero
        want = oracle.collide_time_batch(a, ka, b, kb)
        rejected = passed == 0
        assert np.all(np.isposinf(want[rejected])), f"{name}: the filter dropped {np.count_nonzero(~np.isposinf(want[rejected]))} pairs that collide"
        assert np.array_equal(merged.view(np.uint64), want.view(np.uint64))
        assert np.array_equal(plain.view(np.uint64), want.view(np.uint64))
        assert rejected.mean() > min_rejected, f"{name}: only {rejected.mean():.2f} of the pairs rejected"
        # and it is tight where it can be: with equal durations everything that passes has touching swept bounds, so a
        # passing pair with an infinite time is one whose shapes miss each other inside touching bounds (diagonal passes)
        T += 0.1
        while True:
            t = min(orc.next_time(), T)
            orc.set_time(t)
            while orc.next() is not None:
                pass
            if t >= T:
                break
