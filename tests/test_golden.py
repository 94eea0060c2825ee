"""Committed golden bulk-step fixtures (tests/golden/bulk_steps.json, made by tests/golden/make_golden.py).

CPU leg: the oracle still reproduces them (guards the oracle against regressions).
GPU leg: the CUDA path reproduces them through the C ABI without the oracle in the loop.
"""
import importlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
scenes = importlib.import_module("collider-rs_b200.scenes")

with open(os.path.join(HERE, "golden", "bulk_steps.json")) as f:
    GOLDEN = json.load(f)


def make_scene(case):
    if case["scene"] == "c1":
        return scenes.c1(case["n"])
    if case["scene"] == "c5":
        return scenes.c5(case["n"])
    if case["scene"] == "groups":  # resize velocities, four groups (one HbGroup None, asymmetric interact lists), origin-centred
        return scenes.with_groups_and_resizing(scenes.uniform_density(case["n"], density=case["density"], seed=scenes.SEED_BASE + 6), scenes.SEED_BASE + 6)
    return scenes.uniform_density(case["n"], density=case.get("density", 0.025))


def fx(h):
    return float("inf") if h == "inf" else float.fromhex(h)


def drain_until(orc, t_stop):
    while True:
        t = min(orc.next_time(), t_stop)
        orc.set_time(t)
        while orc.next() is not None:
            pass
        if t >= t_stop:
            return


@pytest.mark.parametrize("entry", GOLDEN["cases"], ids=[c["case"]["name"] for c in GOLDEN["cases"]])
def test_oracle_reproduces_golden(oracle, entry):
    case = entry["case"]
    orc = oracle.Collider(GOLDEN["cell_width"], GOLDEN["padding"])
    orc.add_hitboxes(make_scene(case))
    for step in entry["steps"]:
        T = fx(step["T"])
        drain_until(orc, T)
        a, b = orc.dump_overlaps()
        assert [[int(x), int(y)] for x, y in zip(a, b)] == step["overlaps"]
        orc.bulk_update(end_time=fx(step["end_time"]), record_candidates=True)
        t, idx, kind, ea, eb = orc.dump_events()
        got = [[float(tt).hex(), int(k), int(x), int(y)] for tt, k, x, y in zip(t, kind, ea, eb)]
        assert got == step["events"]
        assert len(orc.bulk_candidates()[0]) == step["n_candidates"]
        assert len(orc.dump_grid()[0]) == step["n_cell_entries"]
        assert float(orc.next_time()).hex() == step["next_time"]


@pytest.mark.gpu
@pytest.mark.parametrize("entry", GOLDEN["cases"], ids=[c["case"]["name"] for c in GOLDEN["cases"]])
def test_device_reproduces_golden(entry):
    cb = importlib.import_module("collider-rs_b200")
    case = entry["case"]
    eng = cb.Engine(GOLDEN["cell_width"], GOLDEN["padding"])
    eng.upload_state(**make_scene(case))
    for step in entry["steps"]:
        ov = np.array(step["overlaps"], dtype=np.uint64).reshape(-1, 2)
        eng.set_overlaps(ov[:, 0], ov[:, 1])
        res = eng.bulk_update(fx(step["T"]), end_time=fx(step["end_time"]))
        ev = eng.fetch_events()
        got = [[float(e["time"]).hex(), int(e["kind"]), int(e["a"]), int(e["b"])] for e in ev]
        assert got == step["events"]
        assert res.n_collide_solves == step["n_candidates"]
        assert res.n_entries == step["n_cell_entries"]
        assert float(res.next_time).hex() == step["next_time"]
