#!/usr/bin/env python
"""Generate tests/golden/bulk_steps.json from the CPU oracle (the reference itself is Rust and cannot run here:
no rustc/cargo — DESIGN.md §2).  The oracle is pinned on the reference's own KATs by tests/test_oracle_kat.py;
this fixture then pins oracle AND device on whole bulk steps: event queue (time as hex bits, kind, a, b),
candidate-pair count and cell-entry count of chained steps on small seeded scenes.

    python tests/golden/make_golden.py        # rewrites bulk_steps.json
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle_py as oracle  # noqa: E402

scenes = importlib.import_module("collider-rs_b200.scenes")

CASES = [
    dict(name="c1_full_cell_sweeps", scene="c1", n=1000, end=None, dt=0.5, steps=2),
    dict(name="uniform_300_dt0.1", scene="uniform", n=300, end=0.1, dt=0.1, steps=3),
    # config-5 fast shapes (speed <= 100) have a cell period of 0.04: keep the window shorter so that no Reiterate
    # (a single-hitbox update) falls between two bulk steps of the chained device state
    dict(name="clustered_2000", scene="c5", n=2000, end=0.03, dt=0.03, steps=3),
    # config 4's density (1 shape per unit^2, ~30 entries per slot run): from its second step on the device takes the
    # dense-run path (k_solve_dense), chosen from the first step's counters
    dict(name="dense_250", scene="uniform", n=250, density=1.0, end=0.1, dt=0.1, steps=3),
    # resizing shapes in several groups
    dict(name="groups_resizing_600", scene="groups", n=600, density=0.2, end=0.1, dt=0.1, steps=3),
]


def make_scene(case):
    if case["scene"] == "c1":
        return scenes.c1(case["n"])
    if case["scene"] == "c5":
        return scenes.c5(case["n"])
    if case["scene"] == "groups":  # resize velocities, four groups (one HbGroup None, asymmetric interact lists), origin-centred
        return scenes.with_groups_and_resizing(scenes.uniform_density(case["n"], density=case["density"], seed=scenes.SEED_BASE + 6), scenes.SEED_BASE + 6)
    return scenes.uniform_density(case["n"], density=case.get("density", 0.025))


def drain_until(orc, t_stop):
    while True:
        t = min(orc.next_time(), t_stop)
        orc.set_time(t)
        while orc.next() is not None:
            pass
        if t >= t_stop:
            return


def run_case(case):
    s = make_scene(case)
    orc = oracle.Collider(scenes.CELL_WIDTH, scenes.PADDING)
    orc.add_hitboxes(s)
    out = []
    T = 0.0
    for _ in range(case["steps"]):
        end = float("inf") if case["end"] is None else T + case["end"]
        overlaps = orc.dump_overlaps()
        orc.bulk_update(end_time=end, record_candidates=True)
        t, idx, kind, a, b = orc.dump_events()
        hi, lo = orc.bulk_candidates()
        gx, gy, gg, gi = orc.dump_grid()
        out.append(dict(
            T=T.hex(), end_time="inf" if case["end"] is None else end.hex(),
            overlaps=[[int(x), int(y)] for x, y in zip(*overlaps)],
            n_candidates=int(len(hi)), n_cell_entries=int(len(gx)), next_time=float(orc.next_time()).hex(),
            events=[[float(tt).hex(), int(k), int(x), int(y)] for tt, k, x, y in zip(t, kind, a, b)],
        ))
        T = T + case["dt"]
        drain_until(orc, T)
    return dict(case=case, steps=out)


if __name__ == "__main__":
    data = dict(generator="tests/golden/make_golden.py", cell_width=scenes.CELL_WIDTH, padding=scenes.PADDING, cases=[run_case(c) for c in CASES])
    with open(os.path.join(HERE, "bulk_steps.json"), "w") as f:
        json.dump(data, f, separators=(",", ":"))
    print({c["case"]["name"]: [len(s["events"]) for s in c["steps"]] for c in data["cases"]})
