"""Test harness: the CPU oracle run as independent spatial tiles on several host threads.

Everything a bulk step leaves behind is a function of single hitboxes (rebased state, Reiterate events, grid cells) or of
PAIRS of hitboxes (overlap at add time, candidate pairs, Collide / Separate events and their times): nothing depends on a
third hitbox except the ORDER of simultaneous events, which is canonical (time, class, hi id, kind, lo id — SURVEY.md 8c)
and therefore re-established by a sort.  So the result of one big oracle equals the union over tiles of small oracles,
where tile t holds every hitbox whose centre lies within `margin` of its core rectangle, and contributes
  * the per-hitbox results of the hitboxes whose centre lies in its core, and
  * the per-pair results of the pairs whose HIGHER id lies in its core
(the partner of such a pair shares a grid cell with it or touches it, so it is well inside the margin).
tests/test_tiled_oracle.py checks this equality against the single oracle on scenes small enough for both; the full-size
GPU parity tests (1M clustered / dense, minutes of single-thread oracle time) then use the tiled form.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

STATE_COLS = ("pos_x", "pos_y", "dim_x", "dim_y", "vel_x", "vel_y", "rsz_x", "rsz_y", "start_time", "end_time", "pub_end_time")


def canonical_order(t, kind, a, b):
    cls = np.where(kind == 0, 0, 1)
    kbit = np.where(kind == 1, 1, 0)
    return np.lexsort((np.where(kind == 0, 0, b), kbit, a, cls, t))


def tiled_bulk_step(oracle, scene: dict, cell_width: float, padding: float, end_time: float, tiles: int = 4, margin: float = 24.0,
                    threads: int | None = None, add_end_time=np.inf):
    """add_hitboxes(scene) at T = 0 followed by ONE bulk_update(end_time) at T = 0, tile by tile.
    Returns a dict: state (columns in ascending id order, as uploaded: the table after add_hitboxes), overlaps (a, b) after
    the adds, and after the step: events (t, kind, a, b) in canonical order, candidates (hi, lo), grid (x, y, id),
    state_after."""
    x, y = scene["pos_x"], scene["pos_y"]
    x0, x1, y0, y1 = float(x.min()), float(x.max()), float(y.min()), float(y.max())
    ex = np.linspace(x0, np.nextafter(x1, np.inf), tiles + 1)
    ey = np.linspace(y0, np.nextafter(y1, np.inf), tiles + 1)
    ids_all = scene["id"]
    jobs = [(i, j) for i in range(tiles) for j in range(tiles)]
    oracle.lib()

    def run(job):
        i, j = job
        inside = (x >= ex[i] - margin) & (x < ex[i + 1] + margin) & (y >= ey[j] - margin) & (y < ey[j + 1] + margin)
        core = (x >= ex[i]) & (x < ex[i + 1]) & (y >= ey[j]) & (y < ey[j + 1])
        if not core.any():
            return None
        sub = {k: v[inside] for k, v in scene.items()}
        core_ids = np.sort(ids_all[core])
        is_core = lambda ids: np.isin(ids, core_ids, assume_unique=False)
        orc = oracle.Collider(cell_width, padding)
        orc.add_hitboxes(sub, end_time=add_end_time)
        st0 = orc.dump_state()
        oa, ob = orc.dump_overlaps()
        keep = is_core(np.maximum(oa, ob))
        oa, ob = oa[keep], ob[keep]
        m0 = is_core(st0["id"])
        st0 = {k: v[m0] for k, v in st0.items()}
        orc.bulk_update(end_time=end_time, record_candidates=True)
        t, idx, kind, a, b = orc.dump_events()
        keep = is_core(a)  # `a` is the creator: the hitbox itself (Reiterate) or the higher id of the pair
        ev = (t[keep], kind[keep], a[keep], b[keep])
        chi, clo = orc.bulk_candidates()
        keep = is_core(chi)
        cand = (chi[keep], clo[keep])
        gx, gy, gg, gi = orc.dump_grid()
        keep = is_core(gi)
        grid = (gx[keep], gy[keep], gi[keep])
        st1 = orc.dump_state()
        m1 = is_core(st1["id"])
        st1 = {k: v[m1] for k, v in st1.items()}
        return dict(st0=st0, ov=(oa, ob), ev=ev, cand=cand, grid=grid, st1=st1, next_time=orc.next_time())

    with ThreadPoolExecutor(threads or min(len(jobs), os.cpu_count() or 1)) as pool:
        parts = [p for p in pool.map(run, jobs) if p is not None]

    def cat_state(key):
        st = {k: np.concatenate([p[key][k] for p in parts]) for k in parts[0][key]}
        o = np.argsort(st["id"], kind="stable")
        return {k: v[o] for k, v in st.items()}

    t = np.concatenate([p["ev"][0] for p in parts])
    kind = np.concatenate([p["ev"][1] for p in parts])
    a = np.concatenate([p["ev"][2] for p in parts])
    b = np.concatenate([p["ev"][3] for p in parts])
    o = canonical_order(t, kind, a, b)
    return dict(
        state=cat_state("st0"),
        overlaps=(np.concatenate([p["ov"][0] for p in parts]), np.concatenate([p["ov"][1] for p in parts])),
        events=(t[o], kind[o], a[o], b[o]),
        candidates=(np.concatenate([p["cand"][0] for p in parts]), np.concatenate([p["cand"][1] for p in parts])),
        grid=tuple(np.concatenate([p["grid"][k] for p in parts]) for k in range(3)),
        state_after=cat_state("st1"),
    )
