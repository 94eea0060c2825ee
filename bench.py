#!/usr/bin/env python
"""bench.py — hitbox updates/s + pair TOI solves/s of the bulk step (BASELINE.json metric), SURVEY.md 8(d) contract.

A "step" = one bulk update of the whole scene through the C ABI: every hitbox rebased to T, re-binned into the grid, every
candidate pair's collide_time solved (plus separate_time for the overlapping pairs), next event min-reduced, and the ordered
event queue of the step materialised and copied to the host.  Workload at N=1: config 3 (1M mixed circles/rects, uniform
density 0.025/unit^2, cell 4, padding 0.01, bulk step every 0.1 time units).  N>1: weak scaling, one 1M-hitbox column strip
per GPU; the same line carries a `config4` block (16,777,216 shapes in 4096^2 split over the N strips — strong scaling; the
N=1 line carries the one-GPU anchor).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n HITBOXES] [--workload c3|c3_dense|c5|c3_groups]

ONE JSON line on stdout (rank 0):
  value / ms_per_step   t_step of SURVEY 8(d): device-resident state, bulk step + queue materialisation + D2H of the sorted
                        event list into pinned memory; median over `repeats` runs of K steps, max over ranks
  device_ms_per_step    the kernels of the bulk step alone (CUDA events on the engine stream)
  e2e                   the same through host buffers: new velocities H2D from pinned memory every step, events D2H
  roofline              B_step = 184 N + 104 E + 16 C + 24 V / device time of the step, vs the measured HBM peak;
                        per-kernel figures beside it
  cpu_baseline          the oracle port (C++ restatement of the reference; no Rust toolchain) on ONE host core, bounded sample
  config1 / config2     event-driven loops through the Collider API, device vs port (N=1 only)
  parity                a down-scaled copy of the workload run through the same code path and checked against the oracle
`--impl reference` times ONE single-threaded oracle collider on the full scene (the reference is single-threaded).
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DT = 0.1  # bulk step period (config 2/3)
METRIC = "hitbox updates/sec (full re-bin + pair TOI solve + next-event min + ordered event list per bulk step)"
UNIT = "hitbox updates/s"
C4_TOTAL = 16_777_216
C4_SIDE = 4096.0
DENSITY = {"c3": 0.025, "c3_dense": 1.0, "c5": 0.025, "c3_groups": 0.025}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_of(args):
    """Identical for both arms (the driver compares them)."""
    return {"workload": f"{args.workload}: {args.n} mixed circles/rects per GPU, uniform density {DENSITY[args.workload]}/unit^2, cell_width 4, padding 0.01, "
                        f"bulk step every {DT} (full re-bin + pair TOI solve + next-event min + ordered event list)",
            "hitboxes_per_gpu": args.n, "scaling_unit": "one column strip of the world per GPU",
            "l2": "inputs larger than L2 (GPU arm): the working set of a step — SoA state, records, slot table, cell entries — is > 300 MB per "
                  "million hitboxes against the 126 MB L2; nothing is flushed between iterations"}


def make_workload(scenes, name: str, n: int, rank: int, world: int):
    """Per-rank scene.  Weak scaling: rank r owns the column strip [r*W, (r+1)*W) of a world `world` strips wide."""
    if name in ("c3", "c3_dense"):
        side = math.sqrt(n / DENSITY[name])
        s = scenes.make_scene(n, side, side, seed=scenes.SEED_BASE + 3 + 1000 * rank, x0=side * rank)
    elif name == "c3_groups":  # f-4: resizing shapes in four groups (single GPU)
        s = scenes.c3_groups(n)
        for k in ("rsz_x", "rsz_y"):  # (the bench runs hundreds of steps: resize rates small enough that no shape shrinks below the padding)
            s[k] = s[k] / 64.0
    elif name == "c5":
        s = scenes.c5(n)
        s["pos_x"] = s["pos_x"] + math.sqrt(n / 0.025) * rank
    elif name == "c4":  # config 4: C4_TOTAL shapes in 4096 x 4096, strip `rank` of `world`
        per = n
        s = scenes.make_scene(per, C4_SIDE / world, C4_SIDE, seed=scenes.SEED_BASE + 4 + 1000 * rank, x0=C4_SIDE / world * rank)
    else:
        raise SystemExit(f"unknown workload {name}")
    s["id"] = s["id"] + np.uint64(rank) * np.uint64(n)
    return s


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False

    def run(self):
        # NVML in this process (nvidia_ml_py) — forking nvidia-smi several times a second takes driver locks and CPU time away
        # from the run it is watching; the subprocess is the fallback when the module is missing
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            while not self.stop_flag:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = int(reasons_fn(h))
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
                time.sleep(0.05)
            return
        except Exception:
            self.source = "nvidia-smi"
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.25)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "source": getattr(self, "source", None)}


# ------------------------------------------------------------------------------------------------ CPU side (oracle port)
def oracle_bulk_timing(scene, steps: int, warmup: int):
    """The reference's CPU implementation of the path (C++ restatement; Rust toolchain absent): ONE collider, the ascending
    set_hitbox_vel loop per step, single-threaded like the reference.  (updates/s, solves/s, ms/step, set-up seconds)."""
    from oracle import oracle_py as oracle

    t_setup = time.perf_counter()
    orc = oracle.Collider(4.0, 0.01)
    orc.add_hitboxes(scene)
    t_setup = time.perf_counter() - t_setup
    n = len(scene["id"])
    T = 0.0
    for _ in range(warmup):
        orc.bulk_update(end_time=T + DT)
        T = min(T + DT, orc.next_time())  # stay within the reference's set_time contract without draining events
        orc.set_time(T)
    orc.stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.bulk_update(end_time=T + DT)
        T = min(T + DT, orc.next_time())
        orc.set_time(T)
    el = time.perf_counter() - t0
    st = orc.stats()
    return n * steps / el, (st["collide_solves"] + st["separate_solves"]) / el, el / steps * 1e3, t_setup


def oracle_tiles_all_threads(scenes, workload: str, n_tile: int, steps: int, threads: int):
    """Separately named row: one independent collider per host thread, each on its own tile of the scene (no cross-tile pairs)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle_py as oracle

    oracle.lib()
    start = threading.Barrier(threads)
    spans = [None] * threads

    def worker(k):
        s = make_workload(scenes, workload, n_tile, k, 1)
        orc = oracle.Collider(4.0, 0.01)
        orc.add_hitboxes(s)
        orc.bulk_update(end_time=DT)
        T = min(DT, orc.next_time())
        orc.set_time(T)
        start.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.bulk_update(end_time=T + DT)
            T = min(T + DT, orc.next_time())
            orc.set_time(T)
        spans[k] = (t0, time.perf_counter())

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(worker, range(threads)))
    el = max(t1 for _, t1 in spans) - min(t0 for t0, _ in spans)
    return threads * n_tile * steps / el


def run_reference(args, out):
    """Reference arm: ONE single-threaded collider on the full scene of this config — same config as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scenes = importlib.import_module("collider-rs_b200.scenes")
    scene = make_workload(scenes, args.workload, args.n, 0, 1)
    warm = max(1, min(args.warmup, 2))
    ups, sps, ms, t_setup = oracle_bulk_timing(scene, args.steps, warm)
    sample = (f"ONE single-threaded collider (the reference is single-threaded) on the full {args.n}-hitbox {args.workload} scene, {warm} warm-up + "
              f"{args.steps} bulk-equivalent steps (ascending set_hitbox_vel loop, end_time = T + {DT}); add_hitbox set-up {t_setup:.1f} s outside the timed region")
    line = {
        "impl": "reference", "metric": METRIC, "value": ups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args),
        "pair_solves_per_s": sps,
        "cpu_baseline": {"value": ups, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "note": "C++17 restatement of the reference (oracle/); the Rust crate cannot be built here (no rustc/cargo)"},
        "e2e": {"value": ups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores_available": os.cpu_count(),
    }
    if not args.no_all_threads:
        threads = max(1, os.cpu_count() or 1)
        n_tile = max(10_000, min(args.n, 200_000) // max(1, threads // 4))
        line["cpu_all_threads_tiled"] = {
            "value": oracle_tiles_all_threads(scenes, args.workload, n_tile, 3, threads), "unit": UNIT, "cores": threads,
            "note": f"NOT the config: {threads} independent colliders side by side, one per host thread, each a {n_tile}-hitbox tile of the same density (no cross-tile pairs)"}
    print(json.dumps(line), file=out)


# ------------------------------------------------------------------------------------------------ GPU side
def main():
    # stdout carries exactly ONE JSON line: route everything else that native libraries print to fd 1 (e.g. NCCL's
    # version banner) to stderr, and keep the real stdout for the result line
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


class Job:
    """One workload on this rank: engine (single or sharded), step function, overlap hand-over."""

    def __init__(self, cb, scenes, sharding, dist, workload, n, rank, world, local_rank, grid="", slab=0, drain=False):
        self.cb, self.rank, self.world, self.n, self.dist = cb, rank, world, n, dist
        self.drain = drain and world == 1
        self.scene = make_workload(scenes, workload, n, rank, world)
        self.sh = None
        if world == 1:
            self.eng = cb.Engine(scenes.CELL_WIDTH, scenes.PADDING, device=local_rank)
            if grid:
                self.eng.set_grid(*[int(v) for v in grid.split(",")])
            self.eng.upload_state(**self.scene)
            if self.drain:
                self.eng.set_window_drain(True)
            self.step = self.eng.bulk_update
        else:
            # spatial sharding: one column strip of the world per rank; halo records and next-event keys are pushed
            # into the peers' memory by the step kernels (CB200_EXCHANGE=nccl: two NCCL all-gathers instead)
            width = C4_SIDE / world if workload == "c4" else math.sqrt(n / DENSITY[workload])
            bounds = sharding.strip_bounds(world, 0.0, width * world, scenes.CELL_WIDTH)
            self.sh = sharding.ShardedEngine(cb, scenes.CELL_WIDTH, scenes.PADDING, local_rank, bounds, n * world, sharding.Exchange(),
                                             direct=os.environ.get("CB200_EXCHANGE", "direct") != "nccl")
            self.eng = self.sh.eng
            if slab:
                self.eng.shard_set_slab(slab)
            if grid:
                self.eng.set_grid(*[int(v) for v in grid.split(",")])
            self.sh.upload_state(**self.scene)
            self.strip = (width * rank, width * (rank + 1))
            self.step = lambda T, end_time, **vel: self.sh.bulk_update(T, end_time, **vel)
        self.clock = DT

    def first_step(self):
        """Pairs already in contact at T = 0 come back as Collide events at T; they become the overlap set (what add_hitbox's
        immediate-overlap list does in the reference, collider.rs:355-365).  Sharded: a pair whose owner cell may drift into
        another strip is handed to every rank (each rank solves only the pairs whose owner column it holds)."""
        cb = self.cb
        res = self.step(0.0, end_time=DT)
        ev = self.eng.fetch_events()
        touching = ev[(ev["kind"] == cb.EV_COLLIDE) & (ev["time"] == 0.0)]
        ta, tb = touching["a"].copy(), touching["b"].copy()
        if self.world > 1:
            lo, hi = np.uint64(self.rank * self.n), np.uint64((self.rank + 1) * self.n)
            px = self.scene["pos_x"]

            def near_edge(ids):
                local = (ids >= lo) & (ids < hi)
                x = px[np.where(local, ids - lo, 0).astype(np.int64)]
                return ~local | (x < self.strip[0] + 16.0) | (x > self.strip[1] - 16.0)

            m = near_edge(ta) | near_edge(tb)
            parts = [None] * self.world
            self.dist.all_gather_object(parts, (ta[m], tb[m]))
            ta = np.concatenate([ta[~m]] + [p[0] for p in parts])
            tb = np.concatenate([tb[~m]] + [p[1] for p in parts])
        self.eng.set_overlaps(ta, tb)
        return res

    def device_step(self):
        T = self.clock
        self.clock = T + DT  # the next step time is exactly the end_time installed by the previous step
        return self.step(T, end_time=self.clock)


def events_digest(ev):
    h = hashlib.sha256()
    for k in ("time", "a", "b", "kind"):
        h.update(np.ascontiguousarray(ev[k]).tobytes())
    return h.hexdigest()


def parity_check(cb, scenes, sharding, dist, workload, rank, world, local_rank, n_small=20_000):
    """A down-scaled copy of the workload through the SAME code path (sharded exchange included), two steps, against the oracle:
    the union of the ranks' sorted event lists must be the oracle's list, bit for bit."""
    from oracle import oracle_py as oracle

    job = Job(cb, scenes, sharding, dist, workload, n_small, rank, world, local_rank)
    full = {k: np.concatenate([make_workload(scenes, workload, n_small, r, world)[k] for r in range(world)]) for k in job.scene}
    ok, n_events = True, 0
    orc = None
    if rank == 0:
        orc = oracle.Collider(scenes.CELL_WIDTH, scenes.PADDING)
        orc.add_hitboxes(full)
    T = 0.0
    n_steps = 2 if world == 1 else 3
    # config 5's fast shapes have a cell period of 0.04: with the benchmark's window of 0.1 a Reiterate — a single-hitbox
    # update the reference runs between two bulk steps (the Collider API's job) — would fall inside it, and this
    # harness chains bare bulk steps.  The check therefore uses a window without one, like tests/golden's clustered case.
    dt = 0.03 if workload == "c5" else DT
    moved = 0
    for step in range(n_steps):
        if rank == 0:
            oa, ob = orc.dump_overlaps()
        else:
            oa = ob = None
        if world > 1:
            box = [(oa, ob)]
            dist.broadcast_object_list(box, src=0)
            oa, ob = box[0]
        job.eng.set_overlaps(oa, ob)
        job.step(T, end_time=T + dt)
        ev = job.eng.fetch_events().copy()
        parts = [ev]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, ev)
            # owner migration between the steps (margin 0: every row whose centre column left its rank's strip changes rank);
            # the following steps run on the migrated tables and are held to the same comparison
            moved += job.sh.migrate(0)[0]
        if rank == 0:
            orc.bulk_update(end_time=T + dt)
            t, idx, kind, a, b = orc.dump_events()
            evs = np.concatenate(parts)
            cls, kbit = np.where(evs["kind"] == 0, 0, 1), np.where(evs["kind"] == 1, 1, 0)
            evs = evs[np.lexsort((evs["b"], kbit, evs["a"], cls, evs["time"]))]
            same = (len(evs) == len(t) and np.array_equal(evs["kind"], kind) and np.array_equal(evs["a"], a) and
                    np.array_equal(evs["b"], np.where(kind == 0, 0, b)) and np.array_equal(evs["time"].view(np.uint64), t.view(np.uint64)))
            ok = ok and bool(same)
            n_events += len(t)
            T += dt
            while True:
                tt = min(orc.next_time(), T)
                orc.set_time(tt)
                while orc.next() is not None:
                    pass
                if tt >= T:
                    break
        else:
            T += dt
    out = {"scene": f"{workload} at {n_small} hitboxes per GPU, {world} GPU(s), {n_steps} chained steps of {dt}, same code path", "events_compared": int(n_events),
           "bit_identical_to_oracle": ok}
    if world > 1:
        counts = [None] * world
        dist.all_gather_object(counts, moved)
        out["rows_migrated_between_the_steps_all_ranks"] = int(sum(counts))
    return out


def timed_passes(job, barrier, steps, repeats, body, tail=None):
    """`repeats` runs of `steps` calls of body() (+ tail(): whatever the last call left in flight); returns (per-run wall
    seconds, accumulated dict of body's counters)."""
    import gc

    walls, acc = [], {}
    for _ in range(repeats):
        gc.collect()
        gc.disable()  # (a collection pause on one rank is paid by every rank of a sharded step)
        try:
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                for k, v in body().items():
                    acc[k] = acc.get(k, 0) + v
            if tail is not None:
                tail()
            barrier()
            walls.append(time.perf_counter() - t0)
        finally:
            gc.enable()
    return walls, acc


class EventPins:
    """Two page-locked buffers for the pipelined event download (one is being filled while the other is read)."""

    def __init__(self, cb):
        self.cb, self.bufs, self.k, self.last_n = cb, [None, None], 0, 0

    def next(self, n_events):
        self.k ^= 1
        b = self.bufs[self.k]
        if b is None or b.array.shape[0] < n_events:
            b = self.bufs[self.k] = self.cb.PinnedArray(max(int(n_events * 1.3) + 1024, 4096), self.cb.EVENT16_DTYPE)
        self.last_n = int(n_events)
        return b.array

    def last(self):
        return self.bufs[self.k].array


def events_digest16(ev):
    h = hashlib.sha256()
    for k in ("time", "a", "b"):
        h.update(np.ascontiguousarray(ev[k]).tobytes())
    return h.hexdigest()


def collider_rows(cb, scenes, device):
    """Configs 1 and 2 through the Collider API (event-driven: README.md:66-80 loop, velocity halving on Collide), device
    Collider vs the oracle port, both driven by a native loop (cb200_collider_run_halving_loop / orc_halving_loop)."""
    from oracle import oracle_py as oracle

    col = importlib.import_module("collider-rs_b200.collider")
    rows = {}
    # config 1: 1,000 shapes to t = 100
    s = scenes.c1()
    orc = oracle.Collider(scenes.CELL_WIDTH, scenes.PADDING)
    orc.add_hitboxes(s)
    orc.stats(reset=True)
    t0 = time.perf_counter()
    ev_o, col_o = orc.halving_loop(100.0)
    t_o = time.perf_counter() - t0
    upd_o = orc.stats()["updates"]
    d = col.Collider(scenes.CELL_WIDTH, scenes.PADDING, device=device)
    d.add_hitboxes(s)
    l0 = d.launch_count()
    t0 = time.perf_counter()
    ev_d, col_d = d.halving_loop(100.0)
    t_d = time.perf_counter() - t0
    rows["config1"] = {"workload": "1,000 mixed circles/squares in 200x200, README loop to t=100, velocity halving on Collide",
                       "events": ev_d, "events_match_port": bool(ev_d == ev_o and col_d == col_o),
                       "device_collider": {"events_per_s": ev_d / t_d, "seconds": t_d, "gpu_launches": d.launch_count() - l0},
                       "cpu_port_single_thread": {"events_per_s": ev_o / t_o, "hitbox_updates_per_s": upd_o / t_o, "seconds": t_o}}
    # config 2: 100k shapes, bulk step every 0.1 with new velocities, events processed in between
    s = scenes.c2()
    n = len(s["id"])
    steps = 5
    orc = oracle.Collider(scenes.CELL_WIDTH, scenes.PADDING)
    orc.add_hitboxes(s, end_time=DT)
    d = col.Collider(scenes.CELL_WIDTH, scenes.PADDING, device=device)
    d.add_hitboxes(s, end_time=DT)
    vel = [(scenes.uniform(300 + k, n, 1, -2.0, 2.0), scenes.uniform(300 + k, n, 2, -2.0, 2.0)) for k in range(steps)]
    pin = [(cb.PinnedArray(n, np.float64), cb.PinnedArray(n, np.float64)) for _ in range(steps)]  # the device arm reads page-locked arrays
    for k in range(steps):
        pin[k][0].array[:] = vel[k][0]
        pin[k][1].array[:] = vel[k][1]
    out = {}
    for name, c in (("cpu_port_single_thread", orc), ("device_collider", d)):
        T, ev_n, t_bulk, t_loop = 0.0, 0, 0.0, 0.0
        for k in range(steps):
            t0 = time.perf_counter()
            vx, vy = (pin[k][0].array, pin[k][1].array) if name == "device_collider" else vel[k]
            c.bulk_update(end_time=T + DT, vx=vx, vy=vy)
            t1 = time.perf_counter()
            e, _ = c.halving_loop(T + DT)
            t2 = time.perf_counter()
            ev_n += e
            t_bulk += t1 - t0
            t_loop += t2 - t1
            T += DT
        out[name] = {"hitbox_updates_per_s": n * steps / (t_bulk + t_loop), "bulk_ms_per_step": t_bulk / steps * 1e3, "events": ev_n,
                     "events_per_s": ev_n / t_loop if t_loop > 0 else None, "event_loop_ms_per_step": t_loop / steps * 1e3}
    # the drop-in bulk entry point at the headline size: cb200_collider_bulk_update on 1M hitboxes (page-locked velocity arrays)
    s1 = scenes.c3(1_000_000)
    d1 = col.Collider(scenes.CELL_WIDTH, scenes.PADDING, device=device)
    d1.add_hitboxes(s1, end_time=DT)
    p1 = (cb.PinnedArray(1_000_000, np.float64), cb.PinnedArray(1_000_000, np.float64))
    p1[0].array[:] = s1["vel_x"]
    p1[1].array[:] = s1["vel_y"]
    for _ in range(5):
        d1.bulk_update(end_time=DT, vx=p1[0].array, vy=p1[1].array)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        d1.bulk_update(end_time=DT, vx=p1[0].array, vy=p1[1].array)
    t_api = (time.perf_counter() - t0) / reps * 1e3
    # the same with the upload of call k+1's velocities staged beside call k (cb200_stage_velocities on the collider's context)
    d1.stage_velocities(vx=p1[0].array, vy=p1[1].array)
    for _ in range(4):  # (warm-up: the step graph is captured once per staging buffer)
        d1.stage_velocities(vx=p1[0].array, vy=p1[1].array)
        d1.bulk_update(end_time=DT)
    t0 = time.perf_counter()
    for _ in range(reps):
        d1.stage_velocities(vx=p1[0].array, vy=p1[1].array)
        d1.bulk_update(end_time=DT)
    t_api_staged = (time.perf_counter() - t0) / reps * 1e3
    d1.bulk_update(end_time=DT)
    rows["collider_api_bulk_update_1m"] = {"workload": "cb200_collider_bulk_update on 1,000,000 hitboxes (config 3 scene), new velocities from page-locked host arrays (16 MB H2D per call), "
                                                       "queue reset to the device-sorted run",
                                           "ms_per_call": t_api, "ms_per_call_velocities_staged_ahead": t_api_staged,
                                           "hitbox_updates_per_s": 1_000_000 / (t_api_staged * 1e-3)}
    del d1
    rows["config2"] = {"workload": f"100k mixed circles/rects, bulk update every {DT} with new velocities (host arrays), events between the steps "
                                   f"processed through next() with velocity halving on Collide, {steps} steps",
                       "events_match_port": bool(out["device_collider"]["events"] == out["cpu_port_single_thread"]["events"]), **out}
    return rows


def config4_block(cb, scenes, sharding, dist, rank, world, local_rank, barrier, steps=5, warmup=3):
    """BASELINE config 4: 16,777,216 shapes in 4096 x 4096 (1 per unit^2), split over the N column strips (N=1: the one-GPU anchor)."""
    import torch

    n = C4_TOTAL // world
    old = os.environ.get("CB200_DENSE_RANK")
    os.environ["CB200_DENSE_RANK"] = "4"  # dense slot runs from the first step (the automatic choice needs one step of history)
    try:
        job = Job(cb, scenes, sharding, dist, "c4", n, rank, world, local_rank, slab=max(131072, n // 32) if world > 1 else 0)
    finally:
        if old is None:
            os.environ.pop("CB200_DENSE_RANK", None)
        else:
            os.environ["CB200_DENSE_RANK"] = old
    job.first_step()
    job.eng.set_stage_timing(False)
    for _ in range(warmup):
        job.device_step()
    job.eng.reset_timing()
    solves = [0]

    def body():
        r = job.device_step()
        solves[0] += r.n_collide_solves + r.n_separate_solves
        return {}

    walls, _ = timed_passes(job, barrier, steps, 1, body)
    dev_ms, _ = job.eng.step_timing()
    ev_bytes = [0]
    pins = EventPins(cb)
    failures = [0]

    def body_fetch():
        r = job.device_step()
        try:
            job.eng.fetch_events_end()
        except cb.ColliderError:
            failures[0] += 1
        ev_bytes[0] += 16 * job.eng.fetch_events_begin(pins.next(r.n_events))
        return {}

    body_fetch()
    walls_t, _ = timed_passes(job, barrier, steps, 1, body_fetch, tail=job.eng.fetch_events_end)
    t = torch.tensor([walls[0], walls_t[0], dev_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([float(solves[0])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    wall, wall_t, dev = t.tolist()
    fallbacks = job.eng.event_sort_fallbacks()
    del job
    return {"workload": f"{C4_TOTAL} shapes in 4096 x 4096 (1 per unit^2), {world} column strip(s) of {n} hitboxes, bulk step every {DT}, dense slot runs through k_solve_dense",
            "n_gpus": world, "hitboxes_total": C4_TOTAL, "steps": steps, "warmup": warmup,
            "device_ms_per_step": dev, "ms_per_step_device_resident": wall / steps * 1e3, "hitbox_updates_per_s": C4_TOTAL * steps / wall,
            "pair_solves_per_s": s.item() / wall, "t_step_ms_with_event_list": wall_t / steps * 1e3, "d2h_bytes_per_step_rank0": int(ev_bytes[0] / (steps + 1)),
            "event_sort_fallbacks": fallbacks, "pipelined_fetch_failures": failures[0], "scaling": "strong (fixed 16,777,216 hitboxes; compare with the n_gpus=1 line's block)"}


def migration_block(cb, scenes, sharding, dist, args, rank, world, local_rank, barrier, steps=64, every=16):
    """N > 1: a run in which the residents follow their strips (SURVEY.md 8e owner migration): every `every` steps each rank
    hands the rows whose centre left its strip by more than one column to the rank that holds them (ShardedEngine.migrate —
    collective, through the host).  Hitboxes keep their velocities, so they drift up to 0.2 units per step."""
    import torch

    job = Job(cb, scenes, sharding, dist, args.workload, args.n, rank, world, local_rank, grid=args.grid)
    job.first_step()
    job.eng.set_stage_timing(False)
    for _ in range(args.warmup):
        job.device_step()
    acc = {"out": 0, "in": 0, "mig_s": 0.0}

    def body():
        for _ in range(every):
            job.device_step()
        t0 = time.perf_counter()
        o, i = job.sh.migrate(1)
        acc["mig_s"] += time.perf_counter() - t0
        acc["out"] += o
        acc["in"] += i
        return {}

    walls, _ = timed_passes(job, barrier, steps // every, 1, body)
    rows = job.eng.shard_rows()
    t = torch.tensor([walls[0], acc["mig_s"]], dtype=torch.float64, device="cuda")
    c = torch.tensor([float(acc["out"]), float(acc["in"]), float(rows["rows"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    wall, mig = t.tolist()
    n_out, n_in, n_rows = c.tolist()
    calls = steps // every
    return {"what": f"device-resident steps with owner migration every {every} steps (margin 1 column): rows whose centre left their rank's strip move to the rank that holds them",
            "steps": calls * every, "migrations": calls, "rows_moved_all_ranks": int(n_out), "rows_received_all_ranks": int(n_in), "rows_resident_all_ranks": int(n_rows),
            "ms_per_migration_max_rank": mig / calls * 1e3, "ms_per_step_including_migration": wall / (calls * every) * 1e3}


def steady_state_block(cb, scenes, sharding, dist, args, rank, world, local_rank, barrier, K):
    """The same workload with the window drain on (cb200_set_window_drain): the pair events inside each step's window are
    processed on the device (chains included) and the overlap set stays device-resident, as in a run whose caller pops every
    event before the next step.  The headline run never processes events, so pairs that came into contact keep returning as
    zero-delay Collide events (12x the event volume); this block is the steady state."""
    job = Job(cb, scenes, sharding, dist, args.workload, args.n, rank, world, local_rank, grid=args.grid, drain=True)
    eng = job.eng
    job.first_step()
    eng.set_stage_timing(False)
    for _ in range(max(args.warmup, 17)):
        job.device_step()
    eng.reset_timing()
    acc = {"events": 0, "solves": 0, "followups": 0}

    def body():
        r = job.device_step()
        acc["events"] += r.n_events
        acc["solves"] += r.n_collide_solves + r.n_separate_solves
        acc["followups"] += eng.last_drain()["followup_events"]
        return {}

    walls, _ = timed_passes(job, barrier, K, 1, body)
    dev_ms, _ = eng.step_timing()
    pins = EventPins(cb)

    def body_t():
        r = job.device_step()
        eng.fetch_events_end()
        eng.fetch_events_begin(pins.next(r.n_events))
        return {}

    body_t()
    walls_t, _ = timed_passes(job, barrier, K, 3, body_t, tail=eng.fetch_events_end)
    walls_t = [sorted(walls_t)[1]]  # (median of three runs: a run in which the overlap list is compacted re-captures the step graph)
    d = eng.last_drain()
    return {"what": "window drain on: events of each step's window processed on the device, overlap set device-resident (cb200_set_window_drain)",
            "device_ms_per_step": dev_ms, "t_step_ms": walls_t[0] / K * 1e3, "hitbox_updates_per_s": args.n * K / walls_t[0],
            "events_per_step": acc["events"] / K, "followup_events_per_step": acc["followups"] / K, "pair_solves_per_step": acc["solves"] / K,
            "overlapping_pairs": d["pairs_now"], "reiterates_inside_windows": d["reiterates_in_windows"], "path": eng.last_step_path()}


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--hitboxes", dest="n", type=int, default=1_000_000, help="hitboxes per GPU")
    ap.add_argument("--workload", default="c3", choices=["c3", "c3_dense", "c5", "c3_groups"])
    ap.add_argument("--repeats", type=int, default=3, help="runs of K steps; the median run is reported")
    ap.add_argument("--ref-sample", type=int, default=500_000, help="hitboxes in the cpu_baseline sample of the GPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-all-threads", action="store_true", help="--impl reference: skip the separately named all-threads row")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-collider-rows", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-migration", action="store_true", help="skip the owner-migration block of the N>1 lines")
    ap.add_argument("--no-steady-state", dest="steady_state", action="store_false", help="skip the window-drain (steady-state) block of the N=1 line")
    ap.add_argument("--drain", action="store_true", help="window drain: pair events inside each step's window processed on the device, overlap set device-resident")
    ap.add_argument("--grid", default="", help="force the device grid period: log2_w,log2_h")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args, out)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cb = importlib.import_module("collider-rs_b200")
    scenes = importlib.import_module("collider-rs_b200.scenes")
    sharding = importlib.import_module("collider-rs_b200.sharding")
    hbm_peak, peak_kind = load_peaks()
    n, K, R = args.n, args.steps, max(1, args.repeats)

    parity = None if args.no_parity else parity_check(cb, scenes, sharding, dist, args.workload, rank, world, local_rank)

    job = Job(cb, scenes, sharding, dist, args.workload, n, rank, world, local_rank, grid=args.grid, drain=args.drain)
    eng = job.eng
    job.first_step()
    eng.set_stage_timing(False)  # the timed region carries two CUDA events per step (start / end), none between the kernels;
    # the engine then replays the step as one CUDA graph — captured during the warm-up (once per column-buffer set: the table
    # re-sort every 16 steps alternates between two)
    for _ in range(max(args.warmup, 17)):
        job.device_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:  # (one sampler per job: eight ranks each forking nvidia-smi ten times a second disturb the very thing they watch)
        sampler.start()

    # (1) device-resident step alone: kernels of the bulk step, CUDA events per step
    launches0 = eng.launch_count()
    eng.reset_timing()

    def body_dev():
        r = job.device_step()
        return {"solves": r.n_collide_solves + r.n_separate_solves, "entries": r.n_entries, "cells": r.n_cells, "events": r.n_events,
                "work": r.n_work_items, "seps": r.n_separate_solves}

    walls_dev, acc = timed_passes(job, barrier, K, R, body_dev)
    launches = (eng.launch_count() - launches0) // R
    dev_ms, _ = eng.step_timing()  # average over all R*K steps

    # (2) t_step of SURVEY 8(d): step + ordered queue (device sort) + D2H of the event list into pinned host memory.  The
    # download of step k's list (16-byte events, own copy stream) runs beside step k+1: cb200_fetch_events_begin / _end
    pins = EventPins(cb)
    digest = [None]
    fetch_failures = [0]

    def fetch_pipelined(n_events):
        try:
            eng.fetch_events_end()  # the previous step's download
        except cb.ColliderError:
            fetch_failures[0] += 1
        return eng.fetch_events_begin(pins.next(n_events))

    phase = {"step": 0.0, "fetch_end": 0.0, "fetch_begin": 0.0}  # host seconds of rank-local phases of the t_step loop

    def body_tstep():
        t0 = time.perf_counter()
        r = job.device_step()
        t1 = time.perf_counter()
        try:
            eng.fetch_events_end()  # the previous step's download
        except cb.ColliderError:
            fetch_failures[0] += 1
        t2 = time.perf_counter()
        got = eng.fetch_events_begin(pins.next(r.n_events))
        t3 = time.perf_counter()
        phase["step"] += t1 - t0
        phase["fetch_end"] += t2 - t1
        phase["fetch_begin"] += t3 - t2
        return {"ev_bytes": got * 16}

    body_tstep()
    eng.fetch_events_end()
    digest[0] = events_digest16(pins.last()[: pins.last_n])
    for k in phase:
        phase[k] = 0.0
    walls_t, acc_t = timed_passes(job, barrier, K, R, body_tstep, tail=eng.fetch_events_end)
    phase_ms = {k: v / (K * R) * 1e3 for k, v in phase.items()}

    # (3) per-kernel durations: K steps with CUDA events between the stages (each event adds ~2 us of stream time, so these are
    # upper bounds and their sum exceeds device_ms_per_step)
    eng.set_stage_timing(True)
    for _ in range(K):
        job.device_step()
    _, stage_ms = eng.step_timing()
    eng.set_stage_timing(False)

    # (4) e2e: new velocities arrive from pinned host arrays every step, sorted events are read back.  Both transfers are
    # pipelined against the compute: the upload of step k+1's velocities (cb200_stage_velocities, own stream) and the download of
    # step k-1's events run beside step k
    pvx, pvy = cb.PinnedArray(n, np.float64), cb.PinnedArray(n, np.float64)
    pvx.array[:] = job.scene["vel_x"]
    pvy.array[:] = job.scene["vel_y"]
    eng.stage_velocities(vel_x=pvx.array, vel_y=pvy.array)

    def body_e2e():
        eng.stage_velocities(vel_x=pvx.array, vel_y=pvy.array)  # the NEXT step's velocities
        T = job.clock
        job.clock = T + DT
        r = job.step(T, end_time=job.clock)                      # consumes the set staged one iteration ago
        got = fetch_pipelined(r.n_events)
        return {"d2h": got * 16 + C_sizeof_result(cb)}

    for _ in range(3):
        body_e2e()
    walls_e, acc_e = timed_passes(job, barrier, K, R, body_e2e, tail=eng.fetch_events_end)
    job.step(job.clock, end_time=job.clock + DT)  # (consume the set that is still staged)
    job.clock += DT
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    med = lambda w: float(np.median(w))
    times = torch.tensor([med(walls_dev), med(walls_t), med(walls_e), dev_ms, min(walls_t), max(walls_t), phase_ms["step"], phase_ms["fetch_end"], phase_ms["fetch_begin"]],
                         dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(acc["solves"]), float(acc["entries"]), float(acc["cells"]), float(acc["events"])], dtype=torch.float64, device="cuda")
    digests = [digest[0]]
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        digests = [None] * world
        dist.all_gather_object(digests, digest[0])
    wall_dev, wall_t, wall_e, dev_ms_max, wall_t_min, wall_t_max, ph_step, ph_end, ph_begin = times.tolist()
    solves_all = sums.tolist()[0]

    sort_fallbacks = eng.event_sort_fallbacks()
    steady = None
    if args.steady_state and world == 1:
        del job, eng
        steady = steady_state_block(cb, scenes, sharding, dist, args, rank, world, local_rank, barrier, K)
        job = eng = None
    migration = None
    if world > 1 and not args.no_migration:
        job = eng = None
        migration = migration_block(cb, scenes, sharding, dist, args, rank, world, local_rank, barrier)
    c4 = None
    if not args.no_config4 and args.workload == "c3":
        job = eng = None
        c4 = config4_block(cb, scenes, sharding, dist, rank, world, local_rank, barrier)

    if rank == 0:
        KR = K * R
        E, Cc, V, P, W, Q = acc["entries"] / KR, acc["cells"] / KR, acc["events"] / KR, acc["solves"] / KR, acc["work"] / KR, acc["seps"] / KR
        # B_step = 184 N + 104 E + 16 C + 24 V — state in/out 152 N, rect re-reads of the sort passes 32 N, entry
        # write + read 8 E, records gathered per entry in the pair stage 96 E, cell count/offset 16 C, events 24 V
        b_step = 184.0 * n + 104.0 * E + 16.0 * Cc + 24.0 * V
        t_dev = dev_ms_max * 1e-3
        achieved = b_step / t_dev / 1e9
        stage_ms.setdefault("exchange", 0.0)
        kernel_terms = {  # the same B_step attributed to the kernels that move it
            "stage_a": ("k_stage_a", 152.0 * n),
            "scan": ("k_step_zero + k_scan_slots", 16.0 * Cc),
            "scatter": ("k_scatter", 32.0 * n + 4.0 * E),
            "candidates": ("k_enum_pairs", 4.0 * E),
            "solve": ("k_solve (+ k_solve_dense)", 96.0 * E + 24.0 * V),
        }
        stage_ms["scan"] = stage_ms.get("scan", 0.0) + stage_ms.get("pre", 0.0)
        per_kernel = {name: {"ms": stage_ms[k], "survey_bytes": b, "frac": b / (stage_ms[k] * 1e-3) / 1e9 / hbm_peak}
                      for k, (name, b) in kernel_terms.items() if stage_ms.get(k)}
        dom = max(kernel_terms, key=lambda k: stage_ms.get(k, 0.0))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.workload == "c3" and n == 1_000_000:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get("whole_step_bytes") or sum(v for v in tj.values() if isinstance(v, (int, float)))
        line = {
            "metric": METRIC, "value": n * world * K / wall_t, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": wall_t / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args),
            "timing": {"repeats": R, "reported": "median run of `repeats` runs of K steps, max over ranks",
                       "value": "t_step of SURVEY 8(d): device-resident state; bulk step + ordered event queue (device sort) + D2H of the event list (16-byte "
                                "events) into pinned host memory, the download of step k overlapping step k+1; wall clock around the C-ABI calls, bracketed by "
                                "barrier + cuda sync",
                       "t_step_ms_min_max_over_runs": [wall_t_min / K * 1e3, wall_t_max / K * 1e3],
                       "t_step_host_phases_ms_max_rank": {"bulk_step_until_result": ph_step, "fetch_events_end": ph_end, "fetch_events_begin": ph_begin},
                       "l2": "working set per step (SoA state + records + slot table + entries, > 300 MB) exceeds the 126 MB L2; no flush needed"},
            "device_ms_per_step": dev_ms_max,
            "device_resident_step_only": {"value": n * world * K / wall_dev, "unit": UNIT, "ms_per_step": wall_dev / K * 1e3,
                                          "note": "the bulk step without materialising / copying the event list (next event and counters only)"},
            "pair_solves_per_s": solves_all / R / wall_t,
            "stage_ms": stage_ms,
            "per_step": {"cell_entries": E, "cells": Cc, "work_items": W, "pair_solves": P, "events": V, "event_list_bytes": acc_t["ev_bytes"] / KR},
            "gpu_launches": int(launches),
            "e2e": {"value": n * world * K / wall_e, "unit": UNIT, "h2d_bytes_per_step": int(16 * n), "d2h_bytes_per_step": int(acc_e["d2h"] / KR),
                    "ms_per_step": wall_e / K * 1e3, "event_sort_fallbacks": sort_fallbacks, "pipelined_fetch_failures": fetch_failures[0],
                    "pipeline": "velocities of step k+1 staged (H2D stream) and events of step k-1 downloaded as 16-byte records (D2H stream) beside step k"},
            "roofline": {"bound": "hbm", "kernel": "bulk step (all kernels of one cb200_bulk_update)", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
                         "bytes_model": "SURVEY 8(d): B_step = 184 N + 104 E + 16 C + 24 V", "b_step_bytes": b_step,
                         "frac_on_t_step": b_step / (wall_t / K) / 1e9 / hbm_peak,
                         "io_lower_bound_bytes": 152.0 * n + 24.0 * V, "io_lower_bound_frac": (152.0 * n + 24.0 * V) / t_dev / 1e9 / hbm_peak,
                         "dominant_kernel": per_kernel.get(kernel_terms[dom][0]) and {"name": kernel_terms[dom][0], **per_kernel[kernel_terms[dom][0]]},
                         "per_kernel": per_kernel},
            "events_sha256_first_timed_step": digests,
            "clocks": sampler.summary(),
        }
        fpath = os.path.join(ROOT, "profiles", "fp64_peak.json")
        if args.workload == "c3_dense" and os.path.exists(fpath):
            # dense scenes: the pair math dominates — state the FP64 bound next to the HBM one.  FP64 arithmetic
            # instructions (DADD + DMUL + 2 DFMA) per pair solve come from the committed ncu capture of this workload
            # (profiles/r2/dense_fp64.txt); the peak is the box's measured separate-DMUL+DADD rate (the engine is built -fmad=false)
            with open(fpath) as f:
                fp = json.load(f)
            flop_per_solve = fp["fp64_flop_per_step_1m_dense"] / 29_154_750.0
            ach = flop_per_solve * P / t_dev / 1e12
            line["roofline_fp64"] = {"bound": "fp64 (DMUL + DADD, no contraction)", "achieved": ach, "peak": fp["fp64_peak_dmul_dadd_tflops"], "unit": "TFLOP/s",
                                     "frac": ach / fp["fp64_peak_dmul_dadd_tflops"], "fp64_flop_per_pair_solve": flop_per_solve,
                                     "note": "neither roofline binds the dense step: the FP64 pipe is 34 % busy (comparisons, min/max, conversions and the "
                                             "division / sqrt sequences share it) and the kernel is issue- and latency-bound (profiles/r2/dense_fp64.txt, DESIGN 4)"}
        if parity is not None:
            line["parity"] = parity
        if world > 1:
            line["config"]["parallelism"] = (f"spatial sharding: {world} column strips; halo records (96 B) and next-event keys "
                                             + ("pushed into the peers' memory over NVLink by the step kernels (cudaIpc), one host sync per step, no collective on the data path"
                                                if os.environ.get("CB200_EXCHANGE", "direct") != "nccl" else "exchanged by two NCCL all-gathers"))
        if steady is not None:
            line["steady_state_window_drain"] = steady
        if migration is not None:
            line["owner_migration"] = migration
        if c4 is not None:
            line["config4"] = c4
        if not args.no_cpu_baseline and world == 1:
            n_sample = min(n, args.ref_sample)
            ups, sps, ms, _ = oracle_bulk_timing(make_workload(scenes, args.workload, n_sample, 0, 1), 5, 1)
            line["cpu_baseline"] = {"value": ups, "unit": UNIT, "cores": 1, "kind": "port", "pair_solves_per_s": sps,
                                    "sample": f"{n_sample} hitboxes of the same scene density, 5 bulk-equivalent steps (ascending set_hitbox_vel loop), single thread; "
                                              f"`--impl reference` runs the full {n}-hitbox scene",
                                    "host_cores_available": os.cpu_count()}
        if not args.no_collider_rows and world == 1:
            line.update(collider_rows(cb, scenes, local_rank))
        print(json.dumps(line), file=out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def C_sizeof_result(cb):
    import ctypes

    return ctypes.sizeof(cb.StepResult)


if __name__ == "__main__":
    main()
