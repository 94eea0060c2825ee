#!/usr/bin/env python
"""bench.py — hitbox updates/s + pair TOI solves/s of the bulk step (BASELINE.json metric).

A "step" = one cb200_bulk_update over the whole scene: every hitbox rebased to T, re-binned into the grid,
every candidate pair's collide_time solved (plus separate_time for the overlapping pairs), next event
min-reduced.  Workload at N=1: config 3 (1M mixed circles/rects, uniform density 0.025/unit^2, cell 4,
padding 0.01, bulk step every 0.1 time units).  N>1: weak scaling, one 1M-hitbox column strip per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n HITBOXES] [--workload c3|c3_dense|c5]

Prints ONE JSON line (rank 0).  `value` = device-resident whole-job throughput; `e2e` = the same metric with
new velocities coming from pinned HOST arrays each step and the sorted event list read back (H2D + D2H inside
the timed region); `roofline` = dominant kernel vs the measured HBM peak; `cpu_baseline` = the oracle (C++
restatement of the reference; Rust toolchain absent) on one host core, bounded sample.  `--impl reference` times that
oracle on ALL host threads (one independent Collider per thread: the reference itself is single-threaded).
"""
from __future__ import annotations

import argparse
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
METRIC = "hitbox updates/sec (full re-bin + pair TOI solve + next-event min per bulk step)"
UNIT = "hitbox updates/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_workload(scenes, name: str, n: int, rank: int, world: int):
    """Per-rank scene.  Weak scaling: rank r owns the column strip [r*W, (r+1)*W) of a world `world` strips wide."""
    if name == "c3":
        side = math.sqrt(n / 0.025)
        s = scenes.make_scene(n, side, side, seed=scenes.SEED_BASE + 3 + 1000 * rank, x0=side * rank)
    elif name == "c3_dense":
        side = math.sqrt(n / 1.0)
        s = scenes.make_scene(n, side, side, seed=scenes.SEED_BASE + 3 + 1000 * rank, x0=side * rank)
    elif name == "c5":
        s = scenes.c5(n)
        s["pos_x"] = s["pos_x"] + math.sqrt(n / 0.025) * rank
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
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def oracle_timing(scene_fn, n_sample: int, steps: int, warmup: int):
    """The reference's CPU implementation of the path (C++ restatement; Rust toolchain absent): the ascending
    set_hitbox_vel loop per step, single-threaded like the reference.  Returns (updates/s, solves/s, ms/step, stats)."""
    from oracle import oracle_py as oracle

    s = scene_fn(n_sample)
    orc = oracle.Collider(4.0, 0.01)
    orc.add_hitboxes(s)
    T = 0.0
    for _ in range(warmup):
        orc.bulk_update(end_time=T + DT)
        T = min(T + DT, orc.next_time())
        orc.set_time(T)
    orc.stats(reset=True)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        orc.bulk_update(end_time=T + DT)
        done += 1
        T = min(T + DT, orc.next_time())  # stay within the reference's set_time contract without draining events
        orc.set_time(T)
    el = time.perf_counter() - t0
    st = orc.stats()
    return n_sample * done / el, (st["collide_solves"] + st["separate_solves"]) / el, el / done * 1e3, st


def oracle_timing_threads(scene_fn, n_sample: int, steps: int, warmup: int, threads: int):
    """The same, on every host thread: the reference is single-threaded, so "all the host threads it can use" is one
    independent Collider per thread, each on its own tile (n_sample hitboxes, same density, different seed offset).  The C
    calls release the GIL.  Returns whole-host (updates/s, solves/s, ms per step of one collider)."""
    from concurrent.futures import ThreadPoolExecutor
    import threading as th

    from oracle import oracle_py as oracle

    oracle.lib()  # build / load once, before the threads start
    start = th.Barrier(threads)
    spans = [None] * threads

    def worker(k):
        s = scene_fn(n_sample, k)
        orc = oracle.Collider(4.0, 0.01)
        orc.add_hitboxes(s)
        T = 0.0
        for _ in range(warmup):
            orc.bulk_update(end_time=T + DT)
            T = min(T + DT, orc.next_time())
            orc.set_time(T)
        orc.stats(reset=True)
        start.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.bulk_update(end_time=T + DT)
            T = min(T + DT, orc.next_time())
            orc.set_time(T)
        t1 = time.perf_counter()
        st = orc.stats()
        spans[k] = (t0, t1, st["collide_solves"] + st["separate_solves"])

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(worker, range(threads)))
    el = max(t1 for _, t1, _ in spans) - min(t0 for t0, _, _ in spans)
    return threads * n_sample * steps / el, sum(sv for _, _, sv in spans) / el, el / steps * 1e3


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scenes = importlib.import_module("collider-rs_b200.scenes")
    threads = max(1, args.ref_threads or (os.cpu_count() or 1))
    # bounded: every thread steps its own tile of the scene; the tile shrinks with the thread count so that the run (scene
    # set-up included) stays within a few minutes on any host
    n_sample = max(10_000, min(args.n, args.ref_sample) // max(1, threads // 4))
    ups, sps, ms = oracle_timing_threads(lambda m, k: make_workload(scenes, args.workload, m, k, 1), n_sample, args.steps, max(1, min(args.warmup, 1)), threads)
    sample = (f"{threads} independent colliders side by side, one per host thread (the reference is single-threaded), each a tile of {n_sample} hitboxes of the "
              f"{args.workload} scene (same density), {args.steps} bulk-equivalent steps (ascending set_hitbox_vel loop)")
    line = {
        "impl": "reference", "metric": METRIC, "value": ups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {args.n} mixed circles/rects, cell_width 4, padding 0.01, bulk step every {DT}", "cpu_sample_hitboxes": n_sample * threads,
                   "cpu_threads": threads},
        "pair_solves_per_s": sps,
        "cpu_baseline": {"value": ups, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++17 restatement of the reference (oracle/); the Rust crate cannot be built here (no rustc/cargo)"},
        "e2e": {"value": ups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores_available": os.cpu_count(),
    }
    print(json.dumps(line), file=out)


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


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--hitboxes", dest="n", type=int, default=1_000_000, help="hitboxes per GPU")
    ap.add_argument("--workload", default="c3", choices=["c3", "c3_dense", "c5"])
    ap.add_argument("--ref-sample", type=int, default=200_000, help="hitboxes in the CPU sample")
    ap.add_argument("--ref-threads", type=int, default=0, help="--impl reference: host threads (0 = all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    hbm_peak, peak_kind = load_peaks()

    n = args.n
    scene = make_workload(scenes, args.workload, n, rank, world)
    halo = {"sent": 0, "received": 0, "steps": 0}
    if world == 1:
        eng = cb.Engine(scenes.CELL_WIDTH, scenes.PADDING, device=local_rank)
        if args.grid:
            eng.set_grid(*[int(v) for v in args.grid.split(",")])
        eng.upload_state(**scene)
        step_fn = eng.bulk_update
    else:
        # spatial sharding: one column strip of the world per rank, NCCL halo exchange between the two
        # phases of the step, global next event by an all-gather + min of the per-rank keys
        sharding = importlib.import_module("collider-rs_b200.sharding")
        side = math.sqrt(n / (1.0 if args.workload == "c3_dense" else 0.025))
        bounds = sharding.strip_bounds(world, 0.0, side * world, scenes.CELL_WIDTH)
        sh = sharding.ShardedEngine(cb, scenes.CELL_WIDTH, scenes.PADDING, local_rank, bounds, n * world, sharding.Exchange(),
                                    direct=os.environ.get("CB200_EXCHANGE", "direct") != "nccl")
        eng = sh.eng
        if args.grid:
            eng.set_grid(*[int(v) for v in args.grid.split(",")])
        sh.upload_state(**scene)

        def step_fn(T, end_time, **vel):
            return sh.bulk_update(T, end_time, **vel)

    # first step: pairs already in contact at T = 0 come back as Collide events at T; they become the overlap set
    # (what add_hitbox's immediate-overlap list does in the reference, collider.rs:355-365)
    res = step_fn(0.0, end_time=DT)
    ev = eng.fetch_events()
    touching = ev[(ev["kind"] == cb.EV_COLLIDE) & (ev["time"] == 0.0)]
    ta, tb = touching["a"].copy(), touching["b"].copy()
    # sharded: every rank keeps the overlapping pairs it owns (pairs whose owner cell drifts into another strip would
    # have to be handed over by the host; over this run's 0.1-unit steps that is a vanishing fraction)
    eng.set_overlaps(ta, tb)
    clock = [DT]  # the next step time is exactly the end_time installed by the previous step

    def device_step():
        T = clock[0]
        clock[0] = T + DT
        return step_fn(T, end_time=clock[0])

    eng.set_stage_timing(False)  # the timed region carries two CUDA events per step (start / end), none between the kernels;
    # the engine then replays the step as one CUDA graph — captured during the warm-up (once per column-buffer set: the table
    # re-sort every 16 steps alternates between two)
    for _ in range(max(args.warmup, 17)):
        res = device_step()
    launches0 = eng.launch_count()
    eng.reset_timing()
    sampler = ClockSampler(local_rank)
    sampler.start()
    solves = 0
    entries = cells = events = work = seps = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = device_step()
        solves += res.n_collide_solves + res.n_separate_solves
        entries += res.n_entries
        cells += res.n_cells
        events += res.n_events
        work += res.n_work_items
        seps += res.n_separate_solves
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    dev_ms, _ = eng.step_timing()
    # per-kernel durations: the same steps again with CUDA events between the stages (each event adds ~2 us of stream time,
    # so these are upper bounds and their sum exceeds device_ms_per_step)
    eng.set_stage_timing(True)
    for _ in range(args.steps):
        device_step()
    _, stage_ms = eng.step_timing()
    eng.set_stage_timing(False)

    # e2e: new velocities arrive from pinned host arrays every step, sorted events are read back
    pvx, pvy = cb.PinnedArray(n, np.float64), cb.PinnedArray(n, np.float64)
    pvx.array[:] = scene["vel_x"]
    pvy.array[:] = scene["vel_y"]

    def e2e_step():
        T = clock[0]
        clock[0] = T + DT
        r = step_fn(T, end_time=clock[0], vel_x=pvx.array, vel_y=pvy.array)
        evs = eng.fetch_events(pinned=True)  # sorted queue of the step, D2H into page-locked memory
        return r, evs

    for _ in range(3):
        e2e_step()
    d2h = 0
    e2e_each = []
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        ts = time.perf_counter()
        r, evs = e2e_step()
        d2h += evs.nbytes + C_sizeof_result(cb)
        e2e_each.append((time.perf_counter() - ts) * 1e3)
    barrier()
    e2e_wall = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join(timeout=2)

    times = torch.tensor([wall, e2e_wall, dev_ms], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(solves), float(entries), float(cells), float(events)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    wall_max, e2e_max, dev_ms_max = times.tolist()
    solves_all, entries_all, cells_all, events_all = sums.tolist()

    if rank == 0:
        K = args.steps
        E, Cc, V, P, W, Q = entries / K, cells / K, events / K, solves / K, work / K, seps / K
        lw, lh = eng.get_grid()
        n_slots = 1 << (lw + lh)
        # algorithmic bytes per launch of each kernel (DESIGN.md "bytes per unit")
        stage_ms.setdefault("exchange", 0.0)
        alg = {
            "pre": 4.0 * n_slots,                              # k_step_zero
            "stage_a": 262.0 * n + 8.0 * E,
            "scan": 8.0 * n_slots,
            "scatter": 16.0 * n + 16.0 * E,
            "candidates": 8.0 * E + 16.0 * W,                  # k_enum_pairs
            "solve": 208.0 * W + 200.0 * Q + 16.0 * V,         # test + solve + events
        }
        kernel_of = {"pre": "k_step_zero", "stage_a": "k_stage_a", "scan": "k_scan_slots", "scatter": "k_scatter", "candidates": "k_enum_pairs",
                     "solve": "k_solve"}  # (dense_path: the solve stage is k_solve_dense + k_solve)
        # Dense scenes (c3_dense, c5): long slot runs are solved by k_solve_dense from shared memory and never become work items,
        # so W no longer counts them; that kernel is issue-bound (f64), not HBM-bound — its pairs are reported, its bytes are not
        # part of this model (DESIGN.md "Dense cells").
        dense_path = W + Q < 0.9 * P
        dom = max((k for k in alg if k != "pre"), key=lambda k: stage_ms[k])
        achieved = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9
        # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this same command (profiles/)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.workload == "c3" and n == 1_000_000:
            with open(tpath) as f:
                traffic = json.load(f).get(kernel_of[dom])
        step_bytes = sum(alg.values())
        line = {
            "metric": METRIC, "value": n * world * K / wall_max, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": wall_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: {n} mixed circles/rects per GPU, uniform density, cell_width 4, padding 0.01, bulk step every {DT} "
                            f"(full re-bin + pair TOI solve + next-event min)",
                "hitboxes_per_gpu": n, "grid_slots": n_slots, "l2": "working set per step (SoA state + HbRec + slot table + entries) > 126 MB L2; no flush",
                "timing": "wall clock around K synchronous C-ABI calls bracketed by barrier+cuda sync, max over ranks; device_ms_per_step = CUDA events on the engine "
                          "stream around each whole step; stage_ms = a second pass of K steps with events between the kernels",
            },
            "pair_solves_per_s": solves_all / wall_max,
            "device_ms_per_step": dev_ms_max,
            "stage_ms": stage_ms,
            "per_step": {"cell_entries": E, "cells": Cc, "work_items": W, "pair_solves": P, "events": V},
            "dense_path": {"active": bool(dense_path), "pairs_per_step": max(0.0, P - Q - W) if dense_path else 0.0,
                           "note": "pairs of long slot runs solved by k_solve_dense (issue-bound; outside the HBM byte model)"},
            "gpu_launches": int(launches),
            "e2e": {"value": n * world * K / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(16 * n), "d2h_bytes_per_step": int(d2h / K),
                    "ms_per_step": e2e_max / K * 1e3, "ms_per_step_median_rank0": float(np.median(e2e_each)), "ms_per_step_max_rank0": float(np.max(e2e_each)),
                    "event_sort_fallbacks": eng.event_sort_fallbacks()},
            "roofline": {"bound": "hbm", "kernel": kernel_of[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "peak_kind": peak_kind,
                         "per_kernel": {kernel_of[k]: {"ms": stage_ms[k], "bytes": alg[k], "frac": alg[k] / (stage_ms[k] * 1e-3) / 1e9 / hbm_peak}
                                        for k in alg if stage_ms.get(k)},
                         "whole_step": {"bytes": step_bytes, "achieved": step_bytes / (dev_ms_max * 1e-3) / 1e9,
                                        "frac": step_bytes / (dev_ms_max * 1e-3) / 1e9 / hbm_peak,
                                        # SURVEY 8d's strict I/O lower bound of a step: 152 B of state in + out per hitbox, 24 B per event
                                        "io_lower_bound_bytes": 152.0 * n + 24.0 * V,
                                        "io_lower_bound_frac": (152.0 * n + 24.0 * V) / (dev_ms_max * 1e-3) / 1e9 / hbm_peak}},
            "clocks": sampler.summary(),
        }
        if world > 1:
            if sh.direct:
                line["config"]["parallelism"] = (f"spatial sharding: {world} column strips; halo records (96 B) and next-event keys pushed into the peers' "
                                                 f"memory over NVLink by the step kernels (cudaIpc), one host sync per step, no collective on the data path")
            else:
                line["config"]["parallelism"] = f"spatial sharding: {world} column strips, NCCL all-gather of halo slabs (96-B records) + all-gather min"
                hs, hr = sh.halo_counts()
                line["halo_records_last_step_rank0"] = {"sent": hs, "received": hr}
            line["config"]["exchange"] = "direct" if sh.direct else "nccl"
        if not args.no_cpu_baseline:
            n_sample = min(n, args.ref_sample)
            ups, sps, ms, st = oracle_timing(lambda m: make_workload(scenes, args.workload, m, 0, 1), n_sample, 3, 1)
            line["cpu_baseline"] = {"value": ups, "unit": UNIT, "cores": 1, "kind": "port", "pair_solves_per_s": sps,
                                    "sample": f"{n_sample} hitboxes of the same scene density, 3 bulk-equivalent steps (ascending set_hitbox_vel loop), single thread",
                                    "host_cores_available": os.cpu_count()}
        print(json.dumps(line), file=out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def C_sizeof_result(cb):
    import ctypes

    return ctypes.sizeof(cb.StepResult)


if __name__ == "__main__":
    main()
