"""Synthetic scenes of SURVEY.md 8(d) / BASELINE.json configs, as SoA numpy columns.

Deterministic and language-agnostic: every random number is SplitMix64(seed + stream) so a Rust or C
harness can regenerate the same scene.  Used by bench.py and tests/ (inputs only — no algorithm here).
"""
from __future__ import annotations

import math

import numpy as np

CELL_WIDTH = 4.0
PADDING = 0.01
SEED_BASE = 0xC0111DE5
_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n outputs of SplitMix64 started at seed + stream * 2^32 (vectorised: output i uses state seed + (i+1)*gamma)."""
    with np.errstate(over="ignore"):
        s = np.uint64((seed + (stream << 32)) & 0xFFFFFFFFFFFFFFFF)
        z = s + _GAMMA * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, n: int, stream: int, lo: float, hi: float) -> np.ndarray:
    u = (splitmix64(seed, n, stream) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


def make_scene(n: int, world_w: float, world_h: float | None = None, seed: int = SEED_BASE, squares: bool = False, x0: float = 0.0,
               speed: float = 2.0, size_lo: float = 1.0, size_hi: float = 2.0) -> dict:
    """Uniform-density mixed circles/rects: kind ~ Bernoulli(1/2), size U[1,2], velocity U[-2,2]^2, group 0."""
    world_h = world_w if world_h is None else world_h
    kind = (splitmix64(seed, n, 0) >> np.uint64(63)).astype(np.uint8)  # 0 circle, 1 rect
    w = uniform(seed, n, 3, size_lo, size_hi)
    h = uniform(seed, n, 4, size_lo, size_hi)
    h = np.where((kind == 0) | squares, w, h)
    return dict(
        id=np.arange(n, dtype=np.uint64),
        pos_x=x0 + uniform(seed, n, 1, 0.0, world_w),
        pos_y=uniform(seed, n, 2, 0.0, world_h),
        dim_x=w,
        dim_y=h,
        vel_x=uniform(seed, n, 5, -speed, speed),
        vel_y=uniform(seed, n, 6, -speed, speed),
        kind=kind,
    )


def c1(n: int = 1000) -> dict:
    """Config 1: 1,000 mixed circles/squares in a 200x200 box."""
    return make_scene(n, 200.0, seed=SEED_BASE + 1, squares=True)


def uniform_density(n: int, density: float = 0.025, seed: int = SEED_BASE + 3) -> dict:
    """Configs 2/3: uniform density (default C1's 0.025 per unit^2) -> square world of side sqrt(n / density)."""
    return make_scene(n, math.sqrt(n / density), seed=seed)


def c2() -> dict:
    return uniform_density(100_000, seed=SEED_BASE + 2)


def c3(n: int = 1_000_000) -> dict:
    return uniform_density(n, seed=SEED_BASE + 3)


def c3_dense(n: int = 1_000_000) -> dict:
    return uniform_density(n, density=1.0, seed=SEED_BASE + 3)


def with_groups_and_resizing(s: dict, seed: int) -> dict:
    """The rows the reference's own tests leave out: resize velocities (circles keep rsz_x == rsz_y),
    four groups with asymmetric interact_groups lists and one of them HbGroup None, coordinates centred on the origin."""
    n = len(s["id"])
    s["pos_x"] = s["pos_x"] - 0.5 * (s["pos_x"].max() + s["pos_x"].min())
    s["pos_y"] = s["pos_y"] - 0.5 * (s["pos_y"].max() + s["pos_y"].min())
    r = uniform(seed, n, 20, -0.05, 0.3)
    s["rsz_x"] = r
    s["rsz_y"] = np.where(s["kind"] == 0, r, uniform(seed, n, 21, -0.05, 0.3))
    g = (splitmix64(seed, n, 22) % np.uint64(4)).astype(np.uint32)
    s["group"] = np.where(g == 3, np.uint32(0xFFFFFFFF), g)  # CB200_GROUP_NONE
    s["interact_mask"] = np.array([0b011, 0b101, 0b110, 0b111], dtype=np.uint32)[g]
    return s


def c3_groups(n: int = 1_000_000) -> dict:
    """Config 3's density with resizing shapes in several groups (f-4 at the benchmark's size)."""
    return with_groups_and_resizing(uniform_density(n, seed=SEED_BASE + 6), SEED_BASE + 6)


def c4(n: int = 16_777_216) -> dict:
    """Config 4: n shapes at 1 per unit^2 (16M -> 4096 x 4096)."""
    return uniform_density(n, density=1.0, seed=SEED_BASE + 4)


def c5(n: int = 1_000_000, hot_spots: int = 64, fast_fraction: float = 0.1) -> dict:
    """Config 5: clustered (Gaussian hot spots, ~50x density skew) with 10% fast small shapes."""
    seed = SEED_BASE + 5
    world = math.sqrt(n / 0.025)
    s = make_scene(n, world, seed=seed)
    # half of the shapes are pulled into hot spots: sigma chosen so the peak cell holds ~50x the mean 0.4/cell
    centres_x = uniform(seed, hot_spots, 10, 0.1 * world, 0.9 * world)
    centres_y = uniform(seed, hot_spots, 11, 0.1 * world, 0.9 * world)
    which = (splitmix64(seed, n, 12) % np.uint64(hot_spots)).astype(np.int64)
    clustered = uniform(seed, n, 13, 0.0, 1.0) < 0.5
    # Box-Muller
    u1 = np.maximum(uniform(seed, n, 14, 0.0, 1.0), 1e-300)
    u2 = uniform(seed, n, 15, 0.0, 1.0)
    r = np.sqrt(-2.0 * np.log(u1))
    sigma = math.sqrt((0.5 * n / hot_spots) * CELL_WIDTH * CELL_WIDTH / (2.0 * math.pi * 50.0 * 0.4))
    gx = centres_x[which] + sigma * r * np.cos(2.0 * math.pi * u2)
    gy = centres_y[which] + sigma * r * np.sin(2.0 * math.pi * u2)
    s["pos_x"] = np.where(clustered, np.clip(gx, 0.0, world), s["pos_x"])
    s["pos_y"] = np.where(clustered, np.clip(gy, 0.0, world), s["pos_y"])
    fast = uniform(seed, n, 16, 0.0, 1.0) < fast_fraction
    small = uniform(seed, n, 17, 0.05, 0.2)
    spd = uniform(seed, n, 18, 20.0, 100.0)
    ang = uniform(seed, n, 19, 0.0, 2.0 * math.pi)
    s["dim_x"] = np.where(fast, small, s["dim_x"])
    s["dim_y"] = np.where(fast, small, s["dim_y"])
    s["vel_x"] = np.where(fast, spd * np.cos(ang), s["vel_x"])
    s["vel_y"] = np.where(fast, spd * np.sin(ang), s["vel_y"])
    return s


def strip(scene: dict, rank: int, world_size: int, world_w: float) -> dict:
    """Column-strip shard of a scene by pos_x (multi-GPU spatial sharding)."""
    lo, hi = world_w * rank / world_size, world_w * (rank + 1) / world_size
    m = (scene["pos_x"] >= lo) & ((scene["pos_x"] < hi) | (rank == world_size - 1))
    return {k: v[m] for k, v in scene.items()}
