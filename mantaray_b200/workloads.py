"""Synthetic workloads C1-C5: the five configurations of BASELINE.json.

Every field is closed-form and deterministic.  Fields are
built the way the reference would see them after ``open``: bathymetry
coordinates f32, current coordinates f64, data f64, flat ``[y][x]``.  The
gravity constant of the *initial-condition* formulas is the notebooks' 9.81
(notebooks/canonical_examples/utils.py:8); the solver's stays 9.8.

Rays are generated per index range so that a rank of a sharded run only builds
its own block.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Tuple

import numpy as np

from .fields import CartesianCurrent, CartesianNetcdf3

G_IC = 9.81

#: algorithmic FP64 work per ray-step: gridded fields / constant fields
FLOP_PER_RAY_STEP_GRID = 601
FLOP_PER_RAY_STEP_CONST = 441
#: algorithmic HBM bytes per stored row (x, y, kx, ky as f64)
BYTES_PER_ROW = 32


def period2wavenumber(T):
    """Deep-water k for period T (notebooks/canonical_examples/utils.py:11-26)."""
    return (2.0 * math.pi) ** 2 / (G_IC * np.asarray(T, dtype=np.float64) ** 2)


def deep_group_velocity(k):
    return 0.5 * np.sqrt(G_IC / k)


@dataclass
class Workload:
    name: str
    description: str
    bathymetry: CartesianNetcdf3
    current: CartesianCurrent
    n_rays: int
    duration: float
    dt: float
    stride: int
    output: str                               # "full" | "final"
    rays: Callable[[int, int], Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]]
    flop_per_ray_step: int = FLOP_PER_RAY_STEP_GRID
    t0: float = 0.0
    extra: dict = field(default_factory=dict)

    @property
    def n_steps(self) -> int:
        return int(math.ceil((self.duration - self.t0) / self.dt))

    @property
    def n_rows(self) -> int:
        return self.n_steps // self.stride + 1

    def all_rays(self):
        return self.rays(0, self.n_rays)


def _grid_xy(nx, ny, dx, dy, x_first=0.0, y_first=0.0):
    x = x_first + dx * np.arange(nx, dtype=np.float64)
    y = y_first + dy * np.arange(ny, dtype=np.float64)
    return x, y


# ---- C1 ------------------------------------------------------------------------------------
def c1_canonical(n_rays: int = 1000, n_steps: int = 10_000) -> Workload:
    """Constant depth 4000 m, zero current, 200x100 @ 1 km; 1 000 rays, dt 2.5 s, 10 000 steps."""
    nx, ny, d = 200, 100, 1000.0
    x, y = _grid_xy(nx, ny, d, d)
    bathy = CartesianNetcdf3(x, y, np.full((ny, nx), 4000.0))
    cur = CartesianCurrent(x, y, np.zeros((ny, nx)), np.zeros((ny, nx)))
    k0 = float(period2wavenumber(10.0))

    def rays(lo, hi):
        i = np.arange(lo, hi, dtype=np.float64)
        y0 = i * (99_000.0 / max(n_rays - 1, 1))
        m = hi - lo
        return np.full(m, 10.0), y0, np.full(m, k0), np.zeros(m)

    dt = 2.5
    return Workload("C1-canonical", "constant-depth deep water, zero current, 200x100 @ 1 km",
                    bathy, cur, n_rays, dt * n_steps, dt, 1, "full", rays,
                    flop_per_ray_step=FLOP_PER_RAY_STEP_CONST)


# ---- C2 ------------------------------------------------------------------------------------
def c2_sea_mount(n_rays: int = 100_000, n_steps: int = 2000, half: int = 1000) -> Workload:
    """Linear sea-mount (support/linear_sea_mount.py shape): 2001x2001 @ 10 m on [-10 km, 10 km],
    h = 0.1 R - 50 for R <= 8 km else 750 m; the island R < 500 m has h <= 0."""
    d = 10.0
    n = 2 * half + 1
    x = (np.arange(-half, half + 1, dtype=np.float64) * d).astype(np.float32)
    X, Y = np.meshgrid(x.astype(np.float64), x.astype(np.float64))       # [y][x]
    R = np.sqrt(X * X + Y * Y)
    r_out = 0.8 * half * d
    h = np.where(R <= r_out, 0.1 * R - 50.0, 0.1 * r_out - 50.0)
    bathy = CartesianNetcdf3(x, x, h)
    cur = CartesianCurrent(x.astype(np.float64), x.astype(np.float64), np.zeros((n, n)), np.zeros((n, n)))
    k0 = float(period2wavenumber(10.0))
    ext = half * d

    def rays(lo, hi):
        i = np.arange(lo, hi, dtype=np.float64)
        y0 = -0.9 * ext + i * (1.8 * ext / max(n_rays - 1, 1))
        m = hi - lo
        return np.full(m, -ext + d), y0, np.full(m, k0), np.zeros(m)

    dt = d / float(deep_group_velocity(k0))
    return Workload("C2-sea-mount", "linear sea-mount 2001x2001 @ 10 m, shoaling + refraction, termination at shore",
                    bathy, cur, n_rays, dt * n_steps, dt, 1, "full", rays)


# ---- C3 ------------------------------------------------------------------------------------
def c3_shear_jet(n_rays: int = 1_000_000, n_steps: int = 6000, nx: int = 1024) -> Workload:
    """Snell's-law shear current: 1024x1024 @ 50 m, depth 10 km, v = 2 m/s for columns >= nx/2
    (notebooks/theoretical_comparison/data_generation.ipynb); rays at 15 degrees, dt 1 s."""
    d = 50.0
    x, y = _grid_xy(nx, nx, d, d)
    bathy = CartesianNetcdf3(x, y, np.full((nx, nx), 10_000.0))
    v = np.zeros((nx, nx))
    v[:, nx // 2:] = 2.0
    cur = CartesianCurrent(x, y, np.zeros((nx, nx)), v)
    k0 = float(period2wavenumber(10.0))
    phi = math.radians(15.0)
    kx0, ky0 = k0 * math.cos(phi), k0 * math.sin(phi)
    ymax = 25_000.0 * nx / 1024

    def rays(lo, hi):
        i = np.arange(lo, hi, dtype=np.float64)
        y0 = 500.0 + i * ((ymax - 500.0) / max(n_rays - 1, 1))
        m = hi - lo
        return np.full(m, 50.0), y0, np.full(m, kx0), np.full(m, ky0)

    dt = 1.0
    return Workload("C3-shear-jet", "shear current jet on 1024x1024 @ 50 m, current-induced refraction",
                    bathy, cur, n_rays, dt * n_steps, dt, 1, "final", rays)


# ---- C4 ------------------------------------------------------------------------------------
def _ring_eddy(X, Y, xc, yc, L, U_max, core_ratio=0.25):
    """Parabolic ring eddy (shape of utils.generate_parabolic_ring_eddy): azimuthal flow,
    zero inside the core, parabolic profile across the ring."""
    dx, dy = X - xc, Y - yc
    r = np.hypot(dx, dy)
    r_outer = 0.5 * L
    r_core = core_ratio * r_outer
    r_mid, w = 0.5 * (r_core + r_outer), 0.5 * (r_outer - r_core)
    s = (r - r_mid) / w
    ut = np.where(np.abs(s) <= 1.0, U_max * (1.0 - s * s), 0.0)
    rs = np.where(r > 0, r, 1.0)
    return -ut * dy / rs, ut * dx / rs


def c4_agulhas(n_points: int = 1000, n_dirs: int = 1000, n_steps: int = 2048, nx: int = 2048,
               seed: int = 20261017) -> Workload:
    """Agulhas-like eddy field + variable bathymetry on 2048x2048 @ 500 m, f64; 1M rays
    (1000 start points x 1000 directions), full trajectory output."""
    d = 500.0
    L = d * nx
    x, y = _grid_xy(nx, nx, d, d)
    X, Y = np.meshgrid(x, y)
    rng = np.random.default_rng(seed)
    # westward meandering Gaussian jet
    yc = 0.5 * L + 0.078125 * L * np.sin(2.0 * np.pi * X / (0.5859375 * L))
    u = -1.5 * np.exp(-(((Y - yc) / (0.05859375 * L)) ** 2))
    v = np.zeros_like(u)
    for _ in range(6):
        Le = rng.uniform(120e3, 320e3) * (L / 1.024e6)
        Um = rng.uniform(0.5, 1.5) * rng.choice([-1.0, 1.0])
        xc, yc_ = rng.uniform(0.1 * L, 0.9 * L, size=2)
        du, dv = _ring_eddy(X, Y, xc, yc_, Le, Um)
        u += du
        v += dv
    depth = 4000.0 - 3800.0 / (1.0 + np.exp(-(Y - 0.87890625 * L) / (0.0390625 * L)))
    for _ in range(3):
        xc, yc_ = rng.uniform(0.1 * L, 0.9 * L, size=2)
        depth -= 1500.0 * np.exp(-((X - xc) ** 2 + (Y - yc_) ** 2) / (2.0 * (0.029296875 * L) ** 2))
    depth = np.maximum(depth, 50.0)
    bathy = CartesianNetcdf3(x, y, depth)
    cur = CartesianCurrent(x, y, u, v)
    k0 = float(period2wavenumber(10.0))
    n_rays = n_points * n_dirs
    sc = L / 1.024e6

    def rays(lo, hi):
        i = np.arange(lo, hi, dtype=np.int64)
        p = (i // n_dirs).astype(np.float64)
        q = (i % n_dirs).astype(np.float64)
        fp = p / max(n_points - 1, 1)
        x0 = (10e3 + 390e3 * fp) * sc
        y0 = (200e3 - 100e3 * fp) * sc
        th = np.radians(45.0 + 30.0 * q / max(n_dirs - 1, 1))
        return x0, y0, k0 * np.cos(th), k0 * np.sin(th)

    dt = d / float(deep_group_velocity(k0))
    return Workload("C4-agulhas", "Agulhas-like eddy current + variable bathymetry 2048x2048 f64, full trajectory output",
                    bathy, cur, n_rays, dt * n_steps, dt, 1, "full", rays,
                    extra={"n_points": n_points, "n_dirs": n_dirs})


# ---- C5 ------------------------------------------------------------------------------------
def c5_nazare(n_periods: int = 64, n_dirs: int = 64, n_points: int = 16_384, n_steps: int = 4096,
              nx: int = 4096, stride: int = 64) -> Workload:
    """Nazare-style canyon on 4096x4096 @ 25 m + alongshore jet; 64 periods x 64 directions x
    16 384 start points = 2^26 rays, decimated (stride 64) output."""
    d = 25.0
    L = d * nx                                   # 102.4 km
    x, y = _grid_xy(nx, nx, d, d)
    X, Y = np.meshgrid(x, y)
    s = L / 102.4e3
    y_axis = 0.5 * L + 8e3 * s * np.sin(2.0 * np.pi * X / (80e3 * s))
    depth = 0.004 * (100e3 * s - X) + 600.0 * np.exp(-(((Y - y_axis) / (1.5e3 * s)) ** 2)) * np.clip(
        (X - 20e3 * s) / (60e3 * s), 0.0, 1.0)
    u = np.zeros_like(depth)
    v = 0.3 * np.exp(-(((X - 90e3 * s) / (5e3 * s)) ** 2))
    bathy = CartesianNetcdf3(x, y, depth)
    cur = CartesianCurrent(x, y, u, v)
    periods = np.linspace(8.0, 18.0, n_periods)
    ks = period2wavenumber(periods)
    thetas = np.radians(np.linspace(-30.0, 30.0, n_dirs))
    n_rays = n_periods * n_dirs * n_points

    def rays(lo, hi):
        i = np.arange(lo, hi, dtype=np.int64)
        pt = (i % n_points).astype(np.float64)
        pd = i // n_points
        k = ks[pd // n_dirs]
        th = thetas[pd % n_dirs]
        y0 = (10e3 + 82e3 * pt / max(n_points - 1, 1)) * s
        return np.full(hi - lo, 2e3 * s), y0, k * np.cos(th), k * np.sin(th)

    dt = 2.0
    return Workload("C5-nazare", "Nazare-style canyon 4096x4096 @ 25 m + coastal jet, frequency/direction ensemble, stride-64 output",
                    bathy, cur, n_rays, dt * n_steps, dt, stride, "full", rays,
                    extra={"tile": n_points, "n_periods": n_periods, "n_dirs": n_dirs})


WORKLOADS = {
    "C1": c1_canonical,
    "C2": c2_sea_mount,
    "C3": c3_shear_jet,
    "C4": c4_agulhas,
    "C5": c5_nazare,
}


def shard_range(n: int, rank: int, world: int, align: int = 128) -> Tuple[int, int]:
    """Contiguous block of rays of `rank`: ceil(n/world) rounded up to whole
    thread blocks, in input order, so that the gather is a concatenation along `ray`."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard_tiles(n: int, rank: int, world: int, tile: int):
    """Tiles of `tile` consecutive rays dealt round-robin: rank r gets tiles r, r + world, r + 2 world, ...
    (interleaved assignment where termination is correlated with the ray index).  In C5 a tile is one
    (period, direction) pair over all start points, so every rank gets every period — a contiguous block would be a
    band of periods, and long periods run ashore early.  The gather is a strided concatenation of the tiles.
    Returns the list of (lo, hi) index ranges."""
    tile = max(int(tile), 1)
    n_tiles = -(-n // tile)
    return [(t * tile, min((t + 1) * tile, n)) for t in range(rank, n_tiles, world)]
