"""Seeded fuzzing of the CUDA path against the oracle: random grid shapes, spacings and origins (affine and
not, ascending and descending), partially overlapping bathymetry / current domains, non-finite and non-positive
depth nodes, rays starting on nodes, grid lines, edges and outside, degenerate wavenumbers.  Same bar as
everywhere: rows, len and NaN patterns bit-exact, values to 1e-9 (conftest.assert_parity)."""

import os

import numpy as np
import pytest

from conftest import assert_parity
from mantaray_b200 import (MR_MATH_FAST, MR_MATH_STRICT, ArrayDepth, CartesianCurrent, CartesianNetcdf3,
                           ConstantCurrent, ConstantDepth, ConstantSlope, Fields, trace_many)

pytestmark = pytest.mark.gpu

SPACINGS = [1.0, 0.5, 37.3, 500.0, 1e-3, 1e4, 3.0, 12.5]


def axis(rng, n, affine, descending=False, step=None):
    step = float(rng.choice(SPACINGS)) if step is None else step
    first = float(rng.choice([0.0, 0.0, -step * rng.integers(0, n), rng.uniform(-50, 50) * step]))
    a = first + step * np.arange(n)
    if not affine and n > 2:
        a = a + rng.uniform(-0.2, 0.2, n) * step
        a[1] = a[0] + step                      # the spacing the reference uses is |a[1] - a[0]|
        a = np.sort(a)
    return (a[::-1].copy() if descending else a), step


def make_case(seed):
    rng = np.random.default_rng(1000 + seed)
    kind = seed % 6
    nx, ny = int(rng.integers(2, 40)), int(rng.integers(2, 40))
    affine = seed % 3 != 1
    bx, sx = axis(rng, nx, affine, descending=(seed % 11 == 7))
    by, sy = axis(rng, ny, affine, step=sx * float(rng.choice([0.25, 0.5, 1.0, 1.0, 2.0, 4.0])))
    cell = min(sx, sy)
    # relief and current speed scale with the cell, so that slopes and shears stay those of an ocean
    # (|grad h| <~ 1, |grad U| <~ 0.05 /s): with metre-per-second shears across millimetre cells the ray
    # equations amplify one ulp to 1e-9 within a few hundred steps, and no two evaluation orders agree
    relief = min(50.0, 3.0 * cell)
    speed = min(0.6, 0.05 * cell)
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny))
    base = 60.0 if seed % 2 else 0.85 * relief          # even seeds: shoals that fall dry (h <= 0) here and there
    depth = base + relief * np.sin(X / 3.0 + rng.uniform(0, 6)) * np.cos(Y / 4.0) + rng.normal(0, 0.04 * relief, X.shape)
    if seed % 2 == 0:                           # exact zeros and non-finite nodes
        depth[rng.random(depth.shape) < 0.01] = 0.0
        depth[rng.random(depth.shape) < 0.01] = np.inf
        depth[rng.random(depth.shape) < 0.01] = -np.inf
        depth[rng.random(depth.shape) < 0.01] = np.nan
    if kind == 4:
        bathy = ConstantSlope(40.0, float(bx[0]), float(by[0]), 0.03, -0.02)
    elif kind == 5:
        bathy = ConstantDepth(float(rng.choice([5.0, 200.0, 4000.0])))
    else:
        bathy = CartesianNetcdf3(bx.astype(np.float32), by.astype(np.float32), depth)
    if kind == 3:
        cur = ConstantCurrent(float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)))
    else:
        # the current grid covers most, not all, of the bathymetry's extent, at its own resolution
        cnx, cny = int(rng.integers(2, 50)), int(rng.integers(2, 50))
        lo_x, hi_x = float(min(bx[0], bx[-1])), float(max(bx[0], bx[-1]))
        lo_y, hi_y = float(by[0]), float(by[-1])
        ex, ey = (hi_x - lo_x) or sx, (hi_y - lo_y) or sy
        cx = np.linspace(lo_x - rng.uniform(-0.1, 0.3) * ex, hi_x + rng.uniform(-0.1, 0.3) * ex, cnx)
        cy = np.linspace(lo_y - rng.uniform(-0.1, 0.3) * ey, hi_y + rng.uniform(-0.1, 0.3) * ey, cny)
        if seed % 3 == 1 and cnx > 2:
            cx[2:] += rng.uniform(0, 0.3, cnx - 2) * (cx[1] - cx[0])
            cx = np.sort(cx)
        CX, CY = np.meshgrid(np.arange(cnx), np.arange(cny))
        u = speed * (np.sin(CY / 2.5 + rng.uniform(0, 6)) + rng.normal(0, 0.08, CX.shape))
        v = speed * (np.cos(CX / 3.5 + rng.uniform(0, 6)) + rng.normal(0, 0.08, CX.shape))
        if seed % 4 == 2:
            u[rng.random(u.shape) < 0.01] = np.nan
            v[rng.random(v.shape) < 0.01] = np.inf
        cur = CartesianCurrent(cx, cy, u, v)

    n = 512
    lo_x, hi_x = float(min(bx[0], bx[-1])), float(max(bx[0], bx[-1]))
    lo_y, hi_y = float(by[0]), float(by[-1])
    ex, ey = (hi_x - lo_x) or sx, (hi_y - lo_y) or sy
    x0 = rng.uniform(lo_x - 0.05 * ex, hi_x + 0.05 * ex, n)
    y0 = rng.uniform(lo_y - 0.05 * ey, hi_y + 0.05 * ey, n)
    m = n // 8
    x0[:m] = rng.choice(bx, m)                                  # on grid lines ...
    y0[m // 2:m] = rng.choice(by, m - m // 2)                   # ... and nodes
    x0[m:m + 4] = [bx[0], bx[-1], bx[0], bx[-1]]                # the four corners of the domain
    y0[m:m + 4] = [by[0], by[-1], by[-1], by[0]]
    kmag = 10.0 ** rng.uniform(-2.5, 0.5, n) / max(cell, 1e-2) ** 0.5
    th = rng.uniform(0, 2 * np.pi, n)
    kx0, ky0 = kmag * np.cos(th), kmag * np.sin(th)
    kx0[m + 4:m + 12] = [0.0, -0.0, 0.3, 0.0, np.nan, 1e-300, 1e6, -0.3]
    ky0[m + 4:m + 12] = [0.3, -0.3, 0.0, 0.0, 0.1, 0.0, 1e6, -0.0]
    x0[m + 12], y0[m + 13] = np.nan, np.inf
    # about a third of a cell per step for a typical ray
    k_typ = 10.0 ** -1.0 / max(cell, 1e-2) ** 0.5
    cg = 0.5 * np.sqrt(9.8 / k_typ)
    dt = 0.3 * cell / cg
    return bathy, cur, (x0, y0, kx0, ky0), dt, int(rng.choice([60, 60, 150, 400]))


@pytest.mark.parametrize("math", [MR_MATH_FAST, MR_MATH_STRICT], ids=["fast", "strict"])
@pytest.mark.parametrize("seed", range(int(os.environ.get("MR_FUZZ_SEEDS", "120"))))   # more: MR_FUZZ_SEEDS=1000
def test_fuzz(oracle, gpu, seed, math):
    bathy, cur, rays, dt, steps = make_case(seed)
    stride = 1 if seed % 5 else 7
    ref = oracle.trace_many(bathy, cur, *rays, 0.0, dt * steps, dt, stride=stride)
    with Fields(bathy, cur, devices=[0]) as f:
        res = trace_many(f, *rays, 0.0, dt * steps, dt, math=math, final_state=True, stride=stride,
                         chunk_rays=(0 if seed % 4 else 192), env=(math == MR_MATH_FAST))
    assert_parity(res, ref, what=f"fuzz seed {seed}")
    if stride == 1:
        final_rows_check(res)
    if res.depth is not None:                   # the environment planes, bit-exact at the stored states
        want = oracle.sample_fields(bathy, cur, res.x, res.y)
        for got, w, name in zip((res.depth, res.u, res.v), want, ("depth", "u", "v")):
            np.testing.assert_array_equal(np.isnan(got), np.isnan(w), err_msg=f"seed {seed}: NaN pattern of {name}")
            ok = ~np.isnan(w)
            np.testing.assert_array_equal(got[ok], w[ok], err_msg=f"seed {seed}: {name}")


def final_rows_check(res):
    """final_state is the last NaN-free row of the same call, bit for bit"""
    valid = np.nonzero(res.len > 0)[0]
    last = res.len[valid] - 1
    for c, plane in enumerate((res.x, res.y, res.kx, res.ky)):
        np.testing.assert_array_equal(res.final_state[c, valid], plane[last, valid])
    assert np.isnan(res.final_state[:, res.len == 0]).all()
