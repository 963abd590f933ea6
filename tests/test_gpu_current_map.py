"""The uniform-current map (include/mantaray_b200.h MR_OPT_CURRENT_MAP / MR_OPT_NO_CURRENT_MAP).

Where a block of 8 x 8 cells of an affine current grid holds one u and one v, the reference's bilinear of four equal
corners returns exactly that value (interpolator.rs:78-83: a10 = a01 = a11 = 0) and its finite differences
(cartesian_current.rs:522-536) are exactly 0; the fast path then reads the block's {u, v} from a small map instead
of the 64-byte cell record.  With the map and without it the results must be bit-identical (up to the sign of an
exact zero), on uniform grids, piecewise-constant ones, grids that are uniform only in places, and around the
blocks' edges; and both agree with the oracle."""

import numpy as np
import pytest

from conftest import assert_parity
from mantaray_b200 import MR_MATH_FAST, CartesianCurrent, CartesianNetcdf3, Fields, trace_many
from mantaray_b200 import workloads as W
from mantaray_b200._abi import (MR_OPT_CURRENT_MAP, MR_OPT_DEEP_MAP, MR_OPT_NO_CURRENT_MAP, MR_OPT_NO_DEEP_MAP,
                                MR_OPT_NO_SAME_GRID)

pytestmark = pytest.mark.gpu


def assert_identical(a, b, what):
    for name in ("rows", "len", "x", "y", "kx", "ky", "final_state"):
        # assert_array_equal: NaN == NaN and -0 == +0, everything else bit for bit
        np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=f"{what}: {name}")


def run_all(bathy, cur, rays, t_end, dt, stride=1):
    out = {}
    with Fields(bathy, cur, devices=[0]) as f:
        # (without either map the library would take the same-grid shortcut by itself where the grids coincide: the
        # map is compared with the plain lookups)
        nsg = MR_OPT_NO_SAME_GRID
        for name, flags in (("off", MR_OPT_NO_CURRENT_MAP | MR_OPT_NO_DEEP_MAP | nsg), ("on", MR_OPT_CURRENT_MAP | MR_OPT_NO_DEEP_MAP | nsg),
                            ("off+depth map", MR_OPT_NO_CURRENT_MAP | MR_OPT_DEEP_MAP | nsg), ("on+depth map", MR_OPT_CURRENT_MAP | MR_OPT_DEEP_MAP | nsg),
                            ("default", 0)):
            out[name] = trace_many(f, *rays, 0.0, t_end, dt, stride=stride, math=MR_MATH_FAST, final_state=True, flags=flags)
    return out


@pytest.mark.parametrize("name,make", [
    ("C2: zero current", lambda: W.c2_sea_mount(3000, 1500)),
    ("C3: two constant halves", lambda: W.c3_shear_jet(4096, 900, nx=256)),
    ("C4: no uniform block", lambda: W.c4_agulhas(32, 32, 500, nx=512)),
    ("C5: u = 0, v nowhere constant", lambda: W.c5_nazare(4, 4, 32, 1200, nx=1024)),
])
def test_named_workloads_with_the_current_map(oracle, gpu, name, make):
    wl = make()
    rays = wl.all_rays()
    ref = oracle.trace_many(wl.bathymetry, wl.current, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride)
    res = run_all(wl.bathymetry, wl.current, rays, wl.duration, wl.dt, wl.stride)
    assert_identical(res["on"], res["off"], name)
    assert_identical(res["on+depth map"], res["off+depth map"], name + " (depth-floor map)")
    for k, r in res.items():
        assert_parity(r, ref, what=f"{name}: current map {k}")


def test_patchwork_current_and_block_edges(oracle, gpu):
    """Uniform patches of different values, a varying patch, single odd nodes, NaN / inf nodes, and rays started on
    and around the block edges (multiples of 8 cells), the grid edges and outside the grid."""
    nx, ny, d = 67, 45, 20.0
    x = (np.arange(nx) * d).astype(np.float32)
    y = (np.arange(ny) * d).astype(np.float32)
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny))
    rng = np.random.default_rng(3)
    u = np.where(X < 24, 0.0, np.where(X < 48, 0.7, -0.3)).astype(np.float64)
    v = np.where(Y < 16, 0.25, 0.0).astype(np.float64)
    u[20:30, 30:42] = 0.4 * np.sin(X[20:30, 30:42] / 3.0)          # a varying patch
    v[5, 9] = 0.2500000000000001                                   # one node off by an ulp: its blocks are NOT uniform
    u[33, 60], v[12, 50], u[40, 5] = np.nan, np.inf, -0.0          # non-finite nodes; a -0 among +0
    depth = 30.0 + 10.0 * np.sin(X / 6.0) * np.cos(Y / 5.0)
    bathy = CartesianNetcdf3(x, y, depth)
    cur = CartesianCurrent(x.astype(np.float64), y.astype(np.float64), u, v)
    m = 3000
    x0 = rng.uniform(-2 * d, (nx + 1) * d, m)
    y0 = rng.uniform(-2 * d, (ny + 1) * d, m)
    edges = np.arange(0, nx, 8) * d
    x0[:400] = rng.choice(edges, 400) + rng.choice([-1e-9, 0.0, 1e-9, -d * 2.0 ** -20, d * 2.0 ** -20], 400)
    y0[200:600] = rng.choice(np.arange(0, ny, 8) * d, 400) + rng.choice([-1e-9, 0.0, 1e-9], 400)
    k = 10.0 ** rng.uniform(-1.2, 0.2, m)
    th = rng.uniform(0, 2 * np.pi, m)
    rays = (x0, y0, k * np.cos(th), k * np.sin(th))
    dt, steps = 0.5, 400
    ref = oracle.trace_many(bathy, cur, *rays, 0.0, dt * steps, dt)
    res = run_all(bathy, cur, rays, dt * steps, dt)
    assert_identical(res["on"], res["off"], "patchwork")
    assert_identical(res["on+depth map"], res["off+depth map"], "patchwork (depth-floor map)")
    for kname, r in res.items():
        assert_parity(r, ref, what=f"patchwork: current map {kname}")
