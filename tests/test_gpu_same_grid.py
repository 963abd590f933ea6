"""The same-grid shortcut (include/mantaray_b200.h MR_OPT_SAME_GRID / MR_OPT_NO_SAME_GRID).

When the current is given on the bathymetry's own grid the fast path derives the current's cell from the
bathymetry's f32 fractional index (cartesian_netcdf3.rs:289) instead of forming the f64 index of
cartesian_current.rs:246 — wherever the f32 index is further from a grid line than the two can disagree — and runs
the two separate lookups elsewhere.  The claim is exactness of the CELLS: every looked-up value is identical, and
the results agree to the few ulp by which the kernel variants round the sum of the advection terms differently
(a wrong cell changes the piecewise-constant current gradients and shows up at 1e-6 or more).  The danger zone is a position within a few f32 ulps of a grid line, where `x as f32` rounds across the line and the two
indices name different cells; these tests sit rays exactly there, on long axes (the index error grows with the
index), with zero, negative and large origins, and compare with the separate lookups bit for bit and with the
oracle."""

import numpy as np
import pytest

from conftest import assert_parity
from mantaray_b200 import MR_MATH_FAST, CartesianCurrent, CartesianNetcdf3, Fields, trace_many
from mantaray_b200 import workloads as W
from mantaray_b200._abi import MR_OPT_DEEP_MAP, MR_OPT_NO_DEEP_MAP, MR_OPT_NO_SAME_GRID, MR_OPT_SAME_GRID

pytestmark = pytest.mark.gpu


def assert_same_cells(a, b, what, tol=1e-11):
    """rows / len / NaN pattern identical; values within `tol` of the ray's scale"""
    np.testing.assert_array_equal(a.rows, b.rows, err_msg=f"{what}: rows")
    np.testing.assert_array_equal(a.len, b.len, err_msg=f"{what}: len")
    with np.errstate(invalid="ignore"):
        pos = np.nanmax(np.maximum(np.abs(b.x), np.abs(b.y)), axis=0, initial=0.0)
        ksc = np.nanmax(np.hypot(b.kx, b.ky), axis=0, initial=0.0)
        pos, ksc = np.where(pos > 0, pos, 1.0), np.where(ksc > 0, ksc, 1.0)
        for name, sc in (("x", pos), ("y", pos), ("kx", ksc), ("ky", ksc)):
            u, v = getattr(a, name), getattr(b, name)
            np.testing.assert_array_equal(np.isnan(u), np.isnan(v), err_msg=f"{what}: NaN pattern of {name}")
            err = float(np.nanmax(np.abs(u - v) / sc[None, :], initial=0.0))
            assert err <= tol, f"{what}: {name} differs by {err:.2e} of the ray's scale"


def grid(nx, ny, x_first, y_first, d, seed, deep=False):
    """Bathymetry and current on ONE grid; coordinates chosen so that the f32 grid is exactly affine."""
    x = (x_first + d * np.arange(nx)).astype(np.float32)
    y = (y_first + d * np.arange(ny)).astype(np.float32)
    assert np.all(x.astype(np.float64) == x_first + d * np.arange(nx)) and np.all(y.astype(np.float64) == y_first + d * np.arange(ny))
    rng = np.random.default_rng(seed)
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny))
    base = 3000.0 if deep else 40.0
    depth = base + 0.3 * base * np.sin(X / 5.0 + 1.0) * np.cos(Y / 3.0) + rng.normal(0, 0.02 * base, X.shape)
    u = 0.4 * np.sin(Y / 2.5 + 0.3) + rng.normal(0, 0.03, X.shape)
    v = 0.4 * np.cos(X / 3.5 + 0.7) + rng.normal(0, 0.03, X.shape)
    return CartesianNetcdf3(x, y, depth), CartesianCurrent(x.astype(np.float64), y.astype(np.float64), u, v)


def rays_around_grid_lines(bathy, n_lines, seed):
    """Positions on, just below and just above grid lines and nodes (a few f32 and f64 ulps away, both sides), at
    the domain edges and just outside, plus ordinary interior points."""
    rng = np.random.default_rng(seed)
    bx, by = bathy.x.astype(np.float64), bathy.y.astype(np.float64)
    xs, ys = [], []
    ulps32 = np.array([-3, -2, -1, -0.5, -0.25, 0, 0.25, 0.5, 1, 2, 3]) * 2.0 ** -24
    ulps64 = np.array([-2, -1, 1, 2]) * 2.0 ** -52
    for _ in range(n_lines):
        i, j = int(rng.integers(0, bx.size)), int(rng.integers(0, by.size))
        gx, gy = bx[i], by[j]
        for rel in np.concatenate([ulps32, ulps64]):
            # near a vertical grid line, anywhere along it; near a horizontal one; near the node itself
            xs += [gx + abs(gx) * rel + (rel * 1e-3 if gx == 0 else 0.0), rng.uniform(bx[0], bx[-1]), gx + abs(gx) * rel]
            ys += [rng.uniform(by[0], by[-1]), gy + abs(gy) * rel + (rel * 1e-3 if gy == 0 else 0.0), gy + abs(gy) * rel]
    d = bx[1] - bx[0]
    for ex in (bx[0], bx[-1]):                                   # the edges and just outside / inside
        for off in (-1e-3 * d, -1e-9 * d, 0.0, 1e-9 * d, 1e-3 * d):
            xs.append(ex + off); ys.append(rng.uniform(by[0], by[-1]))
            xs.append(rng.uniform(bx[0], bx[-1])); ys.append((by[0] if ex == bx[0] else by[-1]) + off)
    m = 512
    xs += list(rng.uniform(bx[0], bx[-1], m)); ys += list(rng.uniform(by[0], by[-1], m))
    x0, y0 = np.array(xs), np.array(ys)
    k = 10.0 ** rng.uniform(-1.5, 0.0, x0.size) / np.sqrt(max(d, 1e-2))
    th = rng.uniform(0, 2 * np.pi, x0.size)
    return x0, y0, k * np.cos(th), k * np.sin(th)


CASES = [
    # nx, ny, x_first, y_first, spacing
    (64, 48, 0.0, 0.0, 500.0),
    (4096, 6, 0.0, 0.0, 25.0),             # a long axis: index error ~ 3 * 4096 * 2^-24
    (6, 4096, 0.0, 0.0, 25.0),
    (2001, 9, -10000.0, -40.0, 10.0),      # C2's origin: |x0|/s = 1000 adds to the bound
    (300, 200, 4096.0, -8192.0, 0.5),      # large origin against the spacing
    (129, 65, -64.0, 0.0, 1.0),
]


@pytest.mark.parametrize("nx,ny,xf,yf,d", CASES)
@pytest.mark.parametrize("deep", [False, True])
def test_same_grid_shortcut_names_the_same_cells_near_grid_lines(oracle, gpu, nx, ny, xf, yf, d, deep):
    bathy, cur = grid(nx, ny, xf, yf, d, seed=nx + ny, deep=deep)
    rays = rays_around_grid_lines(bathy, 40, seed=7 * nx + ny)
    # a short step: most rays stay within a cell or two of where they were put, i.e. near the line for several stages
    dt = 0.02 * d / 5.0
    t_end = 12 * dt
    ref = oracle.trace_many(bathy, cur, *rays, 0.0, t_end, dt)
    with Fields(bathy, cur, devices=[0]) as f:
        sep = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_NO_SAME_GRID)
        sg = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_SAME_GRID)
        sg_map = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_DEEP_MAP | MR_OPT_SAME_GRID)
        auto = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_SAME_GRID)
    assert_same_cells(sg, sep, "same-grid vs separate")
    for res, what in ((sg, "same-grid"), (sg_map, "same-grid + depth-floor map"), (auto, "same-grid, map by default")):
        assert_parity(res, ref, what=f"{what} {nx}x{ny} @ {d} from ({xf},{yf})")


@pytest.mark.parametrize("name,make", [
    ("C2", lambda: W.c2_sea_mount(4000, 600)),
    ("C3", lambda: W.c3_shear_jet(4096, 500, nx=512)),
    ("C4", lambda: W.c4_agulhas(64, 64, 600, nx=1024)),
    ("C5", lambda: W.c5_nazare(8, 8, 64, 1500, nx=2048)),
])
def test_named_workloads_take_the_shortcut_and_agree(oracle, gpu, name, make):
    """Every named shape has its current on the bathymetry's grid.  Longer runs than above: the rays cross
    thousands of grid lines."""
    wl = make()
    rays = wl.all_rays()
    ref = oracle.trace_many(wl.bathymetry, wl.current, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride)
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        sep = trace_many(f, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_NO_SAME_GRID)
        sg = trace_many(f, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_SAME_GRID)
        auto = trace_many(f, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride, final_state=True, flags=MR_OPT_SAME_GRID)
    assert_same_cells(sg, sep, f"{name}: same-grid vs separate")
    assert_parity(sg, ref, what=f"{name} same-grid")
    assert_parity(auto, ref, what=f"{name} same-grid, map by default")


def test_grids_that_only_look_alike_do_not_take_the_shortcut(oracle, gpu):
    """Same shape but the current's coordinates are shifted by a fraction of a cell, or its spacing differs in the
    last bits: the shortcut must not apply (the results would be wrong by whole cells), and nothing changes."""
    bathy, cur = grid(64, 48, 0.0, 0.0, 500.0, seed=3)
    rays = rays_around_grid_lines(bathy, 10, seed=5)
    dt, t_end = 5.0, 200.0
    for cx, cy in ((cur.x + 125.0, cur.y), (cur.x, cur.y * (1.0 + 2.0 ** -40)), (cur.x[:-1], cur.y)):
        c2 = CartesianCurrent(cx, cy, cur.u.reshape(cur.y.size, cur.x.size)[:cy.size, :cx.size].copy(),
                              cur.v.reshape(cur.y.size, cur.x.size)[:cy.size, :cx.size].copy())
        ref = oracle.trace_many(bathy, c2, *rays, 0.0, t_end, dt)
        with Fields(bathy, c2, devices=[0]) as f:
            a = trace_many(f, *rays, 0.0, t_end, dt, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_SAME_GRID)
            b = trace_many(f, *rays, 0.0, t_end, dt, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_NO_SAME_GRID)
        for nm in ("rows", "len", "x", "y", "kx", "ky"):
            np.testing.assert_array_equal(getattr(a, nm), getattr(b, nm), err_msg=nm)
        assert_parity(a, ref, what="look-alike grids")
