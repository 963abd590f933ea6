"""The depth-floor map (include/mantaray_b200.h): the fast path skips the depth lookup wherever a
per-block lower bound of the depth proves kh >= 22.  It is on by default where a quarter of the grid's blocks are
deep for a 10 s wave; MR_OPT_DEEP_MAP forces it on, MR_OPT_NO_DEEP_MAP off.  These tests hold the path WITH the
map to the oracle (same bar as everywhere) and to the path WITHOUT it — same rows, same len, same NaN pattern,
values to a few ulp (a lane the map proves deep evaluates cg cos(theta) as (sqrt(G)/2) k^-3/2 kx from one fourth
root instead of forming k, 1/k and 1/sqrt(G k): the same function, rounded differently) — on every grid that has
a map, whatever its deep share.  The same-grid shortcut (MR_OPT_SAME_GRID switches it on) only changes which
index names the cell: the cells, hence every looked-up value, are identical; the kernel variants differ in how
they round the sum of the advection terms, so the comparison is to the same few ulp (a wrong cell would show up
at 1e-6)."""

import os

import numpy as np
import pytest

from conftest import assert_parity
from mantaray_b200 import MR_MATH_FAST, CartesianCurrent, CartesianNetcdf3, ConstantCurrent, Fields, trace_many
from mantaray_b200 import workloads as W
from mantaray_b200._abi import MR_OPT_DEEP_MAP, MR_OPT_NO_DEEP_MAP, MR_OPT_NO_SAME_GRID, MR_PLAN_DEEP_MAP, MR_PLAN_SAME_GRID
from test_gpu_fuzz import make_case

pytestmark = pytest.mark.gpu


def both(f, rays, t_end, dt, **kw):
    # (plain = the separate lookups: without a map the library would take the same-grid shortcut by itself)
    plain = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_NO_SAME_GRID, **kw)
    mapped = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_DEEP_MAP, **kw)
    return plain, mapped


#: map against no map: the ray equations amplify a last-bit difference in cg along the ray, nothing more
ULP_TOL = 1e-11


def assert_same(mapped, plain, what, exact=False):
    np.testing.assert_array_equal(mapped.rows, plain.rows, err_msg=f"{what}: rows")
    np.testing.assert_array_equal(mapped.len, plain.len, err_msg=f"{what}: len")
    if exact:
        for name in ("x", "y", "kx", "ky", "final_state"):
            # assert_array_equal: NaN == NaN and -0 == +0, everything else bit for bit
            np.testing.assert_array_equal(getattr(mapped, name), getattr(plain, name), err_msg=f"{what}: {name}")
        return
    with np.errstate(invalid="ignore"):
        pos = np.nanmax(np.maximum(np.abs(plain.x), np.abs(plain.y)), axis=0, initial=0.0)
        ksc = np.nanmax(np.hypot(plain.kx, plain.ky), axis=0, initial=0.0)
        pos, ksc = np.where(pos > 0, pos, 1.0), np.where(ksc > 0, ksc, 1.0)
        for name, sc in (("x", pos), ("y", pos), ("kx", ksc), ("ky", ksc)):
            a, b = getattr(mapped, name), getattr(plain, name)
            np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f"{what}: NaN pattern of {name}")
            err = float(np.nanmax(np.abs(a - b) / sc[None, :], initial=0.0))
            assert err <= ULP_TOL, f"{what}: {name} differs by {err:.2e} of the ray's scale with the map"
        np.testing.assert_array_equal(np.isnan(mapped.final_state), np.isnan(plain.final_state), err_msg=f"{what}: final state")


def same_grid_pair(f, rays, t_end, dt, **kw):
    """(separate lookups, same-grid shortcut), both without the map"""
    from mantaray_b200._abi import MR_OPT_SAME_GRID
    sep = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_NO_DEEP_MAP | MR_OPT_NO_SAME_GRID, **kw)
    sg = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True,
                    flags=MR_OPT_NO_DEEP_MAP | MR_OPT_SAME_GRID, **kw)
    return sep, sg


@pytest.mark.parametrize("name,make", [
    ("C2", lambda: W.c2_sea_mount(1000, 2000)),                 # 750 m plateau (deep at T = 10 s), then the shoal
    ("C3", lambda: W.c3_shear_jet(1024, 600, nx=256)),          # 10 km everywhere: every lookup skipped
    ("C4", lambda: W.c4_agulhas(32, 32, 700, nx=512)),          # deep basin, shelf and seamounts
    ("C5", lambda: W.c5_nazare(8, 8, 16, 1200, nx=1024)),       # 400 m and shoaling: mostly not deep
])
def test_workloads_with_the_depth_floor_map(oracle, gpu, name, make):
    wl = make()
    rays = wl.all_rays()
    ref = oracle.trace_many(wl.bathymetry, wl.current, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride)
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        plain, mapped = both(f, rays, wl.duration, wl.dt, stride=wl.stride)
        sep, sg = same_grid_pair(f, rays, wl.duration, wl.dt, stride=wl.stride)
    assert_parity(mapped, ref, what=f"{name} with the depth-floor map")
    assert_same(mapped, plain, name)
    assert_same(sg, sep, f"{name}: same-grid shortcut")


@pytest.mark.parametrize("seed", range(int(os.environ.get("MR_FUZZ_SEEDS", "120"))))
def test_fuzz_with_the_depth_floor_map(oracle, gpu, seed):
    bathy, cur, rays, dt, steps = make_case(seed)
    stride = 1 if seed % 5 else 7
    ref = oracle.trace_many(bathy, cur, *rays, 0.0, dt * steps, dt, stride=stride)
    with Fields(bathy, cur, devices=[0]) as f:
        plain, mapped = both(f, rays, dt * steps, dt, stride=stride, chunk_rays=(0 if seed % 4 else 192))
        sep, sg = same_grid_pair(f, rays, dt * steps, dt, stride=stride)
    assert_parity(mapped, ref, what=f"fuzz seed {seed} with the depth-floor map")
    assert_same(mapped, plain, f"fuzz seed {seed}")
    assert_same(sg, sep, f"fuzz seed {seed}: same-grid shortcut")


def test_blocks_with_dry_and_non_finite_nodes_fall_back_to_the_lookup(oracle, gpu):
    """A deep basin (2 km) with a dry node, a +inf, a -inf and a NaN node, each in a different 8 x 8 block, and a
    shoal whose block bound is too low for the shorter waves: rays through all of them, every wavenumber from
    'deep everywhere' to 'deep nowhere', degenerate ones included."""
    n, d = 80, 50.0
    x = (np.arange(n) * d).astype(np.float32)
    X, Y = np.meshgrid(np.arange(n), np.arange(n))
    depth = 2000.0 + 100.0 * np.sin(X / 7.0) * np.cos(Y / 5.0)
    depth[20, 20], depth[20, 44], depth[44, 20], depth[44, 44] = 0.0, np.inf, -np.inf, np.nan
    depth[60:70, 10:30] = 12.0                                   # a shoal
    bathy = CartesianNetcdf3(x, x, depth)
    u = 0.3 * np.sin(Y / 9.0)
    v = 0.2 * np.cos(X / 11.0)
    for cur in (CartesianCurrent(x.astype(np.float64), x.astype(np.float64), u, v), ConstantCurrent(0.1, -0.2)):
        rng = np.random.default_rng(7)
        m = 4096
        x0, y0 = rng.uniform(-50, n * d, m), rng.uniform(-50, n * d, m)
        x0[:64] = rng.choice(x, 64)                              # on grid lines and nodes
        y0[32:96] = rng.choice(x, 64)
        kmag = 10.0 ** rng.uniform(-3.2, 0.3, m)                 # kh from 1 to 4000 over the basin
        th = rng.uniform(0, 2 * np.pi, m)
        kx0, ky0 = kmag * np.cos(th), kmag * np.sin(th)
        kx0[100:108] = [0.0, -0.0, 1e200, 1e-200, np.nan, np.inf, 1e6, 3e19]
        ky0[100:108] = [0.0, 0.5, 1e200, 0.0, 0.1, 0.0, -1e6, 3e19]
        dt, steps = 2.0, 300
        ref = oracle.trace_many(bathy, cur, x0, y0, kx0, ky0, 0.0, dt * steps, dt)
        with Fields(bathy, cur, devices=[0]) as f:
            plain, mapped = both(f, (x0, y0, kx0, ky0), dt * steps, dt)
            sep, sg = same_grid_pair(f, (x0, y0, kx0, ky0), dt * steps, dt)
        assert_parity(mapped, ref, what="dry / non-finite blocks with the depth-floor map")
        assert_same(mapped, plain, "dry / non-finite blocks")
        assert_same(sg, sep, "dry / non-finite blocks: same-grid shortcut")


def test_default_follows_the_deep_share_of_the_grid(gpu):
    """flags = 0: the map is used on C4's deep basin (share >= 1/4) and not on C5's 400 m shelf — there, with no map
    in use and the current on the bathymetry's grid, the same-grid shortcut is; the library says so (mr_trace_plan)
    and the results are those of the forced variant, bit for bit.  MR_OPT_NO_DEEP_MAP wins over MR_OPT_DEEP_MAP."""
    from mantaray_b200 import depth_floor_map
    from mantaray_b200._abi import MR_OPT_SAME_GRID
    for wl, deep in ((W.c4_agulhas(16, 16, 300, nx=256), True), (W.c5_nazare(4, 4, 16, 600, nx=512), False)):
        _, share, affine = depth_floor_map(wl.bathymetry)
        assert affine and (share >= 0.25) == deep
        rays = wl.all_rays()
        with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
            plan = f.plan()
            assert bool(plan & MR_PLAN_DEEP_MAP) == deep and bool(plan & MR_PLAN_SAME_GRID) == (not deep)
            assert f.plan(flags=MR_OPT_NO_SAME_GRID | MR_OPT_NO_DEEP_MAP) & (MR_PLAN_DEEP_MAP | MR_PLAN_SAME_GRID) == 0
            assert f.plan(math=1) == 0                                   # MR_MATH_STRICT has no specialisations
            plain, mapped = both(f, rays, wl.duration, wl.dt, stride=wl.stride)
            shortcut = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, stride=wl.stride,
                                  flags=MR_OPT_NO_DEEP_MAP | MR_OPT_SAME_GRID)
            auto = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, stride=wl.stride)
            off = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, stride=wl.stride,
                             flags=MR_OPT_DEEP_MAP | MR_OPT_NO_DEEP_MAP | MR_OPT_NO_SAME_GRID)
        assert_same(auto, mapped if deep else shortcut, "default flags", exact=True)
        assert_same(off, plain, "both flags", exact=True)
        assert_same(mapped, plain, "forced")
        assert_same(shortcut, plain, "same-grid shortcut")


def test_steep_cells_at_large_indices(oracle, gpu):
    """The map's bound must hold where the lookup EXTRAPOLATES: the cell is floor() of an f32 index whose rounding
    error grows with the index, so near column 4000 a point up to ~2e-4 cells outside a cell is still assigned to it
    and the bilinear reaches beyond the corner values by that much of the corner difference.  Steep steps (hundreds
    of metres per cell) straddling kh = 22 at large indices, rays sitting a few f32 ulps either side of the grid
    lines: with the map and without it the rows, len and NaN patterns must be identical and the values agree to
    ulps (a lane wrongly taken for deep would differ at 1e-10 or more: exp(-44) against 0 in tanh)."""
    nx, ny, d = 4096, 12, 25.0
    x = (np.arange(nx) * d).astype(np.float32)
    y = (np.arange(ny) * d).astype(np.float32)
    rng = np.random.default_rng(11)
    k0 = 0.0402                                           # 10 s wave: kh = 22 at h = 547 m
    # columns alternate between just shallower and much deeper than 547 m, with jitter: every 8x8 block mixes them
    col = np.where(np.arange(nx) % 3 == 0, 540.0, 900.0) + rng.uniform(-6.0, 6.0, nx)
    depth = np.tile(col, (ny, 1)) + rng.uniform(-3.0, 3.0, (ny, nx))
    bathy = CartesianNetcdf3(x, y, depth)
    cur = CartesianCurrent(x.astype(np.float64), y.astype(np.float64), 0.05 * np.ones((ny, nx)), np.zeros((ny, nx)))
    m = 6000
    i = rng.integers(3000, nx - 2, m)
    ulps = rng.integers(-4, 5, m) * 2.0 ** -24
    x0 = x[i].astype(np.float64) * (1.0 + ulps) + np.where(rng.random(m) < 0.5, 0.0, rng.uniform(0, d, m))
    y0 = rng.uniform(2 * d, (ny - 3) * d, m)
    kk = k0 * rng.uniform(0.97, 1.03, m)                  # kh between 21 and 37 over these depths
    th = rng.uniform(-0.3, 0.3, m)
    rays = (x0, y0, kk * np.cos(th), kk * np.sin(th))
    dt, steps = 0.25, 40
    ref = oracle.trace_many(bathy, cur, *rays, 0.0, dt * steps, dt)
    with Fields(bathy, cur, devices=[0]) as f:
        plain, mapped = both(f, rays, dt * steps, dt)
    assert_parity(mapped, ref, what="steep cells at large indices, with the map")
    assert_same(mapped, plain, "steep cells at large indices")
