"""MR_OPT_DEEP_MAP (include/mantaray_b200.h): the fast path skips the depth lookup wherever a per-block lower
bound of the depth proves kh >= 22.  The flag is OFF by default this round — it was measured (C4: 63.8 -> 56.9 ms
per 1M-ray launch, identical rows / len / final-state checksums) after the round's GPU time for the full parity
suite had run out — and so are these tests: run them with MR_TEST_DEEP_MAP=1.  They hold the flagged path to
the oracle (same bar as everywhere) and to the unflagged path (identical up to the sign of an exact zero)."""

import os

import numpy as np
import pytest

from conftest import assert_parity
from mantaray_b200 import MR_MATH_FAST, CartesianCurrent, CartesianNetcdf3, ConstantCurrent, Fields, trace_many
from mantaray_b200 import workloads as W
from mantaray_b200._abi import MR_OPT_DEEP_MAP
from test_gpu_fuzz import make_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MR_TEST_DEEP_MAP") != "1",
                                 reason="opt-in: MR_OPT_DEEP_MAP is not enabled by default this round (MR_TEST_DEEP_MAP=1)")]


def both(f, rays, t_end, dt, **kw):
    plain = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, **kw)
    mapped = trace_many(f, *rays, 0.0, t_end, dt, math=MR_MATH_FAST, final_state=True, flags=MR_OPT_DEEP_MAP, **kw)
    return plain, mapped


def assert_same(mapped, plain, what):
    np.testing.assert_array_equal(mapped.rows, plain.rows, err_msg=f"{what}: rows")
    np.testing.assert_array_equal(mapped.len, plain.len, err_msg=f"{what}: len")
    for name in ("x", "y", "kx", "ky", "final_state"):
        # assert_array_equal: NaN == NaN and -0 == +0, everything else bit for bit
        np.testing.assert_array_equal(getattr(mapped, name), getattr(plain, name), err_msg=f"{what}: {name}")


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
    assert_parity(mapped, ref, what=f"{name} with the depth-floor map")
    assert_same(mapped, plain, name)


@pytest.mark.parametrize("seed", range(int(os.environ.get("MR_FUZZ_SEEDS", "120"))))
def test_fuzz_with_the_depth_floor_map(oracle, gpu, seed):
    bathy, cur, rays, dt, steps = make_case(seed)
    stride = 1 if seed % 5 else 7
    ref = oracle.trace_many(bathy, cur, *rays, 0.0, dt * steps, dt, stride=stride)
    with Fields(bathy, cur, devices=[0]) as f:
        plain, mapped = both(f, rays, dt * steps, dt, stride=stride, chunk_rays=(0 if seed % 4 else 192))
    assert_parity(mapped, ref, what=f"fuzz seed {seed} with the depth-floor map")
    assert_same(mapped, plain, f"fuzz seed {seed}")


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
        assert_parity(mapped, ref, what="dry / non-finite blocks with the depth-floor map")
        assert_same(mapped, plain, "dry / non-finite blocks")
