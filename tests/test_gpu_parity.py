"""CUDA path vs the CPU oracle, through the C ABI, on the same seeded inputs.

Bar (BASELINE.json north_star): per-ray termination step identical (bit-exact
``rows`` and ``len``), trajectories within 1e-9 relative in position and
wavenumber (metric in conftest.assert_parity).  Both arithmetic modes are held
to it.
"""

import numpy as np
import pytest

from conftest import assert_parity
from mantaray_b200 import (MR_MATH_FAST, MR_MATH_STRICT, ArrayDepth, CartesianCurrent, CartesianNetcdf3,
                           ConstantCurrent, ConstantDepth, ConstantSlope, Fields, trace_many)
from mantaray_b200 import workloads as W
from mantaray_b200._abi import MR_OPT_DEEP_MAP, MR_OPT_SAME_GRID

pytestmark = pytest.mark.gpu

MODES = [pytest.param(MR_MATH_FAST, id="fast"), pytest.param(MR_MATH_STRICT, id="strict")]


def run_both(oracle, gpu, bathy, cur, rays, t0, t_end, dt, math, **kw):
    x0, y0, kx0, ky0 = rays
    ref = oracle.trace_many(bathy, cur, x0, y0, kx0, ky0, t0, t_end, dt, stride=kw.get("stride", 1))
    with Fields(bathy, cur) as f:
        res = trace_many(f, x0, y0, kx0, ky0, t0, t_end, dt, math=math, final_state=True, **kw)
    return res, ref


@pytest.mark.parametrize("math", MODES)
@pytest.mark.parametrize("name,make", [
    ("C1", lambda: W.c1_canonical(300, 1500)),
    ("C2", lambda: W.c2_sea_mount(1000, 2000)),
    ("C3", lambda: W.c3_shear_jet(500, 3000)),
    ("C4", lambda: W.c4_agulhas(24, 24, 2048)),
    ("C5", lambda: W.c5_nazare(6, 6, 32, 4096, 1024, 64)),
])
def test_workload_parity(oracle, gpu, name, make, math):
    wl = make()
    res, ref = run_both(oracle, gpu, wl.bathymetry, wl.current, wl.all_rays(), 0.0, wl.duration, wl.dt, math,
                        stride=wl.stride)
    worst = assert_parity(res, ref, what=f"{name}")
    # Regression guard, far inside the 1e-9 contract: with the f32 stage value-identical and the f64
    # stage good to a few ulp per evaluation, these shapes agree to ~1e-15.  A packed-f32 build in which
    # ptxas had contracted mul+add pairs of the bilinear still met 1e-9 (errors ~1e-10); this catches it.
    assert worst <= 1e-12, f"{name}: worst relative error {worst:.3e} — the f32 stage is no longer exact"
    np.testing.assert_array_equal(res.t, ref.t)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(np.isnan(res.final_state), np.isnan(ref.final_state))
    print(f"{name}: worst relative error {worst:.3e}, executed ray-steps {int((ref.rows - 1).sum())}")


@pytest.mark.parametrize("math", MODES)
def test_random_grids_parity(oracle, gpu, math):
    """Seeded random smooth fields on grids whose f32 coordinates are NOT exactly
    representable multiples (generic spacing, non-zero origin), rays in every direction."""
    rng = np.random.default_rng(1234)
    nx, ny = 97, 61
    x = (-1234.5 + 37.3 * np.arange(nx)).astype(np.float32)
    y = (987.25 + 41.7 * np.arange(ny)).astype(np.float32)
    X, Y = np.meshgrid(x.astype(np.float64), y.astype(np.float64))
    depth = 30.0 + 25.0 * np.sin(X / 700.0) * np.cos(Y / 500.0) + rng.normal(0, 0.5, X.shape)
    cx = -1300.0 + 41.0 * np.arange(90)
    cy = 900.0 + 43.0 * np.arange(64)
    CX, CY = np.meshgrid(cx, cy)
    u = 0.8 * np.sin(CY / 600.0) + rng.normal(0, 0.01, CX.shape)
    v = 0.5 * np.cos(CX / 800.0) + rng.normal(0, 0.01, CX.shape)
    bathy, cur = CartesianNetcdf3(x, y, depth), CartesianCurrent(cx, cy, u, v)
    n = 3000
    x0 = rng.uniform(x[0], x[-1], n)
    y0 = rng.uniform(y[0], y[-1], n)
    th = rng.uniform(0, 2 * np.pi, n)
    k = rng.uniform(0.02, 0.6, n)
    res, ref = run_both(oracle, gpu, bathy, cur, (x0, y0, k * np.cos(th), k * np.sin(th)), 0.0, 400.0, 0.5, math)
    assert ref.rows.min() < ref.rows.max(), "the case should contain rays that leave the domain"
    assert_parity(res, ref, what="random grids")


@pytest.mark.parametrize("math", MODES)
@pytest.mark.parametrize("bathy", [
    ConstantDepth(10.0), ConstantDepth(2000.0), ConstantDepth(0.0),
    ConstantSlope(100.0, 0.0, 0.0, -0.05, 0.0), ConstantSlope(50.0, 10.0, -5.0, 0.02, -0.03),
    ArrayDepth(np.full((40, 40), 1000.0)),
], ids=["h10", "h2000", "h0", "slope_x", "slope_xy", "array"])
@pytest.mark.parametrize("cur", [ConstantCurrent(0.0, 0.0), ConstantCurrent(0.5, -0.25)], ids=["still", "uv"])
def test_analytic_fields(oracle, gpu, bathy, cur, math):
    rng = np.random.default_rng(7)
    n = 257
    th = rng.uniform(0, 2 * np.pi, n)
    k = rng.uniform(0.01, 1.0, n)
    rays = (rng.uniform(0, 30, n), rng.uniform(0, 30, n), k * np.cos(th), k * np.sin(th))
    res, ref = run_both(oracle, gpu, bathy, cur, rays, 0.0, 60.0, 0.5, math)
    assert_parity(res, ref, what="analytic")


@pytest.mark.parametrize("math", MODES)
@pytest.mark.parametrize("bathy", [
    ConstantDepth(10.0), ConstantSlope(50.0, 10.0, -5.0, 0.02, -0.03), ArrayDepth(np.arange(1.0, 1601.0).reshape(40, 40)),
], ids=["constant", "slope", "array"])
def test_special_inputs_analytic(oracle, gpu, bathy, math):
    """NaN / inf / negative / huge start coordinates on the analytic kinds.  ArrayDepth indexes with a
    saturating cast (array_depth.rs:19-20: NaN -> 0, negative -> 0), so a ray with a NaN x keeps a finite
    depth and wavenumber derivative and keeps integrating with x = NaN: `rows` tells that apart from a stop."""
    x0 = np.array([np.nan, 3.0, -4.0, 5.0, 1e30, np.inf, 2.5, 39.9, 41.0, 0.0])
    y0 = np.array([2.0, np.nan, 3.0, -2.0, 1.0, 1.0, -np.inf, 39.9, 1.0, 0.0])
    n = x0.size
    rays = (x0, y0, np.full(n, 0.05), np.full(n, 0.02))
    res, ref = run_both(oracle, gpu, bathy, ConstantCurrent(0.1, 0.0), rays, 0.0, 30.0, 0.5, math)
    assert_parity(res, ref, what="special analytic")
    assert np.array_equal(res.rows, ref.rows)
    if isinstance(bathy, ArrayDepth):
        assert ref.rows[0] > 3 and ref.len[0] == 0          # NaN x: integrates on (until y leaves the array), never NaN-free
        assert ref.rows[8] == 2                              # x = 41 is outside the 40 x 40 array: NaN depth


@pytest.mark.parametrize("math", MODES)
def test_infinite_depth_nodes(oracle, gpu, math):
    """grid nodes holding +inf / -inf / NaN: cells touching them give a non-finite depth or gradient; rays that
    run into them stop where the reference stops them, with the same rows"""
    x = (100.0 * np.arange(40)).astype(np.float32)
    y = (100.0 * np.arange(30)).astype(np.float32)
    depth = np.full((30, 40), 200.0)
    depth[10, 12] = np.inf
    depth[20, 25] = -np.inf
    depth[5, 30] = np.nan
    depth[15:18, 20:23] = np.inf            # a whole block: inf - inf gradients
    cx = 100.0 * np.arange(40.0)
    cy = 100.0 * np.arange(30.0)
    cur = CartesianCurrent(cx, cy, np.full((30, 40), 0.2), np.full((30, 40), -0.1))
    bathy = CartesianNetcdf3(x, y, depth)
    n = 600
    rng = np.random.default_rng(11)
    th = rng.uniform(-0.5, 0.5, n)
    rays = (np.full(n, 150.0), np.linspace(150.0, 2750.0, n), 0.05 * np.cos(th), 0.05 * np.sin(th))
    res, ref = run_both(oracle, gpu, bathy, cur, rays, 0.0, 600.0, 2.0, math)
    assert_parity(res, ref, what="non-finite nodes")
    assert np.array_equal(res.rows, ref.rows)
    assert ref.rows.min() < 100 and ref.rows.max() > 250
    # rays whose last rows are partially NaN (cg = NaN from an infinite depth, wavenumber still finite)
    assert (ref.rows - ref.len >= 2).sum() > 5


@pytest.mark.parametrize("math", MODES)
def test_special_inputs(oracle, gpu, math):
    """NaN / zero-k / out-of-domain starts / inf, on gridded fields."""
    wl = W.c2_sea_mount(8, 50, half=100)
    nan, inf = np.nan, np.inf
    x0 = np.array([nan, 0.0, 0.0, 0.0, 5000.0, -900.0, inf, -990.0, 0.0, -1000.0, 1000.0])
    y0 = np.array([0.0, nan, 100.0, 100.0, 0.0, 0.0, 0.0, 990.0, 0.0, -1000.0, 1000.0])
    kx = np.array([0.1, 0.1, nan, 0.0, 0.1, 0.1, 0.1, -0.1, 0.1, 0.1, -0.1])
    ky = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.1, 0.0, 0.1, -0.1])
    res, ref = run_both(oracle, gpu, wl.bathymetry, wl.current, (x0, y0, kx, ky), 0.0, 40.0, 1.0, math)
    assert_parity(res, ref, what="special")
    assert ref.rows[0] == 2 and ref.len[0] == 0
    assert ref.rows[3] == 2 and ref.len[3] == 1          # k == 0 -> Err -> NaN row


@pytest.mark.parametrize("math", MODES)
def test_stride_final_and_len_only(oracle, gpu, math):
    wl = W.c2_sea_mount(700, 600, half=300)
    rays = wl.all_rays()
    full, ref = run_both(oracle, gpu, wl.bathymetry, wl.current, rays, 0.0, wl.duration, wl.dt, math)
    assert_parity(full, ref, what="stride1")
    for stride in (3, 64, 601):
        res, r2 = run_both(oracle, gpu, wl.bathymetry, wl.current, rays, 0.0, wl.duration, wl.dt, math, stride=stride)
        assert_parity(res, r2, what=f"stride{stride}")
        # decimation picks rows of the stride-1 run bit for bit
        np.testing.assert_array_equal(res.x, full.x[::stride][: res.x.shape[0]])
        np.testing.assert_array_equal(res.t, full.t[::stride][: res.t.shape[0]])
    with Fields(wl.bathymetry, wl.current) as f:
        lo = trace_many(f, *rays, 0.0, wl.duration, wl.dt, math=math, trajectories=False, final_state=True)
    np.testing.assert_array_equal(lo.rows, full.rows)
    np.testing.assert_array_equal(lo.len, full.len)
    np.testing.assert_array_equal(lo.final_state, full.final_state)
    # final_state is the last NaN-free row
    i = np.arange(full.len.size)
    ok = full.len > 0
    np.testing.assert_array_equal(full.final_state[0][ok], full.x[full.len[ok] - 1, i[ok]])
    np.testing.assert_array_equal(full.final_state[3][ok], full.ky[full.len[ok] - 1, i[ok]])


def test_chunked_host_path_is_identical(gpu):
    """Slabs of rays with the drain overlapped give bit-identical output to one slab."""
    wl = W.c4_agulhas(20, 20, 300)
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current) as f:
        one = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True)
        many = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, chunk_rays=96)
        pin = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, chunk_rays=130, pinned=True)
    for other in (many, pin):
        for name in ("t", "x", "y", "kx", "ky", "rows", "len", "final_state"):
            np.testing.assert_array_equal(getattr(one, name), getattr(other, name), err_msg=name)


def test_empty_and_ragged(oracle, gpu):
    wl = W.c1_canonical(5, 10)
    with Fields(wl.bathymetry, wl.current) as f:
        r = trace_many(f, [], [], [], [], 0.0, 10.0, 1.0)
        assert r.x.shape == (11, 0) and r.rows.size == 0 and r.t.size == 11
        # zip() semantics: truncated to the shortest input (src/ffi.rs:65-70)
        r = trace_many(f, [10.0, 10.0, 10.0], [0.0, 1e3], [0.04, 0.04, 0.04], [0.0, 0.0, 0.0], 0.0, 10.0, 1.0)
        assert r.x.shape == (11, 2)
        # zero steps: one row
        r = trace_many(f, [10.0], [0.0], [0.04], [0.0], 0.0, 0.0, 1.0)
        assert r.x.shape == (1, 1) and r.rows[0] == 1 and r.len[0] == 1
        with pytest.raises(ValueError):
            trace_many(f, [10.0], [0.0], [0.04], [0.0], 0.0, 10.0, 0.0)
        # a negative duration is zero steps too (`as usize` saturates; tests/test_gpu_api.py holds it to the oracle)
        r = trace_many(f, [10.0], [0.0], [0.04], [0.0], 0.0, -10.0, 1.0)
        assert r.x.shape == (1, 1) and r.rows[0] == 1 and r.len[0] == 1


def test_permutation_and_restart_properties_full_size(gpu):
    """Size-independent properties at a BASELINE-size batch (1M rays, C4 fields), results
    kept on the host only as final states:
      * permuting the rays permutes the results bit for bit (rays are independent);
      * the stepper is memoryless: S steps == S/2 steps, then S/2 more from the final state.
    """
    wl = W.c4_agulhas(1000, 1000, 64)
    x0, y0, kx0, ky0 = wl.all_rays()
    n = x0.size
    perm = np.random.default_rng(3).permutation(n)
    with Fields(wl.bathymetry, wl.current) as f:
        a = trace_many(f, x0, y0, kx0, ky0, 0.0, wl.duration, wl.dt, trajectories=False, final_state=True)
        b = trace_many(f, x0[perm], y0[perm], kx0[perm], ky0[perm], 0.0, wl.duration, wl.dt,
                       trajectories=False, final_state=True)
        np.testing.assert_array_equal(a.rows[perm], b.rows)
        np.testing.assert_array_equal(a.final_state[:, perm], b.final_state)
        half = wl.dt * 32
        h1 = trace_many(f, x0, y0, kx0, ky0, 0.0, half, wl.dt, trajectories=False, final_state=True)
        alive = h1.len == 33
        fs = h1.final_state
        h2 = trace_many(f, fs[0], fs[1], fs[2], fs[3], 0.0, half, wl.dt, trajectories=False, final_state=True)
        both = alive & (a.len == 65)
        assert both.sum() > 0.9 * n
        np.testing.assert_array_equal(h2.final_state[:, both], a.final_state[:, both])
        np.testing.assert_array_equal((h1.rows + h2.rows - 1)[alive], a.rows[alive])


@pytest.mark.parametrize("math", [MR_MATH_FAST, MR_MATH_STRICT])
def test_rays_creeping_up_to_the_shoreline(oracle, gpu, math):
    """C5's beach: some long-period rays neither run ashore nor leave — they creep up to a shoreline node a few
    femtometres deep while k grows without bound (1e15 ... 1e23 rad/m after 4096 steps), with kh >= 22 all the way.
    The deep-water shortcut of the fast path must not be taken there: what it drops from dk/dt is
    2 exp(-2kh) sqrt(G k) |grad h| relative to k, and sqrt(G k) is 1e8 per second at such k (this case was off by
    1.3e-9 in kx before the shortcut was limited to k <= 16 rad/m)."""
    wl = W.c5_nazare(4, 4, 64, 4096, nx=512)
    rays = wl.all_rays()
    ref = oracle.trace_many(wl.bathymetry, wl.current, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride)
    kfin = np.hypot(ref.final_state[2], ref.final_state[3])
    assert np.nanmax(kfin) > 1e15 and (ref.rows[kfin > 1e6] == wl.n_steps + 1).all()      # the creepers are still alive
    with Fields(wl.bathymetry, wl.current) as f:
        for flags in ((0, MR_OPT_DEEP_MAP, MR_OPT_SAME_GRID) if math == MR_MATH_FAST else (0,)):
            res = trace_many(f, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride, math=math, final_state=True, flags=flags)
            assert_parity(res, ref, what=f"shoreline creepers math={math} flags={flags}")
            creep = kfin > 1e3
            with np.errstate(invalid="ignore"):
                err = np.abs(res.final_state[2:, creep] - ref.final_state[2:, creep]) / kfin[creep]
            assert np.nanmax(err) <= 1e-12, f"k of the creeping rays off by {np.nanmax(err):.2e} (math={math}, flags={flags})"
