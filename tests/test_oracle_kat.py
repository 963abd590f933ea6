"""Pins the CPU oracle to every known-answer value the reference's own tests hold for the path.

The reference (Rust) cannot be built in this image, so these golden vectors — copied as numbers
from the reference's test modules, with file:line — are what anchors the oracle
(SURVEY.md 8c).  Tolerances are the reference's own.
"""

import math

import numpy as np
import pytest

from mantaray_b200 import (ArrayDepth, CartesianCurrent, CartesianNetcdf3, ConstantCurrent, ConstantDepth,
                           ConstantSlope)
from mantaray_b200.io_utility import create_netcdf3_bathymetry, create_netcdf3_current

F32_EPS = float(np.finfo(np.float32).eps)
F64_EPS = float(np.finfo(np.float64).eps)


# ---- interpolator::bilinear (src/interpolator.rs:86-161) ----------------------------------------------
def _pts(x1, y1, x2, y2, q11, q21, q12, q22, t):
    return [(x1 + t, y1 + t, q11), (x1 + t, y2 + t, q12), (x2 + t, y2 + t, q22), (x2 + t, y1 + t, q21)]


def test_bilinear_golden(oracle):
    t = 1230.0
    ans = oracle.bilinear(_pts(-77.0, -19.0, 123.0, 145.0, 10.0, 20.0, 30.0, 40.0, t), (20.0 + t, 23.0 + t))
    # :91-104, :116-121 — every tuple member is f32 there, so the literal is compared as f32
    assert abs(ans - float(np.float32(19.971951219512192))) < F32_EPS
    assert ans == float(np.float32(19.971951219512192)) == 19.97195053100586


@pytest.mark.parametrize("x,y,val", [(0.0, 0.0, 0.0), (10.0, 0.0, 5.0), (0.0, 10.0, 10.0), (10.0, 10.0, 15.0)])
def test_bilinear_corner_coincidence(oracle, x, y, val):
    ans = oracle.bilinear(_pts(0.0, 0.0, 10.0, 10.0, 0.0, 5.0, 10.0, 15.0, 0.0), (x, y))   # :129-160
    assert abs(ans - val) < F32_EPS


def test_bilinear_degenerate_is_err(oracle):
    with pytest.raises(oracle.Err):                         # det == 0, :65-67
        oracle.bilinear([(0, 0, 1.0), (0, 0, 2.0), (1, 0, 3.0), (1, 0, 4.0)], (0.5, 0.5))


# ---- group velocity / odes (src/wave_ray_path.rs:296-358, 698-768) -----------------------------------
@pytest.mark.parametrize("k,cg", [(1.0, 1.565247584249853), (3.0, 0.9036961141150639), (5.0, 0.7), (10.0, 0.4949747468305833)])
def test_group_velocity(oracle, k, cg):
    assert abs(oracle.group_velocity(k, 1000.0) - cg) < 1.0e-4      # :301-314


def test_group_velocity_bit_value(oracle):
    assert oracle.group_velocity(1.0, 1000.0) == 1.565247584249853  # the reference's own digits (:302, :706)


def test_negative_k_is_err(oracle):
    for k in (-1.0, -12.0):                                         # :319-326
        with pytest.raises(oracle.Err):
            oracle.group_velocity(k, 1000.0)
    assert math.isnan(oracle.group_velocity(1.0, -5.0))             # h <= 0 -> NaN (:178-180)
    assert math.isnan(oracle.group_velocity(-1.0, 0.0))             # ... checked before k


@pytest.mark.parametrize("kx,ky,dx,dy", [(1.0, 0.0, 1.565247584249853, 0.0), (0.0, 1.0, 0.0, 1.565247584249853),
                                         (-1.0, 0.0, -1.565247584249853, 0.0), (0.0, -1.0, 0.0, -1.565247584249853)])
def test_odes_axes(oracle, kx, ky, dx, dy):
    out = oracle.odes(ConstantDepth(1000.0), ConstantCurrent(0, 0), 0.0, 0.0, kx, ky)   # :331-357, :657-694
    assert abs(out[0] - dx) < 1e-4 and abs(out[1] - dy) < 1e-4


@pytest.mark.parametrize("u,v", [(1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0), (1.0, 1.0), (-1.0, -1.0)])
def test_odes_constant_current_superposition(oracle, u, v):
    out = oracle.odes(ConstantDepth(1000.0), ConstantCurrent(u, v), 0.0, 0.0, 1.0, 0.0)  # :704-767
    assert abs(out[0] - (1.565247584249853 + u)) < F64_EPS
    assert abs(out[1] - (0.0 + v)) < F64_EPS


def test_dk_deep_is_zero(oracle):
    a, b = oracle.dkdt_bathy(1000.0, 1000.0, 0.2, 0.2)              # :574-597
    assert abs(a) < F64_EPS and abs(b) < F64_EPS


# ---- one RK4 step (src/wave_ray_path.rs:405-435, 521-548) ------------------------------------------------
def _one_step(oracle, bathy, kx, ky):
    out = oracle.single_ray(bathy, ConstantCurrent(0, 0), 0.0, 0.0, kx, ky, 0.0, 1.0, 1.0)
    assert out.shape == (2, 5)
    return out[-1, 1], out[-1, 2]


@pytest.mark.parametrize("bathy", [ConstantDepth(1000.0), ArrayDepth(np.full((3, 3), 1000.0))], ids=["constant", "array"])
def test_rk4_axis(oracle, bathy):
    s = math.sqrt(9.8) / 2.0
    for kx, ky, xf, yf in [(0.0, 1.0, 0.0, s), (1.0, 0.0, s, 0.0), (0.0, -1.0, 0.0, -s), (-1.0, 0.0, -s, 0.0)]:
        x, y = _one_step(oracle, bathy, kx, ky)
        assert abs(x - xf) < F64_EPS and abs(y - yf) < F64_EPS      # :280-281


def test_rk4_shallow_goldens(oracle):
    """h = 0.1: 16-digit values including the 6.03e-17 / 1.21e-16 cross terms (:524-545)."""
    cases = [(0.0, 1.0, 0.00000000000000006031543168844801, 0.9850257515953494),
             (1.0, 0.0, 0.9850257515953494, 0.0),
             (0.0, -1.0, 0.00000000000000006031543168844801, -0.9850257515953494),
             (-1.0, 0.0, -0.9850257515953494, 0.00000000000000012063086337689602)]
    for kx, ky, xf, yf in cases:
        x, y = _one_step(oracle, ConstantDepth(0.1), kx, ky)
        assert abs(x - xf) < F64_EPS and abs(y - yf) < F64_EPS
    # and to the digit
    x, y = _one_step(oracle, ConstantDepth(0.1), 0.0, 1.0)
    assert (x, y) == (6.031543168844801e-17, 0.9850257515953494)
    x, y = _one_step(oracle, ConstantDepth(0.1), -1.0, 0.0)
    assert (x, y) == (-0.9850257515953494, 1.2063086337689602e-16)


# ---- NaN propagation and early stop (src/wave_ray_path.rs:362-517, 552-626) -------------------------
def test_zero_k_and_zero_h(oracle):
    out = oracle.single_ray(ConstantDepth(1000.0), ConstantCurrent(0, 0), 0, 0, 0.0, 0.0, 0.0, 10.0, 1.0)   # :362-380
    assert np.isnan(out[-1, 1:]).all()
    out = oracle.single_ray(ConstantDepth(0.0), ConstantCurrent(0, 0), 0, 0, 1.0, 1.0, 0.0, 10.0, 1.0)      # :383-401
    assert np.isnan(out[-1, 1:]).all()


@pytest.mark.parametrize("state,cols", [((math.nan, 0.0, 1.0, 0.0), [1]), ((0.0, math.nan, 1.0, 0.0), [2]),
                                        ((0.0, 0.0, math.nan, 0.0), [1, 2]), ((0.0, 0.0, 0.0, math.nan), [1, 2])])
def test_nan_inputs(oracle, state, cols):
    out = oracle.single_ray(ConstantDepth(1000.0), ConstantCurrent(0, 0), *state, 0.0, 1.0, 1.0)           # :439-517
    for c in cols:
        assert math.isnan(out[-1, c])


def test_out_of_range_and_solout(oracle):
    bathy = ArrayDepth(np.full((2, 2), 1000.0))
    out = oracle.single_ray(bathy, ConstantCurrent(0, 0), 0.0, 0.0, 0.0, 1.0, 0.0, 10.0, 1.0)
    assert out.shape[0] == 3                                        # :617 stops long before t = 10
    assert np.isnan(out[-1, 1:]).all()                              # :620-625
    assert not np.isnan(out[1, 1:]).any()
    assert out[-1, 0] == 2.0                                        # the NaN row still carries its time


def test_row_count_is_ceil_plus_one(oracle):
    """python/tests/test_core.py:51 — duration 10, step 2 -> 6 rows; also a non-dividing step."""
    b, c = ConstantDepth(10_000.0), ConstantCurrent(0.01, 0.01)
    assert oracle.single_ray(b, c, -1000, 0, 0.01, 0, 0.0, 10.0, 2.0).shape[0] == 6
    assert oracle.single_ray(b, c, -1000, 0, 0.01, 0, 0.0, 10.0, 3.0).shape[0] == 5       # ceil(10/3) + 1
    out = oracle.single_ray(b, c, -1000, 0, 0.01, 0, 0.0, 10.0, 3.0)
    assert out[-1, 0] == 12.0                                        # the last step is not shortened


def test_len_is_leading_nan_free_rows(oracle):
    """RayResult::from truncates at the first row with ANY NaN (src/ray_result.rs:134-147, test :187-206)."""
    bathy = ArrayDepth(np.full((2, 2), 1000.0))
    r = oracle.trace_many(bathy, ConstantCurrent(0, 0), [0.0, 0.0], [0.0, 0.0], [0.0, math.nan], [1.0, 1.0], 0.0, 10.0, 1.0)
    assert list(r.rows) == [3, 2] and list(r.len) == [2, 0]
    assert np.isnan(r.x[2:, 0]).all() and not np.isnan(r.x[:2, 0]).any()
    np.testing.assert_array_equal(r.final_state[:, 0], [r.x[1, 0], r.y[1, 0], r.kx[1, 0], r.ky[1, 0]])
    assert np.isnan(r.final_state[:, 1]).all()


# ---- analytic fields ----------------------------------------------------------------------------------------
def test_constant_fields(oracle):
    assert oracle.depth_and_gradient(ConstantDepth(1000.0), 5.0, -3.0) == (1000.0, (0.0, 0.0))
    h, (gx, gy) = oracle.depth_and_gradient(ConstantDepth(1000.0), math.nan, 0.0)           # constant_depth.rs:39-45
    assert math.isnan(h) and math.isnan(gx) and math.isnan(gy)
    s = ConstantSlope(100.0, 0.0, 0.0, -0.05, 0.0)
    assert oracle.depth_and_gradient(s, 1000.0, 77.0) == (50.0, (float(np.float32(-0.05)), 0.0))
    assert math.isnan(oracle.depth(s, 0.0, math.nan))                                        # constant_slope.rs:55-57
    (u, v), (du, dv) = oracle.current_and_gradient(ConstantCurrent(0.25, -3.0), 1e9, math.nan)
    assert (u, v, du, dv) == (0.25, -3.0, (0.0, 0.0), (0.0, 0.0))                            # constant_current.rs:69-77


def test_array_depth(oracle):
    a = ArrayDepth(np.arange(9, dtype=np.float32).reshape(3, 3))
    assert oracle.depth(a, 1.9, 2.1) == 5.0                          # truncating index, array[x][y]
    assert oracle.depth(a, -0.5, 0.0) == 0.0                         # negative saturates to 0
    assert math.isnan(oracle.depth(a, 3.0, 0.0)) and math.isnan(oracle.depth(a, 0.0, 3.0))
    assert oracle.depth(a, math.nan, 0.0) == 0.0                     # NaN as usize == 0 (array_depth.rs:27-28)


# ---- CartesianNetcdf3 grid mechanics (src/bathymetry/cartesian_netcdf3.rs:503-835) -------------------
def four_depth_fn(x, y):                                             # :487-501
    if x < 25000.0:
        return 20.0 if y < 12500.0 else 10.0
    return 5.0 if y < 12500.0 else 15.0


@pytest.fixture(scope="module")
def four_depth(tmp_path_factory):
    p = tmp_path_factory.mktemp("nc") / "four_depth.nc"
    create_netcdf3_bathymetry(p, 101, 51, 500.0, 500.0, four_depth_fn)
    return CartesianNetcdf3.open(p)


def test_bathy_vars_and_nearest(oracle, four_depth):
    d = four_depth
    assert abs(d.x[10] - 5000.0) < F32_EPS                                      # :513
    assert round(oracle.bathy_nearest(5499.0, d.x)) == 11.0                     # :528
    for bad in (-1.0, 25_501.0):                                                # :531-532
        with pytest.raises(oracle.Err):
            oracle.bathy_nearest(bad, d.y)
    assert abs(oracle.bathy_nearest(5500.0, d.x) - 11.0) <= F32_EPS             # :535
    assert round(oracle.bathy_nearest(1.0, d.x)) == 0.0 and round(oracle.bathy_nearest(24_999.0, d.y)) == 50.0   # :550-551
    assert oracle.bathy_nearest(0.0, d.x) == 0.0 and oracle.bathy_nearest(25_000.0, d.y) == 50.0                 # :558-559
    assert oracle.bathy_nearest(3.0, [7.0]) == 0.0                              # single element -> 0 (:281-283)


CORNER_CASES = [
    ((0.0, 25_000.0), [(0, 49), (0, 50), (1, 50), (1, 49)]),          # top left corner      :575-577
    ((0.0, 5_500.0), [(0, 11), (0, 12), (1, 12), (1, 11)]),           # left edge
    ((0.0, 0.0), [(0, 0), (0, 1), (1, 1), (1, 0)]),                   # bottom left corner
    ((5_500.0, 25_000.0), [(11, 49), (11, 50), (12, 50), (12, 49)]),  # top edge
    ((5_500.0, 0.0), [(11, 0), (11, 1), (12, 1), (12, 0)]),           # bottom edge
    ((50_000.0, 25_000.0), [(99, 49), (99, 50), (100, 50), (100, 49)]),
    ((50_000.0, 5_500.0), [(99, 11), (99, 12), (100, 12), (100, 11)]),
    ((50_000.0, 0.0), [(99, 0), (99, 1), (100, 1), (100, 0)]),
    ((5_500.0, 5_500.0), [(11, 11), (11, 12), (12, 12), (12, 11)]),   # both on a grid point  :635-638
    ((5_500.0, 5_750.0), [(11, 11), (11, 12), (12, 12), (12, 11)]),   # only x
    ((5_750.0, 5_500.0), [(11, 11), (11, 12), (12, 12), (12, 11)]),   # only y
    ((5_750.0, 5_750.0), [(11, 11), (11, 12), (12, 12), (12, 11)]),   # neither              :653-656
]
OUT_OF_BOUNDS = [(50_001.0, 0.0), (50_000.0, 25_001.0), (-1.0, 0.0), (50_000.0, -1.0)]   # :614-629


@pytest.mark.parametrize("pt,corners", CORNER_CASES)
def test_bathy_four_corners(oracle, four_depth, pt, corners):
    assert oracle.bathy_four_corners(four_depth, *pt) == corners


@pytest.mark.parametrize("pt", OUT_OF_BOUNDS)
def test_bathy_four_corners_out_of_bounds(oracle, four_depth, pt):
    with pytest.raises(oracle.Err):
        oracle.bathy_four_corners(four_depth, *pt)


def test_bathy_quadrant_depths_oob_nan(oracle, four_depth):
    for x, y, h in [(10099.0, 5099.0, 20.0), (30099.0, 5099.0, 5.0), (10099.0, 15099.0, 10.0), (30099.0, 15099.0, 15.0)]:
        assert abs(oracle.depth_and_gradient(four_depth, x, y)[0] - h) < F32_EPS       # :672-688
    for pt in [(-500.1, 500.1), (500.1, -500.1)]:                                         # :692-722
        with pytest.raises(oracle.Err):
            oracle.depth(four_depth, *pt)
    for pt in [(math.nan, math.nan), (10000.0, math.nan), (math.nan, 10000.0)]:           # :725-741
        assert math.isnan(oracle.depth(four_depth, *pt))


@pytest.mark.parametrize("axis", ["x", "y"])
def test_bathy_gradient_sweep(oracle, tmp_path, axis):
    """depth = 0.05 x (or y) on 100x100 @ 1 m: every grid point gives the gradient (0.05, 0) / (0, 0.05)
    and the exact depth, to f32 EPSILON (:747-834)."""
    fn = (lambda x, y: float(x) * 0.05) if axis == "x" else (lambda x, y: float(y) * 0.05)
    p = tmp_path / "g.nc"
    create_netcdf3_bathymetry(p, 100, 100, 1.0, 1.0, fn)
    d = CartesianNetcdf3.open(p)
    want = (0.05, 0.0) if axis == "x" else (0.0, 0.05)
    for x in range(100):
        for y in range(100):
            h, (gx, gy) = oracle.depth_and_gradient(d, float(x), float(y))
            assert abs(gx - want[0]) < F32_EPS and abs(gy - want[1]) < F32_EPS
            assert abs(h - float(np.float32(fn(np.float32(x), np.float32(y))))) < F32_EPS


# ---- CartesianCurrent grid mechanics (src/current/cartesian_current.rs:545-991) ---------------------------
@pytest.fixture(scope="module")
def simple_current(tmp_path_factory):
    p = tmp_path_factory.mktemp("nc") / "cur.nc"
    create_netcdf3_current(p, 101, 51, 500.0, 500.0, lambda x, y: (5.0, 0.0))
    return CartesianCurrent.open(p)


def test_current_nearest(oracle, simple_current):
    d = simple_current
    assert round(oracle.current_nearest(5499.0, d.x)) == 11.0                   # :669
    for bad in (-1.0, 25_501.0):
        with pytest.raises(oracle.Err):
            oracle.current_nearest(bad, d.y)
    assert abs(oracle.current_nearest(5500.0, d.x) - 11.0) <= F64_EPS           # :676


@pytest.mark.parametrize("pt,corners", CORNER_CASES)
def test_current_four_corners(oracle, simple_current, pt, corners):
    assert oracle.current_four_corners(simple_current, *pt) == corners          # :724-824 (same table as the bathymetry)


@pytest.mark.parametrize("pt", OUT_OF_BOUNDS)
def test_current_four_corners_out_of_bounds(oracle, simple_current, pt):
    with pytest.raises(oracle.Err):
        oracle.current_four_corners(simple_current, *pt)


def test_current_constant_field_sweep(oracle, tmp_path):
    p = tmp_path / "c.nc"
    create_netcdf3_current(p, 100, 50, 1.0, 1.0, lambda x, y: (5.0, 0.0))
    d = CartesianCurrent.open(p)
    for i in range(100):
        for j in range(50):
            assert oracle.current_and_gradient(d, float(i), float(j)) == ((5.0, 0.0), ((0.0, 0.0), (0.0, 0.0)))   # :907-920
    for pt in [(50_001.0, 1000.0), (-50_001.0, -1000.0)]:                       # :924-928
        with pytest.raises(oracle.Err):
            oracle.current_and_gradient(d, *pt)


@pytest.mark.parametrize("axis", ["x", "y"])
def test_current_gradient_sweep_exact(oracle, tmp_path, axis):
    """u = v = x (or y): exact equality with (i, i) and gradients (1, 0)/(0, 1) at every grid point (:935-991)."""
    fn = (lambda x, y: (float(x), float(x))) if axis == "x" else (lambda x, y: (float(y), float(y)))
    p = tmp_path / "cg.nc"
    create_netcdf3_current(p, 100, 100, 1.0, 1.0, fn)
    d = CartesianCurrent.open(p)
    g = (1.0, 0.0) if axis == "x" else (0.0, 1.0)
    for i in range(100):
        for j in range(100):
            val = float(i if axis == "x" else j)
            assert oracle.current_and_gradient(d, float(i), float(j)) == ((val, val), (g, g))


def test_nan_position_on_gridded_current_is_err(oracle, simple_current):
    """No NaN pre-check in CartesianCurrent: a NaN index collapses the cell, det == 0, Err."""
    for pt in [(math.nan, 100.0), (100.0, math.nan)]:
        with pytest.raises(oracle.Err):
            oracle.current_and_gradient(simple_current, *pt)


# ---- API-level (python/tests/test_core.py:40-78) on the oracle ------------------------------------------------
def test_core_cases_on_oracle(oracle):
    x = np.array([-1e4, 0.0, 1e4])
    bathy = CartesianNetcdf3(x, x, 10_000.0 * np.ones((3, 3)))
    cx = np.array([-1e8, 0.0, 1e8])
    cur = CartesianCurrent(cx, cx, 0.01 * np.ones((3, 3)), 0.01 * np.ones((3, 3)))
    out = oracle.single_ray(bathy, cur, -1000, 0, 0.01, 0, 0.0, 10.0, 2.0)
    assert out.shape[0] == 6 and (out[:, 3] == 0.01).all() and (out[:, 4] == 0.0).all()
    r = oracle.trace_many(bathy, cur, 3 * [-1000], 3 * [0], 3 * [0.01], 3 * [0], 0.0, 10.0, 2.0)
    assert r.x.shape == (6, 3) and (r.kx == 0.01).all() and (r.ky == 0.0).all()
    # variable length: the ray heading to -x leaves the 20 km domain first (test_core.py:81-101)
    r = oracle.trace_many(bathy, cur, 2 * [-1e3], 2 * [0], [-0.01, 0.01], 2 * [0], 0.0, 1e6, 20.0)
    assert r.rows[0] < r.rows[1] < 50_001


# ---- depth() / current(): the environment columns of Ray{time,state,depth,current} (datatype.rs:165-194) --
def test_sample_fields_matches_the_accessors(oracle):
    O = oracle
    """orc_sample_fields is orc_depth / orc_current point by point, Err -> NaN; current() returns the (u, v)
    of current_and_gradient (cartesian_current.rs:448-467 against :487-542)"""
    rng = np.random.default_rng(3)
    x = (500.0 * np.arange(101)).astype(np.float32)
    y = (500.0 * np.arange(51)).astype(np.float32)
    X, Y = np.meshgrid(x.astype(np.float64), y.astype(np.float64))
    bathy = CartesianNetcdf3(x, y, 0.05 * X + 0.01 * Y)
    cur = CartesianCurrent(x.astype(np.float64), y.astype(np.float64), 1e-4 * X, -2e-4 * Y)
    px = rng.uniform(-5_000.0, 55_000.0, 400)
    py = rng.uniform(-5_000.0, 30_000.0, 400)
    px[:4] = [np.nan, 0.0, 50_000.0, 25_000.0]
    py[:4] = [10.0, 0.0, 25_000.0, np.nan]
    depth, u, v = O.sample_fields(bathy, cur, px, py)
    assert depth.dtype == np.float32
    n_err = 0
    for i in range(px.size):
        try:
            h = O.depth(bathy, px[i], py[i])
        except O.Err:
            h = np.nan
        assert (np.isnan(h) and np.isnan(depth[i])) or np.float32(h) == depth[i]
        try:
            uu, vv = O.current(cur, px[i], py[i])
            (u2, v2), _ = O.current_and_gradient(cur, px[i], py[i])
            assert (uu, vv) == (u2, v2)
        except O.Err:
            uu = vv = np.nan
            n_err += 1
        assert (np.isnan(uu) and np.isnan(u[i])) or (uu == u[i] and vv == v[i])
    assert 0 < n_err < px.size
    # constant kinds: depth is NaN for a NaN point (constant_depth.rs:26-32), the current ignores it (:51-53)
    depth, u, v = O.sample_fields(ConstantDepth(12.0), ConstantCurrent(0.3, -0.1), [1.0, np.nan], [2.0, 0.0])
    assert depth[0] == 12.0 and np.isnan(depth[1]) and (u == 0.3).all() and (v == -0.1).all()
    # quadrant depths of cartesian_netcdf3.rs:663-689 through the batch entry point
    d2, _, _ = O.sample_fields(bathy, ConstantCurrent(0, 0), [10_000.0, 30_000.0], [5_000.0, 20_000.0])
    np.testing.assert_allclose(d2, [0.05 * 10_000 + 0.01 * 5_000, 0.05 * 30_000 + 0.01 * 20_000], rtol=1e-6)
