"""The reference's behavioural ray tests, re-expressed against this repo's back-ends.

Ports of
  * the 22 ``test_single_wave`` cases and ``test_many_waves_ok``  (src/ray.rs:268-1361),
  * the 12-ray fans over constant depth                           (src/tests/test_constant_depth.rs:38-162),
  * the four linear beaches                                       (src/tests/linear_beach.rs:53-416),
with the same inputs and the same assertions (exact constancy ``assert_eq!``, monotonicity,
last-vs-first).  They run on the CPU oracle everywhere and on the CUDA path (both math
modes) on the GPU box; the exact-constancy assertions are the interesting ones for the
restructured f64 stage.
"""

import math

import numpy as np
import pytest

from backends import KX, KY, X, Y, backend_params, decrease, increase, same
from mantaray_b200 import (CartesianCurrent, CartesianNetcdf3, ConstantCurrent, ConstantDepth, ConstantSlope)
from mantaray_b200.io_utility import create_netcdf3_bathymetry, create_netcdf3_current

BACKENDS = backend_params()
f32 = np.float32


# ---- field fixtures (files written like src/io/utility.rs and read back through the library) ----
def two_depth_fn(x, _y):            # src/ray.rs:260-266
    return 20.0 if x >= 50.0 else 50.0


@pytest.fixture(scope="module")
def two_depth(tmp_path_factory):
    p = tmp_path_factory.mktemp("nc") / "two_depth.nc"
    create_netcdf3_bathymetry(p, 100, 100, 1.0, 1.0, two_depth_fn)
    return CartesianNetcdf3.open(p)


GRADIENT_FNS = {
    "dudx": lambda x, y: (float(f32(x) / f32(100.0)), 0.0),                       # src/ray.rs:882-884
    "dudy": lambda x, y: (float(f32(y) / f32(100.0)), 0.0),                       # :971-973
    "dvdy": lambda x, y: (0.0, float(f32(y) / f32(100.0))),                       # :1066-1068
    "dvdx": lambda x, y: (0.0, float(f32(x) / f32(100.0))),                       # :1166-1168
    "all": lambda x, y: (float((f32(x) + f32(y)) / f32(100.0)),) * 2,             # :1259-1261
}


@pytest.fixture(scope="module")
def gradient_currents(tmp_path_factory):
    d = tmp_path_factory.mktemp("nc")
    out = {}
    for name, fn in GRADIENT_FNS.items():
        p = d / f"{name}.nc"
        create_netcdf3_current(p, 100, 100, 1.0, 1.0, fn)
        out[name] = CartesianCurrent.open(p)
    return out


# ---- assertion vocabulary ------------------------------------------------------------------------
def check(data, spec):
    """spec items: ('eq', col, value) every row == value;  ('ge', col) non-decreasing;
    ('le', col) non-increasing;  ('gt_end', col) last > first;  ('lt_end', col) last < first;
    ('x_ge_y',) x >= y on every row.  `nonan` variants skip rows whose x is NaN."""
    for item in spec:
        kind = item[0]
        rows = data
        if kind.endswith("_nonan"):
            kind = kind[: -len("_nonan")]
            rows = data[~np.isnan(data[:, 0])]
        if kind == "eq":
            assert np.all(rows[:, item[1]] == item[2]), f"col {item[1]} not constantly {item[2]}: {rows[:, item[1]]}"
        elif kind == "ge":
            assert np.all(np.diff(rows[:, item[1]]) >= 0), f"col {item[1]} decreases"
        elif kind == "le":
            assert np.all(np.diff(rows[:, item[1]]) <= 0), f"col {item[1]} increases"
        elif kind == "gt_end":
            assert rows[-1, item[1]] > rows[0, item[1]]
        elif kind == "lt_end":
            assert rows[-1, item[1]] < rows[0, item[1]]
        elif kind == "x_ge_y":
            assert np.all(rows[:, X] >= rows[:, Y])
        else:
            raise AssertionError(kind)


# (name, bathymetry, current, ray, (t0, t1, dt), assertions) — src/ray.rs:268-870
ANALYTIC_CASES = [
    ("constant_wave_shallow_x", ConstantDepth(10.0), ConstantCurrent(0, 0), (10.0, 50.0, 0.01, 0.0), (0.0, 8.0, 1.0),
     [("eq", Y, 50.0), ("eq", KX, 0.01), ("eq", KY, 0.0), ("ge", X)]),
    ("constant_wave_shallow_xy", ConstantDepth(10.0), ConstantCurrent(0, 0), (10.0, 10.0, 0.007, 0.007), (0.0, 8.0, 1.0),
     [("eq", KX, 0.007), ("eq", KY, 0.007), ("ge", X), ("ge", Y)]),
    ("constant_wave_deep_x", ConstantDepth(10.0), ConstantCurrent(0, 0), (10.0, 50.0, 1.0, 0.0), (0.0, 18.0, 1.0),
     [("eq", Y, 50.0), ("eq", KX, 1.0), ("eq", KY, 0.0), ("ge", X)]),
    ("constant_wave_deep_xy", ConstantDepth(10.0), ConstantCurrent(0, 0), (10.0, 10.0, 0.7, 0.7), (0.0, 18.0, 1.0),
     [("eq", KX, 0.7), ("eq", KY, 0.7), ("ge", X), ("ge", Y)]),
    ("slope_depth_wave_x", ConstantSlope(), ConstantCurrent(0, 0), (10.0, 1000.0, 0.1, 0.0), (0.0, 100.0, 1.0),
     [("ge_nonan", X), ("gt_end_nonan", KX)]),
    ("constant_depth_zero_current", ConstantDepth(10.0), ConstantCurrent(0, 0), (0.0, 0.0, 0.1, 0.0), (100.0, 102.0, 1.0),
     [("eq", Y, 0.0), ("eq", KX, 0.1), ("eq", KY, 0.0), ("ge", X)]),
    ("constant_depth_and_current", ConstantDepth(10.0), ConstantCurrent(0.5, 0.0), (0.0, 0.0, 0.1, 0.0), (1.0, 10.0, 1.0),
     [("eq", Y, 0.0), ("eq", KX, 0.1), ("eq", KY, 0.0), ("ge", X)]),
    ("positive_v", ConstantDepth(1000.0), ConstantCurrent(0.0, 0.5), (0.0, 0.0, 0.1, 0.0), (1.0, 10.0, 1.0),
     [("eq", KX, 0.1), ("eq", KY, 0.0), ("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y)]),
    ("negative_v", ConstantDepth(1000.0), ConstantCurrent(0.0, -0.5), (0.0, 0.0, 0.1, 0.0), (1.0, 10.0, 1.0),
     [("eq", KX, 0.1), ("eq", KY, 0.0), ("ge", X), ("le", Y), ("gt_end", X), ("lt_end", Y)]),
    ("positive_u", ConstantDepth(1000.0), ConstantCurrent(0.5, 0.0), (0.0, 0.0, 0.0, 0.1), (1.0, 10.0, 1.0),
     [("eq", KX, 0.0), ("eq", KY, 0.1), ("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y)]),
    ("negative_u", ConstantDepth(1000.0), ConstantCurrent(-0.5, 0.0), (0.0, 0.0, 0.0, 0.1), (1.0, 10.0, 1.0),
     [("eq", KX, 0.0), ("eq", KY, 0.1), ("le", X), ("ge", Y), ("lt_end", X), ("gt_end", Y)]),
    ("positive_u_and_v", ConstantDepth(1000.0), ConstantCurrent(0.5, 0.5), (0.0, 0.0, 0.1, 0.0), (1.0, 10.0, 1.0),
     [("eq", KX, 0.1), ("eq", KY, 0.0), ("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y)]),
    ("negative_u_and_v", ConstantDepth(1000.0), ConstantCurrent(-0.5, -0.5), (0.0, 0.0, -0.1, 0.0), (1.0, 10.0, 1.0),
     [("eq", KX, -0.1), ("eq", KY, 0.0), ("le", X), ("le", Y), ("lt_end", X), ("lt_end", Y)]),
]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case", ANALYTIC_CASES, ids=[c[0] for c in ANALYTIC_CASES])
def test_single_wave_analytic(backend, case):
    _, bathy, cur, ray, (t0, t1, dt), spec = case
    _, data = backend.single(bathy, cur, ray, t0, t1, dt)
    assert data.shape[0] == math.ceil((t1 - t0) / dt) + 1 or np.isnan(data[-1]).all()
    check(data, spec)


def test_slope_depth_values(oracle):
    """src/ray.rs:561-562."""
    assert oracle.depth(ConstantSlope(), 10.0, 1000.0) == 49.5
    assert oracle.depth(ConstantSlope(), 300.0, 1000.0) == 35.0


# src/ray.rs:386-548: the two-depth step file
TWO_DEPTH_CASES = [
    ("two_depth_wave_shallow_x", (10.0, 50.0, 0.01, 0.0), (0.0, 5.0, 1.0),
     [("eq", Y, 50.0), ("eq", KY, 0.0), ("ge", X), ("ge", KX), ("gt_end", KX)]),
    ("two_depth_wave_shallow_xy", (10.0, 10.0, 0.007, 0.007), (0.0, 6.8, 0.1),
     [("eq", KY, 0.007), ("ge", X), ("ge", Y), ("ge", KX), ("x_ge_y",), ("gt_end", KX)]),
    ("two_depth_wave_deep_x", (10.0, 50.0, 1.0, 0.0), (0.0, 30.0, 1.0),
     [("eq", KY, 0.0), ("eq", Y, 50.0), ("eq", KX, 1.0), ("ge", X)]),
    ("two_depth_wave_deep_xy", (10.0, 10.0, 0.7, 0.7), (0.0, 40.0, 1.0),
     [("eq", KX, 0.7), ("eq", KY, 0.7), ("ge", X), ("ge", Y)]),
]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case", TWO_DEPTH_CASES, ids=[c[0] for c in TWO_DEPTH_CASES])
def test_single_wave_two_depth(backend, two_depth, case):
    _, ray, (t0, t1, dt), spec = case
    _, data = backend.single(two_depth, ConstantCurrent(0, 0), ray, t0, t1, dt)
    check(data, spec)


# src/ray.rs:872-1316: one current gradient at a time, depth 1000 m, t in [1, 10], dt = 1
GRADIENT_CASES = [
    ("dudx_kx", "dudx", (1.0, 1.0, 0.1, 0.0), [("eq", KY, 0.0), ("eq", Y, 1.0), ("ge", X), ("le", KX), ("lt_end", KX), ("gt_end", X)]),
    ("dudx_ky", "dudx", (1.0, 1.0, 0.0, 0.1), [("eq", KY, 0.1), ("eq", KX, 0.0), ("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y)]),
    ("dudy_kx", "dudy", (1.0, 50.0, 0.1, 0.0), [("eq", KX, 0.1), ("ge", X), ("gt_end", X), ("le", Y), ("le", KY), ("lt_end", Y), ("lt_end", KY)]),
    ("dudy_ky", "dudy", (1.0, 1.0, 0.0, 0.1), [("eq", KX, 0.0), ("eq", KY, 0.1), ("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y)]),
    ("dvdy_kx", "dvdy", (1.0, 1.0, 0.1, 0.0), [("eq", KX, 0.1), ("eq", KY, 0.0), ("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y)]),
    ("dvdy_ky", "dvdy", (1.0, 1.0, 0.0, 0.1), [("eq", KX, 0.0), ("ge", Y), ("gt_end", Y), ("le", KY), ("lt_end", KY)]),
    ("dvdx_kx", "dvdx", (1.0, 1.0, 0.1, 0.0), [("eq", KX, 0.1), ("eq", KY, 0.0), ("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y)]),
    ("dvdx_ky", "dvdx", (50.0, 1.0, 0.0, 0.1), [("eq", KY, 0.1), ("ge", Y), ("gt_end", Y), ("le", X), ("le", KX), ("lt_end", X), ("lt_end", KX)]),
    ("all_gradients", "all", (1.0, 1.0, 0.1, 0.0), [("ge", X), ("ge", Y), ("gt_end", X), ("gt_end", Y), ("le", KX), ("le", KY), ("lt_end", KX), ("lt_end", KY)]),
]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case", GRADIENT_CASES, ids=[c[0] for c in GRADIENT_CASES])
def test_single_wave_current_gradient(backend, gradient_currents, case):
    _, which, ray, spec = case
    _, data = backend.single(ConstantDepth(1000.0), gradient_currents[which], ray, 1.0, 10.0, 1.0)
    assert data.shape[0] == 10
    check(data, spec)


@pytest.mark.parametrize("backend", BACKENDS)
def test_many_waves_ok(backend):
    """src/ray.rs:1333-1361: 9 rays up a default ConstantSlope; all come back."""
    rays = [(10.0, 10.0 * (i + 1), 1.0, 0.0) for i in range(9)]
    res = backend.many(ConstantSlope(), ConstantCurrent(0, 0), rays, 0.0, 100000.0, 1.0)
    assert len(res) == 9
    for t, data in res:
        assert t.shape[0] == data.shape[0] > 100
        assert np.isnan(data[-1]).all()          # each ray runs ashore (h <= 0) and stops on a NaN row


# ---- src/tests/test_constant_depth.rs ------------------------------------------------------------------
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("h", [2000.0, 10.0], ids=["deep", "shallow"])
def test_constant_depth_fans(backend, h):
    """12 rays, 30 degrees apart, from (50 km, 25 km); (kx, ky) stay exactly the initial values and each
    coordinate moves strictly the way its wavenumber component points (:38-162)."""
    rays = [(50_000.0, 25_000.0, 0.05 * math.cos(math.pi * i / 6.0), 0.05 * math.sin(math.pi * i / 6.0)) for i in range(12)]
    res = backend.many(ConstantDepth(h), ConstantCurrent(0, 0), rays, 0.0, 5000.0, 1.0)
    eps = np.finfo(np.float64).eps
    for (x0, y0, kx, ky), (_, data) in zip(rays, res):
        assert data.shape[0] == 5001
        for k, col in ((kx, X), (ky, Y)):
            if abs(k) < eps:
                assert same(data, col)
            elif k > 0:
                assert increase(data, col)
            else:
                assert decrease(data, col)
        assert same(data, KX) and same(data, KY)


# ---- src/tests/linear_beach.rs ------------------------------------------------------------------------------
K = 0.05
PI = math.pi
BEACHES = [
    # name, slope args (h0, x0, y0, dhdx, dhdy), rays, per-ray expectations (x, y, kx, ky)
    ("right", (100.0, 0.0, 0.0, -0.05, 0.0),
     [(0.0, 0.0, K * math.cos(PI / 6), K * math.sin(PI / 6)), (0.0, 0.0, K * math.cos(-PI / 6), K * math.sin(-PI / 6)), (0.0, 0.0, K, 0.0)],
     [("inc", "inc", "inc", "same"), ("inc", "dec", "inc", "same"), ("inc", "same", "inc", "same")]),
    ("left", (0.0, 0.0, 0.0, 0.05, 0.0),
     [(2000.0, 0.0, -K * math.cos(PI / 6), K * math.sin(PI / 6)), (2000.0, 0.0, -K * math.cos(-PI / 6), K * math.sin(-PI / 6)), (2000.0, 100.0, -K, 0.0)],
     [("dec", "inc", "dec", "same"), ("dec", "dec", "dec", "same"), ("dec", "same", "dec", "same")]),
    ("top", (100.0, 0.0, 0.0, 0.0, -0.05),
     [(0.0, 0.0, K * math.cos(4 * PI / 6), K * math.sin(4 * PI / 6)), (0.0, 0.0, K * math.cos(2 * PI / 6), K * math.sin(2 * PI / 6)), (100.0, 0.0, 0.0, K)],
     [("dec", "inc", "same", "inc"), ("inc", "inc", "same", "inc"), ("same", "inc", "same", "inc")]),
    ("bottom", (0.0, 0.0, 0.0, 0.0, 0.05),
     [(0.0, 2000.0, K * math.cos(4 * PI / 6), -K * math.sin(4 * PI / 6)), (0.0, 2000.0, K * math.cos(2 * PI / 6), -K * math.sin(2 * PI / 6)), (100.0, 2000.0, 0.0, -K)],
     [("dec", "dec", "same", "dec"), ("inc", "dec", "same", "dec"), ("same", "dec", "same", "dec")]),
]
_REL = {"inc": increase, "dec": decrease, "same": same}


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("beach", BEACHES, ids=[b[0] for b in BEACHES])
def test_linear_beach(backend, beach):
    _, (h0, x0, y0, dhdx, dhdy), rays, expect = beach
    res = backend.many(ConstantSlope(h0, x0, y0, dhdx, dhdy), ConstantCurrent(0, 0), rays, 0.0, 1000.0, 1.0)
    assert len(res) == 3
    for (_, data), exp in zip(res, expect):
        for col, rel in zip((X, Y, KX, KY), exp):
            assert _REL[rel](data, col), f"column {col} should be {rel}"
