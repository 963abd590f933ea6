"""The environment along rays: depth and current at every stored state — the columns of
the reference's unfilled ``Ray{time, state, depth, current}`` record (src/datatype.rs:165-194) — from
``mr_sample_fields`` / ``mr_trace_many_env`` / ``mr_sample_device``, bit-exact against the oracle's
``depth()`` (src/bathymetry/mod.rs:38) and ``current()`` (src/current/mod.rs:24)."""

import ctypes as C

import numpy as np
import pytest

import mantaray
from mantaray_b200 import (ArrayDepth, CartesianCurrent, CartesianNetcdf3, ConstantCurrent, ConstantDepth,
                           ConstantSlope, Fields, MantarayError, _abi, _capi, trace_many)
from mantaray_b200 import workloads as W
from mantaray_b200.io_utility import write_netcdf3

pytestmark = pytest.mark.gpu


def same_bits(a, b, what):
    """equal values, NaN == NaN (the payload of a NaN is not part of the contract)"""
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, what
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN pattern differs at {np.argwhere(nan_a != nan_b)[:5]}"
    ok = a[~nan_a] == b[~nan_b]
    assert ok.all(), f"{what}: {np.count_nonzero(~ok)} values differ"


def random_fields(rng, affine):
    """affine=False: f32 coordinates that are not i*step (the lookups then follow the reference operation by
    operation); affine=True: the usual i*step grids (the lookups then go through the cell records)"""
    nx, ny = 97, 61
    if affine:
        x = (50.0 * np.arange(nx)).astype(np.float32)
        y = (-1000.0 + 25.0 * np.arange(ny)).astype(np.float32)
        cx = -100.0 + 62.5 * np.arange(90)
        cy = -1100.0 + 31.25 * np.arange(64)
    else:
        x = (-1234.5 + 37.3 * np.arange(nx)).astype(np.float32)
        y = (987.25 + 41.7 * np.arange(ny)).astype(np.float32)
        cx = -1300.0 + 41.0 * np.arange(90)
        cy = 900.0 + 43.0 * np.arange(64)
    X, Y = np.meshgrid(x.astype(np.float64), y.astype(np.float64))
    depth = 30.0 + 25.0 * np.sin(X / 700.0) * np.cos(Y / 500.0) + rng.normal(0, 0.5, X.shape)
    CX, CY = np.meshgrid(cx, cy)
    u = 0.8 * np.sin(CY / 600.0) + rng.normal(0, 0.01, CX.shape)
    v = 0.5 * np.cos(CX / 800.0) + rng.normal(0, 0.01, CX.shape)
    return CartesianNetcdf3(x, y, depth), CartesianCurrent(cx, cy, u, v)


def probe_points(rng, bathy, cur, n=20_000):
    """inside, outside, on grid lines and nodes, on the last node, NaN and inf"""
    bx, by = np.asarray(bathy.x, dtype=np.float64), np.asarray(bathy.y, dtype=np.float64)
    cx, cy = np.asarray(cur.x), np.asarray(cur.y)
    lo_x, hi_x = min(bx[0], cx[0]), max(bx[-1], cx[-1])
    lo_y, hi_y = min(by[0], cy[0]), max(by[-1], cy[-1])
    x = rng.uniform(lo_x - 0.1 * (hi_x - lo_x), hi_x + 0.1 * (hi_x - lo_x), n)
    y = rng.uniform(lo_y - 0.1 * (hi_y - lo_y), hi_y + 0.1 * (hi_y - lo_y), n)
    k = n // 10
    x[:k] = rng.choice(bx, k); y[k // 2:k] = rng.choice(by, k - k // 2)            # bathymetry lines / nodes
    x[k:2 * k] = rng.choice(cx, k); y[k + k // 2:2 * k] = rng.choice(cy, k - k // 2)   # current lines / nodes
    x[2 * k:2 * k + 8] = [bx[0], bx[-1], bx[0], bx[-1], cx[0], cx[-1], np.nan, np.inf]
    y[2 * k:2 * k + 8] = [by[0], by[-1], by[-1], by[0], cy[-1], cy[0], by[3], by[3]]
    x[2 * k + 8:2 * k + 10] = [bx[2], bx[2]]
    y[2 * k + 8:2 * k + 10] = [np.nan, -np.inf]
    # just inside / outside the last node, where the f64 index of the current is an ulp from n-1
    x[2 * k + 10:2 * k + 14] = np.nextafter(cx[-1], [np.inf, -np.inf, np.inf, -np.inf])
    y[2 * k + 10:2 * k + 14] = cy[5]
    return x, y


@pytest.mark.parametrize("affine", [False, True], ids=["generic", "affine"])
def test_sample_fields_gridded(oracle, gpu, affine):
    rng = np.random.default_rng(99)
    bathy, cur = random_fields(rng, affine)
    x, y = probe_points(rng, bathy, cur)
    ref = oracle.sample_fields(bathy, cur, x, y)
    assert np.isnan(ref[0]).any() and np.isfinite(ref[0]).any() and np.isnan(ref[1]).any() and np.isfinite(ref[1]).any()
    with Fields(bathy, cur, devices=[0]) as f:
        got = _capi.sample_fields(f, x, y)
    for g, r, name in zip(got, ref, ("depth", "u", "v")):
        same_bits(g, r, name)


@pytest.mark.parametrize("bathy", [
    ConstantDepth(10.0), ConstantSlope(50.0, 10.0, -5.0, 0.02, -0.03), ArrayDepth(np.arange(1600.0).reshape(40, 40)),
], ids=["constant", "slope", "array"])
@pytest.mark.parametrize("cur", [ConstantCurrent(0.5, -0.25)], ids=["uv"])
def test_sample_fields_analytic(oracle, gpu, bathy, cur):
    rng = np.random.default_rng(5)
    x = rng.uniform(-10.0, 60.0, 5000)
    y = rng.uniform(-10.0, 60.0, 5000)
    x[:3] = [np.nan, 1.0, np.inf]
    y[:3] = [1.0, np.nan, 2.0]
    ref = oracle.sample_fields(bathy, cur, x, y)
    with Fields(bathy, cur, devices=[0]) as f:
        got = _capi.sample_fields(f, x, y)
    for g, r, name in zip(got, ref, ("depth", "u", "v")):
        same_bits(g, r, name)
    # the constant current ignores the point, NaN or not (constant_current.rs:51-53)
    assert (got[1] == 0.5).all() and (got[2] == -0.25).all()


def test_sample_fields_shapes_and_empty(gpu):
    with Fields(ConstantDepth(7.0), ConstantCurrent(0.0, 0.0), devices=[0]) as f:
        d, u, v = _capi.sample_fields(f, np.zeros((3, 5)), np.zeros((3, 5)))
        assert d.shape == u.shape == v.shape == (3, 5) and d.dtype == np.float32 and (d == 7.0).all()
        d, u, v = _capi.sample_fields(f, np.zeros(0), np.zeros(0))
        assert d.size == 0
        with pytest.raises(ValueError):
            _capi.sample_fields(f, np.zeros(3), np.zeros(4))


@pytest.mark.parametrize("stride", [1, 7])
@pytest.mark.parametrize("make", [lambda: W.c2_sea_mount(600, 300, half=150), lambda: W.c4_agulhas(30, 20, 200, nx=256)],
                         ids=["c2", "c4"])
def test_trace_many_env(oracle, gpu, make, stride):
    """the env planes are the oracle's depth()/current() at the states the same call stored; the state
    planes are those of a call without env; rows past a ray's end hold what the accessors return for NaN"""
    wl = make()
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        res = trace_many(f, *rays, wl.t0, wl.duration, wl.dt, stride=stride, env=True)
        plain = trace_many(f, *rays, wl.t0, wl.duration, wl.dt, stride=stride)
    for name in ("x", "y", "kx", "ky"):
        same_bits(getattr(res, name), getattr(plain, name), name)
    assert np.array_equal(res.rows, plain.rows) and np.array_equal(res.len, plain.len)
    ref = oracle.sample_fields(wl.bathymetry, wl.current, res.x, res.y)
    assert res.depth.dtype == np.float32 and res.depth.shape == res.x.shape
    for g, r, name in zip((res.depth, res.u, res.v), ref, ("depth", "u", "v")):
        same_bits(g, r, name)
    finite = np.isfinite(res.x)
    assert np.isfinite(res.u[finite]).mean() > 0.9       # in-domain states have a current
    assert np.isnan(res.depth[~finite]).all() and np.isnan(res.u[~finite]).all()


def test_trace_many_env_chunked_and_partial_planes(gpu):
    """slabs (two device buffers, ragged last slab) give the same planes; any subset of planes may be asked for"""
    wl = W.c4_agulhas(25, 21, 64, nx=128)
    rays = wl.all_rays()
    lib = _capi.load()
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        whole = trace_many(f, *rays, wl.t0, wl.duration, wl.dt, env=True)
        slabs = trace_many(f, *rays, wl.t0, wl.duration, wl.dt, env=True, chunk_rays=128)
        for name in ("x", "depth", "u", "v"):
            same_bits(getattr(slabs, name), getattr(whole, name), name)
        # only v
        n, rows = wl.n_rays, wl.n_rows
        x0, y0, kx0, ky0 = (np.ascontiguousarray(a, dtype=np.float64) for a in rays)
        planes = [np.empty((rows, n)) for _ in range(4)]
        v = np.empty((rows, n))
        env = _abi.EnvPlanes(None, None, v.ctypes.data)
        rc = lib.mr_trace_many_env(f.handle, n, x0.ctypes.data, y0.ctypes.data, kx0.ctypes.data, ky0.ctypes.data,
                                   wl.t0, wl.duration, wl.dt, None, None, *(p.ctypes.data for p in planes),
                                   None, None, None, C.byref(env))
        assert rc == 0, lib.mr_last_error()
        same_bits(v, whole.v, "v alone")
        # env without the state planes is refused
        rc = lib.mr_trace_many_env(f.handle, n, x0.ctypes.data, y0.ctypes.data, kx0.ctypes.data, ky0.ctypes.data,
                                   wl.t0, wl.duration, wl.dt, None, None, None, None, None, None,
                                   None, None, None, C.byref(env))
        assert rc == _abi.MR_ERR_BAD_ARG
        with pytest.raises(ValueError):
            trace_many(f, *rays, wl.t0, wl.duration, wl.dt, env=True, trajectories=False)


def test_sample_device(oracle, gpu):
    """device-resident: mr_trace_device then mr_sample_device on its planes (pitch > n), same stream"""
    import torch

    wl = W.c4_agulhas(20, 20, 50, nx=128)
    n, rows, ld = wl.n_rays, wl.n_rows, wl.n_rays + 24
    lib = _capi.load()
    dev = torch.device("cuda:0")
    ic = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in wl.all_rays()]
    traj = torch.full((4, rows, ld), -1.0, dtype=torch.float64, device=dev)
    depth = torch.full((rows, ld), -1.0, dtype=torch.float32, device=dev)
    u = torch.full((rows, ld), -1.0, dtype=torch.float64, device=dev)
    v = torch.full((rows, ld), -1.0, dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    launches = C.c_int32(0)
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        rc = lib.mr_trace_device(f.handle, 0, stream, n, p(ic[0]), p(ic[1]), p(ic[2]), p(ic[3]),
                                 wl.t0, wl.duration, wl.dt, None, p(traj[0]), p(traj[1]), p(traj[2]), p(traj[3]), ld,
                                 None, None, None, None)
        assert rc == 0, lib.mr_last_error()
        rc = lib.mr_sample_device(f.handle, 0, stream, rows, n, ld, p(traj[0]), p(traj[1]), p(depth), p(u), p(v),
                                  C.byref(launches))
        assert rc == 0, lib.mr_last_error()
        torch.cuda.synchronize()
        assert launches.value == 1
        assert lib.mr_sample_device(f.handle, 0, stream, rows, ld + 1, ld, p(traj[0]), p(traj[1]), p(depth), None, None, None) == _abi.MR_ERR_BAD_ARG
        assert lib.mr_sample_device(f.handle, 1 if _capi.device_count() == 1 else 31, stream, rows, n, ld, p(traj[0]), p(traj[1]),
                                    p(depth), None, None, None) == _abi.MR_ERR_BAD_ARG
    x, y = traj[0, :, :n].cpu().numpy(), traj[1, :, :n].cpu().numpy()
    ref = oracle.sample_fields(wl.bathymetry, wl.current, x, y)
    for g, r, name in zip((depth, u, v), ref, ("depth", "u", "v")):
        g = g.cpu().numpy()
        same_bits(np.ascontiguousarray(g[:, :n]), r, name)
        assert (g[:, n:] == -1.0).all(), "the pitch padding is not written"


def test_ray_tracing_diagnostics(oracle, gpu, tmp_path):
    """mantaray.ray_tracing(..., diagnostics=True): the reference's five variables unchanged, plus depth, u, v
    and the k, theta, sigma the notebooks derive; off by default"""
    x = 50.0 * np.arange(101)
    y = 50.0 * np.arange(51)
    X, Y = np.meshgrid(x, y)
    b = CartesianNetcdf3(x, y, 60.0 - 0.01 * X + 0.002 * Y)        # a beach shoaling from 60 m to 10 m
    c = CartesianCurrent(x, y, np.zeros_like(X), np.zeros_like(X))
    write_netcdf3(tmp_path / "b.nc", [("y", len(b.y)), ("x", len(b.x))],
                  {"x": (["x"], np.asarray(b.x, np.float64)), "y": (["y"], np.asarray(b.y, np.float64)),
                   "depth": (["y", "x"], np.asarray(b.depth).reshape(len(b.y), len(b.x)))})
    write_netcdf3(tmp_path / "c.nc", [("y", len(c.y)), ("x", len(c.x))],
                  {"x": (["x"], np.asarray(c.x)), "y": (["y"], np.asarray(c.y)),
                   "u": (["y", "x"], np.asarray(c.u).reshape(len(c.y), len(c.x))),
                   "v": (["y", "x"], np.asarray(c.v).reshape(len(c.y), len(c.x)))})
    n = 48
    k0 = W.period2wavenumber(10.0)
    th = np.linspace(-0.6, 0.6, n)
    rays = (np.full(n, 100.0), np.linspace(600.0, 1900.0, n), k0 * np.cos(th), k0 * np.sin(th))
    plain = mantaray.ray_tracing(*rays, 700.0, 5.0, tmp_path / "b.nc", tmp_path / "c.nc")
    ds = mantaray.ray_tracing(*rays, 700.0, 5.0, tmp_path / "b.nc", tmp_path / "c.nc", diagnostics=True)
    for name in ("time", "x", "y", "kx", "ky"):
        same_bits(np.asarray(ds[name]), np.asarray(plain[name]), name)
        assert "depth" not in plain
    x, y = np.asarray(ds["x"]), np.asarray(ds["y"])
    ref = oracle.sample_fields(b, c, x, y)
    for r, name in zip(ref, ("depth", "u", "v")):
        same_bits(np.asarray(ds[name]), r, name)
    k = np.asarray(ds["k"])
    ok = np.isfinite(x)
    np.testing.assert_allclose(k[ok], np.hypot(np.asarray(ds["kx"]), np.asarray(ds["ky"]))[ok])
    # shoaling: the absolute frequency sigma + k.U is conserved along a ray (U = 0 here) to the integrator's
    # accuracy, while k grows as the water shoals
    sigma = np.asarray(ds["sigma"])
    inside = np.isfinite(np.asarray(ds["depth"]))
    assert inside.sum() > 0.5 * inside.size and (~inside).any()
    first = sigma[0][None, :] * np.ones_like(sigma)
    assert np.max(np.abs(sigma[inside] / first[inside] - 1.0)) < 1e-3
    assert (k[inside].reshape(-1) >= k0 * (1 - 1e-9)).all() and k[inside].max() > 1.2 * k0
