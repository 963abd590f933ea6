"""The reference-facing surface on the GPU: `mantaray.single_ray` / `mantaray.ray_tracing` through
NetCDF files (ports of python/tests/test_core.py), the file-opening entry point, the exact f32
division self-test, pinned buffers and the multi-device handle."""

import ctypes as C

import numpy as np
import pytest

import mantaray
import mantaray_b200
from conftest import assert_parity
from mantaray_b200 import CartesianCurrent, CartesianNetcdf3, Fields, _abi, _capi, trace_many
from mantaray_b200 import workloads as W
from mantaray_b200.io_utility import write_netcdf3
from tools import mrtools

pytestmark = pytest.mark.gpu


# ---- python/tests/test_core.py:9-37 fixtures: 3x3 grids, data declared (x, y) -------------------------
@pytest.fixture
def island_and_current(tmp_path):
    x = np.array([-1e4, 0.0, 1e4])
    write_netcdf3(tmp_path / "island.nc", [("x", 3), ("y", 3)],
                  {"x": (["x"], x), "y": (["y"], x), "depth": (["x", "y"], 10_000.0 * np.ones((3, 3)))})
    cx = np.array([-1e8, 0.0, 1e8])
    write_netcdf3(tmp_path / "current.nc", [("x", 3), ("y", 3)],
                  {"x": (["x"], cx), "y": (["y"], cx), "u": (["x", "y"], 0.01 * np.ones((3, 3))),
                   "v": (["x", "y"], 0.01 * np.ones((3, 3)))})
    return tmp_path / "island.nc", tmp_path / "current.nc"


def test_single_ray(gpu, island_and_current):
    """test_core.py:40-53"""
    island, current = island_and_current
    ds = mantaray.single_ray(-1000, 0, 0.01, 0, 10, 2, island, current)
    assert ds.sizes["time_step"] == 6
    assert (np.asarray(ds.kx) == 0.01).all()
    assert (np.asarray(ds.ky) == 0.0).all()
    np.testing.assert_array_equal(np.asarray(ds.time), [0.0, 2.0, 4.0, 6.0, 8.0, 10.0])


def test_multiple_rays(gpu, island_and_current):
    """test_core.py:56-78"""
    island, current = island_and_current
    ds = mantaray.ray_tracing(3 * [-1000], 3 * [0], 3 * [0.01], 3 * [0], 10, 2, str(island), str(current))
    assert ds.sizes["time_step"] == 6
    assert ds.sizes["ray"] == 3
    assert (np.asarray(ds.kx) == 0.01).all()
    assert (np.asarray(ds.ky) == 0.0).all()


def test_rays_variable_length(gpu, oracle, island_and_current):
    """test_core.py:81-101 (which only runs it); here also checked against the oracle: the Dataset is
    padded to the longest ray and the shorter ray's tail is NaN, time included."""
    island, current = island_and_current
    ds = mantaray.ray_tracing(2 * [-1e3], 2 * [0], [-0.01, 0.01], 2 * [0], 1e6, 20, str(island), str(current))
    ref = oracle.trace_many(CartesianNetcdf3.open(island), CartesianCurrent.open(current),
                            2 * [-1e3], 2 * [0], [-0.01, 0.01], 2 * [0], 0.0, 1e6, 20.0)
    assert ds.sizes["ray"] == 2 and ds.sizes["time_step"] == int(ref.rows.max()) < 50_001
    short = int(ref.rows.min())
    x = np.asarray(ds.x)
    t = np.asarray(ds.time)
    assert np.isnan(x[short:, 0]).all() and np.isnan(t[short:, 0]).all() and not np.isnan(t[:short, 0]).any()
    L = ds.sizes["time_step"]
    np.testing.assert_allclose(x[: short - 1, 0], ref.x[: short - 1, 0], rtol=1e-9)
    np.testing.assert_allclose(x[: L - 1, 1], ref.x[: L - 1, 1], rtol=1e-9)


def test_file_entry_point_matches_descriptor_entry_point(gpu, tmp_path):
    wl = W.c2_sea_mount(300, 200, half=50)
    b, c = wl.bathymetry, wl.current
    write_netcdf3(tmp_path / "b.nc", [("y", b.y.size), ("x", b.x.size)],
                  {"x": (["x"], b.x), "y": (["y"], b.y), "depth": (["y", "x"], b.depth.reshape(b.y.size, b.x.size))})
    write_netcdf3(tmp_path / "c.nc", [("y", c.y.size), ("x", c.x.size)],
                  {"x": (["x"], c.x), "y": (["y"], c.y), "u": (["y", "x"], c.u.reshape(c.y.size, c.x.size)),
                   "v": (["y", "x"], c.v.reshape(c.y.size, c.x.size))})
    rays = wl.all_rays()
    with Fields(b, c) as f1, Fields.open_netcdf3(tmp_path / "b.nc", tmp_path / "c.nc") as f2:
        r1 = trace_many(f1, *rays, 0.0, wl.duration, wl.dt)
        r2 = trace_many(f2, *rays, 0.0, wl.duration, wl.dt)
    for name in ("x", "y", "kx", "ky", "rows", "len"):
        np.testing.assert_array_equal(getattr(r1, name), getattr(r2, name))
    # defaults when a path is NULL: ConstantDepth 2000 m / ConstantCurrent (0, 0)
    with Fields.open_netcdf3(None, None) as f3, Fields(mantaray_b200.ConstantDepth(2000.0), mantaray_b200.ConstantCurrent(0, 0)) as f4:
        a = trace_many(f3, [0.0], [0.0], [0.05], [0.01], 0.0, 50.0, 1.0)
        bb = trace_many(f4, [0.0], [0.0], [0.05], [0.01], 0.0, 50.0, 1.0)
    np.testing.assert_array_equal(a.x, bb.x)


def test_repeated_api_calls_reuse_the_field_handle(gpu, island_and_current):
    """the second call on the same files skips parse and upload (the reference re-opens both files on
    every call, src/ffi.rs:62-64); a rewritten file is opened again."""
    island, current = island_and_current
    mantaray_b200.clear_cache()
    base = mantaray_b200.cache_info()
    a = mantaray.ray_tracing(3 * [-1000], 3 * [0], 3 * [0.01], 3 * [0], 10, 2, str(island), str(current))
    b = mantaray.ray_tracing(3 * [-1000], 3 * [0], 3 * [0.02], 3 * [0], 10, 2, str(island), str(current))
    c = mantaray.single_ray(-1000, 0, 0.01, 0, 10, 2, island, current)
    info = mantaray_b200.cache_info()
    assert info["misses"] - base["misses"] <= 2 and info["hits"] - base["hits"] >= 1      # ray_tracing twice: one open
    np.testing.assert_array_equal(np.asarray(a.x)[:, 0], np.asarray(c.x))
    assert (np.asarray(b.kx) == 0.02).all()
    # rewrite the bathymetry (shallower): same path, new content -> new handle, new result
    x = np.array([-1e4, 0.0, 1e4])
    import time
    time.sleep(0.02)
    write_netcdf3(island, [("x", 3), ("y", 3)], {"x": (["x"], x), "y": (["y"], x), "depth": (["x", "y"], 5.0 * np.ones((3, 3)))})
    d = mantaray.ray_tracing(3 * [-1000], 3 * [0], 3 * [0.01], 3 * [0], 10, 2, str(island), str(current))
    assert not np.array_equal(np.asarray(d.x), np.asarray(a.x))                  # 5 m of water: slower group velocity
    mantaray_b200.clear_cache()
    assert mantaray_b200.cache_info()["entries"] == 0


def test_file_errors_raise(gpu, tmp_path):
    with pytest.raises(OSError):
        mantaray.single_ray(0, 0, 0.01, 0, 10, 2, tmp_path / "nope.nc", tmp_path / "nope2.nc")
    junk = tmp_path / "junk.nc"
    junk.write_bytes(b"not a netcdf file at all")
    with pytest.raises(_capi.MantarayError):
        Fields.open_netcdf3(junk, None)


@pytest.mark.parametrize("spacing", [500.0, 10.0, 25.0, 50.0, 1000.0, 1.0, 0.1, 3.0, 37.3, 41.7, 1e-3, 12345.678, 0.30000001192092896, 7.0e5])
def test_f32_index_division_is_exact_for_every_float(gpu, spacing):
    """fdiv_const (Markstein, two exact-residual steps) == IEEE divide for ALL 2^31 non-negative finite floats."""
    rc, bad, usable = mrtools.selftest_fdiv(0, spacing)
    assert rc == 0 and usable == 1
    assert bad == 0, f"{bad} floats divide differently by {spacing}"


def test_division_shortcut_refuses_the_excluded_divisors(gpu):
    all_ones = np.frombuffer(np.uint32(0x3fffffff).tobytes(), dtype=np.float32)[0]      # significand all ones
    assert mrtools.selftest_fdiv(0, float(all_ones))[::2] == (0, 0)
    assert mrtools.selftest_fdiv(0, 1e-40)[::2] == (0, 0)


def test_non_affine_grid_takes_the_general_path_and_agrees(oracle, gpu):
    """x = f32(i) * 0.1f is NOT exactly affine in f32: the kernel must load the corner coordinates and divide
    per cell (interpolator.rs:64-72) instead of using launch constants."""
    n = 300
    x = (np.arange(n, dtype=np.float32) * np.float32(0.1)).astype(np.float32)
    assert len(set(np.diff(x.astype(np.float64)))) > 1
    X, Y = np.meshgrid(x.astype(np.float64), x.astype(np.float64))
    bathy = CartesianNetcdf3(x, x, 2.0 + 0.5 * np.sin(X) * np.cos(0.7 * Y))
    cur = CartesianCurrent(x.astype(np.float64), x.astype(np.float64), 0.2 * np.cos(Y), 0.1 * np.sin(X))
    rng = np.random.default_rng(5)
    m = 2000
    th = rng.uniform(0, 2 * np.pi, m)
    rays = (rng.uniform(1, 28, m), rng.uniform(1, 28, m), 1.5 * np.cos(th), 1.5 * np.sin(th))
    ref = oracle.trace_many(bathy, cur, *rays, 0.0, 6.0, 0.01)
    for math in (_abi.MR_MATH_FAST, _abi.MR_MATH_STRICT):
        with Fields(bathy, cur) as f:
            res = trace_many(f, *rays, 0.0, 6.0, 0.01, math=math)
        assert_parity(res, ref, what=f"non-affine math={math}")


def test_rays_on_grid_lines_and_nodes(oracle, gpu):
    """Starts exactly on grid nodes / lines / domain edges: the corner-coincidence early return of
    interpolator.rs:46-50 and the edge rules of four_corners."""
    wl = W.c3_shear_jet(16, 40, nx=64)
    d = 50.0
    xs = np.array([0.0, d, 2 * d, 10 * d, 63 * d, 63 * d, 5 * d, 5.5 * d, 0.0, 31 * d, 32 * d, 32 * d])
    ys = np.array([0.0, d, 7.5 * d, 10 * d, 63 * d, 0.0, 5 * d, 5 * d, 63 * d, 20 * d, 20 * d, 20.5 * d])
    kx = np.array([0.04, 0.04, 0.04, -0.04, -0.04, -0.03, 0.0, 0.04, 0.03, 0.04, 0.04, -0.04])
    ky = np.array([0.01, 0.0, 0.0, 0.0, -0.01, 0.02, 0.04, 0.0, -0.02, 0.0, 0.0, 0.0])
    ref = oracle.trace_many(wl.bathymetry, wl.current, xs, ys, kx, ky, 0.0, 40.0, 1.0)
    for math in (_abi.MR_MATH_FAST, _abi.MR_MATH_STRICT):
        with Fields(wl.bathymetry, wl.current) as f:
            res = trace_many(f, xs, ys, kx, ky, 0.0, 40.0, 1.0, math=math)
        assert_parity(res, ref, what=f"grid lines math={math}")


def test_multi_device_handle_shards_and_gathers(gpu):
    """The devices of a handle share one queue of slabs; the gather is a concatenation of column blocks."""
    ndev = _capi.device_count()
    if ndev < 2:
        pytest.skip("one visible device")
    wl = W.c4_agulhas(40, 40, 200)
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f1, Fields(wl.bathymetry, wl.current, devices=list(range(ndev))) as fn:
        assert fn.device_mask == (1 << ndev) - 1
        a = trace_many(f1, *rays, 0.0, wl.duration, wl.dt, final_state=True, env=True)
        b = trace_many(fn, *rays, 0.0, wl.duration, wl.dt, final_state=True, env=True)
        assert sum(fn.last_split()) == rays[0].size and len(fn.last_split()) == ndev
        # again on the warm handle (cached work buffers on every device), in slabs, and after a trim
        c = trace_many(fn, *rays, 0.0, wl.duration, wl.dt, final_state=True, env=True, chunk_rays=64)
        split = fn.last_split()
        assert sum(split) == rays[0].size and sum(1 for s in split if s > 0) >= 2, split     # 25 slabs: more than one device worked
        fn.trim()
        d = trace_many(fn, *rays, 0.0, wl.duration, wl.dt, final_state=True, env=True)
    for other in (b, c, d):
        for name in ("t", "x", "y", "kx", "ky", "rows", "len", "final_state", "depth", "u", "v"):
            np.testing.assert_array_equal(getattr(a, name), getattr(other, name), err_msg=name)


def test_multi_device_queue_balances_a_lopsided_batch(oracle, gpu):
    """The first half of the batch stops at once (rays started on dry land), the second half runs every step.  A
    split into one contiguous block per device would leave the first device idle; the slab queue
    (src/ray.rs:105-123: rayon's par_iter balances dynamically) gives every device a share of the live rays."""
    ndev = _capi.device_count()
    if ndev < 2:
        pytest.skip("one visible device")
    wl = W.c2_sea_mount(400_000, 400)
    x0, y0, kx0, ky0 = wl.all_rays()
    half = x0.size // 2
    x0 = x0.copy(); y0 = y0.copy()
    x0[:half], y0[:half] = 0.0, np.linspace(-100.0, 100.0, half)          # on the island: h <= 0, one NaN row
    with Fields(wl.bathymetry, wl.current, devices=list(range(ndev))) as fn:
        r = trace_many(fn, x0, y0, kx0, ky0, 0.0, wl.duration, wl.dt, trajectories=False, final_state=True, chunk_rays=8192)
        split = fn.last_split()
    assert sum(split) == x0.size and sum(1 for s in split if s > 0) >= 2, split
    assert (r.rows[:half] == 2).all() and (r.rows[half:] > 2).all()
    sel = np.r_[0:half:997, half:x0.size:997]
    ref = oracle.trace_many(wl.bathymetry, wl.current, x0[sel], y0[sel], kx0[sel], ky0[sel], 0.0, wl.duration, wl.dt,
                            trajectories=False)
    np.testing.assert_array_equal(r.rows[sel], ref.rows)
    np.testing.assert_array_equal(r.len[sel], ref.len)
    # (absolute tolerances from the rays' scales: ky is ~1e-22 where it should be 0)
    np.testing.assert_allclose(r.final_state[:2, sel], ref.final_state[:2], rtol=0, atol=1e-9 * 1e4, equal_nan=True)
    np.testing.assert_allclose(r.final_state[2:, sel], ref.final_state[2:], rtol=0, atol=1e-9 * 0.04, equal_nan=True)


def test_gather_beyond_the_2d_copy_pitch_limit_goes_row_by_row(gpu, monkeypatch):
    """cudaMemcpy2DAsync refuses pitches above cudaDevAttrMaxPitch (2 GiB: 2.68e8 rays per row of doubles).  The
    limit is lowered for this handle so that an ordinary batch takes the row-by-row gather."""
    wl = W.c4_agulhas(30, 30, 150, nx=128)
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        a = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, env=True)
    monkeypatch.setenv("MR_DEBUG_MAX_PITCH", "1024")
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        b = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, env=True)
        c = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True, env=True, chunk_rays=256)
    for other in (b, c):
        for name in ("t", "x", "y", "kx", "ky", "rows", "len", "final_state", "depth", "u", "v"):
            np.testing.assert_array_equal(getattr(a, name), getattr(other, name), err_msg=name)


@pytest.mark.parametrize("math", [_abi.MR_MATH_FAST, _abi.MR_MATH_STRICT])
def test_negative_or_zero_duration_is_the_initial_row_only(oracle, gpu, math):
    """`((x_end - x) / h).ceil() as usize` saturates a negative quotient to 0 steps (ode_solvers Rk4::integrate,
    called from src/ray.rs:207-208): the ray is its initial row, not an error."""
    wl = W.c4_agulhas(4, 4, 10, nx=64)
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        for t_end in (-50.0, 0.0):
            res = trace_many(f, *rays, 0.0, t_end, wl.dt, math=math, final_state=True)
            ref = oracle.trace_many(wl.bathymetry, wl.current, *rays, 0.0, t_end, wl.dt)
            assert res.x.shape == (1, 16) and (res.rows == 1).all() and (res.len == 1).all()
            assert_parity(res, ref, what=f"t_end={t_end}")
            np.testing.assert_array_equal(res.x[0], rays[0])
        one = _capi.single_ray(f, *(a[3] for a in rays), 0.0, -1.0, wl.dt, math=math)
        assert one.shape == (1, 5) and one[0, 0] == 0.0 and one[0, 1] == rays[0][3]
        with pytest.raises(ValueError):
            trace_many(f, *rays, 0.0, 10.0, -1.0)                    # dt <= 0 stays an error
        with pytest.raises(ValueError):
            trace_many(f, *rays, 0.0, float("inf"), 1.0)


def test_work_buffers_are_reused_and_trimmed(gpu):
    """the handle keeps the device work buffers of the host-buffer path between calls (no cudaMalloc/cudaFree per
    call); any mix of sizes, slabs, planes and single rays gives the same results; mr_fields_trim gives them back"""
    import torch

    wl = W.c4_agulhas(24, 24, 300, nx=128)
    rays = wl.all_rays()
    few = tuple(a[:100] for a in rays)

    def same(a, b):
        for name in ("x", "y", "kx", "ky"):
            np.testing.assert_array_equal(getattr(a, name), getattr(b, name))
        assert np.array_equal(a.rows, b.rows) and np.array_equal(a.len, b.len)

    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        idle = torch.cuda.mem_get_info(0)[0]
        first = trace_many(f, *rays, wl.t0, wl.duration, wl.dt)
        held = torch.cuda.mem_get_info(0)[0]
        assert held < idle                                       # something is kept ...
        same(trace_many(f, *rays, wl.t0, wl.duration, wl.dt), first)
        assert torch.cuda.mem_get_info(0)[0] == held             # ... and reused as is
        small = trace_many(f, *few, wl.t0, wl.duration, wl.dt)
        np.testing.assert_array_equal(small.x, first.x[:, :100])
        same(trace_many(f, *rays, wl.t0, wl.duration, wl.dt, chunk_rays=128), first)      # two slabs
        same(trace_many(f, *rays, wl.t0, wl.duration, wl.dt, env=True, final_state=True), first)   # grows
        one = _capi.single_ray(f, *(a[7] for a in rays), wl.t0, wl.duration, wl.dt)
        t, states = one if isinstance(one, tuple) else (one[:, 0], one[:, 1:])
        np.testing.assert_array_equal(np.asarray(states)[:, 0], first.x[:int(first.rows[7]), 7])
        same(trace_many(f, *rays, wl.t0, wl.duration, wl.dt), first)
        f.trim()
        assert torch.cuda.mem_get_info(0)[0] > held - (1 << 20) and torch.cuda.mem_get_info(0)[0] >= idle - (64 << 20)
        same(trace_many(f, *rays, wl.t0, wl.duration, wl.dt), first)
        f.trim(); f.trim()
    # a failed call (out of memory) leaves the handle usable and holds nothing
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        same(trace_many(f, *rays, wl.t0, wl.duration, wl.dt), first)


@pytest.mark.parametrize("math", [_abi.MR_MATH_FAST, _abi.MR_MATH_STRICT])
@pytest.mark.parametrize("stride", [1, 7])
def test_device_entry_point_with_scattered_planes_pitch_and_optional_outputs(oracle, gpu, math, stride):
    """mr_trace_device on caller-owned device buffers: the four planes are separate allocations in arbitrary
    address order (the kernel reaches y, kx, ky through byte offsets from the x plane, negative ones included),
    the pitch exceeds n, n is not a multiple of the block size (the last block's spare threads repeat the last
    ray), and rows / len / final are each optional.  Everything outside the n used columns and the stored rows
    must keep its sentinel; everything inside must match the oracle (some rays leave the domain early, so
    whole warps stop and NaN-fill)."""
    import torch

    wl = W.c2_sea_mount(1000, 700, half=300)
    x0, y0, kx0, ky0 = wl.all_rays()
    n = x0.size
    assert n % 128 != 0
    ref = oracle.trace_many(wl.bathymetry, wl.current, x0, y0, kx0, ky0, 0.0, wl.duration, wl.dt, stride=stride)
    rows_cap = ref.x.shape[0]
    dev = torch.device("cuda", 0)
    ld, guard, sent = n + 37, 3, -777.25
    lib = _capi.load()
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        ic = torch.from_numpy(np.stack([x0, y0, kx0, ky0])).to(dev)
        # allocate ky first and x last: with torch's caching allocator the planes end up in no particular order
        planes = {}
        for name in ("ky", "y", "kx", "x"):
            planes[name] = torch.full((rows_cap + guard, ld), sent, dtype=torch.float64, device=dev)
        d_rows = torch.full((n + 5,), -9, dtype=torch.int32, device=dev)
        d_len = torch.full((n + 5,), -9, dtype=torch.int32, device=dev)
        d_fin = torch.full((4, n), sent, dtype=torch.float64, device=dev)
        opts = _abi.TraceOpts(stride, math, 0, 0)
        st = torch.cuda.current_stream()
        p = lambda t: C.c_void_p(t.data_ptr())

        def run(rows, length, fin):
            rc = lib.mr_trace_device(f.handle, 0, C.c_void_p(st.cuda_stream), n, p(ic[0]), p(ic[1]), p(ic[2]), p(ic[3]),
                                     0.0, wl.duration, wl.dt, C.byref(opts), p(planes["x"]), p(planes["y"]),
                                     p(planes["kx"]), p(planes["ky"]), ld, rows, length, fin, None)
            assert rc == 0, lib.mr_last_error()
            torch.cuda.synchronize()

        run(p(d_rows), p(d_len), p(d_fin))
        got = {k: v.cpu().numpy() for k, v in planes.items()}
        for name in ("x", "y", "kx", "ky"):
            a = got[name]
            assert (a[rows_cap:] == sent).all(), f"{name}: rows past the last stored row were written"
            assert (a[:, n:] == sent).all(), f"{name}: columns past the last ray were written"
        res = oracle.Result(ref.t, got["x"][:rows_cap, :n], got["y"][:rows_cap, :n], got["kx"][:rows_cap, :n],
                            got["ky"][:rows_cap, :n], d_rows.cpu().numpy()[:n], d_len.cpu().numpy()[:n], None, stride)
        assert_parity(res, ref, what=f"device entry point math={math} stride={stride}")
        assert (d_rows.cpu().numpy()[n:] == -9).all() and (d_len.cpu().numpy()[n:] == -9).all()
        fin = d_fin.cpu().numpy()
        assert ref.rows.min() < ref.rows.max(), "the case should contain rays that stop early"
        # final state = the last NaN-free state, stored row or not; with stride 1 it is row len-1 of the trajectory
        np.testing.assert_array_equal(np.isnan(fin), np.isnan(ref.final_state))
        np.testing.assert_allclose(fin[:2], ref.final_state[:2], rtol=0, atol=1e-9 * np.nanmax(np.abs(ref.final_state[:2])))
        np.testing.assert_allclose(fin[2:], ref.final_state[2:], rtol=0, atol=1e-9 * np.nanmax(np.abs(ref.final_state[2:])))
        if stride == 1:
            last = res.len - 1
            for j, name in enumerate(("x", "y", "kx", "ky")):
                np.testing.assert_array_equal(fin[j], got[name][last, np.arange(n)])
        # the same launch without the optional outputs leaves the planes identical
        for k in planes.values():
            k[:rows_cap, :n] = 0.0
        run(None, None, None)
        for name in ("x", "y", "kx", "ky"):
            np.testing.assert_array_equal(planes[name].cpu().numpy(), got[name])
