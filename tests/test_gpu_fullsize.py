"""Parity at BASELINE.json's full sizes.

The CUDA path traces the whole batch of each configuration; the oracle — which needs ~2 us per
ray-step per core — traces a uniformly strided sample of the same rays (rays are independent, so a
ray's result does not depend on its batch), and the sampled columns must agree: termination
bit-exact, trajectories / final states within 1e-9.  Plus batch-level invariants.
"""

import numpy as np
import pytest

from conftest import REL_TOL, assert_parity
from mantaray_b200 import Fields, trace_many
from mantaray_b200 import workloads as W

pytestmark = pytest.mark.gpu


class _Cols:
    """Columns `sel` of a TraceResult, shaped like a result."""

    def __init__(self, r, sel):
        self.rows, self.len = r.rows[sel], r.len[sel]
        self.x = None if r.x is None else r.x[:, sel]
        self.y = None if r.y is None else r.y[:, sel]
        self.kx = None if r.kx is None else r.kx[:, sel]
        self.ky = None if r.ky is None else r.ky[:, sel]
        self.final_state = None if r.final_state is None else r.final_state[:, sel]


def _final_state_close(a, b):
    assert np.array_equal(np.isnan(a), np.isnan(b))
    pos = np.nanmax(np.abs(b[:2]), initial=1.0)
    k = np.nanmax(np.abs(b[2:]), initial=1.0)
    with np.errstate(invalid="ignore"):
        assert np.nanmax(np.abs(a[:2] - b[:2]), initial=0.0) <= REL_TOL * pos
        assert np.nanmax(np.abs(a[2:] - b[2:]), initial=0.0) <= REL_TOL * k


def test_c2_sea_mount_100k_rays_full_trajectories(oracle, gpu):
    """configs[1]: 2001x2001 grid, 100 000 rays x 2000 steps, termination at the island's shore."""
    wl = W.c2_sea_mount()
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current) as f:
        res = trace_many(f, *rays, 0.0, wl.duration, wl.dt, final_state=True)
    assert res.x.shape == (2001, 100_000)
    sel = np.arange(0, wl.n_rays, 50)
    ref = oracle.trace_many(wl.bathymetry, wl.current, *(a[sel] for a in rays), 0.0, wl.duration, wl.dt)
    assert_parity(_Cols(res, sel), ref, what="C2 full size")
    # the island (R < 500 m has h <= 0) stops the rays aimed at it; the others cross the domain
    hit = res.rows < 2001
    assert 1000 < hit.sum() < 20_000 and np.abs(rays[1][hit]).max() < 1500.0
    # (no mirror symmetry in y is asserted: the reference's one-sided finite differences and cell rule
    # are not symmetric, and the oracle agrees with the kernel on that)


def test_c3_shear_jet_1m_rays_final_state(oracle, gpu):
    """configs[2]: 1024x1024 grid, 1M rays x 6000 steps, len + final state only."""
    wl = W.c3_shear_jet()
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current) as f:
        res = trace_many(f, *rays, 0.0, wl.duration, wl.dt, trajectories=False, final_state=True)
    assert res.rows.shape == (1_000_000,) and (res.rows == 6001).all() and (res.len == 6001).all()
    sel = np.arange(0, wl.n_rays, 2500)
    ref = oracle.trace_many(wl.bathymetry, wl.current, *(a[sel] for a in rays), 0.0, wl.duration, wl.dt, trajectories=False)
    np.testing.assert_array_equal(res.rows[sel], ref.rows)
    _final_state_close(res.final_state[:, sel], ref.final_state)
    # Snell across the current front: the field does not depend on y, so ky is conserved — exactly, the
    # y-differences of the grid being exactly zero — and every ray's kx drops by the same amount at the front
    kx = res.final_state[2]
    assert np.ptp(kx) <= 1e-9 * abs(kx[0])
    assert (kx < rays[2]).all()
    np.testing.assert_array_equal(res.final_state[3], rays[3])


def test_c5_nazare_8m_ray_shard_decimated(oracle, gpu):
    """configs[4], one GPU's shard: 4096x4096 grid, 8.4M rays (8 periods x 64 directions x 16 384 points) x 4096
    steps, stride 64; long-period rays run ashore, short-period rays finish."""
    wl = W.c5_nazare(8, 64, 16_384)
    assert wl.n_rays == 8_388_608 and wl.n_rows == 65
    rays = wl.all_rays()
    with Fields(wl.bathymetry, wl.current) as f:
        res = trace_many(f, *rays, 0.0, wl.duration, wl.dt, stride=wl.stride, trajectories=False, final_state=True)
        sel = np.arange(0, wl.n_rays, 16_411)           # 512 rays across every period / direction / start point
        dec = trace_many(f, *(a[sel] for a in rays), 0.0, wl.duration, wl.dt, stride=wl.stride, final_state=True)
    ref = oracle.trace_many(wl.bathymetry, wl.current, *(a[sel] for a in rays), 0.0, wl.duration, wl.dt, stride=wl.stride)
    assert_parity(dec, ref, what="C5 decimated rows")
    np.testing.assert_array_equal(res.rows[sel], ref.rows)
    np.testing.assert_array_equal(res.len[sel], ref.len)
    _final_state_close(res.final_state[:, sel], ref.final_state)
    assert ref.rows.min() < 4097 == ref.rows.max()      # built-in load imbalance: some rays stop early


def test_c1_canonical_through_the_python_api(oracle, gpu, tmp_path):
    """configs[0], the reference's own CPU-runnable case, through `mantaray.ray_tracing` and NetCDF files:
    constant depth 4000 m, zero current on 200x100 @ 1 km, 1 000 rays, dt 2.5 s, 10 000 steps."""
    import mantaray
    from mantaray_b200.io_utility import write_netcdf3

    wl = W.c1_canonical()
    b, c = wl.bathymetry, wl.current
    write_netcdf3(tmp_path / "bathy.nc", [("y", b.y.size), ("x", b.x.size)],
                  {"x": (["x"], b.x), "y": (["y"], b.y), "depth": (["y", "x"], b.depth.reshape(b.y.size, b.x.size))})
    write_netcdf3(tmp_path / "cur.nc", [("y", c.y.size), ("x", c.x.size)],
                  {"x": (["x"], c.x), "y": (["y"], c.y), "u": (["y", "x"], c.u.reshape(c.y.size, c.x.size)),
                   "v": (["y", "x"], c.v.reshape(c.y.size, c.x.size))})
    x0, y0, kx0, ky0 = wl.all_rays()
    ds = mantaray.ray_tracing(x0, y0, kx0, ky0, wl.duration, wl.dt, str(tmp_path / "bathy.nc"), str(tmp_path / "cur.nc"))
    assert ds.sizes["time_step"] == 10_001 and ds.sizes["ray"] == 1000
    kx, ky, x = np.asarray(ds.kx), np.asarray(ds.ky), np.asarray(ds.x)
    assert (kx == kx0[0]).all() and (ky == 0.0).all()               # constant fields: k is conserved bit for bit
    assert not np.isnan(x).any() and (np.diff(x, axis=0) > 0).all()  # nobody leaves the 199 km domain
    np.testing.assert_array_equal(np.asarray(ds.time)[:, 0], np.asarray(ds.time)[:, -1])
    sel = np.arange(0, 1000, 100)
    ref = oracle.trace_many(b, c, x0[sel], y0[sel], kx0[sel], ky0[sel], 0.0, wl.duration, wl.dt)
    scale = np.abs(ref.x).max()
    assert np.abs(x[:, sel] - ref.x).max() <= REL_TOL * scale
    assert np.abs(np.asarray(ds.y)[:, sel] - ref.y).max() <= REL_TOL * scale
    np.testing.assert_array_equal(np.asarray(ds.time)[:, 0], ref.t)


def test_c4_agulhas_1m_rays_full_trajectories(oracle, gpu):
    """configs[3], the benchmarked configuration, laid out exactly as bench.py lays it out: 1M rays x 2049 rows,
    four 16.4 GB planes in one device buffer (ld = 1 000 000, plane offsets beyond 2^34 bytes, 7 813 blocks),
    written by ONE mr_trace_device launch with the library's default flags.  2 049 strided columns plus every
    rows / len / final state come back and are held to the oracle: termination bit-exact, trajectories to 1e-9."""
    import ctypes as C

    import torch

    from mantaray_b200 import _abi, _capi

    free_b, _ = torch.cuda.mem_get_info(0)
    if free_b < 70e9:
        pytest.skip("needs 66 GB of free device memory")
    lib = _capi.load()
    wl = W.c4_agulhas()
    n, rows = wl.n_rays, wl.n_rows
    assert (n, rows) == (1_000_000, 2049)
    x0, y0, kx0, ky0 = wl.all_rays()
    dev = torch.device("cuda", 0)
    ic = torch.from_numpy(np.stack([x0, y0, kx0, ky0])).to(dev)
    traj = torch.empty((4, rows, n), dtype=torch.float64, device=dev)
    traj.fill_(-1.0)                                    # any element the launch fails to write shows up
    d_rows = torch.empty(n, dtype=torch.int32, device=dev)
    d_len = torch.empty(n, dtype=torch.int32, device=dev)
    d_fin = torch.empty((4, n), dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = torch.cuda.current_stream()
    with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
        rc = lib.mr_trace_device(f.handle, 0, C.c_void_p(st.cuda_stream), n, p(ic[0]), p(ic[1]), p(ic[2]), p(ic[3]),
                                 0.0, wl.duration, wl.dt, None, p(traj[0]), p(traj[1]), p(traj[2]), p(traj[3]), n,
                                 p(d_rows), p(d_len), p(d_fin), None)
        assert rc == 0, lib.mr_last_error()
        torch.cuda.synchronize()
    sel = np.unique(np.linspace(0, n - 1, 2049).astype(np.int64))
    sel = np.union1d(sel, [0, 1, 127, 128, n - 129, n - 128, n - 2, n - 1])      # first / last block and warp edges
    sel_t = torch.from_numpy(sel).to(dev)

    class Got:
        pass

    got = Got()
    got.rows, got.len = d_rows[sel_t].cpu().numpy(), d_len[sel_t].cpu().numpy()
    got.x, got.y, got.kx, got.ky = (traj[i][:, sel_t].cpu().numpy() for i in range(4))
    ref = oracle.trace_many(wl.bathymetry, wl.current, x0[sel], y0[sel], kx0[sel], ky0[sel], 0.0, wl.duration, wl.dt)
    worst = assert_parity(got, ref, what="C4 at the benchmark's size and layout")
    assert worst <= REL_TOL
    _final_state_close(d_fin[:, sel_t].cpu().numpy(), ref.final_state)
    # whole-batch invariants: every row of every plane was written (no sentinel left), rows beyond a ray's own
    # are NaN, and the executed ray-steps are the number bench.py divides by
    rows_all = d_rows.to(torch.int64)
    assert int((traj == -1.0).sum().item()) == 0
    last = traj[0][rows - 1]
    assert bool(torch.isnan(last)[rows_all < rows].all().item())
    assert int((rows_all - 1).sum().item()) == int((d_rows.cpu().numpy().astype(np.int64) - 1).sum())
    assert bool((d_len.to(torch.int64) <= rows_all).all().item()) and int(rows_all.max().item()) == rows
