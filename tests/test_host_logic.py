"""Host-side logic that needs no GPU: Dataset assembly (python/mantaray/core.py:114-132), the
RayBundle sequence view of `Vec<Vec<(t,x,y,kx,ky)>>`, workload sharding, and the N>1 bench
reduction logic over gloo."""

import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from mantaray_b200 import _capi, _mantaray, core
from mantaray_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_result():
    nan = np.nan
    t = np.array([0.0, 2.0, 4.0, 6.0])
    x = np.array([[1.0, 10.0, 100.0], [2.0, 20.0, 200.0], [3.0, nan, 300.0], [nan, nan, nan]])
    rows = np.array([4, 3, 3], dtype=np.int32)          # ray 2 ran out of steps... ray 0 stopped on a NaN row
    ln = np.array([3, 2, 3], dtype=np.int32)
    return _capi.TraceResult(t, x, x + 0.5, x * 0 + 0.01, x * 0, rows, ln, None, 1)


def test_ray_bundle_is_a_sequence_of_per_ray_rows():
    b = _mantaray.RayBundle(fake_result())
    assert len(b) == 3
    assert [r.shape for r in b] == [(4, 5), (3, 5), (3, 5)]
    np.testing.assert_array_equal(b[1][:, 0], [0.0, 2.0, 4.0])
    np.testing.assert_array_equal(b[1][:2, 1], [10.0, 20.0])
    assert np.isnan(b[0][3, 1:]).all() and b[0][3, 0] == 6.0     # trailing NaN row keeps its time
    with pytest.raises(IndexError):
        b[3]


def test_ray_tracing_dataset_assembly(monkeypatch):
    """NaN-pad to the LONGEST ray (not to S+1), dims (time_step, ray), time NaN beyond a ray's rows."""
    res = fake_result()
    big = _capi.TraceResult(np.arange(6) * 2.0, *(np.vstack([a, np.full((2, 3), np.nan)]) for a in (res.x, res.y, res.kx, res.ky)),
                            res.rows, res.len, None, 1)
    monkeypatch.setattr(_mantaray, "ray_tracing", lambda *a: _mantaray.RayBundle(big))
    ds = core.ray_tracing([0] * 3, [0] * 3, [0] * 3, [0] * 3, 10.0, 2.0, "b.nc", "c.nc")
    assert dict(ds.sizes) == {"time_step": 4, "ray": 3}
    np.testing.assert_array_equal(np.asarray(ds["time"])[:, 0], [0.0, 2.0, 4.0, 6.0])
    np.testing.assert_array_equal(np.asarray(ds["time"])[:, 1], [0.0, 2.0, 4.0, np.nan])
    np.testing.assert_array_equal(np.asarray(ds["x"]), res.x)
    assert list(np.asarray(ds["time_step"])) == [0, 1, 2, 3] and list(np.asarray(ds["ray"])) == [0, 1, 2]
    assert "date_created" in ds.attrs
    assert (np.asarray(ds.kx)[~np.isnan(np.asarray(ds.kx))] == 0.01).all()


def test_single_ray_dataset_assembly(monkeypatch):
    rows = np.array([[0.0, 1.0, 2.0, 0.01, 0.0], [2.0, 3.0, 4.0, 0.01, 0.0]])
    monkeypatch.setattr(_mantaray, "single_ray", lambda *a: rows)
    ds = core.single_ray(0, 0, 0.01, 0, 2.0, 2.0, "b.nc", "c.nc")
    assert dict(ds.sizes) == {"time_step": 2}
    np.testing.assert_array_equal(np.asarray(ds["x"]), [1.0, 3.0])
    assert (np.asarray(ds.kx) == 0.01).all() and (np.asarray(ds.ky) == 0.0).all()


def test_shard_ranges_partition_the_rays():
    for n, world in [(1_000_000, 8), (1000, 3), (5, 8), (128, 2), (67_108_864, 8)]:
        spans = [W.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))        # contiguous, in input order
        assert all(lo % 128 == 0 for lo, hi in spans if lo < n)           # whole thread blocks
    wl = W.c4_agulhas(4, 8, 16, nx=64)
    whole = np.stack(wl.all_rays())
    parts = np.concatenate([np.stack(wl.rays(*W.shard_range(wl.n_rays, r, 3, align=4))) for r in range(3)], axis=1)
    np.testing.assert_array_equal(whole, parts)                           # a rank builds exactly its block


def test_workload_shapes():
    for name, wl in [("C1", W.c1_canonical(10, 10)), ("C2", W.c2_sea_mount(10, 10, half=20)), ("C3", W.c3_shear_jet(10, 10, nx=32)),
                     ("C4", W.c4_agulhas(3, 3, 10, nx=32)), ("C5", W.c5_nazare(2, 2, 3, 10, nx=32, stride=5))]:
        x0, y0, kx0, ky0 = wl.all_rays()
        assert x0.shape == y0.shape == kx0.shape == ky0.shape == (wl.n_rays,)
        assert wl.bathymetry.x.dtype == np.float32 and wl.current.x.dtype == np.float64
        assert wl.n_steps == 10 and wl.n_rows == 10 // wl.stride + 1, name


GLOO_SCRIPT = textwrap.dedent("""
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from mantaray_b200 import workloads as W
    from oracle import mr_oracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    wl = W.c2_sea_mount(512, 120, half=60)      # 2 x 256: equal shards (gloo all_gather needs equal sizes)
    lo, hi = W.shard_range(wl.n_rays, rank, world)
    r = O.trace_many(wl.bathymetry, wl.current, *wl.rays(lo, hi), 0.0, wl.duration, wl.dt, nthreads=1)
    # what bench.py reduces: executed ray-steps (SUM) and elapsed time (MAX)
    e = torch.tensor([float((r.rows - 1).sum())], dtype=torch.float64); dist.all_reduce(e, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # and the gather is a concatenation along `ray`
    rows = [torch.zeros(W.shard_range(wl.n_rays, k, world)[1] - W.shard_range(wl.n_rays, k, world)[0], dtype=torch.int32) for k in range(world)]
    dist.all_gather(rows, torch.from_numpy(r.rows.copy()))
    if rank == 0:
        full = O.trace_many(wl.bathymetry, wl.current, *wl.all_rays(), 0.0, wl.duration, wl.dt, nthreads=1)
        ok = bool((torch.cat(rows).numpy() == full.rows).all())
        print(json.dumps({"E": e.item(), "E_full": float((full.rows - 1).sum()), "tmax": t.item(), "concat_ok": ok}))
    dist.destroy_process_group()
""")


def test_two_rank_sharding_over_gloo(tmp_path):
    """world_size 2 on CPU: per-rank shards cover the batch, SUM/MAX reductions as bench.py does them."""
    script = tmp_path / "gloo_shard.py"
    script.write_text(GLOO_SCRIPT % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["E"] == r["E_full"] and r["tmax"] == 2.0 and r["concat_ok"]


# ---- the depth-floor map of MR_OPT_DEEP_MAP (host side; the kernel side is tests/test_gpu_deep_map.py) ----------
def _map_cases():
    from mantaray_b200 import CartesianNetcdf3
    from test_gpu_fuzz import make_case

    yield "C4", W.c4_agulhas(4, 4, 10, nx=256).bathymetry
    yield "C2", W.c2_sea_mount(8, 10, half=200).bathymetry
    yield "C5", W.c5_nazare(2, 2, 2, 10, nx=512).bathymetry
    for seed in range(40):                              # even seeds: dry, infinite and NaN nodes
        b = make_case(seed)[0]
        if isinstance(b, CartesianNetcdf3):
            yield f"fuzz{seed}", b


def test_depth_floor_map_is_a_lower_bound_of_the_reference_lookup(oracle):
    """Soundness of the shortcut: wherever the map holds H^2 > 0, every depth the reference's f32 bilinear
    (oracle.sample_fields = BathymetryData::depth) returns in that block is >= H, on nodes, grid lines and
    anywhere in between; blocks touching a NaN, infinite or non-positive node hold 0.  And the kernel's f32 test
    RN(f32(k^2) * H^2) >= 484.01 then implies k * depth >= 22 for the depth the lookup would have returned."""
    from mantaray_b200 import ConstantCurrent
    from mantaray_b200._capi import depth_floor_map

    rng = np.random.default_rng(11)
    checked = n_affine = 0
    for name, b in _map_cases():
        m, frac, affine = depth_floor_map(b)
        nx, ny = b.x.size, b.y.size
        assert m.shape == ((ny - 1 + 7) // 8, (nx - 1 + 7) // 8), name
        assert 0.0 <= frac <= 1.0
        x, y = b.x.astype(np.float64), b.y.astype(np.float64)
        if not affine:
            continue        # perturbed or descending coordinates: the fast path does not consult the map there
        n_affine += 1
        z32 = np.asarray(b.depth, dtype=np.float64).reshape(ny, nx).astype(np.float32)
        # blocks with a bad node hold 0
        for by in range(m.shape[0]):
            for bx in range(m.shape[1]):
                blk = z32[by * 8: min(by * 8 + 8, ny - 1) + 1, bx * 8: min(bx * 8 + 8, nx - 1) + 1]
                if not (np.isfinite(blk).all() and (blk > 0).all()):
                    assert m[by, bx] == 0.0, f"{name}: block ({by},{bx}) has a bad node but a bound"
        # random points, points on grid lines and nodes
        n = 60_000
        px = rng.uniform(x[0], x[-1], n)
        py = rng.uniform(y[0], y[-1], n)
        px[: n // 8] = rng.choice(x, n // 8)
        py[n // 16: n // 5] = rng.choice(y, n // 5 - n // 16)
        depth, _, _ = oracle.sample_fields(b, ConstantCurrent(0.0, 0.0), px, py)
        # the cell the reference picks (cartesian_netcdf3.rs:274-296, 334-390): f32 index, floor, clamped to n-2
        fx = (px.astype(np.float32) - b.x[0]) / np.abs(b.x[1] - b.x[0])
        fy = (py.astype(np.float32) - b.y[0]) / np.abs(b.y[1] - b.y[0])
        inside = (fx >= 0) & (fx <= np.float32(nx - 1)) & (fy >= 0) & (fy <= np.float32(ny - 1))
        cx = np.clip(np.floor(fx).astype(np.int64), 0, nx - 2)
        cy = np.clip(np.floor(fy).astype(np.int64), 0, ny - 2)
        hsq = m[cy >> 3, cx >> 3]
        sel = inside & (hsq > 0)
        H = np.sqrt(hsq[sel].astype(np.float64))
        d = depth[sel].astype(np.float64)
        assert not np.isnan(d).any(), f"{name}: a bounded block produced a failed lookup"
        assert (d >= H).all(), f"{name}: lookup below the block's bound by {np.max(H - d)}"
        # the kernel's test in its own arithmetic
        k = 10.0 ** rng.uniform(-3, 1, sel.sum())
        k2f = (k * k).astype(np.float32)
        deep = (k2f * hsq[sel]).astype(np.float32) >= np.float32(484.01)
        assert (k[deep] * d[deep] >= 22.0).all(), f"{name}: flagged deep with kh = {np.min(k[deep] * d[deep])}"
        checked += int(deep.sum())
    assert checked > 50_000 and n_affine >= 10


def test_interleaved_tiles_cover_the_batch_once_and_mix_the_periods():
    """bench.py's C5 sharding: tiles dealt round-robin.  Every ray belongs to exactly one rank, and
    every rank holds every period of the ensemble (a contiguous block would hold one band of them)."""
    wl = W.c5_nazare(8, 8, 32, 64, nx=64)
    n, tile = wl.n_rays, wl.extra["tile"]
    for world in (1, 2, 3, 8):
        seen = np.zeros(n, dtype=np.int32)
        for r in range(world):
            ranges = W.shard_tiles(n, r, world, tile)
            periods = set()
            for lo, hi in ranges:
                seen[lo:hi] += 1
                periods.add(lo // tile // wl.extra["n_dirs"])
            if world <= 8 and world != 3:
                assert periods == set(range(wl.extra["n_periods"])), (world, r, periods)
        assert (seen == 1).all()
    assert W.shard_tiles(10, 1, 4, 3) == [(3, 6)] and W.shard_tiles(10, 3, 4, 3) == [(9, 10)] and W.shard_tiles(5, 2, 4, 3) == []


def test_field_handle_cache_keys_on_the_files_and_evicts(tmp_path, monkeypatch):
    """repeated API calls on the same files reuse one field handle; a rewritten file, another
    device set or a full cache get a new one; clear_cache() frees everything.  (Handle creation is stubbed: no GPU.)"""
    import time

    made, freed, trimmed = [], [], []

    class FakeFields:
        def __init__(self, key):
            self.key = key
            made.append(key)

        def free(self):
            freed.append(self.key)

        def trim(self):
            trimmed.append(self.key)

    monkeypatch.setattr(_capi.Fields, "open_netcdf3", classmethod(lambda cls, b, c, devices=None: FakeFields((b, c, tuple(devices)))))
    monkeypatch.setenv("MANTARAY_B200_CACHE", "2")
    _mantaray.clear_cache()
    b, c, b2 = tmp_path / "b.nc", tmp_path / "c.nc", tmp_path / "b2.nc"
    for p in (b, c, b2):
        p.write_bytes(b"x" * 10)
    base = _mantaray.cache_info()
    with _mantaray._open_fields(str(b), str(c), [0]) as f1:
        pass
    with _mantaray._open_fields(str(b), str(c), [0]) as f2:
        pass
    assert f1 is f2 and len(made) == 1 and not freed                      # reused, not freed on exit
    info = _mantaray.cache_info()
    assert (info["hits"] - base["hits"], info["misses"] - base["misses"], info["entries"]) == (1, 1, 1)
    with _mantaray._open_fields(str(b), str(c), [0, 1]) as f3:             # another device set: another handle
        pass
    assert f3 is not f1 and trimmed == [f1.key]                            # only the newest keeps its work buffers
    time.sleep(0.01)
    b.write_bytes(b"y" * 11)                                               # rewritten file: new key, oldest entry evicted
    with _mantaray._open_fields(str(b), str(c), [0]) as f4:
        pass
    assert f4 is not f1 and freed == [f1.key] and _mantaray.cache_info()["entries"] == 2
    with pytest.raises(OSError):
        _mantaray._open_fields(str(tmp_path / "missing.nc"), str(c), [0])
    _mantaray.clear_cache()
    assert _mantaray.cache_info()["entries"] == 0 and len(freed) == 3
    monkeypatch.setenv("MANTARAY_B200_CACHE", "0")                         # disabled: a private handle, freed on exit
    with _mantaray._open_fields(str(b2), str(c), [0]) as f5:
        pass
    assert freed[-1] == f5.key and _mantaray.cache_info()["entries"] == 0


def test_uniform_current_map_is_sound_against_the_oracle_lookup():
    """mr_uniform_current_map (host only): in every block it marks uniform, the reference's lookup
    (cartesian_current.rs:487-542, here the oracle's) returns exactly the block's {u, v} and exactly zero gradients
    at any point the lookup assigns to one of the block's cells; blocks with a node off by one ulp, a NaN or an
    infinity are not marked."""
    from mantaray_b200 import CartesianCurrent, uniform_current_map
    from oracle import mr_oracle as O

    O.build()
    nx, ny, d = 61, 37, 25.0
    x, y = np.arange(nx) * d, np.arange(ny) * d
    X, Y = np.meshgrid(np.arange(nx), np.arange(ny))
    u = np.where(X < 24, 0.0, 0.7).astype(np.float64)
    v = np.where(Y < 16, 0.25, -0.1).astype(np.float64)
    u[20:30, 30:42] += 0.3 * np.sin(X[20:30, 30:42])
    v[5, 9] = np.nextafter(0.25, 1.0)
    u[33, 50], v[3, 55] = np.nan, np.inf
    cur = CartesianCurrent(x, y, u, v)
    m, frac, affine = uniform_current_map(cur)
    assert affine and m.shape == ((ny - 1 + 7) // 8, (nx - 1 + 7) // 8, 2) and 0.3 < frac < 0.9
    uni = ~np.isnan(m[..., 0])
    assert not uni[0, 1] and not uni[4, 6] and not uni[0, 6]          # the ulp, the NaN, the infinity
    assert not uni[2, 2] and not uni[2, 3]                            # u steps from 0 to 0.7 at column 24: blocks 2 and 3 touch it
    rng = np.random.default_rng(0)
    n_checked = 0
    for by, bx in zip(*np.nonzero(uni)):
        for _ in range(6):
            px = rng.uniform(bx * 8 * d, min((bx + 1) * 8, nx - 1) * d)
            py = rng.uniform(by * 8 * d, min((by + 1) * 8, ny - 1) * d)
            (cu, cv), ((dudx, dudy), (dvdx, dvdy)) = O.current_and_gradient(cur, px, py)
            assert (np.float32(cu), np.float32(cv)) == (m[by, bx, 0], m[by, bx, 1]) and dudx == dudy == dvdx == dvdy == 0.0
            n_checked += 1
    assert n_checked > 100
    # a zero-current file (what "no current" looks like through the API) is uniform everywhere
    zero = CartesianCurrent(x, y, np.zeros((ny, nx)), np.zeros((ny, nx)))
    mz, fz, _ = uniform_current_map(zero)
    assert fz == 1.0 and (mz == 0.0).all()
