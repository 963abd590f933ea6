"""The native NetCDF-3 reader (mr_nc3_*), which replaces the `netcdf3` crate calls of
CartesianNetcdf3::open / CartesianCurrent::open.  Runs without a GPU."""

import numpy as np
import pytest

from mantaray_b200 import CartesianCurrent, CartesianNetcdf3, _capi
from mantaray_b200.io_utility import create_netcdf3_bathymetry, create_netcdf3_current, write_netcdf3


def test_reads_reference_style_fixture(tmp_path):
    p = tmp_path / "b.nc"
    create_netcdf3_bathymetry(p, 7, 5, 2.5, 10.0, lambda x, y: float(x) + 100.0 * float(y))
    with _capi.Nc3Reader(p) as f:
        assert f.variables() == ["y", "x", "depth"]
        assert f.info("depth") == (6, 35, (5, 7))
        x, y, d = f.read_f32("x"), f.read_f32("y"), f.read_f64("depth")
    np.testing.assert_array_equal(x, np.arange(7, dtype=np.float32) * np.float32(2.5))
    np.testing.assert_array_equal(y, np.arange(5, dtype=np.float32) * np.float32(10.0))
    np.testing.assert_array_equal(d.reshape(5, 7), x[None, :].astype(np.float64) + 100.0 * y[:, None])
    g = CartesianNetcdf3.open(p)
    assert g.x.dtype == np.float32 and g.depth.dtype == np.float64 and g.depth[7 * 3 + 2] == 2 * 2.5 + 100.0 * 30.0


@pytest.mark.parametrize("version", [1, 2])
def test_every_external_type_casts_like_the_reference(tmp_path, version):
    """{i8, u8(char), i16, i32, f32, f64} -> `as f32` / `as f64` (cartesian_netcdf3.rs:171-253,
    cartesian_current.rs:69-211; dtype matrix test :645-657), classic and 64-bit-offset files."""
    p = tmp_path / "t.nc"
    vals = {
        "a_i8": np.array([-128, -1, 0, 127], dtype=np.int8),
        "a_u8": np.array([0, 1, 200, 255], dtype=np.uint8),
        "a_i16": np.array([-32768, -2, 3, 32767], dtype=np.int16),
        "a_i32": np.array([-2**31, -16777217, 16777217, 2**31 - 1], dtype=np.int32),
        "a_f32": np.array([-1.5, 0.1, 3.4e38, np.nan], dtype=np.float32),
        "a_f64": np.array([-1.5, 0.1, 1e300, np.nan], dtype=np.float64),
    }
    write_netcdf3(p, [("n", 4)], {k: (["n"], v) for k, v in vals.items()}, version=version)
    with _capi.Nc3Reader(p) as f:
        for k, v in vals.items():
            with np.errstate(over="ignore"):
                np.testing.assert_array_equal(f.read_f64(k), v.astype(np.float64))
                np.testing.assert_array_equal(f.read_f32(k), v.astype(np.float32))      # 16777217 -> 16777216f, 1e300 -> inf
        assert [f.info(k)[0] for k in vals] == [1, 2, 3, 4, 5, 6]


def test_one_by_one_files_of_all_types_open(tmp_path):
    """cartesian_current.rs:645-657 opens 1x1 files written as f32/f64 and as i16/i8/u8/i32."""
    p = tmp_path / "c.nc"
    create_netcdf3_current(p, 1, 1, 1.0, 1.0, lambda x, y: (5.0, 0.0))
    c = CartesianCurrent.open(p)
    assert (c.x.size, c.y.size, c.u[0], c.v[0]) == (1, 1, 5.0, 0.0)
    write_netcdf3(p, [("y", 1), ("x", 1)], {"y": (["y"], np.zeros(1, np.int8)), "x": (["x"], np.zeros(1, np.int16)),
                                           "u": (["y", "x"], np.full((1, 1), 5, np.uint8)), "v": (["y", "x"], np.zeros((1, 1), np.int32))})
    c = CartesianCurrent.open(p)
    assert c.u.dtype == np.float64 and c.u[0] == 5.0 and c.v[0] == 0.0


def test_dimension_order_is_ignored(tmp_path):
    """A file whose data variable is declared (x, y) — as python/tests/test_core.py:12 and
    support/linear_sea_mount.py:26 write it — is read as the flat buffer and indexed [y][x]."""
    p = tmp_path / "xy.nc"
    x = np.arange(3, dtype=np.float64)
    y = np.arange(4, dtype=np.float64) * 10
    depth_xy = np.arange(12, dtype=np.float64).reshape(3, 4)          # dims ("x", "y")
    write_netcdf3(p, [("x", 3), ("y", 4)], {"x": (["x"], x), "y": (["y"], y), "depth": (["x", "y"], depth_xy)})
    g = CartesianNetcdf3.open(p)
    np.testing.assert_array_equal(g.depth, np.arange(12.0))          # flat, untouched
    assert g.depth[g.x.size * 1 + 2] == 5.0                           # nx*yi + xi lands on depth_xy.flat[5]


def test_agrees_with_scipy_writer(tmp_path):
    scipy_io = pytest.importorskip("scipy.io")
    p = tmp_path / "s.nc"
    rng = np.random.default_rng(0)
    u = rng.normal(size=(6, 9))
    with scipy_io.netcdf_file(str(p), "w", version=2) as f:
        f.createDimension("t", None)                                  # a record dimension (must come first) ...
        f.createDimension("y", 6)
        f.createDimension("x", 9)
        f.createVariable("x", "f8", ("x",))[:] = np.linspace(0, 1, 9)
        f.createVariable("y", "f4", ("y",))[:] = np.linspace(5, 6, 6)
        f.createVariable("u", "f8", ("y", "x"))[:] = u
        f.createVariable("v", "i2", ("y", "x"))[:] = (100 * u).astype(np.int16)
        r = f.createVariable("series", "f4", ("t", "x"))            # ... and two record variables
        for i in range(3):
            r[i] = np.arange(9, dtype=np.float32) + 100 * i
        r2 = f.createVariable("tt", "i4", ("t",))
        r2[:] = np.arange(3)
        f.title = "with attributes"
        f.variables["u"].units = "m/s"
    with _capi.Nc3Reader(p) as f:
        np.testing.assert_array_equal(f.read_f64("x"), np.linspace(0, 1, 9))
        np.testing.assert_array_equal(f.read_f32("y"), np.linspace(5, 6, 6).astype(np.float32))
        np.testing.assert_array_equal(f.read_f64("u").reshape(6, 9), u)
        np.testing.assert_array_equal(f.read_f64("v"), (100 * u).astype(np.int16).ravel().astype(np.float64))
        assert f.info("series") == (5, 27, (3, 9))
        np.testing.assert_array_equal(f.read_f32("series").reshape(3, 9), np.arange(9, dtype=np.float32)[None] + 100 * np.arange(3)[:, None])
        np.testing.assert_array_equal(f.read_f64("tt"), [0.0, 1.0, 2.0])


def test_errors(tmp_path):
    with pytest.raises(OSError):
        _capi.Nc3Reader(tmp_path / "missing.nc")
    bad = tmp_path / "bad.nc"
    bad.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)                 # a NetCDF-4 / HDF5 signature
    with pytest.raises(_capi.MantarayError) as e:
        _capi.Nc3Reader(bad)
    assert e.value.code == -5
    ok = tmp_path / "ok.nc"
    create_netcdf3_bathymetry(ok, 3, 3, 1.0, 1.0, lambda x, y: 1.0)
    with _capi.Nc3Reader(ok) as f:
        with pytest.raises(_capi.MantarayError):
            f.read_f64("nope")
    trunc = tmp_path / "trunc.nc"
    trunc.write_bytes(ok.read_bytes()[:-16])
    with pytest.raises(_capi.MantarayError) as e:       # refused when opened: the header promises bytes the file lacks
        with _capi.Nc3Reader(trunc) as f:
            f.read_f64("depth")
    assert e.value.code == -5


# ---- hostile headers: every size a header can claim is checked against the file before anything is allocated ----
def _patch_dim(path, out, name: str, new_len: int):
    """Rewrite the length of dimension `name` in the header of a classic file, leaving everything else alone."""
    import struct

    raw = bytearray(open(path, "rb").read())
    tag = struct.pack(">I", len(name)) + name.encode() + b"\0" * (-len(name) % 4)
    at = raw.index(tag) + len(tag)
    raw[at:at + 4] = struct.pack(">I", new_len)
    open(out, "wb").write(bytes(raw))


@pytest.mark.parametrize("ylen, xlen", [(65535, 65535), (0xFFFFFFF0, 0xFFFFFFF0), (1 << 31, 8), (3, 1 << 30)])
def test_header_that_claims_more_data_than_the_file_holds_is_a_format_error(tmp_path, ylen, xlen):
    """65535 x 65535 doubles is 34 GB claimed by a 200-byte file; 0xFFFFFFF0^2 * 8 wraps 64 bits.  Both used to reach
    std::vector::resize (bad_alloc / length_error through the C ABI = std::terminate in the caller's process)."""
    good, bad = tmp_path / "good.nc", tmp_path / "bad.nc"
    create_netcdf3_bathymetry(good, 4, 3, 1.0, 1.0, lambda x, y: 10.0)
    _patch_dim(good, bad, "y", ylen)
    _patch_dim(bad, bad, "x", xlen)
    with pytest.raises(_capi.MantarayError) as e:
        _capi.Nc3Reader(bad)
    assert e.value.code == -5 and "larger than the file" in e.value.message
    with pytest.raises(_capi.MantarayError) as e:                # the path ray_tracing takes
        _capi.Fields.open_netcdf3(bad, None)
    assert e.value.code == -5


def test_truncated_data_section_is_a_format_error(tmp_path):
    good, bad = tmp_path / "good.nc", tmp_path / "bad.nc"
    create_netcdf3_bathymetry(good, 40, 30, 1.0, 1.0, lambda x, y: 10.0)
    raw = open(good, "rb").read()
    open(bad, "wb").write(raw[: len(raw) - 1000])
    with pytest.raises(_capi.MantarayError) as e:
        _capi.Nc3Reader(bad)
    assert e.value.code == -5


def test_record_variable_claiming_too_many_records_is_a_format_error(tmp_path):
    """numrecs is a header field too: 2^31 records of a 16-byte record do not fit in a 150-byte file."""
    import struct

    p = tmp_path / "rec.nc"
    name = lambda s: struct.pack(">I", len(s)) + s.encode() + b"\0" * (-len(s) % 4)
    hdr = b"CDF\x01" + struct.pack(">I", 1 << 31)                       # numrecs
    hdr += struct.pack(">II", 0x0A, 2) + name("t") + struct.pack(">I", 0) + name("n") + struct.pack(">I", 2)
    hdr += struct.pack(">II", 0, 0)
    var = name("v") + struct.pack(">I", 2) + struct.pack(">II", 0, 1) + struct.pack(">II", 0, 0) + struct.pack(">II", 6, 16)
    hdr += struct.pack(">II", 0x0B, 1) + var
    begin = len(hdr) + 4
    open(p, "wb").write(hdr + struct.pack(">I", begin) + b"\0" * 32)
    with pytest.raises(_capi.MantarayError) as e:
        _capi.Nc3Reader(p)
    assert e.value.code == -5 and "records" in e.value.message
    # the same file with an honest record count reads
    raw = bytearray(open(p, "rb").read())
    raw[4:8] = struct.pack(">I", 2)
    open(p, "wb").write(bytes(raw))
    with _capi.Nc3Reader(p) as f:
        assert f.info("v") == (6, 4, (2, 2)) and f.read_f64("v").tolist() == [0.0] * 4
