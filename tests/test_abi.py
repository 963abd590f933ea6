"""The C-ABI library loads on a box without a GPU and exports exactly what the header declares."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from mantaray_b200 import _abi, _capi
from mantaray_b200 import ConstantCurrent, ConstantDepth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mantaray_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(_abi.SIGNATURES)


def test_library_exports_every_symbol():
    lib = C.CDLL(_capi.lib_path())
    for name in header_functions():
        assert hasattr(lib, name), f"{name} is declared in include/mantaray_b200.h but not exported"
    assert _capi.load().mr_abi_version() == 1


def test_struct_layouts_match_c():
    """sizeof/offsets as a C compiler lays the header's structs out (LP64)."""
    assert C.sizeof(_abi.BathymetryDesc) == 72 and _abi.BathymetryDesc.x.offset == 16 and _abi.BathymetryDesc.h0.offset == 48
    assert C.sizeof(_abi.CurrentDesc) == 64 and _abi.CurrentDesc.x.offset == 16 and _abi.CurrentDesc.u0.offset == 48
    assert C.sizeof(_abi.TraceOpts) == 16


def test_sizes_without_gpu():
    lib = _capi.load()
    assert lib.mr_num_steps(0.0, 10.0, 2.0) == 5 and lib.mr_num_rows(0.0, 10.0, 2.0, 1) == 6
    assert lib.mr_num_steps(0.0, 10.0, 3.0) == 4                 # ceil
    assert lib.mr_num_rows(0.0, 10.0, 1.0, 4) == 3 and lib.mr_num_rows(0.0, 10.0, 1.0, 0) == 11
    assert lib.mr_num_steps(100.0, 102.0, 1.0) == 2              # t0 != 0 (src/ray.rs tests)
    # a negative duration saturates to 0 steps like the reference's `as usize` cast: the initial row only
    assert lib.mr_num_steps(0.0, -10.0, 1.0) == 0 and lib.mr_num_rows(0.0, -10.0, 1.0, 1) == 1
    assert lib.mr_num_steps(5.0, 5.0, 1.0) == 0
    for bad in [(0.0, 10.0, 0.0), (0.0, 10.0, -1.0), (0.0, float("nan"), 1.0), (0.0, float("inf"), 1.0), (0.0, 1e30, 1e-3)]:
        assert lib.mr_num_steps(*bad) == -1


@pytest.mark.skipif(_capi.device_count() > 0, reason="needs a box WITHOUT a GPU")
def test_compute_fails_loudly_without_gpu():
    """No CPU fallback: every compute entry point reports MR_ERR_CUDA."""
    with pytest.raises(_capi.MantarayError) as e:
        _capi.Fields(ConstantDepth(10.0), ConstantCurrent(0, 0))
    assert e.value.code == _abi.MR_ERR_CUDA and "no CPU fallback" in e.value.message
    p = C.c_void_p()
    assert _capi.load().mr_host_alloc(16, C.byref(p)) == _abi.MR_ERR_CUDA


def test_descriptor_validation_messages():
    lib = _capi.load()
    h = C.c_void_p()
    b = ConstantDepth(10.0).to_desc()
    c = ConstantCurrent(0, 0).to_desc()
    b.kind = 9
    assert lib.mr_fields_create(C.byref(b), C.byref(c), 1, C.byref(h)) == _abi.MR_ERR_BAD_ARG
    assert b"unknown bathymetry kind" in lib.mr_last_error()
    b = _abi.BathymetryDesc(kind=_abi.MR_BATHY_GRID, nx=1, ny=5)
    assert lib.mr_fields_create(C.byref(b), C.byref(c), 1, C.byref(h)) == _abi.MR_ERR_BAD_ARG
    assert b"nx >= 2" in lib.mr_last_error()
    x = np.zeros(3, dtype=np.float32)                           # zero spacing
    d = np.zeros(9)
    b = _abi.BathymetryDesc(kind=_abi.MR_BATHY_GRID, nx=3, ny=3, x=x.ctypes.data_as(_abi.c_float_p),
                            y=x.ctypes.data_as(_abi.c_float_p), depth=d.ctypes.data_as(_abi.c_double_p))
    assert lib.mr_fields_create(C.byref(b), C.byref(c), 1, C.byref(h)) == _abi.MR_ERR_BAD_ARG
    assert lib.mr_fields_create(None, C.byref(c), 1, C.byref(h)) == _abi.MR_ERR_BAD_ARG
    assert lib.mr_trace_many(None, 1, None, None, None, None, 0.0, 1.0, 1.0, None, *([None] * 8)) == _abi.MR_ERR_BAD_ARG
    assert lib.mr_trace_plan(None, None) == _abi.MR_ERR_BAD_ARG
    assert b"NULL handle" in lib.mr_last_error()


def _build_cpp_example(tmp_path):
    """integration/cpp/example.cpp: the C++ host mirror of ManyRays (src/ray.rs:24-127) over the C ABI."""
    import shutil
    import subprocess

    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++ on this box")
    exe = str(tmp_path / "example")
    libdir = os.path.dirname(_capi.lib_path())
    subprocess.run([cxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "integration", "cpp", "example.cpp"),
                    "-L" + libdir, "-lmantaray_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


@pytest.mark.skipif(_capi.device_count() > 0, reason="needs a box WITHOUT a GPU")
def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    r = _build_cpp_example(tmp_path)
    assert r.returncode == 0 and "error %d" % _abi.MR_ERR_CUDA in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_runs_the_beach(tmp_path, oracle):
    """The same two rays through the Python mirror's oracle back-end: row counts and the last finite x agree."""
    from mantaray_b200 import ConstantSlope

    r = _build_cpp_example(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    got = re.findall(r"(\d+) rows, last finite x = ([-0-9.]+)", r.stdout)
    assert len(got) == 2, r.stdout
    k = 0.05
    ref = oracle.trace_many(ConstantSlope(100.0, 0.0, 0.0, -0.05, 0.0), ConstantCurrent(0.0, 0.0),
                            np.zeros(2), np.zeros(2), np.array([k * np.cos(np.pi / 6), k]), np.array([k * np.sin(np.pi / 6), 0.0]),
                            0.0, 1000.0, 1.0)
    for i, (rows, x_last) in enumerate(got):
        assert int(rows) == ref.rows[i]
        assert abs(float(x_last) - ref.x[ref.rows[i] - 2, i]) <= 1e-3


def test_missing_library_fails_loudly(monkeypatch):
    """No silent fallback: without the built CUDA extension every entry into the product raises ImportError."""
    import mantaray_b200

    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "lib_path", lambda: os.path.join(ROOT, "mantaray_b200", "no_such_library.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _capi.load()
    with pytest.raises(ImportError):
        mantaray_b200.ray_tracing([0.0], [0.0], [0.1], [0.0], 10.0, 1.0, "bathy.nc", "current.nc")
    with pytest.raises(ImportError):
        _capi.Fields(ConstantDepth(10.0), ConstantCurrent(0, 0))


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under mantaray_b200/ (Python or C++/CUDA) or include/ refers to it."""
    offenders = []
    for base in ("mantaray_b200", "mantaray", "include", "integration"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in os.path.basename(dirpath):
                continue
            for name in files:
                if not name.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs", "Makefile")):
                    continue
                text = open(os.path.join(dirpath, name), errors="replace").read()
                if re.search(r"libmr_oracle|\bfrom oracle\b|\bimport oracle\b|\borc_[a-z_]+\s*\(|mr_oracle\.h", text):
                    offenders.append(os.path.relpath(os.path.join(dirpath, name), ROOT))
    assert offenders == []


def test_header_is_plain_c(tmp_path):
    """include/mantaray_b200.h is the boundary a cgo / Rust-bindgen / ctypes user binds: it has to be C, not C++."""
    import shutil
    import subprocess

    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("no gcc on this box")
    src = tmp_path / "hdr.c"
    src.write_text('#include "mantaray_b200.h"\n'
                   "int main(void) { mr_trace_opts o = {1, MR_MATH_FAST, 0, MR_OPT_NO_DEEP_MAP}; (void)o;\n"
                   "                 return mr_abi_version() == 1 && mr_num_rows(0.0, 10.0, 2.0, 1) == 6 ? 0 : 1; }\n")
    libdir = os.path.dirname(_capi.lib_path())
    exe = str(tmp_path / "hdr")
    subprocess.run([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src),
                    "-L" + libdir, "-lmantaray_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    assert subprocess.run([exe]).returncode == 0
