"""ctypes binding of tools/libmr_tools.so: the FP64 peak probe and the exhaustive f32-division self-test.

Measurement plumbing for bench.py and the GPU tests; not part of the product (`mantaray_b200/` never imports it).
Build: ``make -C tools/csrc`` (``__graft_entry__.build()`` does)."""
import ctypes as C
import os

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmr_tools.so")
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} not found: build it with `make -C tools/csrc`")
        _lib = C.CDLL(_PATH)
        _lib.mrt_measure_fp64_peak.restype = C.c_int
        _lib.mrt_measure_fp64_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.mrt_selftest_fdiv.restype = C.c_int
        _lib.mrt_selftest_fdiv.argtypes = [C.c_int, C.c_float, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    return _lib


def measure_fp64_peak(device: int = 0, millis: int = 200) -> float:
    """Sustained DFMA throughput of `device` in TFLOP/s (2 flop per DFMA)."""
    v = C.c_double()
    rc = load().mrt_measure_fp64_peak(device, millis, C.byref(v))
    if rc != 0:
        raise RuntimeError(f"mrt_measure_fp64_peak failed ({rc})")
    return float(v.value)


def selftest_fdiv(device: int, spacing: float):
    """(status, mismatches, usable) of the exhaustive comparison of fdiv_const with the IEEE divide."""
    bad, usable = C.c_uint64(), C.c_int32()
    rc = load().mrt_selftest_fdiv(device, spacing, C.byref(bad), C.byref(usable))
    return rc, int(bad.value), int(usable.value)
