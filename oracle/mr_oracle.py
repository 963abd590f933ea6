"""ctypes wrapper of the CPU oracle (``oracle/libmr_oracle.so``).

TEST INFRASTRUCTURE.  Imported only by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs — never by
``mantaray_b200``.  Field descriptors are the product's ctypes structs
(``mantaray_b200._abi``), which is a declaration-only module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mantaray_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    B, Cu = C.POINTER(_abi.BathymetryDesc), C.POINTER(_abi.CurrentDesc)
    f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.orc_bilinear.argtypes = [C.POINTER(C.c_float * 3 * 4), C.c_float, C.c_float, f32p]
    lib.orc_bathy_nearest.argtypes = [C.c_float, C.c_void_p, C.c_int, f32p]
    lib.orc_bathy_four_corners.argtypes = [B, C.c_float, C.c_float, C.POINTER(C.c_size_t * 2 * 4)]
    lib.orc_depth.argtypes = [B, C.c_float, C.c_float, f32p]
    lib.orc_depth_and_gradient.argtypes = [B, C.c_float, C.c_float, f32p, f32p, f32p]
    lib.orc_current_nearest.argtypes = [C.c_double, C.c_void_p, C.c_int, f64p]
    lib.orc_current_four_corners.argtypes = [Cu, C.c_double, C.c_double, C.POINTER(C.c_size_t * 2 * 4)]
    lib.orc_current_and_gradient.argtypes = [Cu, C.c_double, C.c_double, f64p, f64p, C.POINTER(C.c_double * 4)]
    lib.orc_current.argtypes = [Cu, C.c_double, C.c_double, f64p, f64p]
    lib.orc_sample_fields.argtypes = [B, Cu, C.c_int64] + [C.c_void_p] * 5
    lib.orc_sample_fields.restype = None
    lib.orc_group_velocity.argtypes = [C.c_double, C.c_double, f64p]
    lib.orc_dkdt_bathy.argtypes = [C.c_double] * 4 + [f64p, f64p]
    lib.orc_dkdt_bathy.restype = None
    lib.orc_odes.argtypes = [B, Cu] + [C.c_double] * 4 + [C.POINTER(C.c_double * 4)]
    lib.orc_num_steps.argtypes = [C.c_double] * 3
    lib.orc_num_steps.restype = C.c_int64
    lib.orc_trace_many.argtypes = [B, Cu, C.c_int64] + [C.c_void_p] * 4 + [C.c_double] * 3 + [C.c_int32, C.c_int32] + [C.c_void_p] * 8
    lib.orc_single_ray.argtypes = [B, Cu] + [C.c_double] * 7 + [C.c_void_p, C.c_int64]
    lib.orc_single_ray.restype = C.c_int64
    _lib = lib
    return lib


class Err(Exception):
    """The reference's ``Err(..)`` at this call."""


def bilinear(points, target) -> float:
    """interpolator::bilinear; points = 4 x (x, y, z)."""
    pts = (C.c_float * 3 * 4)()
    for i, p in enumerate(points):
        for j in range(3):
            pts[i][j] = p[j]
    out = C.c_float()
    if load().orc_bilinear(C.byref(pts), target[0], target[1], C.byref(out)):
        raise Err("bilinear")
    return out.value


def bathy_nearest(target, arr) -> float:
    a = np.ascontiguousarray(arr, dtype=np.float32)
    out = C.c_float()
    if load().orc_bathy_nearest(np.float32(target), a.ctypes.data, a.size, C.byref(out)):
        raise Err("IndexOutOfBounds")
    return out.value


def current_nearest(target, arr) -> float:
    a = np.ascontiguousarray(arr, dtype=np.float64)
    out = C.c_double()
    if load().orc_current_nearest(float(target), a.ctypes.data, a.size, C.byref(out)):
        raise Err("IndexOutOfBounds")
    return out.value


def bathy_four_corners(bathy, x, y):
    d = bathy.to_desc()
    c = (C.c_size_t * 2 * 4)()
    if load().orc_bathy_four_corners(C.byref(d), np.float32(x), np.float32(y), C.byref(c)):
        raise Err("IndexOutOfBounds")
    return [(c[i][0], c[i][1]) for i in range(4)]


def current_four_corners(current, x, y):
    d = current.to_desc()
    c = (C.c_size_t * 2 * 4)()
    if load().orc_current_four_corners(C.byref(d), float(x), float(y), C.byref(c)):
        raise Err("IndexOutOfBounds")
    return [(c[i][0], c[i][1]) for i in range(4)]


def depth(bathy, x, y) -> float:
    d = bathy.to_desc()
    h = C.c_float()
    if load().orc_depth(C.byref(d), np.float32(x), np.float32(y), C.byref(h)):
        raise Err("depth")
    return h.value


def depth_and_gradient(bathy, x, y):
    d = bathy.to_desc()
    h, gx, gy = C.c_float(), C.c_float(), C.c_float()
    if load().orc_depth_and_gradient(C.byref(d), np.float32(x), np.float32(y), C.byref(h), C.byref(gx), C.byref(gy)):
        raise Err("depth_and_gradient")
    return h.value, (gx.value, gy.value)


def current_and_gradient(current, x, y):
    d = current.to_desc()
    u, v = C.c_double(), C.c_double()
    g = (C.c_double * 4)()
    if load().orc_current_and_gradient(C.byref(d), float(x), float(y), C.byref(u), C.byref(v), C.byref(g)):
        raise Err("current_and_gradient")
    return (u.value, v.value), ((g[0], g[1]), (g[2], g[3]))


def current(current, x, y):
    d = current.to_desc()
    u, v = C.c_double(), C.c_double()
    if load().orc_current(C.byref(d), float(x), float(y), C.byref(u), C.byref(v)):
        raise Err("current")
    return u.value, v.value


def sample_fields(bathy, current, x, y):
    """``(depth f32, u, v)`` at points: ``depth()`` / ``current()``, NaN where they return Err."""
    bd, cd = bathy.to_desc(), current.to_desc()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    depth = np.empty(x.shape, dtype=np.float32)
    u, v = np.empty(x.shape), np.empty(x.shape)
    load().orc_sample_fields(C.byref(bd), C.byref(cd), x.size, x.ctypes.data, y.ctypes.data,
                             depth.ctypes.data, u.ctypes.data, v.ctypes.data)
    return depth, u, v


def group_velocity(k, h) -> float:
    out = C.c_double()
    if load().orc_group_velocity(float(k), float(h), C.byref(out)):
        raise Err("ArgumentOutOfBounds")
    return out.value


def dkdt_bathy(k, h, dhdx, dhdy):
    a, b = C.c_double(), C.c_double()
    load().orc_dkdt_bathy(float(k), float(h), float(dhdx), float(dhdy), C.byref(a), C.byref(b))
    return a.value, b.value


def odes(bathy, current, x, y, kx, ky):
    bd, cd = bathy.to_desc(), current.to_desc()
    out = (C.c_double * 4)()
    if load().orc_odes(C.byref(bd), C.byref(cd), float(x), float(y), float(kx), float(ky), C.byref(out)):
        raise Err("odes")
    return tuple(out)


def num_steps(t0, t_end, dt) -> int:
    return int(load().orc_num_steps(float(t0), float(t_end), float(dt)))


class Result:
    def __init__(self, t, x, y, kx, ky, rows, length, final_state, stride):
        self.t, self.x, self.y, self.kx, self.ky = t, x, y, kx, ky
        self.rows, self.len, self.final_state, self.stride = rows, length, final_state, stride


def trace_many(bathy, current, x0, y0, kx0, ky0, t0, t_end, dt, *, stride=1, nthreads=None,
               trajectories=True, final_state=True) -> Result:
    """Same output contract as ``mr_trace_many``."""
    lib = load()
    bd, cd = bathy.to_desc(), current.to_desc()
    x0, y0, kx0, ky0 = (np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (x0, y0, kx0, ky0))
    n = min(a.size for a in (x0, y0, kx0, ky0))
    nsteps = num_steps(t0, t_end, dt)
    if nsteps < 0:
        raise ValueError("bad time arguments")
    stride = max(int(stride), 1)
    rows_cap = nsteps // stride + 1
    t = np.empty(rows_cap)
    if trajectories:
        x, y, kx, ky = (np.empty((rows_cap, n)) for _ in range(4))
    else:
        x = y = kx = ky = None
    rows = np.empty(n, dtype=np.int32)
    length = np.empty(n, dtype=np.int32)
    fin = np.empty((4, n)) if final_state else None
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    p = lambda a: a.ctypes.data if a is not None else None
    rc = lib.orc_trace_many(C.byref(bd), C.byref(cd), n, p(x0), p(y0), p(kx0), p(ky0),
                            float(t0), float(t_end), float(dt), stride, int(nthreads),
                            p(t), p(x), p(y), p(kx), p(ky), p(rows), p(length), p(fin))
    if rc:
        raise ValueError(f"orc_trace_many failed: {rc}")
    return Result(t, x, y, kx, ky, rows, length, fin, stride)


def single_ray(bathy, current, x0, y0, kx0, ky0, t0, t_end, dt) -> np.ndarray:
    """Rows of (t, x, y, kx, ky) as ffi::single_ray returns them."""
    lib = load()
    bd, cd = bathy.to_desc(), current.to_desc()
    cap = num_steps(t0, t_end, dt) + 1
    out = np.empty((cap, 5))
    n = lib.orc_single_ray(C.byref(bd), C.byref(cd), float(x0), float(y0), float(kx0), float(ky0),
                           float(t0), float(t_end), float(dt), out.ctypes.data, cap)
    if n < 0:
        raise ValueError("bad time arguments")
    return out[:n]
