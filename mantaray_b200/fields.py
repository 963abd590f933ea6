"""Host-side mirrors of the reference's field types.

Same names and argument meaning as the Rust structs they stand for:

====================  =====================================================
``ConstantDepth``     src/bathymetry/constant_depth.rs:15-18  (default h = 1000, :16)
``ConstantSlope``     src/bathymetry/constant_slope.rs:28-44  (defaults 50, 0, 0, -0.05, 0)
``CartesianNetcdf3``  src/bathymetry/cartesian_netcdf3.rs:35-43, ``open`` :167-256
``ArrayDepth``        src/bathymetry/array_depth.rs:9-11
``ConstantCurrent``   src/current/constant_current.rs:13-17   (default (0, 0), :10)
``CartesianCurrent``  src/current/cartesian_current.rs:18-27, ``open`` :58-213
====================  =====================================================

They only hold data and lower themselves to the C descriptors of
``include/mantaray_b200.h``; all arithmetic happens on the device.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class ConstantDepth:
    """Constant depth ``h`` [m] everywhere (f32, as in the reference)."""

    def __init__(self, h: float = 1000.0):
        self.h = float(np.float32(h))

    def to_desc(self) -> _abi.BathymetryDesc:
        d = _abi.BathymetryDesc()
        d.kind = _abi.MR_BATHY_CONSTANT
        d.h0 = self.h
        return d


#: src/bathymetry/constant_depth.rs:9
DEFAULT_BATHYMETRY = ConstantDepth(2000.0)


class ConstantSlope:
    """``h = h0 + dhdx*(x-x0) + dhdy*(y-y0)`` evaluated in f32."""

    def __init__(self, h0: float = 50.0, x0: float = 0.0, y0: float = 0.0, dhdx: float = -5e-2, dhdy: float = 0.0):
        self.h0, self.x0, self.y0, self.dhdx, self.dhdy = (float(np.float32(v)) for v in (h0, x0, y0, dhdx, dhdy))

    def to_desc(self) -> _abi.BathymetryDesc:
        d = _abi.BathymetryDesc()
        d.kind = _abi.MR_BATHY_SLOPE
        d.h0, d.x0, d.y0, d.dhdx, d.dhdy = self.h0, self.x0, self.y0, self.dhdx, self.dhdy
        return d


class ArrayDepth:
    """Test aid of the reference: ``array[int(x)][int(y)]``, NaN outside, zero gradient."""

    def __init__(self, array):
        self.array = _f32(array)
        if self.array.ndim != 2:
            raise ValueError("ArrayDepth needs a 2-D array")

    def to_desc(self) -> _abi.BathymetryDesc:
        d = _abi.BathymetryDesc()
        d.kind = _abi.MR_BATHY_ARRAY
        d.nx, d.ny = self.array.shape
        d.array = self.array.ctypes.data_as(_abi.c_float_p)
        d._keep = (self.array,)
        return d


class CartesianNetcdf3:
    """Gridded bathymetry: ``x``, ``y`` as f32, ``depth`` as f64, flat ``[y][x]``.

    ``depth`` may be 2-D ``(ny, nx)`` or already flat; it is used as the flat
    buffer the reference indexes with ``nx*yi + xi``
    (src/bathymetry/cartesian_netcdf3.rs:465-471).
    """

    def __init__(self, x, y, depth):
        self.x = _f32(x).ravel()
        self.y = _f32(y).ravel()
        self.depth = _f64(depth).ravel()
        if self.depth.size != self.x.size * self.y.size:
            raise ValueError("depth must have len(x)*len(y) values")

    @classmethod
    def open(cls, path, xname: str = "x", yname: str = "y", depth_name: str = "depth") -> "CartesianNetcdf3":
        from ._capi import Nc3Reader

        with Nc3Reader(path) as f:
            return cls(f.read_f32(xname), f.read_f32(yname), f.read_f64(depth_name))

    def to_desc(self) -> _abi.BathymetryDesc:
        d = _abi.BathymetryDesc()
        d.kind = _abi.MR_BATHY_GRID
        d.nx, d.ny = self.x.size, self.y.size
        d.x = self.x.ctypes.data_as(_abi.c_float_p)
        d.y = self.y.ctypes.data_as(_abi.c_float_p)
        d.depth = self.depth.ctypes.data_as(_abi.c_double_p)
        d._keep = (self.x, self.y, self.depth)
        return d


class ConstantCurrent:
    """Constant current ``(u, v)`` [m/s], zero gradients."""

    def __init__(self, u: float = 0.0, v: float = 0.0):
        self.u, self.v = float(u), float(v)

    def to_desc(self) -> _abi.CurrentDesc:
        d = _abi.CurrentDesc()
        d.kind = _abi.MR_CURRENT_CONSTANT
        d.u0, d.v0 = self.u, self.v
        return d


#: src/current/constant_current.rs:10
DEFAULT_CURRENT = ConstantCurrent(0.0, 0.0)


class CartesianCurrent:
    """Gridded current snapshot: ``x``, ``y``, ``u``, ``v`` all f64, flat ``[y][x]``."""

    def __init__(self, x, y, u, v):
        self.x = _f64(x).ravel()
        self.y = _f64(y).ravel()
        self.u = _f64(u).ravel()
        self.v = _f64(v).ravel()
        if self.u.size != self.x.size * self.y.size or self.v.size != self.u.size:
            raise ValueError("u and v must have len(x)*len(y) values")

    @classmethod
    def open(cls, path, x_name: str = "x", y_name: str = "y", u_name: str = "u", v_name: str = "v") -> "CartesianCurrent":
        from ._capi import Nc3Reader

        with Nc3Reader(path) as f:
            return cls(f.read_f64(x_name), f.read_f64(y_name), f.read_f64(u_name), f.read_f64(v_name))

    def to_desc(self) -> _abi.CurrentDesc:
        d = _abi.CurrentDesc()
        d.kind = _abi.MR_CURRENT_GRID
        d.nx, d.ny = self.x.size, self.y.size
        d.x = self.x.ctypes.data_as(_abi.c_double_p)
        d.y = self.y.ctypes.data_as(_abi.c_double_p)
        d.u = self.u.ctypes.data_as(_abi.c_double_p)
        d.v = self.v.ctypes.data_as(_abi.c_double_p)
        d._keep = (self.x, self.y, self.u, self.v)
        return d


def as_path(p) -> bytes:
    return os.fsencode(os.fspath(p))
