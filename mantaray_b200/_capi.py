"""ctypes binding of ``libmantaray_b200.so`` and the host-side batch driver.

``ManyRays`` / ``SingleRay`` mirror src/ray.rs:24-214 (same constructor
arguments, ``trace_many(start_time, end_time, step_size)`` /
``trace_individual(...)``), but the integration runs on the GPU through the C
ABI of ``include/mantaray_b200.h``.  There is no CPU fallback: if the shared
library is missing or no CUDA device is present, calls raise.
"""

from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _abi
from .fields import DEFAULT_BATHYMETRY, DEFAULT_CURRENT, as_path

_LIB_NAME = "libmantaray_b200.so"
_lib: Optional[C.CDLL] = None


class MantarayError(RuntimeError):
    """A failed call into the C ABI (status < 0)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


def lib_path() -> str:
    """The in-tree CUDA extension.  ``MANTARAY_B200_LIB`` points at another build of the same ABI (the A/B
    variants of ``make VARIANT=...``); it is still this library's sm_100a code, never a fallback."""
    override = os.environ.get("MANTARAY_B200_LIB")
    if override:
        return os.path.abspath(override)
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load() -> C.CDLL:
    """Load the CUDA extension; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build it with `make -C mantaray_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "mantaray_b200 has no CPU fallback."
            )
        _lib = _abi.declare(C.CDLL(path))
        if _lib.mr_abi_version() != 1:
            raise ImportError(f"{path}: unexpected ABI version {_lib.mr_abi_version()}")
    return _lib


def _check(rc: int) -> None:
    if rc == _abi.MR_OK:
        return
    msg = load().mr_last_error().decode("utf-8", "replace")
    if rc == _abi.MR_ERR_IO:
        raise OSError(f"[{rc}] {msg}")
    if rc == _abi.MR_ERR_OOM:
        raise MemoryError(f"[{rc}] {msg}")
    raise MantarayError(rc, msg)


def device_count() -> int:
    return int(load().mr_device_count())


# ---- pinned host arrays -------------------------------------------------------

def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """A page-locked numpy array (``mr_host_alloc``), freed when collected."""
    lib = load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = C.c_void_p()
    _check(lib.mr_host_alloc(max(nbytes, 1), C.byref(p)))
    buf = (C.c_byte * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=nbytes // dtype.itemsize).reshape(shape)
    weakref.finalize(buf, lib.mr_host_free, p.value)
    return arr


# ---- NetCDF-3 -------------------------------------------------------------------

class Nc3Reader:
    """The library's NetCDF-3 reader (``mr_nc3_*``)."""

    def __init__(self, path):
        self._lib = load()
        self._h = C.c_void_p()
        _check(self._lib.mr_nc3_open(as_path(path), C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            self._lib.mr_nc3_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def variables(self):
        out = []
        buf = C.create_string_buffer(256)
        for i in range(self._lib.mr_nc3_var_count(self._h)):
            _check(self._lib.mr_nc3_var_name(self._h, i, buf, 256))
            out.append(buf.value.decode())
        return out

    def info(self, name: str):
        t = C.c_int32()
        n = C.c_int64()
        nd = C.c_int32()
        dims = (C.c_int64 * _abi.MR_NC3_MAX_DIMS)()
        _check(self._lib.mr_nc3_var_info(self._h, name.encode(), C.byref(t), C.byref(n), C.byref(nd), dims))
        return int(t.value), int(n.value), tuple(int(dims[i]) for i in range(min(nd.value, _abi.MR_NC3_MAX_DIMS)))

    def read_f32(self, name: str) -> np.ndarray:
        _, n, _ = self.info(name)
        out = np.empty(n, dtype=np.float32)
        _check(self._lib.mr_nc3_read_f32(self._h, name.encode(), out.ctypes.data, n))
        return out

    def read_f64(self, name: str) -> np.ndarray:
        _, n, _ = self.info(name)
        out = np.empty(n, dtype=np.float64)
        _check(self._lib.mr_nc3_read_f64(self._h, name.encode(), out.ctypes.data, n))
        return out


# ---- field handle ------------------------------------------------------------------

def _mask(devices) -> int:
    if devices is None:
        return 1
    if isinstance(devices, int):
        return 1 << devices
    m = 0
    for d in devices:
        m |= 1 << int(d)
    return m


class Fields:
    """Both fields resident on the selected devices (``mr_fields``)."""

    def __init__(self, bathymetry=None, current=None, devices=None):
        self._lib = load()
        self._h = C.c_void_p()
        b = (bathymetry if bathymetry is not None else DEFAULT_BATHYMETRY).to_desc()
        c = (current if current is not None else DEFAULT_CURRENT).to_desc()
        _check(self._lib.mr_fields_create(C.byref(b), C.byref(c), _mask(devices), C.byref(self._h)))

    @classmethod
    def open_netcdf3(cls, bathymetry_path, current_path, devices=None) -> "Fields":
        """``CartesianNetcdf3::open(.., "x","y","depth")`` + ``CartesianCurrent::open(.., "x","y","u","v")``
        (src/ffi.rs:36-38)."""
        self = cls.__new__(cls)
        self._lib = load()
        self._h = C.c_void_p()
        bp = as_path(bathymetry_path) if bathymetry_path is not None else None
        cp = as_path(current_path) if current_path is not None else None
        _check(self._lib.mr_fields_open_netcdf3(bp, cp, _mask(devices), C.byref(self._h)))
        return self

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise ValueError("Fields handle already freed")
        return self._h

    @property
    def device_mask(self) -> int:
        return int(self._lib.mr_fields_device_mask(self.handle))

    def last_split(self):
        """Rays each device of the handle took in the last host-buffer call (``mr_fields_last_split``)."""
        out = (C.c_int64 * 32)()
        n = int(self._lib.mr_fields_last_split(self.handle, out, 32))
        if n < 0:
            _check(n)
        return [int(out[i]) for i in range(n)]

    def plan(self, math: int = 0, flags: int = 0) -> int:
        """Which kernel specialisations a trace with these options runs on this handle: ``MR_PLAN_*`` bits
        (``mr_trace_plan``)."""
        opts = _abi.TraceOpts(1, math, 0, flags)
        p = int(self._lib.mr_trace_plan(self.handle, C.byref(opts)))
        if p < 0:
            _check(p)
        return p

    def trim(self) -> None:
        """Give the cached device work buffers of the host-buffer path back (``mr_fields_trim``)."""
        if self._h:
            load().mr_fields_trim(self._h)

    def free(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mr_fields_free(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- results -------------------------------------------------------------------------

@dataclass
class TraceResult:
    """Step-major structure-of-arrays output of one batch.

    ``x, y, kx, ky`` are ``(rows_cap, n)``; rows a ray never reached are NaN.
    ``rows[i]`` is what the reference would have stored for ray i (including the
    trailing NaN row that stops it), ``len[i]`` the leading NaN-free rows
    (src/ray_result.rs:123-151), ``final_state`` the last NaN-free (x, y, kx, ky).
    """

    t: np.ndarray
    x: Optional[np.ndarray]
    y: Optional[np.ndarray]
    kx: Optional[np.ndarray]
    ky: Optional[np.ndarray]
    rows: np.ndarray
    len: np.ndarray
    final_state: Optional[np.ndarray]
    stride: int = 1
    #: the environment at every stored row (``env=True``): depth f32, current f64, ``(rows_cap, n)``
    depth: Optional[np.ndarray] = None
    u: Optional[np.ndarray] = None
    v: Optional[np.ndarray] = None


def num_steps(t0: float, t_end: float, dt: float) -> int:
    n = int(load().mr_num_steps(t0, t_end, dt))
    if n < 0:
        raise ValueError("need step_size > 0 and a finite (end - start)/step_size < 2**31")
    return n


def _opts(stride: int, math: int, chunk_rays: int, flags: int = 0) -> _abi.TraceOpts:
    return _abi.TraceOpts(int(stride), int(math), int(chunk_rays), int(flags))


def trace_many(fields: Fields, x0, y0, kx0, ky0, t0: float, t_end: float, dt: float, *,
               stride: int = 1, math: int = _abi.MR_MATH_FAST, chunk_rays: int = 0,
               trajectories: bool = True, final_state: bool = False, pinned: Optional[bool] = False,
               env: bool = False, flags: int = 0) -> TraceResult:
    """``mr_trace_many`` on numpy arrays; ``env=True`` adds depth, u, v at every stored row (``mr_trace_many_env``);
    ``flags`` are the ``MR_OPT_*`` bits of ``mr_trace_opts``."""
    lib = load()
    x0 = np.ascontiguousarray(x0, dtype=np.float64).ravel()
    y0 = np.ascontiguousarray(y0, dtype=np.float64).ravel()
    kx0 = np.ascontiguousarray(kx0, dtype=np.float64).ravel()
    ky0 = np.ascontiguousarray(ky0, dtype=np.float64).ravel()
    n = min(x0.size, y0.size, kx0.size, ky0.size)       # zip() truncates to the shortest, src/ffi.rs:65-70
    nsteps = num_steps(t0, t_end, dt)
    stride = max(int(stride), 1)
    rows_cap = nsteps // stride + 1
    def alloc(shape, dtype=np.float64):
        # page-locked planes let the device-to-host gather run at PCIe rate and overlap the kernels;
        # pinned=None picks them for outputs large enough to matter and falls back if the driver refuses
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        if pinned or (pinned is None and nbytes >= (32 << 20)):
            try:
                return pinned_empty(shape, dtype)
            except (MantarayError, MemoryError):
                if pinned:
                    raise
        return np.empty(shape, dtype=dtype)

    t = np.empty(rows_cap, dtype=np.float64)
    if trajectories:
        x, y, kx, ky = (alloc((rows_cap, n)) for _ in range(4))
    else:
        x = y = kx = ky = None
    rows = np.empty(n, dtype=np.int32)
    length = np.empty(n, dtype=np.int32)
    fin = np.empty((4, n), dtype=np.float64) if final_state else None
    o = _opts(stride, math, chunk_rays, flags)
    ptr = lambda a: a.ctypes.data if a is not None else None
    if env:
        if not trajectories:
            raise ValueError("env=True needs trajectories=True")
        depth, u, v = alloc((rows_cap, n), np.float32), alloc((rows_cap, n)), alloc((rows_cap, n))
        planes = _abi.EnvPlanes(ptr(depth), ptr(u), ptr(v))
        _check(lib.mr_trace_many_env(fields.handle, n, ptr(x0), ptr(y0), ptr(kx0), ptr(ky0),
                                     float(t0), float(t_end), float(dt), C.byref(o),
                                     ptr(t), ptr(x), ptr(y), ptr(kx), ptr(ky), ptr(rows), ptr(length), ptr(fin),
                                     C.byref(planes)))
        return TraceResult(t, x, y, kx, ky, rows, length, fin, stride, depth, u, v)
    _check(lib.mr_trace_many(fields.handle, n, ptr(x0), ptr(y0), ptr(kx0), ptr(ky0),
                             float(t0), float(t_end), float(dt), C.byref(o),
                             ptr(t), ptr(x), ptr(y), ptr(kx), ptr(ky), ptr(rows), ptr(length), ptr(fin)))
    return TraceResult(t, x, y, kx, ky, rows, length, fin, stride)


def sample_fields(fields: Fields, x, y):
    """``mr_sample_fields``: ``(depth f32, u, v)`` of the fields at points, by the reference's ``depth()``
    (src/bathymetry/mod.rs:38) and ``current()`` (src/current/mod.rs:24); NaN where they return Err."""
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if x.shape != y.shape:
        raise ValueError("x and y must have the same shape")
    depth = np.empty(x.shape, dtype=np.float32)
    u = np.empty(x.shape, dtype=np.float64)
    v = np.empty(x.shape, dtype=np.float64)
    _check(lib.mr_sample_fields(fields.handle, x.size, x.ctypes.data, y.ctypes.data,
                                depth.ctypes.data, u.ctypes.data, v.ctypes.data))
    return depth, u, v


def single_ray(fields: Fields, x0: float, y0: float, kx0: float, ky0: float,
               t0: float, t_end: float, dt: float, *, math: int = _abi.MR_MATH_FAST) -> np.ndarray:
    """``mr_single_ray``: rows of (t, x, y, kx, ky), shape ``(rows, 5)``."""
    lib = load()
    cap = num_steps(t0, t_end, dt) + 1
    out = np.empty((cap, 5), dtype=np.float64)
    n_rows = C.c_int64()
    o = _opts(1, math, 0)
    _check(lib.mr_single_ray(fields.handle, float(x0), float(y0), float(kx0), float(ky0),
                             float(t0), float(t_end), float(dt), C.byref(o),
                             out.ctypes.data, cap, C.byref(n_rows)))
    return out[: n_rows.value]


# ---- the batch driver, mirroring src/ray.rs ------------------------------------------------

class RayState:
    """``RayState<f64>`` (src/datatype.rs:117-138): a point and a wavenumber."""

    __slots__ = ("x", "y", "kx", "ky")

    def __init__(self, x: float, y: float, kx: float, ky: float):
        self.x, self.y, self.kx, self.ky = float(x), float(y), float(kx), float(ky)


class SingleRay:
    """``SingleRay`` (src/ray.rs:130-214)."""

    def __init__(self, bathymetry_data, current_data, initial_ray: RayState, devices=None):
        self._fields = Fields(bathymetry_data, current_data, devices)
        self.initial_ray = initial_ray

    def trace_individual(self, start_time: float, end_time: float, step_size: float, *, math: int = _abi.MR_MATH_FAST):
        """Returns ``(t, states)`` like ``SolverResult::get()``: ``t`` of shape (rows,), ``states`` (rows, 4)."""
        r = self.initial_ray
        out = single_ray(self._fields, r.x, r.y, r.kx, r.ky, start_time, end_time, step_size, math=math)
        return out[:, 0].copy(), out[:, 1:].copy()


class ManyRays:
    """``ManyRays`` (src/ray.rs:24-127)."""

    def __init__(self, bathymetry_data, current_data, initial_rays: Sequence[RayState], devices=None):
        self._fields = Fields(bathymetry_data, current_data, devices)
        self.initial_rays = list(initial_rays)

    def trace_many(self, start_time: float, end_time: float, step_size: float, *, math: int = _abi.MR_MATH_FAST):
        """A list with one ``(t, states)`` pair per ray, each cut to the rows that ray stored
        (``Vec<Option<SolverResult>>``; the integration itself never fails, so no ``None``)."""
        r = self.initial_rays
        res = trace_many(self._fields, [s.x for s in r], [s.y for s in r], [s.kx for s in r], [s.ky for s in r],
                         start_time, end_time, step_size, math=math)
        out = []
        for i in range(len(r)):
            m = int(res.rows[i])
            out.append((res.t[:m].copy(), np.stack([res.x[:m, i], res.y[:m, i], res.kx[:m, i], res.ky[:m, i]], axis=1)))
        return out


def depth_floor_map(bathy):
    """``mr_depth_floor_map``: ``(map[nby, nbx] f32, deep_frac, affine)`` — per block of 8 x 8 cells the square of a
    lower bound of every depth the lookup can return there (0: no bound); ``affine`` tells whether the fast path
    would use it on this grid.  Host only."""
    lib = load()
    bd = bathy.to_desc()
    nbx, nby, frac, aff = C.c_int32(), C.c_int32(), C.c_float(), C.c_int32()
    _check(lib.mr_depth_floor_map(C.byref(bd), None, 0, C.byref(nbx), C.byref(nby), C.byref(frac), C.byref(aff)))
    out = np.empty((nby.value, nbx.value), dtype=np.float32)
    _check(lib.mr_depth_floor_map(C.byref(bd), out.ctypes.data, out.size, C.byref(nbx), C.byref(nby), C.byref(frac),
                                  C.byref(aff)))
    return out, frac.value, bool(aff.value)


def uniform_current_map(current):
    """``mr_uniform_current_map``: ``(map[nby, nbx, 2] f32, uniform_frac, affine)`` — per block of 8 x 8 cells the
    {u, v} every node of the block holds, NaNs where the block is not uniform.  Host only."""
    lib = load()
    cd = current.to_desc()
    nbx, nby, frac, aff = C.c_int32(), C.c_int32(), C.c_float(), C.c_int32()
    _check(lib.mr_uniform_current_map(C.byref(cd), None, 0, C.byref(nbx), C.byref(nby), C.byref(frac), C.byref(aff)))
    out = np.empty((nby.value, nbx.value, 2), dtype=np.float32)
    _check(lib.mr_uniform_current_map(C.byref(cd), out.ctypes.data, out.size, C.byref(nbx), C.byref(nby), C.byref(frac),
                                      C.byref(aff)))
    return out, frac.value, bool(aff.value)
