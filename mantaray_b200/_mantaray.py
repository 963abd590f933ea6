"""Stand-in for the reference's PyO3 extension module ``mantaray._mantaray``.

Same two callables, same argument order and meaning as the ``#[pyfunction]``s
of src/ffi.rs:25-85; the work is done by ``libmantaray_b200.so`` through the C
ABI instead of by the Rust crate.

Differences a caller can observe:

* ``single_ray`` returns an ``ndarray`` of shape ``(rows, 5)`` instead of a list
  of 5-tuples (``np.array(list_of_tuples)`` in python/mantaray/core.py:54 yields
  exactly this array);
* ``ray_tracing`` returns a :class:`RayBundle`, a sequence whose items are the
  per-ray ``(rows_i, 5)`` arrays the reference returns as lists of tuples, and
  which also exposes the step-major arrays directly so the Dataset can be built
  without touching N*S Python objects;
* file errors raise ``OSError`` / ``MantarayError`` instead of a PyO3
  ``PanicException`` (src/ffi.rs:36-37 ``.expect``).
"""

from __future__ import annotations

import os
import threading
from collections import OrderedDict

import numpy as np

from . import _capi
from ._abi import MR_MATH_FAST

# ---- field-handle cache ----------------------------------------------------------------------------------------
# The reference opens and parses both NetCDF files on every call (src/ffi.rs:36-38, :62-64).  Here a call leaves its
# field handle — the grids resident on the devices, their cell records, the work buffers of the host path — in a small
# LRU cache keyed on what identifies the inputs: (real path, size, mtime in ns) of both files and the device set.  A
# repeated call, the usual notebook pattern (same fields, new rays), then skips the parse, the upload and the record
# build.  A file that is rewritten gets a new key.  `MANTARAY_B200_CACHE=<n>` sets the number of handles kept
# (default 2, 0 disables); `clear_cache()` frees them now.  Only the most recent handle keeps its work buffers.
_CACHE: "OrderedDict[tuple, _capi.Fields]" = OrderedDict()
_CACHE_LOCK = threading.Lock()
_CACHE_STATS = {"hits": 0, "misses": 0}


def _cache_capacity() -> int:
    try:
        return max(int(os.environ.get("MANTARAY_B200_CACHE", "2")), 0)
    except ValueError:
        return 2


def _file_key(path):
    if path is None:
        return None
    st = os.stat(path)                       # a missing file raises here as it would at open: FileNotFoundError is an OSError
    return (os.path.realpath(path), st.st_size, st.st_mtime_ns)


def clear_cache() -> None:
    """Free every cached field handle (device memory included)."""
    with _CACHE_LOCK:
        while _CACHE:
            _CACHE.popitem(last=False)[1].free()


def cache_info() -> dict:
    with _CACHE_LOCK:
        return {"entries": len(_CACHE), "capacity": _cache_capacity(), **_CACHE_STATS}


class _Borrowed:
    """Context manager over a cached (not freed on exit) or a private (freed on exit) handle."""

    def __init__(self, fields, owned: bool):
        self.fields, self.owned = fields, owned

    def __enter__(self):
        return self.fields

    def __exit__(self, *exc):
        if self.owned:
            self.fields.free()


def _open_fields(bathymetry_filename: str, current_filename: str, dev) -> _Borrowed:
    cap = _cache_capacity()
    if cap == 0:
        return _Borrowed(_capi.Fields.open_netcdf3(bathymetry_filename, current_filename, devices=dev), True)
    key = (_file_key(bathymetry_filename), _file_key(current_filename), tuple(dev))
    with _CACHE_LOCK:
        f = _CACHE.get(key)
        if f is not None:
            _CACHE.move_to_end(key)
            _CACHE_STATS["hits"] += 1
            return _Borrowed(f, False)
        _CACHE_STATS["misses"] += 1
        f = _capi.Fields.open_netcdf3(bathymetry_filename, current_filename, devices=dev)
        for old in _CACHE.values():
            old.trim()                       # only the newest handle keeps its slabs
        _CACHE[key] = f
        while len(_CACHE) > cap:
            _CACHE.popitem(last=False)[1].free()
        return _Borrowed(f, False)


def _devices():
    """Devices used by the Python API: all visible ones unless MANTARAY_B200_DEVICES says otherwise."""
    env = os.environ.get("MANTARAY_B200_DEVICES")
    if env:
        return [int(s) for s in env.split(",") if s.strip() != ""]
    n = _capi.device_count()
    if n <= 0:
        raise _capi.MantarayError(-3, "no CUDA device available (mantaray_b200 has no CPU fallback)")
    return list(range(n))


class RayBundle:
    """What ``_mantaray.ray_tracing`` returns: ``Vec<Vec<(t, x, y, kx, ky)>>`` backed by SoA arrays."""

    def __init__(self, result: _capi.TraceResult):
        self.result = result

    def __len__(self) -> int:
        return int(self.result.rows.size)

    def __getitem__(self, i: int) -> np.ndarray:
        r = self.result
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        m = int(r.rows[i])
        return np.stack([r.t[:m], r.x[:m, i], r.y[:m, i], r.kx[:m, i], r.ky[:m, i]], axis=1)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def single_ray(x0: float, y0: float, kx0: float, ky0: float, duration: float, step_size: float,
               bathymetry_filename: str, current_filename: str) -> np.ndarray:
    """src/ffi.rs:25-49.  ``t0 = 0`` (:41)."""
    dev = _devices()[:1]
    with _open_fields(str(bathymetry_filename), str(current_filename), dev) as f:
        return _capi.single_ray(f, x0, y0, kx0, ky0, 0.0, duration, step_size, math=MR_MATH_FAST)


def ray_tracing(x0, y0, kx0, ky0, duration: float, step_size: float,
                bathymetry_filename: str, current_filename: str, *, env: bool = False) -> RayBundle:
    """src/ffi.rs:51-85.  ``t0 = 0`` (:72); inputs are zipped to the shortest (:65-70).

    ``env=True`` (extension) also fills ``result.depth/u/v``: the columns of the reference's unfilled
    ``Ray`` record (src/datatype.rs:165-194) at every stored row."""
    n = min(len(x0), len(y0), len(kx0), len(ky0))
    dev = _devices()
    if n < 4096:                         # not worth more than one device
        dev = dev[:1]
    with _open_fields(str(bathymetry_filename), str(current_filename), dev) as f:
        res = _capi.trace_many(f, x0, y0, kx0, ky0, 0.0, duration, step_size, math=MR_MATH_FAST, pinned=None, env=env)
    return RayBundle(res)
