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

import numpy as np

from . import _capi
from ._abi import MR_MATH_FAST


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
    with _capi.Fields.open_netcdf3(str(bathymetry_filename), str(current_filename), devices=dev) as f:
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
    with _capi.Fields.open_netcdf3(str(bathymetry_filename), str(current_filename), devices=dev) as f:
        res = _capi.trace_many(f, x0, y0, kx0, ky0, 0.0, duration, step_size, math=MR_MATH_FAST, pinned=None, env=env)
    return RayBundle(res)
