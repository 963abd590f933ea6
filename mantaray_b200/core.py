"""``mantaray.core`` on the B200 path: same two functions, same Dataset.

Mirrors python/mantaray/core.py:9-132.  The step-major structure-of-arrays
buffers the kernel writes already ARE the ``(time_step, ray)`` variables of the
Dataset the reference assembles (pad every ray with NaN to the longest ray,
stack, transpose; core.py:114-125), so nothing is re-packed here.

``xarray`` is imported lazily; where it is not installed a minimal stand-in with
the same ``sizes`` / variable / ``attrs`` access is returned.
"""

from __future__ import annotations

import datetime

import numpy as np

from . import _mantaray

_VARNAMES = ("time", "x", "y", "kx", "ky")


class RayDataset:
    """Tiny Dataset look-alike used only when xarray is absent."""

    def __init__(self, data_vars, dims, coords=("time", "x", "y"), attrs=None, index=None):
        self.data_vars = dict(data_vars)
        self.dims = tuple(dims)
        self.coord_names = tuple(coords)
        self.attrs = dict(attrs or {})
        first = next(iter(self.data_vars.values()))
        self.sizes = {d: int(s) for d, s in zip(self.dims, first.shape)}
        self.index = dict(index or {})

    def __getitem__(self, name):
        if name in self.data_vars:
            return self.data_vars[name]
        return self.index[name]

    def __getattr__(self, name):
        try:
            return self.__getitem__(name)
        except KeyError:
            raise AttributeError(name) from None

    def __contains__(self, name):
        return name in self.data_vars or name in self.index

    def __repr__(self):
        return f"<RayDataset {self.sizes} vars={list(self.data_vars)}>"


def _dataset(data_vars, dims, index=None):
    attrs = {"date_created": str(datetime.datetime.now())}
    try:
        import xarray as xr
    except ImportError:
        return RayDataset(data_vars, dims, attrs=attrs, index=index)
    ds = xr.Dataset(data_vars={k: (list(dims), v) for k, v in data_vars.items()}, attrs=attrs)
    ds = ds.set_coords(["time", "x", "y"])
    for k, v in (index or {}).items():
        ds[k] = v
    return ds


def single_ray(x0: float, y0: float, kx0: float, ky0: float, duration: float, step_size: float,
               bathymetry: str, current: str):
    """Propagate a single ray.

    Parameters and return value as ``mantaray.single_ray``
    (python/mantaray/core.py:9-65): a Dataset with variables ``time, x, y, kx,
    ky`` over ``time_step``; ``time, x, y`` are coordinates.
    """
    rows = _mantaray.single_ray(x0, y0, kx0, ky0, duration, step_size, str(bathymetry), str(current))
    cols = np.ascontiguousarray(rows.T)
    return _dataset({v: cols[i] for i, v in enumerate(_VARNAMES)}, ("time_step",))


def ray_tracing(x0, y0, kx0, ky0, duration: float, step_size: float, bathymetry: str, current: str, *,
                diagnostics: bool = False):
    """Ray tracing for multiple initial conditions.

    Parameters and return value as ``mantaray.ray_tracing``
    (python/mantaray/core.py:68-132): variables ``time, x, y, kx, ky`` of shape
    ``(time_step, ray)``, every ray NaN-padded to the longest ray.

    ``diagnostics=True`` (extension, off by default) adds the environment along the rays — ``depth``
    (f32), ``u``, ``v``: the columns of the reference's ``Ray`` record, src/datatype.rs:165-194 — and
    what the notebooks derive from it afterwards (notebooks/snells_law_verification.ipynb, cell 9):
    ``k = |(kx, ky)|``, ``theta = atan2(ky, kx)`` and the intrinsic frequency
    ``sigma = sqrt(g k tanh(k depth))`` with the solver's g = 9.8 (src/wave_ray_path.rs:23).
    """
    extra = {"env": True} if diagnostics else {}
    bundle = _mantaray.ray_tracing(x0, y0, kx0, ky0, duration, step_size, str(bathymetry), str(current), **extra)
    r = bundle.result
    longest = int(r.rows.max()) if r.rows.size else 0
    step = np.arange(longest)
    # time of a row is ray-independent; a ray has it only for the rows it stored
    time = np.where(step[:, None] < r.rows[None, :], r.t[:longest, None], np.nan)
    data = {"time": time, "x": r.x[:longest], "y": r.y[:longest], "kx": r.kx[:longest], "ky": r.ky[:longest]}
    if diagnostics:
        kx, ky, depth = data["kx"], data["ky"], r.depth[:longest]
        k = np.hypot(kx, ky)
        with np.errstate(invalid="ignore"):
            sigma = np.sqrt(9.8 * k * np.tanh(k * depth.astype(np.float64)))
        data.update(depth=depth, u=r.u[:longest], v=r.v[:longest], k=k, theta=np.arctan2(ky, kx), sigma=sigma)
    return _dataset(data, ("time_step", "ray"), index={"time_step": step, "ray": np.arange(r.rows.size)})
