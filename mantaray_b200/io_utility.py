"""NetCDF-3 fixture writers.

``create_netcdf3_bathymetry`` / ``create_netcdf3_current`` produce the files the
reference's tests synthesise with src/io/utility.rs:29-90 and :108-181:
dimensions ``y`` then ``x``; ``x[i] = f32(i) * x_step`` and ``y`` likewise as
f32 variables; data variables f64 with dims ``[y][x]``; classic format.

The writer is a small self-contained NetCDF-3 classic encoder so that every
external type {i8, char/u8, i16, i32, f32, f64} can be written for the dtype
tests (src/current/cartesian_current.rs:645-657).
"""

from __future__ import annotations

import struct
from typing import Callable, Dict, Sequence, Tuple

import numpy as np

_NC_TYPES = {"i1": 1, "S1": 2, "u1": 2, "i2": 3, "i4": 4, "f4": 5, "f8": 6}


def _name(s: str) -> bytes:
    b = s.encode()
    return struct.pack(">I", len(b)) + b + b"\0" * (-len(b) % 4)


def write_netcdf3(path, dims: Sequence[Tuple[str, int]], variables: Dict[str, Tuple[Sequence[str], np.ndarray]],
                  version: int = 1) -> None:
    """Write fixed-size variables.  ``variables[name] = (dim_names, array)``; the
    array's dtype picks the external type."""
    dim_ids = {n: i for i, (n, _) in enumerate(dims)}
    header = b"CDF" + bytes([version]) + struct.pack(">I", 0)
    header += struct.pack(">II", 0x0A, len(dims)) if dims else struct.pack(">II", 0, 0)
    for n, l in dims:
        header += _name(n) + struct.pack(">I", l)
    header += struct.pack(">II", 0, 0)                      # no global attributes
    header += struct.pack(">II", 0x0B, len(variables))
    entries = []
    for name, (vdims, arr) in variables.items():
        arr = np.asarray(arr)
        code = arr.dtype.str[1:]
        nc_type = _NC_TYPES[code]
        raw = arr.astype(arr.dtype.newbyteorder(">")).tobytes()
        raw += b"\0" * (-len(raw) % 4)
        e = _name(name) + struct.pack(">I", len(vdims)) + b"".join(struct.pack(">I", dim_ids[d]) for d in vdims)
        e += struct.pack(">II", 0, 0)                       # no attributes
        e += struct.pack(">II", nc_type, len(raw))
        entries.append((e, raw))
    off_size = 8 if version == 2 else 4
    total_header = len(header) + sum(len(e) + off_size for e, _ in entries)
    begin = total_header
    blob = b""
    for e, raw in entries:
        header += e + (struct.pack(">Q", begin) if version == 2 else struct.pack(">I", begin))
        blob += raw
        begin += len(raw)
    with open(path, "wb") as f:
        f.write(header + blob)


def create_netcdf3_bathymetry(path, x_num: int, y_num: int, x_step: float, y_step: float,
                              depth_fn: Callable[[np.float32, np.float32], float]) -> None:
    """src/io/utility.rs:29-90."""
    x = (np.arange(x_num, dtype=np.float32) * np.float32(x_step)).astype(np.float32)
    y = (np.arange(y_num, dtype=np.float32) * np.float32(y_step)).astype(np.float32)
    depth = np.array([[depth_fn(xi, yi) for xi in x] for yi in y], dtype=np.float64)
    write_netcdf3(path, [("y", y_num), ("x", x_num)],
                  {"y": (["y"], y), "x": (["x"], x), "depth": (["y", "x"], depth)})


def create_netcdf3_current(path, x_num: int, y_num: int, x_step: float, y_step: float,
                           current_fn: Callable[[np.float32, np.float32], Tuple[float, float]]) -> None:
    """src/io/utility.rs:108-181."""
    x = (np.arange(x_num, dtype=np.float32) * np.float32(x_step)).astype(np.float32)
    y = (np.arange(y_num, dtype=np.float32) * np.float32(y_step)).astype(np.float32)
    uv = np.array([[current_fn(xi, yi) for xi in x] for yi in y], dtype=np.float64)
    write_netcdf3(path, [("y", y_num), ("x", x_num)],
                  {"y": (["y"], y), "x": (["x"], x), "u": (["y", "x"], uv[..., 0]), "v": (["y", "x"], uv[..., 1])})
