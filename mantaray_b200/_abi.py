"""ctypes declarations of ``include/mantaray_b200.h``.

Pure declarations: importing this module loads no shared library.  The struct
layouts are shared with the CPU oracle's test wrapper (``oracle/mr_oracle.py``),
which takes the same descriptors.
"""

from __future__ import annotations

import ctypes as C

MR_OK = 0
MR_ERR_IO = -1
MR_ERR_BAD_ARG = -2
MR_ERR_CUDA = -3
MR_ERR_OOM = -4
MR_ERR_FORMAT = -5

MR_BATHY_CONSTANT = 0
MR_BATHY_SLOPE = 1
MR_BATHY_GRID = 2
MR_BATHY_ARRAY = 3

MR_CURRENT_CONSTANT = 0
MR_CURRENT_GRID = 1

MR_MATH_FAST = 0
MR_MATH_STRICT = 1

MR_NC3_MAX_DIMS = 8

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class BathymetryDesc(C.Structure):
    """``mr_bathymetry_desc``"""

    _fields_ = [
        ("kind", C.c_int32),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("x", c_float_p),
        ("y", c_float_p),
        ("depth", c_double_p),
        ("array", c_float_p),
        ("h0", C.c_float),
        ("x0", C.c_float),
        ("y0", C.c_float),
        ("dhdx", C.c_float),
        ("dhdy", C.c_float),
    ]


class CurrentDesc(C.Structure):
    """``mr_current_desc``"""

    _fields_ = [
        ("kind", C.c_int32),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("x", c_double_p),
        ("y", c_double_p),
        ("u", c_double_p),
        ("v", c_double_p),
        ("u0", C.c_double),
        ("v0", C.c_double),
    ]


MR_OPT_DEEP_MAP = 1
MR_OPT_NO_DEEP_MAP = 2
MR_OPT_SAME_GRID = 4
MR_OPT_NO_SAME_GRID = 32
MR_OPT_CURRENT_MAP = 8
MR_OPT_NO_CURRENT_MAP = 16
MR_PLAN_AFFINE, MR_PLAN_DEEP_MAP, MR_PLAN_SAME_GRID, MR_PLAN_CURRENT_MAP = 1, 2, 4, 8


class TraceOpts(C.Structure):
    """``mr_trace_opts``"""

    _fields_ = [
        ("stride", C.c_int32),
        ("math", C.c_int32),
        ("chunk_rays", C.c_int32),
        ("flags", C.c_int32),
    ]


class EnvPlanes(C.Structure):
    """``mr_env_planes``"""

    _fields_ = [("depth", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p)]


#: every symbol the header declares: name -> (restype, argtypes)
SIGNATURES = {
    "mr_abi_version": (C.c_int, []),
    "mr_device_count": (C.c_int, []),
    "mr_last_error": (C.c_char_p, []),
    "mr_fields_create": (
        C.c_int,
        [C.POINTER(BathymetryDesc), C.POINTER(CurrentDesc), C.c_uint32, C.POINTER(C.c_void_p)],
    ),
    "mr_fields_open_netcdf3": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "mr_fields_free": (None, [C.c_void_p]),
    "mr_fields_device_mask": (C.c_uint32, [C.c_void_p]),
    "mr_fields_trim": (None, [C.c_void_p]),
    "mr_num_steps": (C.c_int64, [C.c_double, C.c_double, C.c_double]),
    "mr_num_rows": (C.c_int64, [C.c_double, C.c_double, C.c_double, C.c_int32]),
    "mr_trace_many": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_double, C.c_double, C.c_double, C.POINTER(TraceOpts),
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mr_trace_many_env": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_double, C.c_double, C.c_double, C.POINTER(TraceOpts),
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(EnvPlanes)],
    ),
    "mr_sample_fields": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "mr_sample_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_int32_p],
    ),
    "mr_single_ray": (
        C.c_int,
        [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
         C.c_double, C.c_double, C.c_double, C.POINTER(TraceOpts),
         C.c_void_p, C.c_int64, c_int64_p],
    ),
    "mr_trace_device": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_double, C.c_double, C.c_double, C.POINTER(TraceOpts),
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
         C.c_void_p, C.c_void_p, C.c_void_p, c_int32_p],
    ),
    "mr_nc3_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "mr_nc3_close": (None, [C.c_void_p]),
    "mr_nc3_var_count": (C.c_int, [C.c_void_p]),
    "mr_nc3_var_name": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]),
    "mr_nc3_var_info": (
        C.c_int,
        [C.c_void_p, C.c_char_p, c_int32_p, c_int64_p, c_int32_p, c_int64_p],
    ),
    "mr_nc3_read_f32": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "mr_nc3_read_f64": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "mr_fields_last_split": (C.c_int, [C.c_void_p, c_int64_p, C.c_int32]),
    "mr_trace_plan": (C.c_int, [C.c_void_p, C.POINTER(TraceOpts)]),
    "mr_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "mr_host_free": (None, [C.c_void_p]),
    "mr_uniform_current_map": (C.c_int, [C.POINTER(CurrentDesc), C.c_void_p, C.c_size_t, c_int32_p, c_int32_p,
                                         C.POINTER(C.c_float), c_int32_p]),
    "mr_depth_floor_map": (C.c_int, [C.POINTER(BathymetryDesc), C.c_void_p, C.c_size_t, c_int32_p, c_int32_p,
                                     C.POINTER(C.c_float), c_int32_p]),
}


def declare(lib: C.CDLL) -> C.CDLL:
    """Attach restype/argtypes for every exported symbol; raises if one is missing."""
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    return lib
