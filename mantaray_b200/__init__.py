"""mantaray_b200 — B200-native batch ray tracing behind mantaray's API.

``single_ray`` and ``ray_tracing`` have the signatures of
``mantaray.core`` (python/mantaray/core.py); the field types and the batch
driver (``ManyRays`` / ``SingleRay``) mirror the Rust crate's.  The integration
runs as hand-written sm_100a CUDA kernels behind the C ABI of
``include/mantaray_b200.h``; importing the package does not load the shared
library, calling into it does (and fails loudly when it is missing).
"""

from ._abi import (MR_MATH_FAST, MR_MATH_STRICT, MR_OPT_CURRENT_MAP, MR_OPT_DEEP_MAP, MR_OPT_NO_CURRENT_MAP,
                   MR_OPT_NO_DEEP_MAP, MR_OPT_NO_SAME_GRID, MR_OPT_SAME_GRID, MR_PLAN_AFFINE, MR_PLAN_CURRENT_MAP,
                   MR_PLAN_DEEP_MAP, MR_PLAN_SAME_GRID)
from ._capi import (Fields, ManyRays, MantarayError, RayState, SingleRay, TraceResult, depth_floor_map, sample_fields, uniform_current_map,
                    trace_many)
from ._mantaray import cache_info, clear_cache
from .core import ray_tracing, single_ray
from .fields import (ArrayDepth, CartesianCurrent, CartesianNetcdf3, ConstantCurrent, ConstantDepth,
                     ConstantSlope)

__all__ = [
    "single_ray", "ray_tracing", "clear_cache", "cache_info",
    "ManyRays", "SingleRay", "RayState", "Fields", "TraceResult", "trace_many", "sample_fields", "depth_floor_map", "uniform_current_map",
    "MantarayError",
    "ConstantDepth", "ConstantSlope", "CartesianNetcdf3", "ArrayDepth", "ConstantCurrent", "CartesianCurrent",
    "MR_MATH_FAST", "MR_MATH_STRICT", "MR_OPT_DEEP_MAP", "MR_OPT_NO_DEEP_MAP", "MR_OPT_SAME_GRID", "MR_OPT_NO_SAME_GRID", "MR_PLAN_AFFINE", "MR_PLAN_DEEP_MAP", "MR_PLAN_SAME_GRID", "MR_PLAN_CURRENT_MAP", "MR_OPT_CURRENT_MAP", "MR_OPT_NO_CURRENT_MAP",
]
