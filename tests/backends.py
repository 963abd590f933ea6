"""Back-ends the ported reference tests run against.

``oracle``      the CPU restatement (runs everywhere)
``cuda-fast``   the product path, MR_MATH_FAST   (GPU box only)
``cuda-strict`` the product path, MR_MATH_STRICT (GPU box only)

Each back-end offers the reference's internal seams as plain functions:
``single(bathy, current, (x, y, kx, ky), t0, t1, dt) -> (t, states)`` like
``SingleRay::trace_individual(..).get()`` (src/ray.rs:198-213) and
``many(bathy, current, rays, t0, t1, dt) -> [(t, states), ...]`` like
``ManyRays::trace_many`` (src/ray.rs:98-127).
"""

import numpy as np
import pytest


class OracleBackend:
    name = "oracle"

    def single(self, bathy, current, ray, t0, t1, dt):
        from oracle import mr_oracle as O

        out = O.single_ray(bathy, current, *ray, t0, t1, dt)
        return out[:, 0].copy(), out[:, 1:].copy()

    def many(self, bathy, current, rays, t0, t1, dt):
        from oracle import mr_oracle as O

        r = np.asarray(rays, dtype=np.float64).reshape(-1, 4)
        res = O.trace_many(bathy, current, r[:, 0], r[:, 1], r[:, 2], r[:, 3], t0, t1, dt)
        out = []
        for i in range(r.shape[0]):
            m = int(res.rows[i])
            out.append((res.t[:m].copy(), np.stack([res.x[:m, i], res.y[:m, i], res.kx[:m, i], res.ky[:m, i]], axis=1)))
        return out


class CudaBackend:
    def __init__(self, math, name):
        self.math = math
        self.name = name

    def single(self, bathy, current, ray, t0, t1, dt):
        from mantaray_b200 import RayState, SingleRay

        return SingleRay(bathy, current, RayState(*ray)).trace_individual(t0, t1, dt, math=self.math)

    def many(self, bathy, current, rays, t0, t1, dt):
        from mantaray_b200 import ManyRays, RayState

        return ManyRays(bathy, current, [RayState(*r) for r in rays]).trace_many(t0, t1, dt, math=self.math)


def backend_params():
    from mantaray_b200 import MR_MATH_FAST, MR_MATH_STRICT

    return [
        pytest.param(OracleBackend(), id="oracle"),
        pytest.param(CudaBackend(MR_MATH_FAST, "cuda-fast"), id="cuda-fast", marks=pytest.mark.gpu),
        pytest.param(CudaBackend(MR_MATH_STRICT, "cuda-strict"), id="cuda-strict", marks=pytest.mark.gpu),
    ]


# ---- the helpers of src/tests/helper/mod.rs:13-47 (rows whose x is NaN are skipped) --------
def _finite_rows(data):
    return data[~np.isnan(data[:, 0])]


def increase(data, index):
    d = _finite_rows(data)[:, index]
    first = data[0, index]
    seq = np.concatenate([[first], d[1:]])
    return bool(np.all(np.diff(seq) > 0))


def decrease(data, index):
    d = _finite_rows(data)[:, index]
    first = data[0, index]
    seq = np.concatenate([[first], d[1:]])
    return bool(np.all(np.diff(seq) < 0))


def same(data, index):
    d = _finite_rows(data)[:, index]
    first = data[0, index]
    seq = np.concatenate([[first], d[1:]])
    return bool(np.all(seq == first))


X, Y, KX, KY = 0, 1, 2, 3
