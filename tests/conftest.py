import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from mantaray_b200 import _capi

        return _capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def oracle():
    from oracle import mr_oracle

    mr_oracle.build()
    mr_oracle.load()
    return mr_oracle


@pytest.fixture(scope="session")
def gpu():
    from mantaray_b200 import _capi

    assert _capi.device_count() > 0, "no CUDA device: the gpu tests must run on the B200 box"
    return _capi


# ---- parity metric ---------------------------------------------------------------------------
#: BASELINE.json north_star: trajectory relative error <= 1e-9 in position and wavenumber
REL_TOL = 1e-9


def assert_parity(res, ref, rel_tol=REL_TOL, what=""):
    """len / rows bit-exact; x, y, kx, ky within rel_tol of the oracle.

    The error of a position component is taken relative to the ray's position scale
    (max |x|, |y| along its reference trajectory) and that of a wavenumber component
    relative to the ray's max |k|: a component that is legitimately ~0 (e.g. y of a
    ray travelling along x) has no meaningful relative error of its own.
    NaN patterns must coincide exactly.
    """
    np.testing.assert_array_equal(res.rows, ref.rows, err_msg=f"{what}: rows differ")
    np.testing.assert_array_equal(res.len, ref.len, err_msg=f"{what}: len (termination step) differs")
    if ref.x is None or res.x is None:
        return 0.0
    worst = 0.0
    pos_scale = np.nanmax(np.maximum(np.abs(ref.x), np.abs(ref.y)), axis=0, initial=0.0)
    k_scale = np.nanmax(np.hypot(ref.kx, ref.ky), axis=0, initial=0.0)
    pos_scale = np.where(pos_scale > 0, pos_scale, 1.0)
    k_scale = np.where(k_scale > 0, k_scale, 1.0)
    for name, scale in (("x", pos_scale), ("y", pos_scale), ("kx", k_scale), ("ky", k_scale)):
        a, b = getattr(res, name), getattr(ref, name)
        assert a.shape == b.shape, f"{what}: {name} shape {a.shape} vs {b.shape}"
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f"{what}: NaN pattern of {name} differs")
        with np.errstate(invalid="ignore"):
            err = np.abs(a - b) / scale[None, :]
        e = float(np.nanmax(err, initial=0.0))
        assert e <= rel_tol, f"{what}: {name} relative error {e:.3e} > {rel_tol:g}"
        worst = max(worst, e)
    return worst
