"""Frozen oracle trajectories (tests/golden/oracle_trajectories.npz, made by tests/golden/make_golden.py).

Not reference outputs (the Rust reference cannot run here) but a second, out-of-process anchor: the
oracle must reproduce them on any box (libm differences stay at the 1e-15 level), and the CUDA path is
held to the same file."""

import os

import numpy as np
import pytest

from conftest import assert_parity
from mantaray_b200 import MR_MATH_FAST, MR_MATH_STRICT

HERE = os.path.dirname(os.path.abspath(__file__))


class _Frozen:
    def __init__(self, z, name):
        for k in ("t", "x", "y", "kx", "ky", "rows", "len", "final_state"):
            setattr(self, k, z[f"{name}_{k}"])


def _cases():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CASES


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "oracle_trajectories.npz"))


@pytest.mark.parametrize("name", ["c2", "c4", "c5"])
def test_oracle_reproduces_frozen_trajectories(oracle, golden, name):
    wl = _cases()[name]()
    r = oracle.trace_many(wl.bathymetry, wl.current, *wl.all_rays(), 0.0, wl.duration, wl.dt, stride=wl.stride)
    worst = assert_parity(r, _Frozen(golden, name), rel_tol=1e-12, what=f"oracle vs golden {name}")
    np.testing.assert_array_equal(r.t, golden[f"{name}_t"])
    assert worst <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("math", [MR_MATH_FAST, MR_MATH_STRICT], ids=["fast", "strict"])
@pytest.mark.parametrize("name", ["c2", "c4", "c5"])
def test_cuda_matches_frozen_trajectories(gpu, golden, name, math):
    from mantaray_b200 import Fields, trace_many

    wl = _cases()[name]()
    with Fields(wl.bathymetry, wl.current) as f:
        r = trace_many(f, *wl.all_rays(), 0.0, wl.duration, wl.dt, stride=wl.stride, math=math, final_state=True)
    assert_parity(r, _Frozen(golden, name), what=f"cuda vs golden {name}")
