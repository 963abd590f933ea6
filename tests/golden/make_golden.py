#!/usr/bin/env python
"""Regenerates tests/golden/oracle_trajectories.npz.

The reference (Rust) cannot be run in this image, so these are NOT reference outputs: they are the
CPU oracle's trajectories on small fixed inputs, frozen so that (a) a change in the oracle, the
compiler flags or the host libm that moves its results is noticed, and (b) the CUDA path is also
compared against numbers that were not produced in the same process.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mantaray_b200 import workloads as W  # noqa: E402
from oracle import mr_oracle as O  # noqa: E402

CASES = {
    "c2": lambda: W.c2_sea_mount(24, 400, half=80),
    "c4": lambda: W.c4_agulhas(5, 5, 256, nx=128),
    "c5": lambda: W.c5_nazare(2, 3, 4, 512, nx=256, stride=8),
}


def main():
    out = {}
    for name, make in CASES.items():
        wl = make()
        r = O.trace_many(wl.bathymetry, wl.current, *wl.all_rays(), 0.0, wl.duration, wl.dt, stride=wl.stride, nthreads=1)
        for k in ("t", "x", "y", "kx", "ky", "rows", "len", "final_state"):
            out[f"{name}_{k}"] = getattr(r, k)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_trajectories.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
