"""Diagnostic: the C5 rays of bench.py --inproc's parity sample (full 64-period ensemble, strided), traced with the
fast path, the strict path and the oracle; prints where the final states differ most."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mantaray_b200 import Fields, trace_many, MR_MATH_FAST, MR_MATH_STRICT, workloads as W
from oracle import mr_oracle as O

O.build()
wl = W.c5_nazare(64, 64, 16384)
n, tot = wl.n_rays, 8388608
step = n // tot
x0, y0, kx0, ky0 = wl.rays(0, n)
sel0 = slice(0, step * tot, step)
x0, y0, kx0, ky0 = x0[sel0], y0[sel0], kx0[sel0], ky0[sel0]
s2 = np.unique(np.linspace(0, tot - 1, 1464).astype(np.int64))
r = (x0[s2], y0[s2], kx0[s2], ky0[s2])
ref = O.trace_many(wl.bathymetry, wl.current, *r, 0.0, wl.duration, wl.dt, stride=64, trajectories=False)


def report(name, res):
    f, g = ref.final_state, res.final_state
    pos = np.maximum(np.abs(f[0]), np.abs(f[1])); ksc = np.hypot(f[2], f[3])
    errs = np.array([np.abs(f[0] - g[0]) / pos, np.abs(f[1] - g[1]) / pos, np.abs(f[2] - g[2]) / ksc, np.abs(f[3] - g[3]) / ksc])
    print(name, "rows equal", np.array_equal(res.rows, ref.rows), "len equal", np.array_equal(res.len, ref.len),
          "max err per component", np.nanmax(errs, axis=1))
    worst = np.argsort(-np.nanmax(errs, axis=0))[:6]
    for i in worst:
        print("   ray", int(s2[i]), "rows", int(ref.rows[i]), "len", int(ref.len[i]), "errs", errs[:, i], "final ref", f[:, i], "got", g[:, i])


with Fields(wl.bathymetry, wl.current, devices=[0]) as fld:
    for name, kw in (("fast", dict(math=MR_MATH_FAST)), ("strict", dict(math=MR_MATH_STRICT)), ("fast+same-grid", dict(math=MR_MATH_FAST, flags=4))):
        res = trace_many(fld, *r, 0.0, wl.duration, wl.dt, stride=64, trajectories=False, final_state=True, **kw)
        report(name, res)
