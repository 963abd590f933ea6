"""Time the environment kernel (mr_sample_device) on the planes of a C4-shaped trace, CUDA events on the
launching stream.  Usage: python tools/envbench.py [--rays N] [--steps S]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mantaray_b200 import Fields, _capi
from mantaray_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=2048)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lib", default=None)
a = ap.parse_args()
side = int(round(a.rays ** 0.5))
wl = W.c4_agulhas(side, side, a.steps)
n, rows = wl.n_rays, wl.n_rows
if a.lib:
    _capi.lib_path = lambda: os.path.abspath(a.lib)
lib = _capi.load()
dev = torch.device("cuda:0")
ic = [torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev) for v in wl.all_rays()]
traj = torch.empty((4, rows, n), dtype=torch.float64, device=dev)
depth = torch.empty((rows, n), dtype=torch.float32, device=dev)
u = torch.empty((rows, n), dtype=torch.float64, device=dev)
v = torch.empty((rows, n), dtype=torch.float64, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
st = torch.cuda.current_stream()
s = C.c_void_p(st.cuda_stream)
with Fields(wl.bathymetry, wl.current, devices=[0]) as f:
    rc = lib.mr_trace_device(f.handle, 0, s, n, p(ic[0]), p(ic[1]), p(ic[2]), p(ic[3]), wl.t0, wl.duration, wl.dt, None,
                             p(traj[0]), p(traj[1]), p(traj[2]), p(traj[3]), n, None, None, None, None)
    assert rc == 0, lib.mr_last_error()
    for what, args in (("depth+u+v", (p(depth), p(u), p(v))), ("depth", (p(depth), None, None)), ("u+v", (None, p(u), p(v)))):
        ms = []
        for _ in range(a.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            rc = lib.mr_sample_device(f.handle, 0, s, rows, n, n, p(traj[0]), p(traj[1]), *args, None)
            assert rc == 0, lib.mr_last_error()
            e1.record(st)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = float(np.median(ms[1:]))
        out_b = (4 if args[0] else 0) + (8 if args[1] else 0) + (8 if args[2] else 0)
        gb = rows * n * (16 + out_b) / 1e9
        print(json.dumps({"lib": os.path.basename(a.lib or "default"), "planes": what, "rows": rows, "rays": n, "ms": t, "alg_GB": gb, "GB_per_s": gb / (t * 1e-3),
                          "points_per_s": rows * n / (t * 1e-3)}))
