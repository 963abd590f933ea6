#!/usr/bin/env python
"""Kernel A/B harness: time the C4 trace kernel of one or more builds of the library.

    python tools/kbench.py [--rays 262144] [--steps 1024] lib1.so lib2.so ...

Every library is loaded in its own subprocess (the ABI is identical), runs the same
device-resident launch 1 + 3 times, and prints the best CUDA-event time and a checksum of
rows / final states so that variants can be compared for equality.
"""
import argparse, ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(lib_path, rays, steps, workload, math, notraj=False, nx=2048, flags=0, order="start"):
    import numpy as np, torch
    from mantaray_b200 import _abi, _capi, workloads as W
    _capi.lib_path = lambda: lib_path
    lib = _capi.load()
    side = int(rays ** 0.5)
    wl = {"C4": lambda: W.c4_agulhas(side, side, steps, nx), "C2": lambda: W.c2_sea_mount(rays, steps),
          "C3": lambda: W.c3_shear_jet(rays, steps), "C5": lambda: W.c5_nazare(8, 8, rays // 64, steps)}[workload]()
    x0, y0, kx0, ky0 = wl.all_rays()
    n = x0.size
    if order != "start":
        # The same rays in another order (C4: ray = start point * side + direction).  Which rays share a warp, a
        # block and a wave of resident blocks decides how often a cell record is fetched again;
        # rows_sum / len_sum / fin_checksum are sums over rays and must not move.
        assert workload == "C4", "--order is defined for C4's (start point, direction) lattice"
        s_, d_ = np.divmod(np.arange(n), side)
        if order == "dir":                       # direction-major: a warp is 32 neighbouring start points, one direction
            key = d_ * side + s_
        elif order == "tile":                    # a block is 128 neighbouring directions of one start point, blocks run
            key = ((d_ // 128) * side + s_) * 128 + d_ % 128       # across all start points before the next 128 directions
        elif order == "tile32":                  # the same with warp-sized direction groups
            key = ((d_ // 32) * side + s_) * 32 + d_ % 32
        else:
            raise SystemExit(f"unknown --order {order}")
        perm = np.argsort(key, kind="stable")
        x0, y0, kx0, ky0 = x0[perm], y0[perm], kx0[perm], ky0[perm]
    if os.environ.get("KB_ANALYTIC"):      # same rays, analytic fields: no record loads at all
        from mantaray_b200 import ConstantDepth, ConstantCurrent, ConstantSlope
        mode = os.environ["KB_ANALYTIC"]
        wl.bathymetry = ConstantDepth(4000.0) if mode in ("deep", "deepcur") else ConstantSlope(300.0, 0.0, 0.0, -1e-4, -1e-4)
        if mode in ("deep", "slope"):
            wl.current = ConstantCurrent(0.1, -0.05)
    if os.environ.get("KB_SHIFT"):         # same fields on coordinates with a non-representable origin:
        from mantaray_b200 import CartesianNetcdf3, CartesianCurrent   # the f32 grids are then not affine
        sh = float(os.environ["KB_SHIFT"])
        b, c = wl.bathymetry, wl.current
        wl.bathymetry = CartesianNetcdf3(np.asarray(b.x, np.float64) + sh, np.asarray(b.y, np.float64) + sh, b.depth)
        wl.current = CartesianCurrent(c.x + sh, c.y + sh, c.u, c.v)
        x0, y0 = x0 + sh, y0 + sh
    dev = torch.device("cuda", 0)
    f = _capi.Fields(wl.bathymetry, wl.current, devices=[0])
    ic = torch.from_numpy(np.stack([x0, y0, kx0, ky0])).to(dev)
    rows_cap = wl.n_rows
    traj = torch.empty((4, rows_cap, n), dtype=torch.float64, device=dev) if not notraj else None
    d_rows = torch.empty(n, dtype=torch.int32, device=dev)
    d_len = torch.empty(n, dtype=torch.int32, device=dev)
    d_fin = torch.empty((4, n), dtype=torch.float64, device=dev)
    opts = _abi.TraceOpts(wl.stride, math, 0, flags)
    st = torch.cuda.current_stream()
    p = lambda t: C.c_void_p(t.data_ptr())
    tp = (lambda i: p(traj[i])) if traj is not None else (lambda i: None)
    def launch():
        rc = lib.mr_trace_device(f.handle, 0, C.c_void_p(st.cuda_stream), n, p(ic[0]), p(ic[1]), p(ic[2]), p(ic[3]),
                                 0.0, wl.duration, wl.dt, C.byref(opts), tp(0), tp(1), tp(2), tp(3), n,
                                 p(d_rows), p(d_len), p(d_fin), None)
        assert rc == 0, lib.mr_last_error()
    launch(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    E = int((d_rows.to(torch.int64) - 1).sum().item())
    fin = d_fin.cpu().numpy()
    print(json.dumps({"lib": os.path.basename(lib_path), "flags": flags, "order": order, "workload": workload, "notraj": notraj, "nx": nx, "ms": best, "ray_steps_per_s": E / best * 1e3, "E": E,
                      "rows_sum": int(d_rows.sum().item()), "len_sum": int(d_len.sum().item()),
                      "fin_checksum": float(np.nansum(np.abs(fin[:2])) + 1e6 * np.nansum(np.abs(fin[2:])))}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--math", type=int, default=0)
    ap.add_argument("--child", default=None)
    ap.add_argument("--notraj", action="store_true")
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--flags", type=int, default=0, help="mr_trace_opts.flags: 0 = library default, 1 / 2 = depth-floor map on / off, 4 = same-grid shortcut, 8 / 16 = uniform-current map on / off")
    ap.add_argument("--order", default="start", choices=["start", "dir", "tile", "tile32"],
                    help="C4 only: order of the same rays (start-point-major as benchmarked, direction-major, or tiled)")
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        child(a.child, a.rays, a.steps, a.workload, a.math, a.notraj, a.nx, a.flags, a.order)
    else:
        for lib in a.libs or [os.path.join(ROOT, "mantaray_b200", "libmantaray_b200.so")]:
            subprocess.run([sys.executable, __file__, "--child", os.path.abspath(lib), "--rays", str(a.rays),
                            "--steps", str(a.steps), "--workload", a.workload, "--math", str(a.math), "--nx", str(a.nx), "--flags", str(a.flags), "--order", a.order] + (["--notraj"] if a.notraj else []), check=False)
