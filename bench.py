#!/usr/bin/env python
"""bench.py — ray-steps/s of the batch ray-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C4]

One "step" is one pass of the hot path over one batch: every ray of the workload
integrated for all its RK4 steps (one kernel launch).  The default workload is
C4 (BASELINE.json configs[3]: Agulhas-like eddy current + variable bathymetry on a
2048x2048 f64 grid, 1M rays, 2048 RK4 steps, full trajectory output) — the
"1M-ray current+bathymetry config" the north_star quotes its target on; it fits
one GPU (65.6 GB of trajectories).  With N GPUs every rank traces its own
contiguous block of 1M rays of an N-times denser ensemble (weak scaling; rays are
independent, so there is no data-path collective).

Printed by rank 0: ONE JSON line (see README / DESIGN.md for the keys).
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64_ray_steps_per_sec"
UNIT = "ray-steps/s"


# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL announces its version on the
# first communicator), so file descriptor 1 is pointed at stderr for the whole run and the line is written
# to the saved descriptor.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--rays-per-gpu", type=int, default=0, help="override the per-GPU batch (testing)")
    ap.add_argument("--math", default="fast", choices=["fast", "strict"])
    ap.add_argument("--deep-map", default="auto", choices=["auto", "on", "off"],
                    help="depth-floor map of the fast path (mr_trace_opts.flags): the library's own choice, forced on, forced off")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def make_workload(name: str, world: int, rays_per_gpu: int):
    """The workload at `world` GPUs: per-GPU work fixed (weak scaling)."""
    from mantaray_b200 import workloads as W

    if name == "C1":
        wl = W.c1_canonical((rays_per_gpu or 1000) * world)
    elif name == "C2":
        wl = W.c2_sea_mount((rays_per_gpu or 100_000) * world)
    elif name == "C3":
        wl = W.c3_shear_jet((rays_per_gpu or 1_000_000) * world)
    elif name == "C4":
        if rays_per_gpu:
            side = max(int(math.isqrt(rays_per_gpu)), 1)
            wl = W.c4_agulhas(side * world, side)
        else:
            wl = W.c4_agulhas(1000 * world, 1000)
    else:
        if rays_per_gpu:
            wl = W.c5_nazare(8 * world, 8, max(rays_per_gpu // 64, 1))
        else:
            wl = W.c5_nazare(8 * world, 64, 16_384)      # 8.4M rays per GPU, 64M at 8 GPUs
    return wl


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t_lo, t_hi):
        sm, smax, reasons = [], [], set()
        for t, line in self.samples:
            if not (t_lo <= t <= t_hi):
                continue
            f = [s.strip() for s in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def strided_sample(wl, n_sample: int):
    """A uniformly strided subsample of the workload's rays (same variety as the full batch)."""
    n = wl.n_rays
    n_sample = min(n_sample, n)
    step = max(n // n_sample, 1)
    x0, y0, kx0, ky0 = wl.rays(0, n)
    sel = slice(0, step * n_sample, step)
    return x0[sel], y0[sel], kx0[sel], ky0[sel]


def cpu_leg(wl, seconds_per_step: float = 12.0, steps: int = 1, warmup: int = 0):
    """The oracle (op-for-op C restatement of the reference's Rust path) on all host threads,
    on a bounded strided sample of the workload's rays."""
    from oracle import mr_oracle as O

    O.build()
    cores = os.cpu_count() or 1
    # calibrate: the probe runs twice and the second pass counts (the first pass over a set of rays pays ~1 s of
    # cold misses on the field arrays); it is large enough (~0.1 s or more) that thread start-up does not dominate
    def probe_pass(n_probe):
        rays_p = strided_sample(wl, n_probe)
        t = time.perf_counter()
        r = O.trace_many(wl.bathymetry, wl.current, *rays_p, wl.t0, wl.duration, wl.dt, stride=wl.stride,
                         nthreads=cores, trajectories=False, final_state=False)
        return time.perf_counter() - t, float((r.rows - 1).sum())

    probe_pass(64 * cores)
    t_probe, e_probe = probe_pass(64 * cores)
    rate = max(e_probe / max(t_probe, 1e-6), 1.0)
    n_sample = int(min(wl.n_rays, max(8 * cores, rate * seconds_per_step / max(wl.n_steps, 1))))
    rays = strided_sample(wl, n_sample)
    n_sample = rays[0].size
    times, E = [], 0
    for it in range(warmup + steps):
        t = time.perf_counter()
        r = O.trace_many(wl.bathymetry, wl.current, *rays, wl.t0, wl.duration, wl.dt, stride=wl.stride,
                         nthreads=cores, trajectories=(wl.output == "full"), final_state=True)
        el = time.perf_counter() - t
        if it >= warmup:
            times.append(el)
            E = int((r.rows - 1).sum())
    total = sum(times)
    return {
        "value": E * len(times) / total,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{n_sample} rays (uniform stride over the {wl.n_rays}-ray batch) x {wl.n_steps} RK4 steps, "
                  f"{E} executed ray-steps per pass; C oracle = op-for-op restatement of the reference's Rust path "
                  f"without its per-RHS heap allocations, so faster than the real reference",
        "ms_per_step": 1e3 * total / len(times),
    }


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate
    cannot be built in this image (no cargo/rustc), so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(args.workload, max(args.gpus, 1), args.rays_per_gpu)
    cb = cpu_leg(wl, seconds_per_step=10.0, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "rays": wl.n_rays, "rk4_steps": wl.n_steps,
                   "grid": [wl.bathymetry.x.size, wl.bathymetry.y.size], "stride": wl.stride},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "warmup_passes_run": min(args.warmup, 1),   # a CPU pass takes ~10 s: at most one is spent untimed
    }
    emit(line)


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from mantaray_b200 import CartesianNetcdf3, _abi, _capi
    from mantaray_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mantaray_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lib = _capi.load()
    math_mode = _abi.MR_MATH_STRICT if args.math == "strict" else _abi.MR_MATH_FAST
    wl = make_workload(args.workload, world, args.rays_per_gpu)
    lo, hi = W.shard_range(wl.n_rays, rank, world)
    n = hi - lo
    x0, y0, kx0, ky0 = wl.rays(lo, hi)
    rows_cap = wl.n_rows
    full = wl.output == "full"

    fields = _capi.Fields(wl.bathymetry, wl.current, devices=[local])
    # ---- resident buffers --------------------------------------------------------------
    ic = torch.from_numpy(np.stack([x0, y0, kx0, ky0])).to(dev)
    traj = torch.empty((4, rows_cap, n), dtype=torch.float64, device=dev) if full else None
    d_rows = torch.empty(n, dtype=torch.int32, device=dev)
    d_len = torch.empty(n, dtype=torch.int32, device=dev)
    d_fin = torch.empty((4, n), dtype=torch.float64, device=dev)
    flags = {"auto": 0, "on": _abi.MR_OPT_DEEP_MAP, "off": _abi.MR_OPT_NO_DEEP_MAP}[args.deep_map]
    opts = _abi.TraceOpts(wl.stride, math_mode, 0, flags)
    # what the library does with those flags on this grid (include/mantaray_b200.h): the map exists on affine
    # gridded bathymetry, and the default uses it when a quarter of its blocks are deep for a 10 s wave
    deep_map_used, deep_share = False, None
    if args.math == "fast" and isinstance(wl.bathymetry, CartesianNetcdf3):
        _, deep_share, affine = _capi.depth_floor_map(wl.bathymetry)
        deep_map_used = bool(affine) and args.deep_map != "off" and (args.deep_map == "on" or deep_share >= 0.25)
    stream = torch.cuda.current_stream()
    launches = C.c_int32(0)

    def launch():
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        rc = lib.mr_trace_device(fields.handle, local, C.c_void_p(stream.cuda_stream), n,
                                 p(ic[0]), p(ic[1]), p(ic[2]), p(ic[3]),
                                 wl.t0, wl.duration, wl.dt, C.byref(opts),
                                 p(traj[0]) if full else None, p(traj[1]) if full else None,
                                 p(traj[2]) if full else None, p(traj[3]) if full else None, n,
                                 p(d_rows), p(d_len), p(d_fin), C.byref(launches))
        if rc != 0:
            raise RuntimeError(lib.mr_last_error().decode())
        return launches.value

    for _ in range(max(args.warmup, 3)):
        launch()
    torch.cuda.synchronize()
    E_local = float((d_rows.to(torch.int64) - 1).sum().item())      # executed ray-steps per pass
    E = allsum(E_local)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_lo = time.perf_counter()
    n_launch = 0
    ev[0].record(stream)
    for k in range(args.steps):
        n_launch += launch()
        ev[k + 1].record(stream)
    barrier()
    t_hi = time.perf_counter()
    ms_total = ev[0].elapsed_time(ev[-1])
    ms_each = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    ms_total = allmax(ms_total)
    clocks = sampler.summary(t_lo, t_hi) if sampler else None
    value = E * args.steps / (ms_total * 1e-3)
    n_launch_total = int(allsum(float(n_launch)))          # trace kernels launched in the timed region, all ranks
    kernel_ms = float(np.mean(ms_each))           # one kernel per step: its average launch duration

    # ---- roofline of the dominant (only) kernel ----------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp64_peak = _capi.measure_fp64_peak(local, 200)
    # algorithmic HBM bytes of one launch: every stored row of every ray (NaN padding included,
    # it has to be written too) + initial state, rows/len and final state per ray
    alg_bytes = (float(rows_cap) * n * W.BYTES_PER_ROW if full else 0.0) + n * (32 + 8 + 32)
    alg_flop = E_local * wl.flop_per_ray_step
    ach_tf = alg_flop / (kernel_ms * 1e-3) / 1e12
    ach_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None                                  # DRAM bytes of one launch, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "r1", "traffic.json")) as f:
            tr = json.load(f).get(wl.name)
        if tr and abs(tr["rays_per_gpu"] - n) <= 0.001 * n and tr["rk4_steps"] == wl.n_steps and args.math == "fast":
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "kernel": "mr::trace_kernel<GRID,GRID,%s%s>" % (args.math, ",depth-floor map" if deep_map_used else ""),
        "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak if fp64_peak else None,
        "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write); algorithmic bytes per launch = %d" % int(alg_bytes),
        "traffic_note": ("captured on the kernel without the depth-floor map (profiles/r1/m_*); with the map the writes are "
                         "the same and the record reads can only be fewer") if (traffic and deep_map_used) else None,
        "flop_per_ray_step": wl.flop_per_ray_step,
        "peak_source": "DFMA probe kernel timed in this run (mr_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
        "kernel_ms": kernel_ms,
    }
    roofline_hbm = {
        "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
        "traffic": traffic, "algorithmic_bytes": int(alg_bytes), "bytes_per_stored_row": W.BYTES_PER_ROW, "peak_source": hbm_src,
    }

    # ---- end to end through the C ABI with host buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil

        avail = psutil.virtual_memory().available
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        per_ray = 32.0 * rows_cap if full else 0.0
        budget = min(0.30 * avail / max(local_world, 1), 70e9)
        n_e2e = n if per_ray == 0 else int(min(n, max(budget // per_ray, 1024)))
        n_e2e = max(n_e2e // 128 * 128, min(n, 128))
        sel = slice(0, n_e2e)
        hx0 = _capi.pinned_empty((4, n_e2e))
        hx0[0], hx0[1], hx0[2], hx0[3] = x0[sel], y0[sel], kx0[sel], ky0[sel]
        h_t = np.empty(rows_cap)
        h_traj = _capi.pinned_empty((4, rows_cap, n_e2e)) if full else None
        h_rows = _capi.pinned_empty((n_e2e,), np.int32)
        h_len = _capi.pinned_empty((n_e2e,), np.int32)
        h_fin = _capi.pinned_empty((4, n_e2e))
        pp = lambda a: a.ctypes.data if a is not None else None

        def e2e_call():
            rc = lib.mr_trace_many(fields.handle, n_e2e, pp(hx0[0]), pp(hx0[1]), pp(hx0[2]), pp(hx0[3]),
                                   wl.t0, wl.duration, wl.dt, C.byref(opts), pp(h_t),
                                   pp(h_traj[0]) if full else None, pp(h_traj[1]) if full else None,
                                   pp(h_traj[2]) if full else None, pp(h_traj[3]) if full else None,
                                   pp(h_rows), pp(h_len), pp(h_fin))
            if rc != 0:
                raise RuntimeError(lib.mr_last_error().decode())

        # free the resident trajectories first: the host path allocates its own slabs
        del traj
        torch.cuda.empty_cache()
        e2e_call()                                              # warm-up
        e2e_steps = max(min(args.steps, 3), 1)
        barrier()
        t0_ = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_call()
        barrier()
        el = allmax(time.perf_counter() - t0_)
        E_e2e = allsum(float((h_rows.astype(np.int64) - 1).sum()))
        e2e = {
            "value": E_e2e * e2e_steps / el, "unit": UNIT,
            "h2d_bytes_per_step": int(32 * n_e2e * world),
            "d2h_bytes_per_step": int((per_ray + 8 + 32) * n_e2e * world),
            "rays_per_step": int(n_e2e * world), "steps": e2e_steps, "ms_per_step": 1e3 * el / e2e_steps,
            "api": "mr_trace_many (C ABI, pinned host buffers, H2D of the ray states and D2H of every stored row inside the timed region)",
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_leg(wl)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if sampler:
        sampler.stop()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl.name, "description": wl.description, "rays": wl.n_rays, "rays_per_gpu": n,
                "rk4_steps": wl.n_steps, "grid": [int(wl.bathymetry.x.size), int(wl.bathymetry.y.size)],
                "stride": wl.stride, "output": wl.output, "math": args.math,
                "deep_map": {"flag": args.deep_map, "used": deep_map_used, "deep_share_of_blocks": deep_share},
                "executed_ray_steps_per_pass": E, "parallelism": f"rays sharded x{world}, fields replicated, no collective",
                "l2": "no flush needed: each pass writes %.1f GB of trajectories per GPU, far larger than the 126 MB L2" % (
                    rows_cap * n * 32 / 1e9) if full else "final-state only: inputs (fields %.0f MB) re-read each pass" % (
                    wl.bathymetry.depth.nbytes * 3 / 1e6),
            },
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": n_launch_total, "clocks": clocks,
        }
        emit(line)
    fields.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
