#!/usr/bin/env python
"""bench.py — ray-steps/s of the batch ray-tracing hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C4]
                    [--shard block|interleave] [--no-extra] [--inproc]

One "step" is one pass of the hot path over one batch: every ray of the workload
integrated for all its RK4 steps (one kernel launch).  The default workload is
C4 (BASELINE.json configs[3]: Agulhas-like eddy current + variable bathymetry on a
2048x2048 f64 grid, 1M rays, 2048 RK4 steps, full trajectory output) — the
"1M-ray current+bathymetry config" the north_star quotes its target on; it fits
one GPU (65.6 GB of trajectories).  With N GPUs every rank traces its own share
of an N-times larger ensemble (weak scaling; rays are independent, so there is
no data-path collective).

After the headline workload the other named shapes (C1, C2, C3 and the C5 shard)
run at their BASELINE.json sizes, a few passes each, and are reported in
"workloads"; every workload's TIMED output buffers are then checked against the
CPU oracle on a strided sample of rays ("parity").

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
PARITY_TOL = 1e-9                  # BASELINE.json north_star: relative error in position and wavenumber
FP64_FLOP_PER_CLK_PER_SM = 128     # 64 DFMA lanes x 2 flop


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
    ap.add_argument("--same-grid", default="auto", choices=["auto", "off", "on"],
                    help="same-grid shortcut of the fast path (auto: the library's rule — where no map is in use; MR_OPT_[NO_]SAME_GRID)")
    ap.add_argument("--shard", default="auto", choices=["auto", "block", "interleave"],
                    help="how the ensemble is shared out over the ranks: one contiguous block each, or tiles dealt round-robin "
                         "(auto: interleave for C5, whose period bands differ in work; block otherwise)")
    ap.add_argument("--no-extra", action="store_true", help="skip the other named workloads")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--inproc", action="store_true",
                    help="ONE process driving --gpus devices through one field handle (mr_trace_many with host buffers): "
                         "the library's own multi-device path, what mantaray.ray_tracing takes on this box")
    ap.add_argument("--total-rays", type=int, default=0, help="--inproc: rays of the batch (0 = what the host RAM holds)")
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
    # the whole run is bounded to about a minute and a half of CPU passes whatever --steps is: ~10 s per pass for a
    # few steps, shorter passes (never below 2 s: 16 threads need that to amortise their start) for many
    steps = max(args.steps, 1)
    cb = cpu_leg(wl, seconds_per_step=min(10.0, max(2.0, 75.0 / steps)), steps=steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "rays": wl.n_rays,
                   "rays_per_gpu": wl.n_rays // max(args.gpus, 1), "rk4_steps": wl.n_steps,
                   "grid": [int(wl.bathymetry.x.size), int(wl.bathymetry.y.size)], "stride": wl.stride, "output": wl.output},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "warmup_passes_run": min(args.warmup, 1),   # a CPU pass takes ~10 s: at most one is spent untimed
    }
    emit(line)


# ---- parity of what was timed ------------------------------------------------------------------------------------
def parity_of(wl, rays, got, cores: int):
    """The oracle on `rays` against `got`, the same rays' columns pulled out of the buffers the timed launches
    wrote: rows / len bit-exact, trajectories (or final states) to PARITY_TOL, position errors relative to the ray's
    position scale and wavenumber errors to its max |k| (the metric of tests/conftest.py)."""
    from oracle import mr_oracle as O

    full = got.get("x") is not None
    ref = O.trace_many(wl.bathymetry, wl.current, *rays, wl.t0, wl.duration, wl.dt, stride=wl.stride, nthreads=cores,
                       trajectories=full, final_state=True)
    equal = bool(np.array_equal(ref.rows, got["rows"]) and np.array_equal(ref.len, got["len"]))
    worst, nan_equal = 0.0, True
    with np.errstate(invalid="ignore"):
        if full:
            pos = np.nanmax(np.maximum(np.abs(ref.x), np.abs(ref.y)), axis=0, initial=0.0)
            ksc = np.nanmax(np.hypot(ref.kx, ref.ky), axis=0, initial=0.0)
            pos, ksc = np.where(pos > 0, pos, 1.0), np.where(ksc > 0, ksc, 1.0)
            for name, sc in (("x", pos), ("y", pos), ("kx", ksc), ("ky", ksc)):
                a, b = got[name], getattr(ref, name)
                nan_equal = nan_equal and bool(np.array_equal(np.isnan(a), np.isnan(b)))
                worst = max(worst, float(np.nanmax(np.abs(a - b) / sc[None, :], initial=0.0)))
        f, g = ref.final_state, got["fin"]
        nan_equal = nan_equal and bool(np.array_equal(np.isnan(f), np.isnan(g)))
        pos = np.maximum(np.maximum(np.abs(f[0]), np.abs(f[1])), 1e-300)
        ksc = np.maximum(np.hypot(f[2], f[3]), 1e-300)
        for c, sc in ((0, pos), (1, pos), (2, ksc), (3, ksc)):
            worst = max(worst, float(np.nanmax(np.abs(f[c] - g[c]) / sc, initial=0.0)))
    return {"rays": int(rays[0].size), "rows_len_equal": equal, "nan_pattern_equal": nan_equal, "max_rel_err": worst}


def parity_sample_size(wl, n, full, rows_cap, share=1):
    budget = (6e6 if wl.flop_per_ray_step > 500 else 2e6) / max(share, 1)       # oracle ray-steps
    n_par = int(min(n, max(64, budget // max(wl.n_steps, 1))))
    if full:
        n_par = int(min(n_par, max(64, 2.5e8 // (32 * rows_cap))))             # host copy of the sampled columns
    return n_par


def shard_ids(wl, name, rank, world, mode):
    """The rays of this rank as a list of (lo, hi) index ranges of the ensemble."""
    from mantaray_b200 import workloads as W

    if mode == "auto":
        mode = "interleave" if name == "C5" else "block"
    if world == 1 or mode == "block":
        return [W.shard_range(wl.n_rays, rank, world)], "block"
    return W.shard_tiles(wl.n_rays, rank, world, wl.extra.get("tile", 16_384)), "interleave"


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from mantaray_b200 import CartesianCurrent, CartesianNetcdf3, _abi, _capi
    from mantaray_b200 import workloads as W
    from tools import mrtools

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mantaray_b200 has no CPU fallback")
    if args.inproc:
        if world > 1:
            raise SystemExit("--inproc is ONE process driving all the devices: run it without torchrun")
        return run_inproc(args)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allred(v: float, op) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    allmax = lambda v: allred(v, dist.ReduceOp.MAX)
    allmin = lambda v: allred(v, dist.ReduceOp.MIN)
    allsum = lambda v: allred(v, dist.ReduceOp.SUM)

    def allgather(v: float):
        if world == 1:
            return [v]
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    lib = _capi.load()
    math_mode = _abi.MR_MATH_STRICT if args.math == "strict" else _abi.MR_MATH_FAST
    cores = max((os.cpu_count() or 1) // max(local_world, 1), 1)
    stream = torch.cuda.current_stream()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp64_peak = mrtools.measure_fp64_peak(local, 200)
    hw_counters = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r2", "hw_counters.json")) as f:
            hw_counters = json.load(f)
    except Exception:
        pass
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    flush_buf = [None]
    flags = {"auto": 0, "on": _abi.MR_OPT_DEEP_MAP, "off": _abi.MR_OPT_NO_DEEP_MAP}[args.deep_map]
    flags |= {"auto": 0, "on": _abi.MR_OPT_SAME_GRID, "off": _abi.MR_OPT_NO_SAME_GRID}[args.same_grid]

    def run_workload(name, steps, warmup, rays_per_gpu, sampler=None):
        """Times `steps` device-resident passes of workload `name` on this rank's share, after `warmup` untimed ones.
        Returns (summary dict, state for the end-to-end leg)."""
        wl = make_workload(name, world, rays_per_gpu)
        ranges, shard_mode = shard_ids(wl, name, rank, world, args.shard)
        parts = [wl.rays(lo, hi) for lo, hi in ranges]
        x0, y0, kx0, ky0 = (np.concatenate([p[i] for p in parts]) for i in range(4))
        n = x0.size
        rows_cap = wl.n_rows
        full = wl.output == "full"
        fields = _capi.Fields(wl.bathymetry, wl.current, devices=[local])
        ic = torch.from_numpy(np.stack([x0, y0, kx0, ky0])).to(dev)
        traj = torch.empty((4, rows_cap, n), dtype=torch.float64, device=dev) if full else None
        d_rows = torch.empty(n, dtype=torch.int32, device=dev)
        d_len = torch.empty(n, dtype=torch.int32, device=dev)
        d_fin = torch.empty((4, n), dtype=torch.float64, device=dev)
        opts = _abi.TraceOpts(wl.stride, math_mode, 0, flags)
        # what the library does with those flags on this grid (mr_trace_plan; include/mantaray_b200.h): the depth-floor
        # map when a quarter of the bathymetry's blocks are deep for a 10 s wave, the uniform-current map when half of
        # the current's blocks are uniform, the same-grid shortcut where the grids coincide and neither map is in use
        plan = fields.plan(math_mode, flags)
        deep_map_used, cur_map_used = bool(plan & _abi.MR_PLAN_DEEP_MAP), bool(plan & _abi.MR_PLAN_CURRENT_MAP)
        same_grid_used = bool(plan & _abi.MR_PLAN_SAME_GRID)
        deep_share, cur_share = None, None
        if args.math == "fast" and isinstance(wl.bathymetry, CartesianNetcdf3):
            _, deep_share, _ = _capi.depth_floor_map(wl.bathymetry)
        if args.math == "fast" and isinstance(wl.current, CartesianCurrent):
            _, cur_share, _ = _capi.uniform_current_map(wl.current)
        launches = C.c_int32(0)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None

        def launch():
            rc = lib.mr_trace_device(fields.handle, local, C.c_void_p(stream.cuda_stream), n,
                                     p(ic[0]), p(ic[1]), p(ic[2]), p(ic[3]),
                                     wl.t0, wl.duration, wl.dt, C.byref(opts),
                                     p(traj[0]) if full else None, p(traj[1]) if full else None,
                                     p(traj[2]) if full else None, p(traj[3]) if full else None, n,
                                     p(d_rows), p(d_len), p(d_fin), C.byref(launches))
            if rc != 0:
                raise RuntimeError(lib.mr_last_error().decode())
            return launches.value

        # L2: a pass that writes far more than the 126 MB L2 evicts everything by itself; the others get an explicit
        # flush (a 256 MB buffer written) before every timed pass, outside the timed interval
        out_bytes = (float(rows_cap) * n * W.BYTES_PER_ROW if full else 0.0) + n * 72.0
        need_flush = out_bytes < 1e9
        if need_flush and flush_buf[0] is None:
            flush_buf[0] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(max(warmup, 3)):
            launch()
        torch.cuda.synchronize()
        E_local = float((d_rows.to(torch.int64) - 1).sum().item())      # executed ray-steps per pass
        E = allsum(E_local)
        if sampler:
            sampler.start()
            time.sleep(0.3)
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        t_lo = time.perf_counter()
        n_launch = 0
        for k in range(steps):
            if need_flush:
                flush_buf[0].zero_()
            ev0[k].record(stream)
            n_launch += launch()
            ev1[k].record(stream)
        barrier()
        t_hi = time.perf_counter()
        ms_each = [ev0[k].elapsed_time(ev1[k]) for k in range(steps)]
        # the timed region: back to back from the first event to the last when nothing is flushed in between,
        # else the sum of the per-pass intervals
        ms_local = sum(ms_each) if need_flush else ev0[0].elapsed_time(ev1[-1])
        per_rank_ms = [m / steps for m in allgather(ms_local)]
        ms_total = max(per_rank_ms) * steps
        value = E * steps / (ms_total * 1e-3)
        kernel_ms = float(np.mean(ms_each))           # one kernel per step: its average launch duration
        clocks = sampler.summary(t_lo, t_hi) if sampler else None

        # ---- roofline of the dominant (only) kernel ----------------------------------------
        # algorithmic HBM bytes of one launch: every stored row of every ray (NaN padding included,
        # it has to be written too) + initial state, rows/len and final state per ray
        alg_bytes = (float(rows_cap) * n * W.BYTES_PER_ROW if full else 0.0) + n * (32 + 8 + 32)
        alg_flop = E_local * wl.flop_per_ray_step
        ach_tf = alg_flop / (kernel_ms * 1e-3) / 1e12
        ach_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
        variant = [args.math] + (["depth-floor map"] if deep_map_used else []) + (["uniform-current map"] if cur_map_used else []) + \
                  (["same-grid"] if same_grid_used else [])
        kname = "mr::trace_kernel<GRID,GRID,%s>" % ",".join(variant)
        hw, traffic = None, None
        cap = hw_counters.get(wl.name)
        if cap and args.math == "fast" and bool(cap.get("deep_map")) == deep_map_used and bool(cap.get("same_grid")) == same_grid_used:
            # the hardware's own count of the same kernel, from the committed ncu capture (a smaller launch of the same
            # shape; per-cycle rates do not depend on the launch size)
            fpc = cap["dadd_per_cycle"] + cap["dmul_per_cycle"] + 2.0 * cap["dfma_per_cycle"]
            hw = {
                "fp64_flop_per_cycle": fpc, "fp64_flop_peak_per_cycle": FP64_FLOP_PER_CLK_PER_SM * sm_count,
                "frac_of_fp64_flop_peak": fpc / (FP64_FLOP_PER_CLK_PER_SM * sm_count),
                "pipe_fp64_busy": cap["pipe_fp64_pct"] / 100.0,
                "issue_active": cap.get("issue_active_pct", 0.0) / 100.0,
                "l1_data_pipe_busy": cap.get("l1_lsu_wavefronts_pct", 0.0) / 100.0,
                "instructions_per_rhs": cap.get("instr_per_rhs"),
                "kernel": cap.get("kernel"), "capture": cap.get("capture"), "capture_rays": cap.get("rays"),
                "note": "ncu smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on per cycle: dadd + dmul + 2 dfma over "
                        "148 SM x 128 flop/clk; pipe_fp64 = sm__inst_executed_pipe_fp64 % of peak",
            }
            if cap.get("traffic_rays") and abs(cap["traffic_rays"] - n) <= 0.001 * n and cap.get("rk4_steps") == wl.n_steps:
                traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
        roofline = {
            "bound": "fp64", "kernel": kname,
            "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak if fp64_peak else None,
            "convention": "algorithmic: %d weighted FP64 flop per ray-step (SURVEY.md 8d: add/mul 1, div/sqrt 4, transcendental 8) "
                          "x executed ray-steps / kernel time; the hardware-counted figure is in `hw`" % wl.flop_per_ray_step,
            "hw": hw,
            "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write); algorithmic bytes per launch = %d" % int(alg_bytes),
            "flop_per_ray_step": wl.flop_per_ray_step,
            "peak_source": "builder-probed: register-only DFMA kernel timed in this run (tools/libmr_tools.so, mrt_measure_fp64_peak); "
                           "MEASURED_PEAKS.json has no FP64 entry",
            "kernel_ms": kernel_ms,
        }
        roofline_hbm = {
            "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
            "traffic": traffic, "algorithmic_bytes": int(alg_bytes), "bytes_per_stored_row": W.BYTES_PER_ROW, "peak_source": hbm_src,
        }

        # ---- parity: a strided sample of the buffers the timed launches wrote, against the oracle -----------
        parity = None
        if not args.no_parity:
            n_par = parity_sample_size(wl, n, full, rows_cap, world)
            sel = np.unique(np.linspace(0, n - 1, n_par).astype(np.int64))
            sel_t = torch.from_numpy(sel).to(dev)
            got = {"rows": d_rows[sel_t].cpu().numpy(), "len": d_len[sel_t].cpu().numpy(), "fin": d_fin[:, sel_t].cpu().numpy()}
            if full:
                for i, nm in enumerate(("x", "y", "kx", "ky")):
                    got[nm] = traj[i][:, sel_t].cpu().numpy()
            pr = parity_of(wl, (x0[sel], y0[sel], kx0[sel], ky0[sel]), got, cores)
            pr["rays"] = int(allsum(float(pr["rays"])))
            pr["rows_len_equal"] = bool(allmin(1.0 if pr["rows_len_equal"] else 0.0) > 0.5)
            pr["nan_pattern_equal"] = bool(allmin(1.0 if pr["nan_pattern_equal"] else 0.0) > 0.5)
            pr["max_rel_err"] = allmax(pr["max_rel_err"])
            pr["tol"] = PARITY_TOL
            pr["ok"] = bool(pr["rows_len_equal"] and pr["nan_pattern_equal"] and pr["max_rel_err"] <= PARITY_TOL)
            pr["what"] = ("columns of the timed trajectory planes" if full else "rows / len / final states of the timed launches") + \
                         " on a uniform stride of each rank's rays vs the C oracle (rows and len bit-exact)"
            parity = pr

        summary = {
            "workload": wl.name, "value": value, "unit": UNIT, "ms_per_step": ms_total / steps, "steps": steps,
            "warmup": max(warmup, 3), "rays": wl.n_rays, "rays_per_gpu": n, "rk4_steps": wl.n_steps,
            "grid": [int(wl.bathymetry.x.size), int(wl.bathymetry.y.size)], "stride": wl.stride, "output": wl.output,
            "executed_ray_steps_per_pass": E, "shard": shard_mode,
            "per_rank_ms": per_rank_ms, "imbalance_max_over_mean": max(per_rank_ms) / (sum(per_rank_ms) / len(per_rank_ms)),
            "deep_map": {"flag": args.deep_map, "used": deep_map_used, "deep_share_of_blocks": deep_share},
            "current_map": {"used": cur_map_used, "uniform_share_of_blocks": cur_share},
            "same_grid": {"flag": args.same_grid, "used": same_grid_used},
            "l2": ("each pass writes %.1f GB per GPU, far more than the 126 MB L2: no flush needed" % (out_bytes / 1e9)) if not need_flush
                  else "256 MB buffer written before every timed pass (outside the timed interval)",
            "roofline": roofline, "roofline_hbm": roofline_hbm, "parity": parity, "gpu_launches": int(allsum(float(n_launch))),
        }
        state = dict(wl=wl, fields=fields, n=n, x0=x0, y0=y0, kx0=kx0, ky0=ky0, full=full, rows_cap=rows_cap, clocks=clocks,
                     opts=opts)
        return summary, state

    # ---- the headline workload ---------------------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    head, st = run_workload(args.workload, args.steps, args.warmup, args.rays_per_gpu, sampler)
    wl, fields, n = st["wl"], st["fields"], st["n"]
    full, rows_cap, opts = st["full"], st["rows_cap"], st["opts"]
    x0, y0, kx0, ky0 = st["x0"], st["y0"], st["kx0"], st["ky0"]
    clocks = st["clocks"]
    torch.cuda.empty_cache()            # the resident trajectories are gone: the host path allocates its own slabs

    # ---- end to end through the C ABI with host buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil

        avail = psutil.virtual_memory().available
        per_ray = 32.0 * rows_cap if full else 0.0
        budget = min(0.50 * avail / max(local_world, 1), 70e9)
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

        e2e_call()                                              # warm-up
        e2e_steps = max(min(args.steps, 3), 1)
        barrier()
        t0_ = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_call()
        barrier()
        el = allmax(time.perf_counter() - t0_)
        E_e2e = allsum(float((h_rows.astype(np.int64) - 1).sum()))
        n_e2e_all = int(allsum(float(n_e2e)))
        e2e = {
            "value": E_e2e * e2e_steps / el, "unit": UNIT,
            "h2d_bytes_per_step": int(32 * n_e2e_all),
            "d2h_bytes_per_step": int((per_ray + 8 + 32) * n_e2e_all),
            "rays_per_step": n_e2e_all, "rays_per_step_per_rank": int(n_e2e), "steps": e2e_steps, "ms_per_step": 1e3 * el / e2e_steps,
            "d2h_GB_per_s": (per_ray + 8 + 32) * n_e2e_all * e2e_steps / el / 1e9,
            "batch_note": ("the per-rank batch is what 50 %% of the free host RAM / %d ranks holds as pinned planes "
                           "(%.1f GB per rank): smaller than the %d rays per GPU of the kernel line when several ranks share the host" % (
                               local_world, per_ray * n_e2e / 1e9, n)) if n_e2e < n else "the whole per-GPU batch",
            "api": "mr_trace_many (C ABI, pinned host buffers, H2D of the ray states and D2H of every stored row inside the timed region)",
        }
        del h_traj
    fields.free()
    st = None

    # ---- the other named shapes, at their BASELINE.json sizes ----------------------------------------------------
    workloads = []
    if not args.no_extra and not args.rays_per_gpu:
        for name in ("C1", "C2", "C3", "C4", "C5"):
            if name == args.workload:
                continue
            try:
                s, st_x = run_workload(name, 2, 3, 0)
                st_x["fields"].free()
                st_x = None
                torch.cuda.empty_cache()
                workloads.append(s)
            except Exception as e:           # an extra never takes the headline line down with it
                workloads.append({"workload": name, "error": f"{type(e).__name__}: {e}"})
            barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_leg(wl)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if sampler:
        sampler.stop()
    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl.name, "description": wl.description, "rays": wl.n_rays, "rays_per_gpu": n,
                "rk4_steps": wl.n_steps, "grid": head["grid"],
                "stride": wl.stride, "output": wl.output, "math": args.math,
                "deep_map": head["deep_map"], "current_map": head["current_map"], "same_grid": head["same_grid"],
                "executed_ray_steps_per_pass": head["executed_ray_steps_per_pass"],
                "parallelism": f"rays sharded x{world} ({head['shard']}), fields replicated, no collective",
                "l2": head["l2"],
            },
            "roofline": head["roofline"], "roofline_hbm": head["roofline_hbm"],
            "parity": head["parity"],
            "per_rank_ms": head["per_rank_ms"], "imbalance_max_over_mean": head["imbalance_max_over_mean"],
            "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": head["gpu_launches"], "clocks": clocks,
            "workloads": workloads,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_inproc(args):
    """ONE process, one field handle on --gpus devices, mr_trace_many with pinned host buffers: the library's own
    multi-device path (a shared queue of slabs, one worker thread per device).  End to end by construction: the
    ray states go up and every stored row comes down inside the timed region."""
    import psutil

    from mantaray_b200 import _abi, _capi

    lib = _capi.load()
    G = max(args.gpus, 1)
    if _capi.device_count() < G:
        raise SystemExit(f"--gpus {G} but {_capi.device_count()} device(s) visible")
    # the ensemble of the N-GPU weak-scaling run; --total-rays fixes the batch instead (strong scaling over --gpus)
    wl = make_workload(args.workload, 8 if args.total_rays else G, args.rays_per_gpu)
    rows_cap, full = wl.n_rows, wl.output == "full"
    per_ray = 32.0 * rows_cap if full else 0.0
    avail = psutil.virtual_memory().available
    n = wl.n_rays if per_ray == 0 else int(min(wl.n_rays, max(0.55 * avail // per_ray, 1024)))
    if args.total_rays:
        n = min(n, args.total_rays)
    # a uniform stride over the ensemble, so that a shortened batch keeps the variety of the whole one
    step = max(wl.n_rays // n, 1)
    x0a, y0a, kx0a, ky0a = wl.rays(0, wl.n_rays)
    sel = slice(0, step * n, step)
    hx0 = _capi.pinned_empty((4, n))
    hx0[0], hx0[1], hx0[2], hx0[3] = x0a[sel], y0a[sel], kx0a[sel], ky0a[sel]
    del x0a, y0a, kx0a, ky0a
    h_t = np.empty(rows_cap)
    h_traj = _capi.pinned_empty((4, rows_cap, n)) if full else None
    h_rows = _capi.pinned_empty((n,), np.int32)
    h_len = _capi.pinned_empty((n,), np.int32)
    h_fin = _capi.pinned_empty((4, n))
    math_mode = _abi.MR_MATH_STRICT if args.math == "strict" else _abi.MR_MATH_FAST
    flags = {"auto": 0, "on": _abi.MR_OPT_DEEP_MAP, "off": _abi.MR_OPT_NO_DEEP_MAP}[args.deep_map]
    opts = _abi.TraceOpts(wl.stride, math_mode, 0, flags)
    pp = lambda a: a.ctypes.data if a is not None else None
    fields = _capi.Fields(wl.bathymetry, wl.current, devices=list(range(G)))

    def call():
        rc = lib.mr_trace_many(fields.handle, n, pp(hx0[0]), pp(hx0[1]), pp(hx0[2]), pp(hx0[3]),
                               wl.t0, wl.duration, wl.dt, C.byref(opts), pp(h_t),
                               pp(h_traj[0]) if full else None, pp(h_traj[1]) if full else None,
                               pp(h_traj[2]) if full else None, pp(h_traj[3]) if full else None,
                               pp(h_rows), pp(h_len), pp(h_fin))
        if rc != 0:
            raise RuntimeError(lib.mr_last_error().decode())

    sampler = ClockSampler(0)
    call()                                                      # warm-up: allocates the slabs, pages the planes in
    sampler.start()
    time.sleep(0.3)
    steps = max(min(args.steps, 3), 1)
    t_lo = time.perf_counter()
    for _ in range(steps):
        call()
    t_hi = time.perf_counter()
    sampler.stop()
    el = t_hi - t_lo
    E = float((h_rows.astype(np.int64) - 1).sum())
    split = fields.last_split()
    parity = None
    if not args.no_parity:
        n_par = parity_sample_size(wl, n, full, rows_cap)
        s2 = np.unique(np.linspace(0, n - 1, n_par).astype(np.int64))
        got = {"rows": h_rows[s2].copy(), "len": h_len[s2].copy(), "fin": h_fin[:, s2].copy()}
        if full:
            for i, nm in enumerate(("x", "y", "kx", "ky")):
                got[nm] = h_traj[i][:, s2].copy()
        parity = parity_of(wl, tuple(hx0[i][s2].copy() for i in range(4)), got, os.cpu_count() or 1)
        parity["tol"] = PARITY_TOL
        parity["ok"] = bool(parity["rows_len_equal"] and parity["nan_pattern_equal"] and parity["max_rel_err"] <= PARITY_TOL)
    fields.free()
    d2h = (per_ray + 8 + 32) * n
    val = E * steps / el
    emit({
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": G, "steps": steps, "warmup": 1,
        "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "strong" if args.total_rays else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "mode": "inproc",
        "config": {"workload": wl.name, "description": wl.description, "rays": n, "rays_of_ensemble": wl.n_rays, "rk4_steps": wl.n_steps,
                   "grid": [int(wl.bathymetry.x.size), int(wl.bathymetry.y.size)], "stride": wl.stride, "output": wl.output,
                   "math": args.math, "parallelism": f"one process, one handle on {G} device(s), shared slab queue, no collective"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": int(32 * n), "d2h_bytes_per_step": int(d2h),
                "d2h_GB_per_s": d2h * steps / el / 1e9, "rays_per_step": n,
                "api": "mr_trace_many (C ABI, host buffers; every device of the handle works on the same call)"},
        "rays_taken_per_device_last_call": split,
        "split_max_over_mean": (max(split) / (sum(split) / len(split))) if split and sum(split) else None,
        "parity": parity, "gpu_launches": None, "clocks": sampler.summary(t_lo, t_hi),
        "executed_ray_steps_per_pass": E,
    })


if __name__ == "__main__":
    main()
