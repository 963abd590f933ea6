"""D2H bandwidth of this box, from one device and from all of them at once: the ceiling of the host-buffer path.

    python tools/pciebench.py [--gib 4] [--devices N]

mr_trace_many drains every stored row to pinned host planes; on one GPU that runs at the PCIe link rate, on
eight the aggregate is whatever the host side (root complexes, memory controllers, NUMA placement) takes.
This measures that ceiling without the library: N concurrent contiguous copies into separate pinned buffers,
the same into column blocks of ONE pinned plane (the gather's 2-D pattern), and the single-device figures.
Prints one JSON line per case; wall clock around a full synchronize of every device."""
import argparse
import json
import subprocess
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gib", type=float, default=4.0, help="GiB copied per device and repetition")
ap.add_argument("--devices", type=int, default=0, help="devices to use (0 = all visible)")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
G = a.devices or torch.cuda.device_count()
n = int(a.gib * (1 << 30)) // 8                 # doubles per device
rows = 2049
w = n // rows                                    # doubles per row and device in the 2-D case


def sync_all():
    for g in range(G):
        torch.cuda.synchronize(g)


def timed(label, fn, nbytes, **extra):
    best = 0.0
    for _ in range(a.reps):
        sync_all()
        t = time.perf_counter()
        fn()
        sync_all()
        best = max(best, nbytes / (time.perf_counter() - t) / 1e9)
    print(json.dumps({"case": label, "devices": extra.pop("devices", G), "GB_per_s": round(best, 2),
                      "GB_per_s_per_device": round(best / extra.get("active", G), 2), **extra}), flush=True)
    return best


try:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout)
    print(subprocess.run(["bash", "-c", "lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)'; free -g | head -2"],
                         capture_output=True, text=True, timeout=20).stdout)
except Exception as e:                           # informational only
    print("topology not available:", e)

dev_buf = [torch.empty(n, dtype=torch.float64, device=f"cuda:{g}") for g in range(G)]
streams = [torch.cuda.Stream(device=g) for g in range(G)]

# ---- separate pinned buffers, contiguous ---------------------------------------------------------------------
host_sep = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(G)]


def copy_sep(devs):
    def run():
        for g in devs:
            with torch.cuda.stream(streams[g]):
                host_sep[g].copy_(dev_buf[g], non_blocking=True)
    return run


for g in range(G):
    timed(f"contiguous, device {g} alone", copy_sep([g]), n * 8, active=1, devices=1)
for k in sorted({2, 4, G} & set(range(2, G + 1))):
    timed(f"contiguous, {k} devices at once, separate pinned buffers", copy_sep(range(k)), k * n * 8, active=k, devices=k)
del host_sep

# ---- one pinned plane, each device its column block (the gather of mr_trace_many: cudaMemcpy2DAsync) ----------
from cuda.bindings import runtime as cudart

plane = torch.empty((rows, G * w), dtype=torch.float64).pin_memory()
D2H = cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost


def copy_cols(devs):
    def run():
        for g in devs:
            torch.cuda.set_device(g)
            (err,) = cudart.cudaMemcpy2DAsync(plane.data_ptr() + g * w * 8, G * w * 8, dev_buf[g].data_ptr(), w * 8, w * 8, rows,
                                              D2H, streams[g].cuda_stream)
            assert err == cudart.cudaError_t.cudaSuccess, err
    return run


timed("2-D column block (cudaMemcpy2DAsync), device 0 alone", copy_cols([0]), rows * w * 8, active=1, devices=1, row_bytes=w * 8)
for k in sorted({2, 4, G} & set(range(2, G + 1))):
    timed(f"2-D column blocks of one pinned plane, {k} devices at once", copy_cols(range(k)), k * rows * w * 8, active=k,
          devices=k, row_bytes=w * 8)
