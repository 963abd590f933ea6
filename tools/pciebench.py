"""D2H bandwidth of this box: contiguous vs the strided 2-D pattern of mr_trace_many's drain."""
import torch, time
n = 1 << 30          # 8 GiB of f64
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()
for name, fn in (("contiguous 8 GiB", lambda: h.copy_(d, non_blocking=True)),):
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
        print(name, n * 8 / (time.perf_counter() - t) / 1e9, "GB/s")
# 2-D: 2049 rows of 227328 doubles into a host plane with pitch 961152
rows, w, pitch = 2049, 227328, 961152
d2 = torch.empty((rows, w), dtype=torch.float64, device="cuda")
h2 = torch.empty((rows, pitch), dtype=torch.float64).pin_memory()
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); h2[:, :w].copy_(d2, non_blocking=True); torch.cuda.synchronize()
    print("2-D 2049 x 1.8 MB rows", rows * w * 8 / (time.perf_counter() - t) / 1e9, "GB/s")
# two streams at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n // 2
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s1): h[:half].copy_(d[:half], non_blocking=True)
    with torch.cuda.stream(s2): h[half:].copy_(d[half:], non_blocking=True)
    torch.cuda.synchronize()
    print("two streams", n * 8 / (time.perf_counter() - t) / 1e9, "GB/s")
