"""Probe: page-locked host-to-device bandwidth, contiguous vs 2-D copies with 8 KB / 4 KB rows (the K-panel slices of
a row-major A that the pipelined host GEMM uploads), and concurrent device-to-host traffic."""
import torch

n = 8192
a = torch.empty((n, n), dtype=torch.float64).pin_memory()
d = torch.empty((n, n), dtype=torch.float64, device="cuda")
out = torch.empty((n // 2, n), dtype=torch.float64).pin_memory()
s2 = torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: d.copy_(a, non_blocking=True))
print(f"contiguous 512 MB: {a.numel() * 8 / ms / 1e6:.1f} GB/s")
for cols in (1024, 512, 256):
    ms = timed(lambda: d[:, :cols].copy_(a[:, :cols], non_blocking=True))
    print(f"2-D, {cols * 8} B rows x {n}: {n * cols * 8 / ms / 1e6:.1f} GB/s")


def both():
    with torch.cuda.stream(s2):
        out.copy_(d[: n // 2], non_blocking=True)
    d.copy_(a, non_blocking=True)


ms = timed(both)
print(f"contiguous H2D 512 MB with a concurrent 256 MB D2H: {ms:.2f} ms -> H2D >= {a.numel() * 8 / ms / 1e6:.1f} GB/s")
