"""Probe: cudaMemcpy2DAsync page-locked host -> device, by row width (what the pipelined host GEMM issues for the
K-panel slices of a row-major A)."""
import ctypes as C
import glob
import os

import torch

torch.cuda.init()
rt = None
for pat in ("libcudart.so*",):
    for d in (os.path.join(os.path.dirname(torch.__file__), "lib"), "/usr/local/cuda/lib64"):
        for f in sorted(glob.glob(os.path.join(d, pat))):
            try:
                rt = C.CDLL(f)
                break
            except OSError:
                pass
        if rt:
            break
import nvidia.cuda_runtime as _n  # noqa: E402  (wheel layout: site-packages/nvidia/cuda_runtime/lib)
if rt is None:
    rt = C.CDLL(glob.glob(os.path.join(os.path.dirname(_n.__file__), "lib", "libcudart.so*"))[0])
n = 8192
a = torch.empty((n, n), dtype=torch.float64).pin_memory()
d = torch.empty((n, n), dtype=torch.float64, device="cuda")
H2D = 1


def copy2d(width_bytes, rows):
    rc = rt.cudaMemcpy2DAsync(C.c_void_p(d.data_ptr()), C.c_size_t(n * 8), C.c_void_p(a.data_ptr()), C.c_size_t(n * 8),
                              C.c_size_t(width_bytes), C.c_size_t(rows), C.c_int(H2D), C.c_void_p(0))
    assert rc == 0, rc


for width in (65536, 32768, 16384, 8192, 4096):
    copy2d(width, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        copy2d(width, n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"cudaMemcpy2DAsync rows of {width} B (pitch 65536) x {n}: {width * n / ms / 1e6:.1f} GB/s")
