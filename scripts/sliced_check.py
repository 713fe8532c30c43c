"""f64 GEMM through int8 slices (eml_gemm_f64_sliced_dev): contract check against an exact reference, guard behaviour,
timing against the DMMA kernel."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

vp = lambda t: C.c_void_p(t.data_ptr())


def sliced(a, b, slices=0):
    m, k = a.shape
    n = b.shape[1]
    c = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float64)
    path = C.c_int(-1)
    check(lib.eml_gemm_f64_sliced_dev(vp(a), vp(b), vp(c), m, n, k, a.stride(0), b.stride(0), c.stride(0), slices,
                                      C.byref(path), None))
    torch.cuda.synchronize()
    return c, path.value


def dmma(a, b):
    m, k = a.shape
    n = b.shape[1]
    c = torch.empty((m, n), device="cuda", dtype=torch.float64)
    check(lib.eml_gemm_f64_dev(vp(a), vp(b), vp(c), m, n, k, a.stride(0), b.stride(0), c.stride(0), None))
    torch.cuda.synchronize()
    return c


def exact_rows(a, b, rows):
    """rows of A*B in ~106-bit arithmetic (double-double via splitting into hi/lo float64 of the products is overkill
    here: use Python fractions on a few rows would be slow; long double on the host is enough to judge 1e-12)."""
    ah = a[rows].cpu().numpy().astype(np.longdouble)
    bh = b.cpu().numpy().astype(np.longdouble)
    return ah @ bh, np.abs(ah) @ np.abs(bh)


def report(name, a, b, slices=0):
    c, path = sliced(a, b, slices)
    rows = [0, 1, a.shape[0] // 2, a.shape[0] - 1]
    ref, bound = exact_rows(a, b, rows)
    err = np.abs(c[rows].cpu().numpy().astype(np.longdouble) - ref)
    ratio = float((err / (1e-12 * bound + 1e-300)).max())
    cd = dmma(a, b)
    errd = np.abs(cd[rows].cpu().numpy().astype(np.longdouble) - ref)
    ratiod = float((errd / (1e-12 * bound + 1e-300)).max())
    print(f"{name}: path={path} max err / (1e-12 sum|ab|) = {ratio:.3e}   (DMMA kernel: {ratiod:.3e})", flush=True)
    return path, ratio


def main():
    g = torch.Generator(device="cuda").manual_seed(3)
    u = lambda *s: torch.rand(s, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    report("uniform 512x768x640", u(512, 640), u(640, 768))
    report("uniform 1000x1300x900 (ragged)", u(1000, 900), u(900, 1300))
    report("uniform 2048^3", u(2048, 2048), u(2048, 2048))
    report("normal 1024^3 x 1e5 scale rows", torch.randn(1024, 1024, generator=g, device="cuda", dtype=torch.float64)
           * torch.logspace(-5, 5, 1024, device="cuda", dtype=torch.float64)[:, None], u(1024, 1024))
    ai = torch.randint(-8, 9, (512, 512), generator=g, device="cuda").double()
    bi = torch.randint(-8, 9, (512, 512), generator=g, device="cuda").double()
    c, path = sliced(ai, bi)
    print("small integers exact:", path, bool(torch.equal(c, ai @ bi)))
    # guard must refuse: one tiny element per row against a column that only "sees" it
    a = u(512, 512)
    b = torch.zeros(512, 512, device="cuda", dtype=torch.float64)
    a[:, 0] = 1e-30
    b[0, :] = 1.0
    b[1:, 1:] = u(511, 511)
    path, ratio = report("adversarial (column 0 only meets the 1e-30 entries)", a, b)
    print("guard refused:", path == 0)
    a = u(512, 512); a[5, 7] = float("inf")
    print("inf refused:", sliced(a, u(512, 512))[1] == 0)
    for s in (6, 7, 8, 9, 10):
        report(f"uniform 1024^3, {s} slices", u(1024, 1024), u(1024, 1024), s)

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for n in (4096, 8192):
        a, b = u(n, n), u(n, n)
        c = torch.empty((n, n), device="cuda", dtype=torch.float64)
        path = C.c_int(-1)
        for s in (7, 8):
            ms = timeit(lambda: check(lib.eml_gemm_f64_sliced_dev(vp(a), vp(b), vp(c), n, n, n, n, n, n, s, C.byref(path), None)))
            print(f"n={n} slices={s}: {ms:.2f} ms = {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s (path {path.value})", flush=True)
        ms = timeit(lambda: check(lib.eml_gemm_f64_dev(vp(a), vp(b), vp(c), n, n, n, n, n, n, None)))
        print(f"n={n} DMMA: {ms:.2f} ms = {2 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
