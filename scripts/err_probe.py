"""Relative error of the f32 GEMM on same-sign data (bias) and mixed signs, through ctypes on a given library."""
import ctypes as C, sys
import numpy as np
lib = C.CDLL(sys.argv[1])
lib.eml_gemm_f32.argtypes = [C.c_void_p] * 3 + [C.c_int64] * 6
def gemm(a, b):
    m, k = a.shape; n = b.shape[1]
    c = np.empty((m, n), np.float32)
    rc = lib.eml_gemm_f32(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, n, k, k, n, n)
    assert rc == 0
    return c
for (m, n, k) in [(256, 384, 16384), (128, 256, 8192), (300, 100, 4096), (512, 512, 512), (2048, 2048, 2048)]:
    r = np.random.default_rng(1200 + m + n + k)
    a = r.uniform(0, 1, (m, k)).astype(np.float32); b = r.uniform(0, 1, (k, n)).astype(np.float32)
    got = gemm(a, b)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    rel = (got.astype(np.float64) - exact) / exact
    a2 = r.uniform(-1, 1, (m, k)).astype(np.float32); b2 = r.uniform(-1, 1, (k, n)).astype(np.float32)
    got2 = gemm(a2, b2)
    exact2 = a2.astype(np.float64) @ b2.astype(np.float64)
    bound = np.abs(a2).astype(np.float64) @ np.abs(b2).astype(np.float64)
    print(f"{m}x{n}x{k}: same sign mean rel {rel.mean():+.3e} max |rel| {np.abs(rel).max():.3e};  mixed signs max |err| / sum|ab| {(np.abs(got2 - exact2) / bound).max():.3e}")
