import ctypes as C, time, sys
sys.path.insert(0, ".")
import numpy as np, torch
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib
n = 8192
r = np.random.default_rng(0)
a = r.uniform(-1, 1, (n, n)); b = r.uniform(-1, 1, (n, n)); c = np.empty((n, n))
p = lambda x: x.ctypes.data_as(C.c_void_p)
for kind in ("pageable", "pinned"):
    if kind == "pinned":
        A = torch.from_numpy(a).pin_memory(); B = torch.from_numpy(b).pin_memory(); Cc = torch.empty((n, n), dtype=torch.float64).pin_memory()
        pa, pb, pc = C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(Cc.data_ptr())
    else:
        pa, pb, pc = p(a), p(b), p(c)
    check(lib.eml_gemm_f64(pa, pb, pc, n, n, n, n, n, n))
    t0 = time.perf_counter()
    for _ in range(3):
        check(lib.eml_gemm_f64(pa, pb, pc, n, n, n, n, n, n))
    dt = (time.perf_counter() - t0) / 3
    print(kind, f"{dt*1e3:.1f} ms", f"{2*n**3/dt/1e12:.1f} TF")
