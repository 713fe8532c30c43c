"""One int8-sliced f64 GEMM (for ncu): python scripts/sliced_one.py N [slices]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

n = int(sys.argv[1])
s = int(sys.argv[2]) if len(sys.argv) > 2 else 8
a = torch.rand((n, n), device="cuda", dtype=torch.float64) * 2 - 1
b = torch.rand((n, n), device="cuda", dtype=torch.float64) * 2 - 1
c = torch.empty((n, n), device="cuda", dtype=torch.float64)
vp = lambda t: C.c_void_p(t.data_ptr())
path = C.c_int(-1)
for _ in range(2):
    check(lib.eml_gemm_f64_sliced_dev(vp(a), vp(b), vp(c), n, n, n, n, n, n, s, C.byref(path), None))
torch.cuda.synchronize()
print(eml.last_kernel(), path.value, float(c[0, 0]))
