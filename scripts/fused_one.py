"""One launch of the f32 GEMM at a given shape (for ncu): python scripts/fused_one.py M N K [batch]"""
import ctypes as C
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
eml = importlib.import_module("easy-ml_b200")
lib = eml._ffi.lib
m, n, k = (int(x) for x in sys.argv[1:4])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
a = torch.randn(batch, m, k, device="cuda")
b = torch.randn(batch, k, n, device="cuda")
c = torch.empty(batch, m, n, device="cuda")
s3 = lambda t: (C.c_int64 * 3)(*t.stride())
for _ in range(2):
    rc = lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), batch, m,
                                  n, k, s3(a), s3(b), s3(c), 0, None)
    assert rc == 0
torch.cuda.synchronize()
print(eml.last_kernel(), float(c[0, 0, 0]))
