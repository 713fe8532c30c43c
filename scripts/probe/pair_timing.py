"""In-kernel timeline of the CTA-pair product (library built with -DEML_PAIR_TIMING=<block>): NN forward and dW1 shapes."""
import ctypes as C, sys, os
sys.path.insert(0, ".")
os.environ["EML_SGEMM_FUSED"] = "0"
import torch
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib
def s3(b, r, c): return (C.c_int64 * 3)(b, r, c)
def contract(a, b, c, M, N, K, sa, sb):
    check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, M, N, K, sa, sb, s3(0, N, 1), 0, None))
for K in (32, 256, 784):
    x = torch.rand(8192, K, device="cuda"); w = torch.rand(K, 512, device="cuda") - 0.5; z = torch.empty(8192, 512, device="cuda")
    print("forward K", K, flush=True)
    for _ in range(3): contract(x, w, z, 8192, 512, K, s3(0, K, 1), s3(0, 512, 1)); torch.cuda.synchronize()
x = torch.rand(8192, 784, device="cuda"); dz = torch.rand(8192, 512, device="cuda") - 0.5; dw = torch.empty(784, 512, device="cuda")
print("dW1", flush=True)
for _ in range(3): contract(x, dz, dw, 784, 512, 8192, s3(0, 1, 784), s3(0, 512, 1)); torch.cuda.synchronize()
