"""K sweep of the CTA-pair f32 product at the NN step's two shapes: fixed overhead vs per-k-block cost (run under ncu)."""
import ctypes as C, sys
sys.path.insert(0, ".")
import torch
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

def s3(b, r, c): return (C.c_int64 * 3)(b, r, c)
def contract(a, b, c, M, N, K, sa, sb):
    check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, M, N, K, sa, sb, s3(0, N, 1), 0, None))
torch.manual_seed(0)
for K in (32, 256, 512, 784, 1568, 3136):
    x = torch.rand(8192, K, device="cuda"); w = torch.rand(K, 512, device="cuda") - 0.5; z = torch.empty(8192, 512, device="cuda")
    for _ in range(3): contract(x, w, z, 8192, 512, K, s3(0, K, 1), s3(0, 512, 1))
torch.cuda.synchronize()
# dW1 = X^T dZ: M = 784, N = 512, K = batch
for Kb in (1024, 4096, 8192, 16384):
    x = torch.rand(Kb, 784, device="cuda"); dz = torch.rand(Kb, 512, device="cuda") - 0.5; dw = torch.empty(784, 512, device="cuda")
    for _ in range(3): contract(x, dz, dw, 784, 512, Kb, s3(0, 1, 784), s3(0, 512, 1))
torch.cuda.synchronize()
for M in (768, 1024):
    x = torch.rand(8192, M, device="cuda"); dz = torch.rand(8192, 512, device="cuda") - 0.5; dw = torch.empty(M, 512, device="cuda")
    for _ in range(3): contract(x, dz, dw, M, 512, 8192, s3(0, 1, M), s3(0, 512, 1))
torch.cuda.synchronize()
