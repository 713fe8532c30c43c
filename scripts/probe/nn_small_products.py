"""The NN step's three skinny products on their own, L2-hot, back to back (time per call incl. launch):
Y = H W2 (8192x512 . 512x10), dW2 = H^T dY (512x8192 . 8192x10), dH = dY W2^T (8192x10 . 10x512)."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import torch
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib
def s3(*v): return (C.c_int64 * 3)(*v)
def call(a, b, c, M, N, K, sa, sb):
    check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, M, N, K, sa, sb, s3(0, N, 1), 0, None))
def timeit(f, reps=200):
    for _ in range(10): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
torch.manual_seed(0)
n, h, o = 8192, 512, 10
H = torch.rand(n, h, device="cuda"); W2 = torch.rand(h, o, device="cuda") - 0.5; dY = torch.rand(n, o, device="cuda") - 0.5
Y = torch.empty(n, o, device="cuda"); dW2 = torch.empty(h, o, device="cuda"); dH = torch.empty(n, h, device="cuda")
ref = {}
t = timeit(lambda: call(H, W2, Y, n, o, h, s3(0, h, 1), s3(0, o, 1))); print(f"Y = H W2      {t:6.2f} us  {eml.last_kernel()}  err {(Y.double() - H.double() @ W2.double()).abs().max().item():.2e}")
t = timeit(lambda: call(H, dY, dW2, h, o, n, s3(0, 1, h), s3(0, o, 1))); print(f"dW2 = H^T dY  {t:6.2f} us  {eml.last_kernel()}  err {(dW2.double() - H.double().t() @ dY.double()).abs().max().item():.2e}")
t = timeit(lambda: call(dY, W2, dH, n, h, o, s3(0, o, 1), s3(0, 1, o))); print(f"dH = dY W2^T  {t:6.2f} us  {eml.last_kernel()}  err {(dH.double() - dY.double() @ W2.double().t()).abs().max().item():.2e}")
import hashlib
print("bits", hashlib.sha256(Y.cpu().numpy().tobytes() + dW2.cpu().numpy().tobytes() + dH.cpu().numpy().tobytes()).hexdigest()[:16])
