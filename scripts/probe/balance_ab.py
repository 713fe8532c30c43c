"""Balanced tiling of the CTA-pair f32 kernel (narrow N tiles / transposed product) vs the plain tiling (EML_SGEMM_BALANCE=0):
error against the f64 product in units of the 1e-5 sum|a b| contract, and time."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import torch
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib
def s3(*v): return (C.c_int64 * 3)(*v)
def call(a, b, c, M, N, K, sa, sb):
    check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, M, N, K, sa, sb, s3(0, N, 1), 0, None))
def timeit(f, reps=50):
    for _ in range(5): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
os.environ["EML_SGEMM_FUSED"] = "0"
shapes = [("fwd 8192x512x784", 8192, 512, 784, False), ("dW1 784x512x8192 (A^T)", 784, 512, 8192, True), ("1000x300x2048", 1000, 300, 2048, False),
          ("300x1000x512 (A^T)", 300, 1000, 512, True), ("520x260x300", 520, 260, 300, False), ("2000x2000x256", 2000, 2000, 256, False),
          ("4096x784x512", 4096, 784, 512, False), ("392x392x4096 (A^T)", 392, 392, 4096, True),
          ("4096^3", 4096, 4096, 4096, False), ("2048^3", 2048, 2048, 2048, False), ("3072x3072x1024", 3072, 3072, 1024, False)]
for name, M, N, K, at in shapes:
    torch.manual_seed(3)
    if at:
        a = torch.rand(K, M, device="cuda") - 0.3; sa = s3(0, 1, M); a64 = a.double().t()
    else:
        a = torch.rand(M, K, device="cuda") - 0.3; sa = s3(0, K, 1); a64 = a.double()
    b = torch.rand(K, N, device="cuda") - 0.5; sb = s3(0, N, 1)
    ref = a64 @ b.double(); bound = 1e-5 * (a64.abs() @ b.double().abs())
    res = {}
    for mode in ("0", "1", "0", "1"):
        os.environ["EML_SGEMM_BALANCE"] = mode
        c = torch.full((M, N), float("nan"), device="cuda")
        call(a, b, c, M, N, K, sa, sb)
        torch.cuda.synchronize()
        err = ((c.double() - ref).abs() / bound).max().item()
        us = timeit(lambda: call(a, b, c, M, N, K, sa, sb))
        res[mode] = (err, min(us, res[mode][1]) if mode in res else us, c.clone())
    print(f"{name:28s} balanced: err/bound {res['1'][0]:.3f} {res['1'][1]:7.1f} us | plain: {res['0'][0]:.3f} {res['0'][1]:7.1f} us | same bits {torch.equal(res['1'][2], res['0'][2])}", flush=True)
