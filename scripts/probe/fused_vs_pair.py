"""One-wave f32 products: the in-kernel-split kernel (what the dispatcher picks for small outputs without cached planes)
against the CTA-pair kernel reading the operands in place (+ two small-part passes), which can use narrow tiles."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import torch
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib
def s3(*v): return (C.c_int64 * 3)(*v)
def call(a, b, c, M, N, K):
    check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, M, N, K, s3(0, K, 1), s3(0, N, 1), s3(0, N, 1), 0, None))
def timeit(f, reps=50):
    for _ in range(5): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for M, N, K in [(1000, 300, 2048), (4096, 784, 1024), (2000, 2000, 1024), (2048, 2048, 2048), (1024, 1024, 1024), (512, 512, 4096), (3072, 3072, 1024), (1000, 1000, 1000), (1536, 1536, 8192)]:
    a = torch.rand(M, K, device="cuda") - 0.3; b = torch.rand(K, N, device="cuda") - 0.5; c = torch.empty(M, N, device="cuda")
    out = []
    for mode in (None, "0", None, "0"):
        if mode is None: os.environ.pop("EML_SGEMM_FUSED", None)
        else: os.environ["EML_SGEMM_FUSED"] = mode
        out.append(timeit(lambda: call(a, b, c, M, N, K)))
    print(f"{M}x{N}x{K}: default {min(out[0], out[2]):7.1f} us   pair (FUSED=0) {min(out[1], out[3]):7.1f} us", flush=True)
