"""A/B of the f32 8192^2 product: operands read in place (EML_SGEMM_DIRECT=1) vs packed planes (=0), alternating."""
import ctypes as C, os, sys, time
sys.path.insert(0, ".")
import torch
import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda"); c = torch.empty(n, n, device="cuda")
s3 = lambda t: (C.c_int64 * 3)(0, *t.stride())
def run(reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, n, n, n, s3(a), s3(b), s3(c), 0, None))
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, n, n, n, s3(a), s3(b), s3(c), 0, None))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2 * n ** 3 / ms / 1e9
for rnd in range(3):
    for mode in ("1", "0"):
        os.environ["EML_SGEMM_DIRECT"] = mode
        ms, tf = run()
        print(f"round {rnd} direct={mode}: {ms:.3f} ms {tf:.1f} TFLOP/s")
