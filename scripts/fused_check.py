"""Compare the in-kernel-split f32 GEMM (EML_SGEMM_FUSED=1) with the packed path: bitwise, then timing."""
import ctypes as C
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
eml = importlib.import_module("easy-ml_b200")
lib = eml._ffi.lib


def vp(t):
    return C.c_void_p(t.data_ptr())


def s3(*s):
    return (C.c_int64 * 3)(*s)


def contract(a, b, ta, tb, acc=None):
    # a: (batch, M, K) logical; stored transposed when ta.  b: (batch, K, N) logical; stored transposed when tb
    batch, m, k = a.shape
    n = b.shape[2]
    am = a.transpose(1, 2).contiguous().transpose(1, 2) if ta else a.contiguous()
    bm = b.transpose(1, 2).contiguous().transpose(1, 2) if tb else b.contiguous()
    c = torch.zeros((batch, m, n), device="cuda", dtype=torch.float32) if acc is None else acc.clone()
    rc = lib.eml_contract_f32_dev(vp(am), vp(bm), vp(c), batch, m, n, k, s3(*am.stride()), s3(*bm.stride()),
                                  s3(*c.stride()), 0 if acc is None else 1, None)
    assert rc == 0, eml._ffi.last_error()
    torch.cuda.synchronize()
    return c


def main():
    torch.manual_seed(1)
    cases = [(1, 256, 256, 32), (1, 256, 256, 512), (1, 512, 768, 1024), (1, 300, 500, 700), (3, 260, 384, 100),
             (1, 1024, 1024, 1024), (2, 1024, 1024, 2048), (1, 4096, 4096, 544), (1, 132, 4100, 36)]
    bad = 0
    for (batch, m, n, k) in cases:
        for ta in (False, True):
            for tb in (False, True):
                # keep TMA alignment: leading dimensions multiples of 4
                if (ta and m % 4) or (not ta and k % 4) or (tb and k % 4) or (not tb and n % 4):
                    continue
                a = torch.randn(batch, m, k, device="cuda")
                b = torch.randn(batch, k, n, device="cuda")
                os.environ["EML_SGEMM_FUSED"] = "0"
                ref = contract(a, b, ta, tb)
                os.environ["EML_SGEMM_FUSED"] = "1"
                got = contract(a, b, ta, tb)
                same = torch.equal(ref, got)
                err = (ref - got).abs().max().item()
                print(f"batch={batch} M={m} N={n} K={k} ta={ta} tb={tb}: {'bitwise' if same else 'DIFF'} maxdiff={err:.3e}",
                      flush=True)
                bad += not same
    print("mismatching cases:", bad)

    def bench(batch, m, n, k, mode, reps=10):
        os.environ["EML_SGEMM_FUSED"] = mode
        a = torch.randn(batch, m, k, device="cuda")
        b = torch.randn(batch, k, n, device="cuda")
        c = torch.empty(batch, m, n, device="cuda")
        args = (vp(a), vp(b), vp(c), batch, m, n, k, s3(*a.stride()), s3(*b.stride()), s3(*c.stride()), 0, None)
        for _ in range(3):
            lib.eml_contract_f32_dev(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.eml_contract_f32_dev(*args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        return 2.0 * batch * m * n * k / ms / 1e9, ms

    for shape in [(64, 1024, 1024, 1024), (1, 2048, 2048, 2048), (1, 4096, 4096, 4096), (1, 8192, 8192, 8192),
                  (1, 8192, 8192, 512), (16, 2048, 2048, 256)]:
        p, pms = bench(*shape, "0")
        f, fms = bench(*shape, "1")
        print(f"{shape}: packed {p:.1f} TF ({pms:.3f} ms)   fused {f:.1f} TF ({fms:.3f} ms)", flush=True)


if __name__ == "__main__":
    main()
