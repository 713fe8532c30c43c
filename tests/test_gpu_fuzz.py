"""Randomised shapes / strides / alignments through the device-resident contraction entry points.

Every kernel family has layout preconditions (TMA needs 16-byte aligned rows and extents divisible by 4, the
vector paths need aligned bases, the pair kernel needs M, N > 128, split-K needs divisible K ...); the
dispatcher must route everything else to a kernel that can take it.  Each case is checked against a float64
product within the north star's bound  tol * sum_k |a_k b_k|  (tol 1e-12 f64 / 1e-5 f32).
"""
import ctypes as C

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def vp(t):
    return C.c_void_p(t.data_ptr())


def s3(*s):
    return (C.c_int64 * 3)(*s)


def make_operand(r, g, batch, rows, cols, dtype, transposed, pad, offset):
    """A (batch, rows, cols) logical operand inside a larger buffer: optional transposed storage, leading-dimension
    padding and an element offset of the base pointer.  Returns (view tensor, strides (b, r, c))."""
    pr, pc = (cols, rows) if transposed else (rows, cols)
    ld = pc + pad
    buf = torch.rand(offset + batch * pr * ld + 8, generator=g, device="cuda", dtype=dtype) * 2 - 1
    phys = buf[offset:offset + batch * pr * ld].view(batch, pr, ld)[:, :, :pc]
    if transposed:
        return phys.transpose(1, 2), (pr * ld, 1, ld)
    return phys, (pr * ld, ld, 1)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("seed", range(6))
def test_contract_fuzz(dtype, seed):
    r = np.random.default_rng(0xF00D + seed + (100 if dtype == torch.float32 else 0))
    g = torch.Generator(device="cuda").manual_seed(int(r.integers(1 << 30)))
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    f = lib.eml_contract_f64_dev if dtype == torch.float64 else lib.eml_contract_f32_dev
    kernels = set()
    for case in range(40):
        big = r.random() < 0.35
        hi = 700 if big else 150
        m, n = int(r.integers(1, hi)), int(r.integers(1, hi))
        k = int(r.integers(1, 1200 if big else 200))
        if r.random() < 0.3:  # make the tensor-core layouts reachable: multiples of 4
            m, n, k = (m + 3) // 4 * 4, (n + 3) // 4 * 4, (k + 3) // 4 * 4
        batch = int(r.choice([1, 1, 1, 2, 5]))
        ta, tb = bool(r.integers(2)), bool(r.integers(2))
        pads = [int(r.choice([0, 0, 1, 2, 4, 6])) for _ in range(3)]
        offs = [int(r.choice([0, 0, 1, 2, 3, 4])) for _ in range(3)]
        acc = bool(r.integers(2))
        a, sa = make_operand(r, g, batch, m, k, dtype, ta, pads[0], offs[0])
        b, sb = make_operand(r, g, batch, k, n, dtype, tb, pads[1], offs[1])
        c, sc = make_operand(r, g, batch, m, n, dtype, False, pads[2], offs[2])
        c0 = c.clone()
        untouched = c.storage_offset()  # noqa: F841  (views share storage; padding is checked below)
        check(f(vp(a), vp(b), vp(c), batch, m, n, k, s3(*sa), s3(*sb), s3(*sc), 1 if acc else 0, None))
        torch.cuda.synchronize()
        kernels.add(eml.last_kernel())
        want = torch.matmul(a.double(), b.double()) + (c0.double() if acc else 0)
        bound = tol * torch.matmul(a.abs().double(), b.abs().double()) + (tol * c0.abs().double() if acc else 0) + 1e-300
        err = (c.double() - want).abs()
        worst = float((err / bound).max())
        assert worst <= 1.0, (f"case {case}: {worst:.3g}x the bound; batch={batch} m={m} n={n} k={k} ta={ta} tb={tb} "
                              f"pads={pads} offs={offs} acc={acc} kernel={eml.last_kernel()}")
    assert len(kernels) >= 2, kernels  # the fuzz really crossed several kernel families


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_contract_fuzz_tensor_core_sizes(dtype):
    """Larger random shapes (the TMA / DMMA, cp.async / DMMA, CTA-pair tcgen05 and single-CTA tcgen05 kernels with
    their split-K variants), ragged in every dimension, all orientations, odd leading dimensions and bases."""
    r = np.random.default_rng(0xABCD + (1 if dtype == torch.float32 else 0))
    g = torch.Generator(device="cuda").manual_seed(99)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    f = lib.eml_contract_f64_dev if dtype == torch.float64 else lib.eml_contract_f32_dev
    kernels = set()
    for case in range(24):
        m, n, k = int(r.integers(100, 1800)), int(r.integers(100, 1800)), int(r.integers(64, 4000))
        if case % 3 == 0:
            m, n, k = m // 4 * 4, n // 4 * 4, k // 4 * 4
        if case % 8 == 7:
            m = int(r.integers(16, 128))  # short outputs: the single-CTA / split-K paths
        batch = 1 if case % 5 else 3
        ta, tb = bool(r.integers(2)), bool(r.integers(2))
        if case % 3 == 0:  # 16-byte aligned rows and bases: the TMA-staged kernels are reachable
            pads = [int(r.choice([0, 4, 8])) for _ in range(3)]
            offs = [int(r.choice([0, 4, 8])) for _ in range(3)]
        else:
            pads = [int(r.choice([0, 0, 2, 3, 4])) for _ in range(3)]
            offs = [int(r.choice([0, 0, 1, 2, 4])) for _ in range(3)]
        acc = bool(r.integers(2))
        a, sa = make_operand(r, g, batch, m, k, dtype, ta, pads[0], offs[0])
        b, sb = make_operand(r, g, batch, k, n, dtype, tb, pads[1], offs[1])
        c, sc = make_operand(r, g, batch, m, n, dtype, False, pads[2], offs[2])
        c0 = c.clone()
        check(f(vp(a), vp(b), vp(c), batch, m, n, k, s3(*sa), s3(*sb), s3(*sc), 1 if acc else 0, None))
        torch.cuda.synchronize()
        kernels.add(eml.last_kernel())
        want = torch.matmul(a.double(), b.double()) + (c0.double() if acc else 0)
        bound = tol * torch.matmul(a.abs().double(), b.abs().double()) + (tol * c0.abs().double() if acc else 0) + 1e-300
        worst = float(((c.double() - want).abs() / bound).max())
        assert worst <= 1.0, (f"case {case}: {worst:.3g}x the bound; batch={batch} m={m} n={n} k={k} ta={ta} tb={tb} "
                              f"pads={pads} offs={offs} acc={acc} kernel={eml.last_kernel()}")
    if dtype == torch.float64:
        assert {"dmma_f64_tma", "dmma_f64"} <= kernels, kernels
    else:
        assert "tf32x3_tcgen05" in kernels, kernels


def test_host_gemm_fuzz(orc):
    """Host-pointer entry with leading dimensions; the padding columns of C must stay untouched."""
    r = np.random.default_rng(0xBEEF)
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        for _ in range(15):
            m, n, k = int(r.integers(1, 400)), int(r.integers(1, 400)), int(r.integers(1, 500))
            la, lb, lc = k + int(r.integers(0, 5)), n + int(r.integers(0, 5)), n + int(r.integers(0, 5))
            a = r.uniform(-1, 1, (m, la)).astype(dtype)
            b = r.uniform(-1, 1, (k, lb)).astype(dtype)
            c = np.full((m, lc), 9.0, dtype)
            fn = lib.eml_gemm_f64 if dtype == np.float64 else lib.eml_gemm_f32
            check(fn(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), m, n, k,
                     la, lb, lc))
            av, bv = a[:, :k], b[:, :n]
            want = av.astype(np.float64) @ bv.astype(np.float64)
            bound = tol * (np.abs(av).astype(np.float64) @ np.abs(bv).astype(np.float64)) + 1e-300
            assert np.all(np.abs(c[:, :n].astype(np.float64) - want) <= bound), (dtype, m, n, k, la, lb, lc)
            assert np.all(c[:, n:] == 9.0)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_broadcast_strides_reach_a_kernel_that_can_take_them(dtype):
    """Zero strides (an operand broadcast along the batch, its rows or its columns) are legal views for the C ABI
    but not for TMA descriptors: every such case must be routed to a kernel that reads plain strides."""
    g = torch.Generator(device="cuda").manual_seed(5)
    bt, m, n, k = 3, 256, 320, 288
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    f = lib.eml_contract_f64_dev if dtype == torch.float64 else lib.eml_contract_f32_dev
    a = torch.rand((bt, m, k), generator=g, device="cuda", dtype=dtype) * 2 - 1
    b = torch.rand((bt, k, n), generator=g, device="cuda", dtype=dtype) * 2 - 1
    cases = {
        "A shared by every batch": ((0, k, 1), (k * n, n, 1), a[0].expand(bt, m, k), b),
        "B shared by every batch": ((m * k, k, 1), (0, n, 1), a, b[0].expand(bt, k, n)),
        "every row of A the same": ((m * k, 0, 1), (k * n, n, 1), a[:, :1].expand(bt, m, k), b),
        "every column of B the same": ((m * k, k, 1), (k * n, n, 0), a, b[:, :, :1].expand(bt, k, n)),
        "every row of B the same": ((m * k, k, 1), (k * n, 0, 1), a, b[:, :1].expand(bt, k, n)),
    }
    for name, (sa, sb, av, bv) in cases.items():
        c = torch.empty((bt, m, n), device="cuda", dtype=dtype)
        check(f(vp(a), vp(b), vp(c), bt, m, n, k, s3(*sa), s3(*sb), s3(m * n, n, 1), 0, None))
        torch.cuda.synchronize()
        exact = av.double() @ bv.double()
        bound = tol * (av.double().abs() @ bv.double().abs()) + 1e-300
        assert bool(((c.double() - exact).abs() <= bound).all()), name
