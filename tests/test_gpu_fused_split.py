"""The f32 tensor-core GEMM has two operand pipelines: a split/pack pass followed by TMA loads of the packed planes
(gemm_tf32x3.cu), and an in-kernel split of raw f32 tiles (gemm_tf32x3_fused.cu) that the dispatcher picks for
small outputs.  Both use the same split (tf32_split.cuh), issue the same MMAs in the same order and chunk the
accumulation identically, so their results must agree BIT FOR BIT; each is also checked against the north star's
bound  tol * sum_k |a_k b_k|  (tol 1e-5).
"""
import ctypes as C
import os

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = 1e-5


def vp(t):
    return C.c_void_p(t.data_ptr())


def s3(*s):
    return (C.c_int64 * 3)(*s)


class forced:
    """EML_SGEMM_FUSED=1 (in-kernel split wherever it can run) / =0 (always pack) for the calls inside."""

    def __init__(self, value):
        self.value = value

    def __enter__(self):
        self.old = os.environ.get("EML_SGEMM_FUSED")
        os.environ["EML_SGEMM_FUSED"] = self.value

    def __exit__(self, *exc):
        if self.old is None:
            del os.environ["EML_SGEMM_FUSED"]
        else:
            os.environ["EML_SGEMM_FUSED"] = self.old


def stored(x, transposed):
    # same logical (batch, rows, cols) tensor, physically transposed when asked
    return x.transpose(1, 2).contiguous().transpose(1, 2) if transposed else x.contiguous()


def contract(a, b, c0=None):
    batch, m, k = a.shape
    n = b.shape[2]
    c = torch.zeros((batch, m, n), device="cuda", dtype=torch.float32) if c0 is None else c0.clone()
    check(lib.eml_contract_f32_dev(vp(a), vp(b), vp(c), batch, m, n, k, s3(*a.stride()), s3(*b.stride()),
                                   s3(*c.stride()), 0 if c0 is None else 1, None))
    torch.cuda.synchronize()
    return c


SHAPES = [(1, 512, 512, 64), (1, 300, 500, 700), (3, 260, 384, 200), (2, 1024, 1024, 2048), (1, 132, 4100, 36),
          (1, 200, 136, 8192),  # few tiles, long K: deterministic split-K inside both kernels
          (1, 1024, 768, 1028)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_in_kernel_split_matches_packed_bitwise(shape, ta, tb):
    batch, m, n, k = shape
    if (m if ta else k) % 4 or (k if tb else n) % 4:
        pytest.skip("TMA reads raw operands in place only when rows are 16-byte multiples")
    g = torch.Generator(device="cuda").manual_seed(hash((shape, ta, tb)) & 0xFFFF)
    a = stored(torch.randn(batch, m, k, generator=g, device="cuda"), ta)
    b = stored(torch.randn(batch, k, n, generator=g, device="cuda"), tb)
    with forced("0"):
        ref = contract(a, b)
    with forced("1"):
        got = contract(a, b)
    assert eml.last_kernel() == "tf32x3_tcgen05"
    assert torch.equal(ref, got)
    exact = a.double() @ b.double()
    bound = TOL * (a.double().abs() @ b.double().abs())
    assert bool(((got.double() - exact).abs() <= bound).all())


def test_in_kernel_split_accumulates_into_c():
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(1, 512, 320, generator=g, device="cuda")
    b = torch.randn(1, 320, 384, generator=g, device="cuda")
    c0 = torch.randn(1, 512, 384, generator=g, device="cuda")
    with forced("0"):
        ref = contract(a, b, c0)
    with forced("1"):
        got = contract(a, b, c0)
    assert torch.equal(ref, got)
    assert not torch.equal(got, contract(a, b))


def test_in_kernel_split_special_values():
    """inf / NaN / FLT_MAX-sized inputs take the split's slow path; both pipelines must treat them alike."""
    g = torch.Generator(device="cuda").manual_seed(12)
    a = torch.randn(1, 256, 64, generator=g, device="cuda")
    b = torch.randn(1, 64, 256, generator=g, device="cuda")
    fmax = float(np.finfo(np.float32).max)
    a[0, 3, 5] = fmax
    a[0, 7, 1] = -fmax
    a[0, 9, 0] = float("nan")
    b[0, 2, 200] = 3.0e38
    b[0, 4, 17] = float("inf")
    with forced("0"):
        ref = contract(a, b)
    with forced("1"):
        got = contract(a, b)
    assert torch.equal(torch.isnan(ref), torch.isnan(got))
    ok = ~torch.isnan(ref)
    assert torch.equal(ref[ok], got[ok])
    assert bool(torch.isnan(got[0, 9]).all())          # the NaN row stays NaN
    assert bool(torch.isfinite(got[0, 0]).sum() >= 254)  # ordinary rows are untouched except the inf / 3e38 columns


def test_dispatcher_picks_in_kernel_split_for_small_batched_outputs():
    """BASELINE config 3 (batch x 1024^3) is where the pack pass costs a third of the time: the default dispatch
    must produce the in-kernel-split result there (same bits either way, so ask the library which kernel ran)."""
    a = torch.randn(8, 1024, 1024, device="cuda")
    b = torch.randn(8, 1024, 1024, device="cuda")
    os.environ.pop("EML_SGEMM_FUSED", None)
    before = eml.kernel_launches()
    c = contract(a, b)
    launches = eml.kernel_launches() - before
    # the fused product + the early-exit repair of non-finite inputs; the packed path would need two pack passes more
    assert launches == 2, f"expected the fused launch and its repair, saw {launches} (pack passes ran)"
    with forced("0"):
        assert torch.equal(c, contract(a, b))


def test_broadcast_operand_falls_back_to_the_packed_path():
    """`batch row k, k col -> batch row col`: B has no batch dimension (batch stride 0).  TMA cannot describe a
    zero stride, so the in-kernel-split kernel must decline and the packed path (whose pack pass reads any strides)
    must produce the product -- whatever the dispatcher's cost model preferred."""
    g = torch.Generator(device="cuda").manual_seed(31)
    bt, m, k, n = 6, 512, 1024, 512
    a = torch.randn(bt, m, k, generator=g, device="cuda")
    w = torch.randn(k, n, generator=g, device="cuda")
    for mode in (None, "1"):
        if mode is None:
            os.environ.pop("EML_SGEMM_FUSED", None)
        else:
            os.environ["EML_SGEMM_FUSED"] = mode
        c = torch.empty(bt, m, n, device="cuda")
        check(lib.eml_contract_f32_dev(vp(a), vp(w), vp(c), bt, m, n, k, s3(*a.stride()), s3(0, n, 1), s3(*c.stride()),
                                       0, None))
        torch.cuda.synchronize()
        exact = a.double() @ w.double()
        bound = TOL * (a.double().abs() @ w.double().abs())
        assert bool(((c.double() - exact).abs() <= bound).all())
    os.environ.pop("EML_SGEMM_FUSED", None)
    # f64: the TMA kernel declines the zero stride too (the cp.async kernel takes it)
    c64 = torch.empty(bt, m, n, device="cuda", dtype=torch.float64)
    a64, w64 = a.double(), w.double()
    check(lib.eml_contract_f64_dev(vp(a64), vp(w64), vp(c64), bt, m, n, k, s3(*a64.stride()), s3(0, n, 1),
                                   s3(*c64.stride()), 0, None))
    torch.cuda.synchronize()
    assert eml.last_kernel() == "dmma_f64"
    assert bool(((c64 - exact).abs() <= 1e-12 * (a64.abs() @ w64.abs())).all())
    # the named-dimension front end reaches the same case
    x = eml.Tensor([("batch", bt), ("row", m), ("k", k)], a.cpu().numpy())
    y = eml.Tensor([("k", k), ("col", n)], w.cpu().numpy())
    out = eml.Einsum.with_2(x, y).to(["batch", "row", "col"])
    assert out.shape() == [("batch", bt), ("row", m), ("col", n)]
    assert np.allclose(out.data, exact.cpu().numpy(), rtol=0, atol=float(bound.max()))


class direct:
    """EML_SGEMM_DIRECT=1 (the CTA-pair kernel reads operands TMA can describe in place: raw matrix as the big part,
    a small-part plane with the matrix's own layout) / =0 (packed big + small planes) for the calls inside."""

    def __init__(self, value):
        self.value = value

    def __enter__(self):
        self.old = os.environ.get("EML_SGEMM_DIRECT")
        os.environ["EML_SGEMM_DIRECT"] = self.value

    def __exit__(self, *exc):
        if self.old is None:
            del os.environ["EML_SGEMM_DIRECT"]
        else:
            os.environ["EML_SGEMM_DIRECT"] = self.old


@pytest.mark.parametrize("shape", SHAPES + [(1, 2048, 2304, 520), (1, 784, 512, 8192)])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
def test_operands_read_in_place_match_packed_planes_bitwise(shape, ta, tb):
    """The packed path's big plane holds trunc(x) -- exactly what kind::tf32 makes of the raw word -- so reading the
    caller's matrix itself as the big operand (K-major or MN-major tiles by its contiguous index) gives the same bits,
    with one small-part plane per operand instead of two padded, possibly transposed planes.  Ragged shapes (zero fill
    by TMA), batches, split-K, and rows TMA cannot read in place (not 16-byte multiples: that operand stays packed)."""
    batch, m, n, k = shape
    g = torch.Generator(device="cuda").manual_seed(hash((shape, ta, tb, "d")) & 0xFFFF)
    a = stored(torch.randn(batch, m, k, generator=g, device="cuda"), ta)
    b = stored(torch.randn(batch, k, n, generator=g, device="cuda"), tb)
    with forced("0"):
        with direct("0"):
            before = eml.kernel_launches()
            ref = contract(a, b)
            packed_launches = eml.kernel_launches() - before
        with direct("1"):
            before = eml.kernel_launches()
            got = contract(a, b)
            direct_launches = eml.kernel_launches() - before
    assert eml.last_kernel() == "tf32x3_tcgen05"
    assert torch.equal(ref, got)
    assert direct_launches <= packed_launches
    exact = a.double() @ b.double()
    bound = TOL * (a.double().abs() @ b.double().abs())
    assert bool(((got.double() - exact).abs() <= bound).all())


def test_operands_read_in_place_special_values_and_accumulate():
    g = torch.Generator(device="cuda").manual_seed(13)
    a = torch.randn(1, 512, 320, generator=g, device="cuda")
    b = torch.randn(1, 320, 384, generator=g, device="cuda")
    c0 = torch.randn(1, 512, 384, generator=g, device="cuda")
    fmax = float(np.finfo(np.float32).max)
    a[0, 3, 5] = fmax
    a[0, 9, 0] = float("nan")
    b[0, 4, 17] = float("inf")
    b[0, 2, 200] = -3.0e38
    with forced("0"):
        with direct("0"):
            ref, ref_acc = contract(a, b), contract(a, b, c0)
        with direct("1"):
            got, got_acc = contract(a, b), contract(a, b, c0)
    for r, x in ((ref, got), (ref_acc, got_acc)):
        assert torch.equal(torch.isnan(r), torch.isnan(x))
        ok = ~torch.isnan(r)
        assert torch.equal(r[ok], x[ok])
