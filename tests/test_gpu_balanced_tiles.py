"""Tile width of the f32 CTA-pair kernel (csrc/gemm_tf32x3.cu, launch_gemm_tf32x3_f32): the N tile of a work item is
whatever the MMA's N field says (a multiple of 16 up to 256), so outputs of a few waves pick the width with the least MMA
work on the busiest CTA pair -- 300 columns as two tiles of 160 instead of 256 + 44 -- provided the split-K partition
stays what it was.  Every element's accumulation chain is then untouched: the bits of the plain tiling
(EML_SGEMM_BALANCE=0), also when the product accumulates into C, on ragged edges, with operands TMA cannot read in place,
and with the non-finite repair riding in the narrow tile's epilogue (compared with the reference's textbook loop,
src/matrices/operations.rs:165-196).  Everything is also held to the north star's bound tol * sum_k |a_k b_k| (1e-5).
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


def s3(*s):
    return (C.c_int64 * 3)(*s)


class env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def contract(a, b, c0=None):
    m, k = a.shape
    n = b.shape[1]
    c = torch.full((m, n), float("nan"), device="cuda") if c0 is None else c0.clone()
    check(lib.eml_contract_f32_dev(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()), 1, m, n, k,
                                   s3(0, *a.stride()), s3(0, *b.stride()), s3(0, n, 1), 0 if c0 is None else 1, None))
    torch.cuda.synchronize()
    return c


# (m, n, k, A stored transposed, what the dispatcher does on 74 CTA pairs)
CASES = [
    (1000, 300, 2048, False, "narrow"),   # 300 columns: two tiles of 160 instead of 256 + 44
    (4096, 784, 512, False, "narrow"),    # 784 columns: four tiles of 208 instead of three of 256 + 16
    (1002, 298, 1030, True, "narrow"),    # ragged everywhere, A read through transposed strides, packed planes for B
    (1024, 768, 1028, False, "plain"),    # a narrower width would change the split-K partition: not taken
    (2048, 2048, 256, False, "plain"),
]


@pytest.mark.parametrize("m,n,k,at,plan", CASES)
@pytest.mark.parametrize("accumulate", [False, True])
def test_balanced_tiles_against_the_plain_tiling_and_the_bound(m, n, k, at, plan, accumulate):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    a = torch.rand(k, m, generator=g, device="cuda").t() - 0.3 if at else torch.rand(m, k, generator=g, device="cuda") - 0.3
    b = torch.rand(k, n, generator=g, device="cuda") - 0.5
    c0 = torch.randn(m, n, generator=g, device="cuda") if accumulate else None
    with env(EML_SGEMM_FUSED=0, EML_SGEMM_BALANCE=1):
        got = contract(a, b, c0)
        assert eml.last_kernel() == "tf32x3_tcgen05"
    with env(EML_SGEMM_FUSED=0, EML_SGEMM_BALANCE=0):
        plain = contract(a, b, c0)
    ref = a.double() @ b.double() + (c0.double() if accumulate else 0)
    bound = TOL * (a.double().abs() @ b.double().abs()) + (1e-6 * c0.double().abs() if accumulate else 0)
    assert not torch.isnan(got).any() and not torch.isnan(plain).any()
    assert ((got.double() - ref).abs() <= bound).all()
    assert ((plain.double() - ref).abs() <= bound).all()
    assert torch.equal(got, plain), "narrower tiles must leave every accumulation chain alone"


def test_non_finite_inputs_with_narrow_tiles():
    """An infinite operand entry gives +-inf like the reference's plain products (not the NaN of 0 * inf in a cross term of
    the split), also when the repair rides in the epilogue of a narrow tile."""
    m, n, k = 1000, 300, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.rand(m, k, generator=g, device="cuda") + 0.1
    b = torch.rand(k, n, generator=g, device="cuda") + 0.1
    a[7, 100] = float("inf")
    b[33, 200] = float("-inf")
    a[250, 3] = float("nan")
    with env(EML_SGEMM_FUSED=0, EML_SGEMM_BALANCE=1):
        got = contract(a, b).cpu().numpy()
    with env(EML_SGEMM_FUSED=0, EML_SGEMM_BALANCE=0):
        plain = contract(a, b).cpu().numpy()
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    for (i, j) in [(7, 5), (7, 200), (100, 200), (250, 9), (0, 0), (999, 299), (500, 159), (500, 160)]:
        want = np.float32(0)
        with np.errstate(all="ignore"):
            acc = an[i, 0] * bn[0, j]
            for kk in range(1, k):
                acc = np.float32(acc + an[i, kk] * bn[kk, j])
            want = acc
        for name, c in (("narrow", got), ("plain", plain)):
            if np.isnan(want):
                assert np.isnan(c[i, j]), (name, i, j)
            elif np.isinf(want):
                assert c[i, j] == want, (name, i, j, c[i, j], want)
            else:
                assert abs(c[i, j] - want) <= 1e-4 * abs(want), (name, i, j)
