"""Parity at BASELINE.json's FULL sizes (configs 2-5), where the oracle cannot evaluate the whole product:
sampled rows against the oracle's sequential sum, plus size-independent properties -- a checksum of
checksums (C*1 == A*(B*1)), linearity in B, row-shard invariance (bit-exact), run-to-run bit identity, and
for the NN the closed-form tape length / indexes and a float64 re-derivation of loss and gradients.
Tolerances are the north star's: |gpu - ref| <= tol * sum_k |a_k b_k|, tol = 1e-12 (f64) / 1e-5 (f32).
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = {torch.float64: 1e-12, torch.float32: 1e-5}


def vp(t):
    return C.c_void_p(t.data_ptr())


def s3(*s):
    return (C.c_int64 * 3)(*s)


def dev_gemm(a, b, c=None, accumulate=False):
    m, k = a.shape
    n = b.shape[1]
    if c is None:
        c = torch.empty((m, n), device=a.device, dtype=a.dtype)
    f = lib.eml_contract_f64_dev if a.dtype == torch.float64 else lib.eml_contract_f32_dev
    check(f(vp(a), vp(b), vp(c), 1, m, n, k, s3(0, a.stride(0), 1), s3(0, b.stride(0), 1), s3(0, c.stride(0), 1),
            1 if accumulate else 0, None))
    torch.cuda.synchronize()
    return c


def uniform(shape, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(0x5EED0000 + seed)
    return torch.rand(shape, generator=g, device="cuda", dtype=dtype) * 2 - 1


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_config2_8192_square(orc, dtype):
    n = 8192
    a, b = uniform((n, n), dtype, 2), uniform((n, n), dtype, 3)
    c = dev_gemm(a, b)
    assert eml.last_kernel() == ("dmma_f64_tma" if dtype == torch.float64 else "tf32x3_tcgen05")
    tol = TOL[dtype]
    # sampled rows against the oracle (reference operation order)
    bh = b.cpu().numpy()
    for i in (0, 127, 128, 4095, 6000, n - 1):
        ah = a[i:i + 1].cpu().numpy()
        ref = orc.matrix_mul(ah, bh, threads=8).astype(np.float64)
        bound = tol * (np.abs(ah).astype(np.float64) @ np.abs(bh).astype(np.float64))
        assert np.all(np.abs(c[i:i + 1].cpu().numpy().astype(np.float64) - ref) <= bound), f"row {i}"
    # checksum of checksums in f64: C*1 == A*(B*1); each entry of C*1 sums n outputs, so the bound does too
    ones = torch.ones((n, 1), device="cuda", dtype=torch.float64)
    lhs = c.double() @ ones
    rhs = a.double() @ (b.double() @ ones)
    bound = tol * (a.abs().double() @ (b.abs().double() @ ones))
    assert bool(((lhs - rhs).abs() <= bound).all())
    # linearity in B on a row slab: A*(B + B2) == A*B + A*B2 within the bound of the three products
    b2 = uniform((n, n), dtype, 4)
    rows = slice(1024, 1024 + 512)
    both = dev_gemm(a[rows], b + b2)
    second = dev_gemm(a[rows], b2)
    bound = 3 * tol * (a[rows].abs().double() @ (b.abs() + b2.abs()).double())
    assert bool(((both.double() - (c[rows].double() + second.double())).abs() <= bound).all())
    # run-to-run bit identity, and (f64) invariance to row partitioning: a slab launch gives the same bits
    again = dev_gemm(a, b)
    assert hashlib.sha256(again.cpu().numpy().tobytes()).digest() == hashlib.sha256(c.cpu().numpy().tobytes()).digest()
    if dtype == torch.float64:
        assert torch.equal(dev_gemm(a[3072:5120], b), c[3072:5120])


def test_config3_batched_named_contraction(orc):
    """f32 ["batch","row","k"] x ["batch","k","col"], batch 64, 1024^3, through the named-dimension host API;
    plus a permuted-name variant (strides instead of a transpose)."""
    bt, m = 64, 1024
    r = np.random.default_rng(0x5EED0003)
    x = r.uniform(-1, 1, (bt, m, m)).astype(np.float32)
    y = r.uniform(-1, 1, (bt, m, m)).astype(np.float32)
    X = eml.Tensor([("batch", bt), ("row", m), ("k", m)], x)
    Y = eml.Tensor([("batch", bt), ("k", m), ("col", m)], y)
    out = eml.Einsum.with_2(X, Y).to(["batch", "row", "col"])
    assert eml.last_kernel() == "tf32x3_tcgen05" and out.shape() == [("batch", bt), ("row", m), ("col", m)]
    for b_, i in ((0, 0), (17, 511), (63, 1023)):
        ref = orc.einsum2(x[b_:b_ + 1, i:i + 1], ["batch", "row", "k"], y[b_:b_ + 1], ["batch", "k", "col"],
                          ["batch", "row", "col"])[0]
        bound = 1e-5 * (np.abs(x[b_, i:i + 1]).astype(np.float64) @ np.abs(y[b_]).astype(np.float64))
        assert np.all(np.abs(out.data[b_, i:i + 1].astype(np.float64) - ref.astype(np.float64)) <= bound)
    # checksum per batch in f64
    lhs = out.data.astype(np.float64).sum(axis=2)
    rhs = np.einsum("brk,bk->br", x.astype(np.float64), y.astype(np.float64).sum(axis=2))
    bound = 1e-5 * np.einsum("brk,bk->br", np.abs(x).astype(np.float64), np.abs(y).astype(np.float64).sum(axis=2))
    assert np.all(np.abs(lhs - rhs) <= bound)
    # permuted dimension order of the left operand and of the output: same numbers, no transpose pass
    Xp = eml.Tensor([("row", m), ("batch", bt), ("k", m)], np.ascontiguousarray(x.transpose(1, 0, 2)))
    outp = eml.Einsum.with_2(Xp, Y).to(["row", "batch", "col"])
    assert outp.names() == ["row", "batch", "col"]
    assert np.array_equal(outp.data.transpose(1, 0, 2), out.data)


def test_config4_nn_step_batch_8192():
    """784-512-10, batch 8192, f32 RecordTensor: tape length and indexes by closed form (bit-exact), loss and
    weight gradients against a float64 re-derivation of the same network."""
    batch = 8192
    f32 = np.float32
    r = np.random.default_rng(0x5EED0004)
    x = r.uniform(0, 1, (batch, 784)).astype(f32)
    labels = np.zeros((batch, 10), f32)
    labels[np.arange(batch), r.integers(0, 10, batch)] = 1.0
    w1 = r.uniform(-0.05, 0.05, (784, 512)).astype(f32)
    w2 = r.uniform(-0.05, 0.05, (512, 10)).astype(f32)
    T = lambda names, a: eml.Tensor(list(zip(names, a.shape)), a)
    hist = eml.WengertList(f32)
    W1 = eml.RecordTensor.variables(hist, T(["i", "h"], w1))
    W2 = eml.RecordTensor.variables(hist, T(["h", "o"], w2))
    X = eml.RecordTensor.constants(T(["n", "i"], x), hist)
    Y = eml.RecordTensor.constants(T(["n", "o"], labels), hist)
    L = eml.RecordTensor.constants(T(["u", "n"], np.ones((1, batch), f32)), hist)
    R = eml.RecordTensor.constants(T(["o", "v"], np.ones((10, 1), f32)), hist)
    sigmoid = lambda z: ((-z).exp() + 1.0).div_swapped(1.0)
    n0 = len(hist)
    assert n0 == 784 * 512 + 512 * 10 and W2.indexes()[0, 0] == 784 * 512
    z1 = X * W1
    # a tracked matmul appends M*N*(2K-1) entries; C[i,j] sits at base + (i*N+j)*(2K-1) + 2K-2
    assert len(hist) == n0 + batch * 512 * (2 * 784 - 1)
    idx = z1.indexes()
    assert idx[0, 0] == n0 + 2 * 784 - 2 and idx[batch - 1, 511] == n0 + (batch * 512 - 1) * (2 * 784 - 1) + 2 * 784 - 2
    out = sigmoid(sigmoid(z1) * W2)
    diff = out - Y
    loss = ((L * diff.elementwise_multiply(diff)) * R) / float(batch)
    expected_len = (n0 + batch * 512 * (2 * 784 - 1) + 4 * batch * 512 + batch * 10 * (2 * 512 - 1) + 4 * batch * 10
                    + 2 * batch * 10 + 10 * (2 * batch - 1) + (2 * 10 - 1) + 1)
    assert len(hist) == expected_len  # ~6.7e9 entries the reference would have to store (158 GB)
    d = loss.derivatives_for([0, 0])
    g1, g2 = d.at_tensor(W1).data, d.at_tensor(W2).data
    # float64 re-derivation
    X64, W164, W264, Y64 = (v.astype(np.float64) for v in (x, w1, w2, labels))
    sg = lambda v: 1.0 / (1.0 + np.exp(-v))
    h = sg(X64 @ W164)
    o = sg(h @ W264)
    want_loss = ((o - Y64) ** 2).sum() / batch
    do = 2.0 * (o - Y64) / batch * o * (1 - o)
    want_g2 = h.T @ do
    dh = (do @ W264.T) * h * (1 - h)
    want_g1 = X64.T @ dh
    assert abs(float(loss.numbers().data[0, 0]) - want_loss) <= 1e-5 * want_loss
    # the reference's constant-operand quirk routes the constants' would-be gradient into derivatives[0],
    # i.e. into w1[0,0]'s slot (container_operations.rs:1291-1296): that element is checked against its own closed form
    mask = np.ones_like(g1, dtype=bool)
    mask[0, 0] = False
    scale1 = np.abs(X64).T @ np.abs(dh)
    assert np.all(np.abs(g1.astype(np.float64) - want_g1)[mask] <= (2e-5 * scale1 + 1e-9)[mask])
    # derivatives[0] = dW1[0,0]
    #   + sum_ik (dZ1 W1^T)[i,k]           X is a constant operand of X * W1: its whole would-be gradient
    #   + sum_kj (dS^T sq)...= loss        L (ones row) is a constant operand of L * sq: sum_j dS[j] colsum(sq)[j], dS = 1/batch
    #   + sum_k  S[k] dT = loss            R (ones column) is a constant operand of S * R: dT = 1/batch
    want_d0 = want_g1[0, 0] + float((dh @ W164.T).sum()) + 2.0 * want_loss
    scale_d0 = scale1[0, 0] + float((np.abs(dh) @ np.abs(W164).T).sum()) + 2.0 * want_loss
    assert abs(float(g1[0, 0]) - want_d0) <= 2e-5 * scale_d0
    assert abs(float(d.at_index(0)) - want_d0) <= 2e-5 * scale_d0
    assert abs(want_d0 - want_g1[0, 0]) > 100 * 2e-5 * scale1[0, 0]   # the quirk is far outside dW1[0,0]'s own tolerance
    assert np.all(np.abs(g2.astype(np.float64) - want_g2) <= 2e-5 * (np.abs(h).T @ np.abs(do)) + 1e-9)


def test_config5_slab_of_32768(orc):
    """One rank's row slab of the 32768^2 product (the per-GPU unit of config 5) on sampled rows, plus the
    K-panelled accumulation the sharded step uses: bit-identical to the single launch."""
    n, rows = 32768, 512
    a, b = uniform((rows, n), torch.float64, 5), uniform((n, n), torch.float64, 6)
    c = dev_gemm(a, b)
    bh = b.cpu().numpy()
    for i in (0, rows - 1):
        ah = a[i:i + 1].cpu().numpy()
        ref = orc.matrix_mul(ah, bh, threads=8)
        bound = 1e-12 * (np.abs(ah) @ np.abs(bh))
        assert np.all(np.abs(c[i:i + 1].cpu().numpy() - ref) <= bound), f"row {i}"
    from easy_ml_b200.sharded import k_panels

    c2 = torch.empty_like(c)
    for j, (k0, k1) in enumerate(k_panels(n, 6)):
        dev_gemm(a[:, k0:k1], b[k0:k1], c2, accumulate=j > 0)
    assert torch.equal(c, c2)
