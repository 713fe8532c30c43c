"""The single-launch sweep kernels (csrc/tape_kernels.cu) against the oracle's literal scalar tape.

  * thin products (M <= 4 rows): forward, both derivative products and the reference's constant-operand sums
    (a constant container carries Index 0 and record_scalar_product records it as a parent, reference
    container_operations.rs:1291-1296, so its would-be gradient is added into derivatives[0]);
  * the same sums for ordinary shapes (column sums / row sums in one launch each);
  * short contractions with a wide output (dA = G * B^T of an output layer), written as the first contribution;
  * tall-skinny A^T * B with its ticketed single-pass reduction.
Tolerance: the north star's |gpu - ref| <= tol * sum |a b| (1e-12 f64, 1e-5 f32), here through allclose with the
bound scaled by the largest derivative; tape length and every Index are compared exactly.
"""
import numpy as np
import pytest

import easy_ml_b200 as eml

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]


def T(names, a):
    return eml.Tensor(list(zip(names, a.shape)), a)


def close(got, want, dtype, scale=None):
    rtol = 2e-4 if dtype == np.float32 else 1e-11
    scale = float(np.abs(want).max()) if scale is None else scale
    return np.allclose(got, want, rtol=rtol, atol=rtol * max(scale, 1e-30))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,k,n,a_const,b_const", [
    (1, 300, 10, True, False),    # ones-row times a tracked matrix: how a container is summed (K >= 256: thin forward)
    (1, 10, 1, False, True),      # tracked row times a ones-column
    (3, 700, 37, False, False),   # both tracked
    (4, 64, 256, False, False),   # widest thin output
    (2, 513, 5, True, False),
    (4, 9, 3, False, True),
])
def test_thin_products_forward_and_whole_backward(orc, dtype, m, k, n, a_const, b_const):
    r = np.random.default_rng(100 + m * k + n)
    a = r.uniform(-1, 1, (m, k)).astype(dtype)
    b = r.uniform(-1, 1, (k, n)).astype(dtype)
    w0 = r.uniform(-1, 1, (2, 3)).astype(dtype)  # the first variable ever created owns tape index 0
    hist = eml.WengertList(dtype)
    tape = orc.Tape(dtype)
    W0, ow0 = eml.RecordTensor.variables(hist, T("pq", w0)), tape.variables(w0)
    if a_const:
        A, oa = eml.RecordTensor.constants(T("mk", a), hist), orc.Tape.constants(a, dtype)
    else:
        A, oa = eml.RecordTensor.variables(hist, T("mk", a)), tape.variables(a)
    if b_const:
        B, ob = eml.RecordTensor.constants(T("kn", b), hist), orc.Tape.constants(b, dtype)
    else:
        B, ob = eml.RecordTensor.variables(hist, T("kn", b)), tape.variables(b)
    Cc = A * B
    oc, oci = tape.matmul(oa, ob)
    assert len(hist) == len(tape) and np.array_equal(Cc.indexes(), oci)
    bound = (1e-5 if dtype == np.float32 else 1e-12) * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))
    assert np.all(np.abs(Cc.numbers().data.astype(np.float64) - oc.astype(np.float64)) <= bound + 1e-300)
    # scale the product so that every output element takes part: loss = sum of squares through elementwise ops
    sq = Cc.elementwise_multiply(Cc)
    osq = tape.binary(34, (oc, oci), (oc, oci), mode=3)
    e = (m - 1, n // 2)
    d = sq.derivatives_for(list(e))
    od = tape.derivatives(osq[1][e])
    if not a_const:
        assert close(d.at_tensor(A).data, od[oa[1]], dtype)
    if not b_const:
        assert close(d.at_tensor(B).data, od[ob[1]], dtype)
    # derivatives[0]: W0[0,0] takes the constant operand's sum (nothing else touches W0)
    assert close(np.array(float(d.at_index(0))), np.array(float(od[0])), dtype, scale=float(np.abs(od).max()))
    assert np.count_nonzero(d.at_tensor(W0).data.ravel()[1:]) == 0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,k,n", [(40, 33, 300), (70, 20, 12), (9, 130, 64), (257, 5, 17)])
def test_constant_operand_sums_single_launch(orc, dtype, m, k, n):
    """constants * variables and variables * constants at ordinary (not thin) shapes: N above and below 256 columns,
    N not dividing 256, more rows than one block takes."""
    r = np.random.default_rng(200 + m + k + n)
    a = r.uniform(-1, 1, (m, k)).astype(dtype)
    b = r.uniform(-1, 1, (k, n)).astype(dtype)
    w0 = r.uniform(-1, 1, (1, 2)).astype(dtype)
    for a_const in (True, False):
        hist = eml.WengertList(dtype)
        tape = orc.Tape(dtype)
        W0, ow0 = eml.RecordTensor.variables(hist, T("pq", w0)), tape.variables(w0)
        if a_const:
            A, oa = eml.RecordTensor.constants(T("mk", a), hist), orc.Tape.constants(a, dtype)
            B, ob = eml.RecordTensor.variables(hist, T("kn", b)), tape.variables(b)
        else:
            A, oa = eml.RecordTensor.variables(hist, T("mk", a)), tape.variables(a)
            B, ob = eml.RecordTensor.constants(T("kn", b), hist), orc.Tape.constants(b, dtype)
        Cc = A * B
        oc, oci = tape.matmul(oa, ob)
        # weight the outputs differently so that the sums are not symmetric: c * c
        sq = Cc.elementwise_multiply(Cc)
        osq = tape.binary(34, (oc, oci), (oc, oci), mode=3)
        # a matmul by ones reduces sq to one number whose derivative reaches every element
        ones_l, ones_r = np.ones((1, m), dtype), np.ones((n, 1), dtype)
        L = eml.RecordTensor.constants(T("um", ones_l), hist)
        R = eml.RecordTensor.constants(T("nv", ones_r), hist)
        total = (L * sq) * R
        ototal = tape.matmul(tape.matmul(orc.Tape.constants(ones_l, dtype), osq), orc.Tape.constants(ones_r, dtype))
        assert len(hist) == len(tape)
        d = total.derivatives_for([0, 0])
        od = tape.derivatives(ototal[1][0, 0])
        tracked, otracked = (B, ob) if a_const else (A, oa)
        assert close(d.at_tensor(tracked).data, od[otracked[1]], dtype)
        assert close(np.array(float(d.at_index(0))), np.array(float(od[0])), dtype, scale=float(np.abs(od).max()))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,k,n", [(128, 512, 10), (200, 96, 3), (64, 65, 16)])
def test_output_layer_backward_short_contraction_and_tall_skinny(orc, dtype, m, k, n):
    """H[m x k] * W[k x n] with n small: dH = G * W^T is a short contraction written as the first contribution to H's
    derivatives (no zero fill), dW = H^T * G the tall-skinny product with its single-pass reduction.  H is itself
    the output of tracked work, so its derivatives are observable."""
    r = np.random.default_rng(300 + m + k + n)
    x = r.uniform(-1, 1, (m, k)).astype(dtype)
    w = r.uniform(-1, 1, (k, n)).astype(dtype)
    hist = eml.WengertList(dtype)
    tape = orc.Tape(dtype)
    X, ox = eml.RecordTensor.variables(hist, T("mk", x)), tape.variables(x)
    W, ow = eml.RecordTensor.variables(hist, T("kn", w)), tape.variables(w)
    H = X * 0.5                       # tracked elementwise work in front of the product
    oh = tape.unary(8, ox, 0.5)
    Y = H * W
    oy = tape.matmul(oh, ow)
    sq = Y.elementwise_multiply(Y)
    osq = tape.binary(34, oy, oy, mode=3)
    ones_l, ones_r = np.ones((1, m), dtype), np.ones((n, 1), dtype)
    total = (eml.RecordTensor.constants(T("um", ones_l), hist) * sq) * eml.RecordTensor.constants(T("nv", ones_r), hist)
    ototal = tape.matmul(tape.matmul(orc.Tape.constants(ones_l, dtype), osq), orc.Tape.constants(ones_r, dtype))
    assert len(hist) == len(tape) and np.array_equal(Y.indexes(), oy[1])
    d = total.derivatives_for([0, 0])
    od = tape.derivatives(ototal[1][0, 0])
    # the constants' sums land in derivatives[0] == X[0,0]: compare everything but that element, then it separately
    gx, wx = d.at_tensor(X).data, od[ox[1]]
    assert close(gx.ravel()[1:], wx.ravel()[1:], dtype)
    assert close(np.array(float(gx[0, 0])), np.array(float(wx[0, 0])), dtype, scale=float(np.abs(od).max()))
    assert close(d.at_tensor(W).data, od[ow[1]], dtype)
    assert close(d.at_tensor(H).data, od[oh[1]], dtype)


def test_short_contraction_entry_point_accumulates_and_handles_ragged_widths():
    """eml_contract_*_dev routes K <= 32, wide dense outputs to the short-contraction kernel: plain and accumulating,
    vector and scalar (N not a multiple of the 16-byte vector) variants, B read through transposed strides."""
    import ctypes as C

    from easy_ml_b200._ffi import check, lib

    torch = pytest.importorskip("torch")
    for dtype, tdt, sfx in ((np.float32, torch.float32, "f32"), (np.float64, torch.float64, "f64")):
        for m, n, k in ((300, 512, 10), (129, 131, 32), (64, 64, 1)):
            r = np.random.default_rng(m + n + k)
            a = r.uniform(-1, 1, (m, k)).astype(dtype)
            bt = r.uniform(-1, 1, (n, k)).astype(dtype)          # B(k, n) = bt[n, k]
            c0 = r.uniform(-1, 1, (m, n)).astype(dtype)
            ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(bt).cuda()
            s3 = lambda *v: (C.c_int64 * 3)(*v)
            vp = lambda t: C.c_void_p(t.data_ptr())
            want = a.astype(np.float64) @ bt.astype(np.float64).T
            bound = (1e-5 if dtype == np.float32 else 1e-12) * (np.abs(a).astype(np.float64) @ np.abs(bt).astype(np.float64).T)
            for accumulate in (0, 1):
                tc = torch.from_numpy(c0.copy()).cuda()
                check(getattr(lib, f"eml_contract_{sfx}_dev")(vp(ta), vp(tb), vp(tc), 1, m, n, k, s3(0, k, 1), s3(0, 1, k),
                                                             s3(0, n, 1), accumulate, None))
                torch.cuda.synchronize()
                assert eml.last_kernel() == "small_k"
                ref = want + (c0.astype(np.float64) if accumulate else 0.0)
                extra = np.abs(c0).astype(np.float64) * (1e-6 if dtype == np.float32 else 1e-15) if accumulate else 0.0
                assert np.all(np.abs(tc.cpu().numpy().astype(np.float64) - ref) <= bound + extra + 1e-300)
