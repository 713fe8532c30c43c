"""Parity of the CUDA path against the CPU oracle, through the C ABI (run with -m gpu on a B200).

Bars (BASELINE.json north_star): shapes / names / tape indexes bit-exact; numbers within
|gpu - ref| <= tol * sum_k |a_k b_k| with tol = 1e-12 (f64) / 1e-5 (f32); small-integer inputs and the
`exact` / generic-einsum kernels must match the oracle bit for bit.
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200 import _ffi
from easy_ml_b200._ffi import check, lib

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}
DTYPES = [np.float32, np.float64]


def rng(seed):
    return np.random.default_rng(0x5EED0000 + seed)


def uniform(r, shape, dtype):
    return r.uniform(-1.0, 1.0, size=shape).astype(dtype)


def ints(r, shape, dtype):
    return r.integers(-8, 9, size=shape).astype(dtype)


def assert_within(got, a, b, ref, dtype, what=""):
    """|got - ref| <= tol * (|a| @ |b|) elementwise (the K-scaled bound of the north star)."""
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    bound = TOL[np.dtype(dtype)] * scale + np.finfo(dtype).tiny
    worst = np.max(err / bound)
    assert worst <= 1.0, f"{what}: error {worst:.3g}x the allowed bound"


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def gemm(a, b, exact=False):
    sfx = "f32" if a.dtype == np.float32 else "f64"
    c = np.empty((a.shape[0], b.shape[1]), a.dtype)
    f = getattr(lib, f"eml_gemm_exact_{sfx}" if exact else f"eml_gemm_{sfx}")
    check(f(ptr(a), ptr(b), ptr(c), a.shape[0], b.shape[1], a.shape[1], a.shape[1], b.shape[1], b.shape[1]))
    return c


# ------------------------------------------------------------------ reference goldens on the GPU
@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_goldens_exact(golden, dtype):
    for g in golden["matmul"]:
        got = eml.Matrix(g["a"], dtype) * eml.Matrix(g["b"], dtype)
        assert got == eml.Matrix(g["expect"], dtype), g["cite"]
    for g in golden["tensor_matmul"]:
        a = eml.Tensor(list(zip(g["a_names"], np.shape(g["a"]))), g["a"], dtype)
        b = eml.Tensor(list(zip(g["b_names"], np.shape(g["b"]))), g["b"], dtype)
        want = eml.Tensor(list(zip(g["expect_names"], np.shape(g["expect"]))), g["expect"], dtype)
        assert a * b == want, g["cite"]
    for g in golden["einsum2"]:
        a = eml.Tensor(list(zip(g["a_names"], np.shape(g["a"]))), g["a"], dtype)
        b = eml.Tensor(list(zip(g["b_names"], np.shape(g["b"]))), g["b"], dtype)
        got = eml.Einsum.with_2(a, b).to(g["out_names"])
        assert got.names() == g["out_names"]
        assert np.array_equal(got.data, np.array(g["expect"], dtype)), g["cite"]
        assert eml.last_kernel() == "einsum_generic"


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_einsum_equivalences(golden, orc, dtype):
    for g in golden["einsum2_equiv"]:
        a = eml.Tensor(list(zip(g["a_names"], np.shape(g["a"]))), g["a"], dtype)
        b = eml.Tensor(list(zip(g["b_names"], np.shape(g["b"]))), g["b"], dtype)
        got = eml.Einsum.with_2(a, b).to(g["out_names"])
        want = orc.einsum2(a.data, g["a_names"], b.data, g["b_names"], g["out_names"])
        assert np.array_equal(got.data, want), g["cite"]  # small integers: exact in any order


# ------------------------------------------------------------------ GEMM vs oracle
SHAPES = [(1, 1, 1), (3, 2, 3), (17, 33, 5), (64, 64, 64), (100, 70, 130), (256, 256, 256), (129, 257, 131),
          (512, 10, 784), (10, 512, 300), (300, 200, 1)]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_gemm_matches_oracle(orc, dtype, m, n, k):
    r = rng(m * 7 + n * 3 + k)
    a, b = uniform(r, (m, k), dtype), uniform(r, (k, n), dtype)
    ref = orc.matrix_mul(a, b, threads=4)
    assert_within(gemm(a, b), a, b, ref, dtype, f"gemm {m}x{n}x{k}")
    # exact mode: the reference's operation order, bit for bit
    assert np.array_equal(gemm(a, b, exact=True), ref)
    # small integers: exact regardless of summation order
    ai, bi = ints(r, (m, k), dtype), ints(r, (k, n), dtype)
    assert np.array_equal(gemm(ai, bi), orc.matrix_mul(ai, bi, threads=4))


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_c1_config_256(orc, dtype):
    """BASELINE config 1: 256x256 (the reference's own CPU-runnable case)."""
    r = rng(1)
    a, b = uniform(r, (256, 256), dtype), uniform(r, (256, 256), dtype)
    got = (eml.Matrix(a) * eml.Matrix(b)).data
    assert_within(got, a, b, orc.matrix_mul(a, b, threads=4), dtype, "C1")
    t = eml.Tensor([("r", 256), ("c", 256)], a) * eml.Tensor([("x", 256), ("y", 256)], b)
    assert t.names() == ["r", "y"] and np.array_equal(t.data, got)  # Matrix and Tensor entry points share the kernel


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_tensor_core_kernels_sampled_rows(orc, dtype):
    """Sizes that dispatch to the tensor-core kernels; the oracle checks a sample of rows."""
    r = rng(2)
    m, n, k = 1024, 768, 1536
    a, b = uniform(r, (m, k), dtype), uniform(r, (k, n), dtype)
    got = gemm(a, b)
    assert eml.last_kernel() in ("dmma_f64_tma", "dmma_f64", "tf32x3_tcgen05", "simt")
    rows = [0, 1, 127, 128, 511, 640, 1023]
    for i in rows:
        ref = orc.matrix_mul(a[i:i + 1], b)
        assert_within(got[i:i + 1], a[i:i + 1], b, ref, dtype, f"row {i}")
    ai, bi = ints(r, (m, k), dtype), ints(r, (k, n), dtype)
    gi = gemm(ai, bi)
    assert np.array_equal(gi, (ai.astype(np.float64) @ bi.astype(np.float64)).astype(dtype))


def contract_strided(a_buf, b_buf, batch, m, n, k, sa, sb, dtype=np.float64):
    """C[b,m,n] = sum_k A[b,m,k] B[b,k,n] through the host-pointer ABI with explicit element strides."""
    sfx = "f32" if dtype == np.float32 else "f64"
    c = np.empty((batch, m, n), dtype)
    s3 = lambda s: (C.c_int64 * 3)(*s)
    check(getattr(lib, f"eml_contract_{sfx}")(ptr(a_buf), ptr(b_buf), ptr(c), batch, m, n, k, s3(sa), s3(sb),
                                              s3((m * n, n, 1))))
    return c


@pytest.mark.parametrize("a_kin", [True, False])
@pytest.mark.parametrize("b_kin", [True, False])
@pytest.mark.parametrize("batch,m,n,k", [(1, 128, 128, 128), (1, 384, 260, 132), (3, 200, 136, 132), (1, 1000, 520, 1028)])
def test_dgemm_tma_all_layouts_ragged(orc, monkeypatch, a_kin, b_kin, batch, m, n, k):
    """The TMA-staged DMMA kernel: every operand orientation (so the backward products dC*B^T and A^T*dC
    run without a pack pass), ragged M/N/K tails zero-filled by TMA, batched.  Checked against the oracle
    on sampled rows and bit for bit against the cp.async DMMA kernel (same k-ordered chain per element)."""
    r = rng(40 + m + n + k)
    a = uniform(r, (batch, m, k), np.float64)
    b = uniform(r, (batch, k, n), np.float64)
    a_buf = np.ascontiguousarray(a if a_kin else a.transpose(0, 2, 1))
    b_buf = np.ascontiguousarray(b.transpose(0, 2, 1) if b_kin else b)
    sa = (m * k, k, 1) if a_kin else (m * k, 1, m)
    sb = (k * n, 1, k) if b_kin else (k * n, n, 1)
    monkeypatch.delenv("EML_DGEMM_KERNEL", raising=False)
    got = contract_strided(a_buf, b_buf, batch, m, n, k, sa, sb)
    assert eml.last_kernel() == "dmma_f64_tma"
    for bi in range(batch):
        for i in sorted({0, 1, m // 2, m - 1}):
            ref = orc.matrix_mul(np.ascontiguousarray(a[bi, i:i + 1]), np.ascontiguousarray(b[bi]))
            assert_within(got[bi, i:i + 1], a[bi, i:i + 1], b[bi], ref, np.float64, f"batch {bi} row {i}")
    assert_within(got.reshape(batch * m, n)[:m], a[0], b[0], a[0] @ b[0], np.float64, "vs numpy")
    monkeypatch.setenv("EML_DGEMM_KERNEL", "cpasync")
    other = contract_strided(a_buf, b_buf, batch, m, n, k, sa, sb)
    assert eml.last_kernel() == "dmma_f64"
    assert np.array_equal(got, other)


@pytest.mark.parametrize("a_kin", [True, False])
@pytest.mark.parametrize("b_kin", [True, False])
@pytest.mark.parametrize("batch,m,n,k", [(1, 256, 256, 256), (1, 384, 260, 196), (3, 300, 264, 224), (1, 1000, 520, 1028)])
@pytest.mark.parametrize("variant", ["pair", "1cta"])
def test_sgemm_tf32x3_all_layouts_ragged(orc, monkeypatch, variant, a_kin, b_kin, batch, m, n, k):
    """The tcgen05 3xTF32 kernel: every operand orientation through the split/pack pass, ragged M/N/K
    (zero padded by the pack), batched, split-K on the small grids.  f32-level accuracy: within
    1e-5 * sum|a_k b_k| of the oracle's sequential sum.  Both kernels: the persistent CTA-pair kernel
    (cta_group::2, double-buffered TMEM) and the single-CTA kernel it supersedes for M, N > 128."""
    if variant == "1cta":
        monkeypatch.setenv("EML_SGEMM_KERNEL", "1cta")
    else:
        monkeypatch.delenv("EML_SGEMM_KERNEL", raising=False)
    r = rng(80 + m + n + k)
    a = uniform(r, (batch, m, k), np.float32)
    b = uniform(r, (batch, k, n), np.float32)
    a_buf = np.ascontiguousarray(a if a_kin else a.transpose(0, 2, 1))
    b_buf = np.ascontiguousarray(b.transpose(0, 2, 1) if b_kin else b)
    sa = (m * k, k, 1) if a_kin else (m * k, 1, m)
    sb = (k * n, 1, k) if b_kin else (k * n, n, 1)
    got = contract_strided(a_buf, b_buf, batch, m, n, k, sa, sb, np.float32)
    assert eml.last_kernel() == "tf32x3_tcgen05"
    for bi in range(batch):
        for i in sorted({0, 1, m // 2, m - 1}):
            ref = orc.matrix_mul(np.ascontiguousarray(a[bi, i:i + 1]), np.ascontiguousarray(b[bi]))
            assert_within(got[bi, i:i + 1], a[bi, i:i + 1], b[bi], ref, np.float32, f"batch {bi} row {i}")
        exact = a[bi].astype(np.float64) @ b[bi].astype(np.float64)
        assert_within(got[bi], a[bi], b[bi], exact, np.float32, f"batch {bi} vs f64")
    # small integers are exact TF32 numbers with a zero small part: any summation order gives the same bits
    ai, bi_ = ints(r, (batch, m, k), np.float32), ints(r, (batch, k, n), np.float32)
    gi = contract_strided(ai, bi_, batch, m, n, k, (m * k, k, 1), (k * n, n, 1), np.float32)
    assert np.array_equal(gi, np.matmul(ai.astype(np.float64), bi_.astype(np.float64)).astype(np.float32))


@pytest.mark.parametrize("m,n,k", [(256, 384, 16384), (128, 256, 8192), (300, 100, 4096)])
def test_sgemm_same_sign_long_k(m, n, k):
    """Worst case for an accumulator that rounds toward zero: all products positive, long K.  The tensor core's
    f32 accumulation truncates, so the kernels bound every TMEM accumulation chain to 512 k and combine the
    partials with IEEE adds; the result must stay within 1e-5 * sum|a_k b_k| (here: of the value itself)."""
    r = rng(1200 + m + n + k)
    a = r.uniform(0.0, 1.0, (m, k)).astype(np.float32)
    b = r.uniform(0.0, 1.0, (k, n)).astype(np.float32)
    got = gemm(a, b)
    assert eml.last_kernel() == "tf32x3_tcgen05"
    exact = a.astype(np.float64) @ b.astype(np.float64)
    assert_within(got, a, b, exact, np.float32, f"same-sign {m}x{n}x{k}")
    # the tight version: f32-level relative error on every element (the sequential f32 reference itself is ~1e-5 off)
    assert np.max(np.abs(got.astype(np.float64) - exact) / exact) < 5e-6


def test_sgemm_tf32x3_wide_tile_accumulate_device():
    """128x256 tiles (enough tiles to fill the machine) and the C += A*B epilogue, device-resident."""
    import torch

    m, n, k = 2048, 2304, 328
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.rand((m, k), generator=g, device="cuda") * 2 - 1
    b = torch.rand((k, n), generator=g, device="cuda") * 2 - 1
    c0 = torch.rand((m, n), generator=g, device="cuda")
    c = c0.clone()
    s3 = lambda *s: (C.c_int64 * 3)(*s)
    vp = lambda t: C.c_void_p(t.data_ptr())
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.eml_contract_f32_dev(vp(a), vp(b), vp(c), 1, m, n, k, s3(0, k, 1), s3(0, n, 1), s3(0, n, 1), 1, sp))
    torch.cuda.synchronize()
    assert eml.last_kernel() == "tf32x3_tcgen05"
    want = c0.double() + a.double() @ b.double()
    bound = 1e-5 * (a.abs().double() @ b.abs().double()) + 1e-6 * c0.abs().double()
    assert bool(((c.double() - want).abs() <= bound).all())
    c2 = torch.empty_like(c)
    check(lib.eml_contract_f32_dev(vp(a), vp(b), vp(c2), 1, m, n, k, s3(0, k, 1), s3(0, n, 1), s3(0, n, 1), 0, sp))
    torch.cuda.synchronize()
    assert bool(((c2.double() - a.double() @ b.double()).abs() <= 1e-5 * (a.abs().double() @ b.abs().double())).all())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,n,k", [(1, 10, 8192), (512, 10, 8192), (8, 8, 4096), (200, 136, 2048), (10, 512, 1000)])
def test_split_k_skinny_outputs(orc, dtype, m, n, k):
    """Few output tiles + long K (the NN's loss reduction and dW2 = H^T dY): deterministic split-K through
    the batch dimension and a fixed-order reduction.  Within tolerance of the oracle, bit-identical run to run."""
    r = rng(300 + m + n + k)
    a, b = uniform(r, (m, k), dtype), uniform(r, (k, n), dtype)
    got = gemm(a, b)
    assert_within(got, a, b, orc.matrix_mul(a, b, threads=4), dtype, f"split-K {m}x{n}x{k}")
    assert np.array_equal(got, gemm(a, b))
    ai, bi = ints(r, (m, k), dtype), ints(r, (k, n), dtype)
    assert np.array_equal(gemm(ai, bi), orc.matrix_mul(ai, bi, threads=4))


@pytest.mark.parametrize("dtype", DTYPES)
def test_host_gemm_pipelined_copies(orc, dtype):
    """Large host-pointer GEMM: the three-stream pipelined path (B in K panels under an accumulating first
    row block, then row panels), with leading dimensions larger than the rows.  f64 must equal the
    device-resident single launch bit for bit (same k-ordered chain per element)."""
    import torch

    r = rng(500)
    m, n, k = 4224, 2048, 2080
    big_a, big_b = uniform(r, (m, k + 16), dtype), uniform(r, (k, n + 8), dtype)
    a, b = big_a[:, :k], big_b[:, :n]
    sfx = "f32" if dtype == np.float32 else "f64"
    c = np.full((m, n + 24), 3.0, dtype)
    check(getattr(lib, f"eml_gemm_{sfx}")(ptr(big_a), ptr(big_b), ptr(c), m, n, k, k + 16, n + 8, n + 24))
    assert np.all(c[:, n:] == 3.0)  # elements outside the C view are untouched
    for i in (0, 1, 1663, 1664, 1800, 2751, 2752, m - 1):
        ref = orc.matrix_mul(np.ascontiguousarray(a[i:i + 1]), np.ascontiguousarray(b))
        assert_within(c[i:i + 1, :n], a[i:i + 1], b, ref, dtype, f"row {i}")
    exact = a.astype(np.float64) @ b.astype(np.float64)
    assert_within(c[:, :n], a, b, exact, dtype, "whole product vs f64")
    if dtype == np.float64:
        ta, tb = torch.from_numpy(np.ascontiguousarray(a)).cuda(), torch.from_numpy(np.ascontiguousarray(b)).cuda()
        tc = torch.empty((m, n), dtype=torch.float64, device="cuda")
        vp = lambda t: C.c_void_p(t.data_ptr())
        check(lib.eml_gemm_f64_dev(vp(ta), vp(tb), vp(tc), m, n, k, k, n, n, None))
        torch.cuda.synchronize()
        assert np.array_equal(tc.cpu().numpy(), c[:, :n])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,n,k", [(8192, 10, 512), (300, 1, 64), (1000, 16, 777), (257, 7, 1500)])
def test_skinny_n_streaming_kernel(orc, dtype, m, n, k):
    """N <= 16 with A contiguous along k (the NN's H * W2): the HBM-bound streaming kernel."""
    r = rng(700 + m + n + k)
    a, b = uniform(r, (m, k), dtype), uniform(r, (k, n), dtype)
    got = gemm(a, b)
    assert eml.last_kernel() == "skinny_n"
    assert_within(got, a, b, orc.matrix_mul(a, b, threads=4), dtype, f"skinny {m}x{n}x{k}")
    assert np.array_equal(got, gemm(a, b))
    ai, bi = ints(r, (m, k), dtype), ints(r, (k, n), dtype)
    assert np.array_equal(gemm(ai, bi), orc.matrix_mul(ai, bi, threads=4))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m,n,k", [(512, 10, 8192), (40, 1, 300), (700, 16, 1000), (33, 7, 4097)])
def test_skinny_tn_chunked_kernel(orc, dtype, m, n, k):
    """N <= 16 with A stored transposed (contiguous along m) and a long K -- the NN's dW2 = H^T dY."""
    r = rng(2300 + m + n + k)
    a, b = uniform(r, (m, k), dtype), uniform(r, (k, n), dtype)
    a_t = np.ascontiguousarray(a.T)  # [k][m]: element (m, k) at a_t[k*m_total + m]
    got = contract_strided(a_t, b, 1, m, n, k, (m * k, 1, m), (k * n, n, 1), dtype)[0]
    assert eml.last_kernel() == "skinny_tn"
    assert_within(got, a, b, orc.matrix_mul(a, b, threads=4), dtype, f"skinny_tn {m}x{n}x{k}")
    again = contract_strided(a_t, b, 1, m, n, k, (m * k, 1, m), (k * n, n, 1), dtype)[0]
    assert np.array_equal(got, again)


@pytest.mark.parametrize("dtype", DTYPES)
def test_gemm_leading_dimensions(orc, dtype):
    r = rng(3)
    big_a, big_b = uniform(r, (150, 200), dtype), uniform(r, (200, 180), dtype)
    a, b = big_a[10:140, 16:176], big_b[16:176, 4:164]  # views with lda/ldb > row length
    sfx = "f32" if dtype == np.float32 else "f64"
    c = np.full((130, 170), 7.0, dtype)
    check(getattr(lib, f"eml_gemm_{sfx}")(ptr(big_a[10:, 16:]), ptr(big_b[16:, 4:]), ptr(c), 130, 160, 160, 200, 180, 170))
    ref = orc.matrix_mul(np.ascontiguousarray(a), np.ascontiguousarray(b))
    assert_within(c[:, :160], a, b, ref, dtype, "ld")
    assert np.all(c[:, 160:] == 7.0)  # elements outside the C view are untouched


@pytest.mark.parametrize("dtype", DTYPES)
def test_run_to_run_determinism(dtype):
    r = rng(4)
    a, b = uniform(r, (640, 512), dtype), uniform(r, (512, 384), dtype)
    digests = {hashlib.sha256(gemm(a, b).tobytes()).hexdigest() for _ in range(4)}
    assert len(digests) == 1


def test_bad_arguments_are_errors_not_crashes():
    a = np.zeros((2, 2), np.float64)
    with pytest.raises(eml.BadArgError):
        check(lib.eml_gemm_f64(ptr(a), ptr(a), ptr(a), 2, 2, 0, 2, 2, 2))
    with pytest.raises(eml.BadArgError):
        check(lib.eml_gemm_f64(ptr(a), ptr(a), ptr(a), 2, 2, 2, 1, 2, 2))
    with pytest.raises(eml.BadArgError):
        check(lib.eml_gemm_f64(None, ptr(a), ptr(a), 2, 2, 2, 2, 2, 2))


# ------------------------------------------------------------------ named contraction
@pytest.mark.parametrize("dtype", DTYPES)
def test_batched_contraction_vs_oracle(orc, dtype):
    r = rng(5)
    bt, m, k, n = 3, 20, 37, 12
    x = eml.Tensor([("batch", bt), ("row", m), ("k", k)], uniform(r, (bt, m, k), dtype))
    y = eml.Tensor([("batch", bt), ("k", k), ("col", n)], uniform(r, (bt, k, n), dtype))
    got = eml.Einsum.with_2(x, y).to(["batch", "row", "col"])
    assert eml.last_kernel() != "einsum_generic"
    ref = orc.einsum2(x.data, x.names(), y.data, y.names(), ["batch", "row", "col"])
    assert got.shape() == [("batch", bt), ("row", m), ("col", n)]
    for i in range(bt):
        assert_within(got.data[i], x.data[i], y.data[i], ref[i], dtype, f"batch {i}")
    # permuted dimension orders need no host transpose: same numbers
    xp = eml.Tensor([("row", m), ("batch", bt), ("k", k)], np.ascontiguousarray(x.data.transpose(1, 0, 2)))
    yp = eml.Tensor([("col", n), ("k", k), ("batch", bt)], np.ascontiguousarray(y.data.transpose(2, 1, 0)))
    for out in (["batch", "row", "col"], ["batch", "col", "row"], ["row", "batch", "col"], ["col", "row", "batch"]):
        g2 = eml.Einsum.with_2(xp, yp).to(out)
        ref2 = orc.einsum2(xp.data, xp.names(), yp.data, yp.names(), out)
        assert g2.names() == out
        back = np.transpose(g2.data, [out.index(d) for d in ("batch", "row", "col")])
        refb = np.transpose(ref2, [out.index(d) for d in ("batch", "row", "col")])
        for i in range(bt):
            assert_within(back[i], x.data[i], y.data[i], refb[i], dtype, f"{out} batch {i}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_generic_einsum_is_bit_exact(orc, dtype):
    r = rng(6)
    a = uniform(r, (5, 7, 3), dtype)
    b = uniform(r, (3, 4, 7), dtype)
    cases = [(["i", "j", "k"], ["k", "l", "j"], ["i", "l"]),          # two summation dims
             (["i", "j", "k"], ["k", "l", "j"], ["i", "j", "l"]),     # batch-like j
             (["i", "j", "k"], ["k", "l", "j"], []),                  # full reduction
             (["i", "j", "k"], ["x", "l", "y"], ["i", "l"]),          # unrelated dims summed on each side
             (["i", "j", "k"], ["k", "l", "j"], ["l", "k", "j", "i"])]  # no summation
    for an, bn, out in cases:
        x, y = eml.Tensor(list(zip(an, a.shape)), a), eml.Tensor(list(zip(bn, b.shape)), b)
        check(0)
        sfx = "f32" if dtype == np.float32 else "f64"
        got = eml.Einsum.with_2(x, y).to(out)
        want = orc.einsum2(a, an, b, bn, out)
        if eml.last_kernel() == "einsum_generic":
            assert np.array_equal(got.data, want), (an, bn, out)
        else:
            assert np.allclose(got.data, want, rtol=1e-4 if sfx == "f32" else 1e-11)


@pytest.mark.parametrize("dtype", DTYPES)
def test_vector_dot(orc, dtype):
    r = rng(7)
    for n in (1, 3, 1000, 100003):
        a, b = uniform(r, n, dtype), uniform(r, n, dtype)
        got = eml.Tensor([("a", n)], a).scalar_product(eml.Tensor([("a", n)], b))
        ref = orc.vector_dot(a, "a", b, "a")
        assert abs(float(got) - float(ref)) <= TOL[np.dtype(dtype)] * float(np.abs(a).astype(np.float64) @ np.abs(b)) + 1e-300


# ------------------------------------------------------------------ autodiff
@pytest.mark.parametrize("dtype", DTYPES)
def test_record_matmul_reference_fixture(orc, golden, dtype):
    """container_record/mod.rs:2718-2903: numbers, every derivative, and tape length."""
    g = golden["record_matmul"][0]
    a, b = np.array(g["a"], dtype), np.array(g["b"], dtype)
    for kind in ("tensor", "matrix"):
        hist = eml.WengertList(dtype)
        if kind == "tensor":
            ra = eml.RecordTensor.variables(hist, eml.Tensor([("r", 4), ("c", 3)], a))
            rb = eml.RecordTensor.variables(hist, eml.Tensor([("r", 3), ("c", 4)], b))
        else:
            ra = eml.RecordMatrix.variables(hist, eml.Matrix(a))
            rb = eml.RecordMatrix.variables(hist, eml.Matrix(b))
        rc = ra * rb
        tape = orc.Tape(dtype)
        oa, ob = tape.variables(a), tape.variables(b)
        oc, oci = tape.matmul(oa, ob)
        assert np.array_equal(rc.numbers().data, np.array(g["expect"], dtype))
        assert len(hist) == len(tape)
        assert np.array_equal(ra.indexes(), oa[1]) and np.array_equal(rb.indexes(), ob[1])
        assert np.array_equal(rc.indexes(), oci)
        for i in range(4):
            for j in range(4):
                d = rc.derivatives_for([i, j]) if kind == "tensor" else rc.derivatives_for(i, j)
                od = tape.derivatives(oci[i, j])
                da = d.at_tensor(ra).data if kind == "tensor" else d.at_matrix(ra).data
                db = d.at_tensor(rb).data if kind == "tensor" else d.at_matrix(rb).data
                assert np.array_equal(da, od[oa[1]]) and np.array_equal(db, od[ob[1]])
                assert len(d) == len(tape)
                if i == 1 and j == 2:
                    assert np.array_equal(d.to_vec(), od)  # Vec::from(Derivatives), interior entries included


@pytest.mark.parametrize("dtype", DTYPES)
def test_record_chain_vs_scalar_tape(orc, dtype):
    """Two-layer network with container ops only: tape length, indexes, loss and gradients vs the oracle's
    literal scalar tape (including the reference's constants-carry-Index-0 behaviour)."""
    r = rng(8)
    n, din, dh, dout = 6, 5, 4, 3
    x = r.uniform(0, 1, (n, din)).astype(dtype)
    w1 = r.uniform(-0.5, 0.5, (din, dh)).astype(dtype)
    w2 = r.uniform(-0.5, 0.5, (dh, dout)).astype(dtype)
    tgt = r.uniform(0, 1, (n, dout)).astype(dtype)
    ones_l, ones_r = np.ones((1, n), dtype), np.ones((dout, 1), dtype)

    hist = eml.WengertList(dtype)
    T = lambda names, a: eml.Tensor(list(zip(names, a.shape)), a)
    W1 = eml.RecordTensor.variables(hist, T(["i", "h"], w1))
    W2 = eml.RecordTensor.variables(hist, T(["h", "o"], w2))
    X = eml.RecordTensor.constants(T(["n", "i"], x), hist)
    Y = eml.RecordTensor.constants(T(["n", "o"], tgt), hist)
    L = eml.RecordTensor.constants(T(["u", "n"], ones_l), hist)
    R = eml.RecordTensor.constants(T(["o", "v"], ones_r), hist)

    def sigmoid(z):
        return ((-z).exp() + 1.0).div_swapped(1.0)

    h = sigmoid(X * W1)
    out = sigmoid(h * W2)
    diff = out - Y
    sq = diff.elementwise_multiply(diff)
    loss = ((L * sq) * R) / float(n)

    tape = orc.Tape(dtype)
    ow1, ow2 = tape.variables(w1), tape.variables(w2)
    ox, oy = orc.Tape.constants(x, dtype), orc.Tape.constants(tgt, dtype)
    ol, orr = orc.Tape.constants(ones_l, dtype), orc.Tape.constants(ones_r, dtype)

    def osig(z):
        return tape.unary(12, tape.unary(6, tape.unary(3, tape.unary(0, z)), 1.0), 1.0)

    oh = osig(tape.matmul(ox, ow1))
    oout = osig(tape.matmul(oh, ow2))
    odiff = tape.binary(33, oout, oy, mode=1)
    osq = tape.binary(34, odiff, odiff, mode=3)
    oloss = tape.unary(9, tape.matmul(tape.matmul(ol, osq), orr), float(n))

    assert len(hist) == len(tape)
    assert np.array_equal(loss.indexes(), oloss[1]) and np.array_equal(out.indexes(), oout[1])
    rtol = 2e-5 if dtype == np.float32 else 1e-12
    assert np.allclose(loss.numbers().data, oloss[0], rtol=rtol)
    d = loss.derivatives_for([0, 0])
    od = tape.derivatives(oloss[1][0, 0])
    atol = (2e-6 if dtype == np.float32 else 1e-13) * max(1.0, float(np.abs(od).max()))
    for rec, (_, oidx) in ((W1, ow1), (W2, ow2)):
        assert np.allclose(d.at_tensor(rec).data, od[oidx], rtol=50 * rtol, atol=atol)
    # the constant operand X carries Index 0 == W1[0,0]: its derivative sum lands there, as in the reference
    assert abs(float(d.at_index(0)) - float(od[0])) <= 50 * rtol * abs(float(od[0])) + atol
    plain = np.zeros_like(w1)  # what W1[0,0]'s gradient would be WITHOUT the quirk differs from od[0]
    assert od[0] != plain[0, 0]

    # SGD update + clear + reset renumbers the variables exactly like the reference
    before = W1.numbers().data.copy()
    eml.sgd_update(W1, d, 0.1)
    assert len(hist) == len(tape) + w1.size
    assert np.allclose(W1.numbers().data, before - d.at_index(0) * 0 - 0.1 * od[ow1[1]], rtol=50 * rtol, atol=atol)
    hist.clear()
    W1.reset()
    W2.reset()
    assert len(hist) == w1.size + w2.size
    assert np.array_equal(W1.indexes().ravel(), np.arange(w1.size))
    assert np.array_equal(W2.indexes().ravel(), w1.size + np.arange(w2.size))


def test_stale_and_untracked_containers():
    hist = eml.WengertList(np.float64)
    w = eml.RecordMatrix.variables(hist, eml.Matrix([[1.0, 2.0], [3.0, 4.0]]))
    c = eml.RecordMatrix.constants(eml.Matrix([[1.0, 0.0], [0.0, 1.0]]), hist)
    cc = c * c
    assert len(hist) == 4 and cc.history() is None and cc.derivatives_for(0, 0) is None
    assert np.array_equal(cc.indexes(), np.zeros((2, 2), np.int64))
    hist.clear()
    with pytest.raises(eml.StateError):
        w * c  # not reset after clear: the reference would index out of bounds
    w.reset()
    y = w * c
    assert len(hist) == 4 + 4 * 3
    with pytest.raises(eml.Panic):
        w * eml.RecordMatrix.constants(eml.Matrix([[1.0, 2.0, 3.0]]), hist)


@pytest.mark.parametrize("dtype", DTYPES)
def test_container_elementwise_rules(orc, dtype):
    r = rng(9)
    x = r.uniform(0.5, 2.0, (7, 9)).astype(dtype)
    y = r.uniform(0.5, 2.0, (7, 9)).astype(dtype)
    hist = eml.WengertList(dtype)
    X = eml.RecordMatrix.variables(hist, eml.Matrix(x))
    Y = eml.RecordMatrix.variables(hist, eml.Matrix(y))
    tape = orc.Tape(dtype)
    ox, oy = tape.variables(x), tape.variables(y)
    exact_ops = [("NEG", 0.0), ("ADD_C", 1.5), ("SUB_C", 0.25), ("MUL_C", 3.0), ("DIV_C", 3.0), ("C_SUB", 2.0),
                 ("C_DIV", 2.0), ("SQRT", 0.0)]
    approx_ops = [("EXP", 0.0), ("LN", 0.0), ("SIN", 0.0), ("COS", 0.0), ("POW_C", 2.5)]
    for name, c in exact_ops + approx_ops:
        z = X._unary(name, c)
        oz, ozi = tape.unary(eml.OPS[name], ox, c)
        assert np.array_equal(z.indexes(), ozi)
        d = z.derivatives_for(3, 4).at_matrix(X).data
        od = tape.derivatives(ozi[3, 4])[ox[1]]
        if (name, c) in exact_ops:
            assert np.array_equal(z.numbers().data, oz), name
            assert np.array_equal(d, od), name
        else:
            assert np.allclose(z.numbers().data, oz, rtol=1e-6 if dtype == np.float32 else 1e-14), name
            assert np.allclose(d, od, rtol=1e-5 if dtype == np.float32 else 1e-13), name
    for name in ("ADD", "SUB", "MUL", "DIV"):
        z = X._binary(name, Y)
        oz, ozi = tape.binary(eml.OPS[name], ox, oy)
        assert np.array_equal(z.numbers().data, oz) and np.array_equal(z.indexes(), ozi)
        dd = z.derivatives_for(6, 8)
        od = tape.derivatives(ozi[6, 8])
        assert np.array_equal(dd.at_matrix(X).data, od[ox[1]]) and np.array_equal(dd.at_matrix(Y).data, od[oy[1]])
    assert len(hist) == len(tape)


# ------------------------------------------------------------------ next row: covariance
@pytest.mark.parametrize("dtype", DTYPES)
def test_covariance_reference_goldens_and_oracle(orc, golden, dtype):
    for g in golden["covariance"]:
        x = np.array(g["x"], dtype)
        got = eml.Matrix(x).covariance_column_features().data if g["feature_axis"] == 1 else \
            eml.Matrix(x).covariance_row_features().data
        if g["exact"]:
            assert np.array_equal(got, np.array(g["expect"], dtype)), g["cite"]
        else:
            assert np.array_equal(np.round(got.astype(np.float64), g["decimals"]), np.array(g["expect"])), g["cite"]
        if "names" in g:
            t = eml.Tensor(list(zip(g["names"], x.shape)), x)
            c = t.covariance(g["feature_dimension"])
            assert c.names() == ["i", "j"] and np.array_equal(c.data, got)
            with pytest.raises(eml.Panic, match="is not present in the input tensor's shape"):
                t.covariance("nope")
    r = rng(900)
    tol = TOL[np.dtype(dtype)]
    for rows, cols in ((300, 40), (2048, 256), (70, 513)):
        x = uniform(r, (rows, cols), dtype)
        for axis in (1, 0):
            got = (eml.Matrix(x).covariance_column_features() if axis == 1 else eml.Matrix(x).covariance_row_features()).data
            ref = orc.covariance(x, axis).astype(np.float64)
            x64 = x.astype(np.float64)
            xc = x64 - x64.mean(axis=0 if axis == 1 else 1, keepdims=True)
            s = rows if axis == 1 else cols
            scale = (np.abs(xc).T @ np.abs(xc) if axis == 1 else np.abs(xc) @ np.abs(xc).T) / s
            # 4x: both sides round the means and the centred values independently
            assert np.all(np.abs(got.astype(np.float64) - ref) <= 4 * tol * scale + 1e-30), (rows, cols, axis)
            assert np.array_equal(got, got.T) or np.allclose(got, got.T, rtol=0, atol=float(4 * tol * scale.max()))


# ------------------------------------------------------------------ re-entrancy (Matrix / Tensor are Send + Sync)
def test_concurrent_host_calls_from_threads(orc):
    """The reference's Matrix<T> / Tensor<T, D> are Send + Sync (src/matrices/mod.rs:2030-2040), so the library can be
    entered from several host threads at once: results must be the single-threaded ones, bit for bit."""
    import threading

    r = rng(1500)
    jobs = []
    for i, dtype in enumerate([np.float32, np.float64, np.float32, np.float64]):
        a, b = uniform(r, (300 + 64 * i, 200), dtype), uniform(r, (200, 260), dtype)
        jobs.append((a, b, gemm(a, b)))
    errors = []

    def worker(a, b, want):
        try:
            for _ in range(8):
                if not np.array_equal(gemm(a, b), want):
                    errors.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ------------------------------------------------------------------ next row: 1 and 3..6 operand einsum
@pytest.mark.parametrize("dtype", DTYPES)
def test_einsum_n_reference_goldens_and_oracle(orc, golden, dtype):
    from test_oracle_golden import _einsum_n_expected

    withs = {1: eml.Einsum.with_1, 3: eml.Einsum.with_3, 4: eml.Einsum.with_4, 5: eml.Einsum.with_5, 6: eml.Einsum.with_6}
    for g in golden["einsum_n"]:
        ops, want = _einsum_n_expected(g, dtype)
        tensors = [eml.Tensor(list(zip(nm, o.shape)), o) for nm, o in zip(g["names"], ops)]
        got = withs[len(ops)](*tensors).to(g["out_names"])
        assert got.names() == g["out_names"] and np.array_equal(got.data, want), g["cite"]
        assert eml.last_kernel() == "einsum_generic"
    # random data, every arity: bit-exact with the oracle's restatement of the reference loop nest
    r = rng(1700)
    a, b, c = uniform(r, (5, 7), dtype), uniform(r, (7, 4, 3), dtype), uniform(r, (3, 5), dtype)
    d, e, f = uniform(r, (4, 6), dtype), uniform(r, (6,), dtype), uniform(r, (5, 6), dtype)
    cases = [([a], [["i", "j"]], ["j", "i"]), ([b], [["j", "k", "l"]], ["l"]),
             ([a, b, c], [["i", "j"], ["j", "k", "l"], ["l", "i"]], ["i", "k"]),
             ([a, b, c, d], [["i", "j"], ["j", "k", "l"], ["l", "i"], ["k", "m"]], ["m"]),
             ([a, b, c, d, e], [["i", "j"], ["j", "k", "l"], ["l", "i"], ["k", "m"], ["m"]], []),
             ([a, b, c, d, e, f], [["i", "j"], ["j", "k", "l"], ["l", "i"], ["k", "m"], ["m"], ["i", "m"]], ["k", "i"])]
    for ops, names, out in cases:
        tensors = [eml.Tensor(list(zip(nm, o.shape)), o) for nm, o in zip(names, ops)]
        got = withs[len(ops)](*tensors).to(out)
        assert np.array_equal(got.data, orc.einsum_n(ops, names, out)), (names, out)
    with pytest.raises(eml.InconsistentDimensionLengthError):
        eml.Einsum.with_3(eml.Tensor([("a", 2), ("b", 3)], a[:2, :3]), eml.Tensor([("b", 4), ("c", 2)], a[:4, :2]),
                          eml.Tensor([("c", 2), ("d", 2)], a[:2, :2])).to(["a", "d"])


def test_pack_cache_of_constants_follows_identity_not_address():
    """Constant containers keep their split/packed TF32 planes across matmuls (the NN packs its inputs once, not
    twice per step).  The cache is keyed by a buffer identity: a new constant that lands on a recycled address must
    not see the previous one's planes, and repeated products must give the same bits."""
    r = rng(2100)
    hist = eml.WengertList(np.float32)
    w = eml.RecordTensor.variables(hist, eml.Tensor([("k", 128), ("n", 160)], uniform(r, (128, 160), np.float32)))
    results = []
    for trial in range(3):
        x = uniform(r, (1024, 128), np.float32)
        X = eml.RecordTensor.constants(eml.Tensor([("m", 1024), ("k", 128)], x), hist)
        first = (X * w).numbers().data
        assert eml.last_kernel() == "tf32x3_tcgen05"
        again = (X * w).numbers().data  # second product: planes of X come from the cache
        assert np.array_equal(first, again)
        want = x.astype(np.float64) @ w.numbers().data.astype(np.float64)
        bound = 1e-5 * (np.abs(x).astype(np.float64) @ np.abs(w.numbers().data).astype(np.float64))
        assert np.all(np.abs(first.astype(np.float64) - want) <= bound), trial
        # backward through the cached operand as well: dW = X^T G
        d = (X * w).derivatives_for([3, 5]).at_tensor(w).data
        # (w[0, 0] is tape index 0, where the reference's constant-operand quirk parks the constant's would-be gradient)
        assert np.allclose(d[:, 5], x[3], rtol=0, atol=1e-6) and np.count_nonzero(d[1:, :5]) == 0
        results.append(first)
        del X  # its buffer returns to the stream cache; the next constant reuses the address
    assert not np.array_equal(results[0], results[1])


# ------------------------------------------------------------------ recorded steps (CUDA graphs)
def _small_net(dtype, seed=2500):
    r = rng(seed)
    batch = 256
    x = uniform(r, (batch, 48), dtype)
    y = uniform(r, (batch, 6), dtype)
    w1, w2 = 0.1 * uniform(r, (48, 32), dtype), 0.1 * uniform(r, (32, 6), dtype)
    T = lambda names, a: eml.Tensor(list(zip(names, a.shape)), a)
    hist = eml.WengertList(dtype)
    net = dict(hist=hist, W1=eml.RecordTensor.variables(hist, T(["i", "h"], w1)),
               W2=eml.RecordTensor.variables(hist, T(["h", "o"], w2)), X=eml.RecordTensor.constants(T(["n", "i"], x), hist),
               Y=eml.RecordTensor.constants(T(["n", "o"], y), hist),
               L=eml.RecordTensor.constants(T(["u", "n"], np.ones((1, batch), dtype)), hist),
               R=eml.RecordTensor.constants(T(["o", "v"], np.ones((6, 1), dtype)), hist), batch=batch,
               slot=eml.PinnedBuffer((1,), dtype))

    def step():
        sig = lambda z: ((-z).exp() + 1.0).div_swapped(1.0)
        out = sig(sig(net["X"] * net["W1"]) * net["W2"])
        diff = out - net["Y"]
        loss = ((net["L"] * diff.elementwise_multiply(diff)) * net["R"]) / float(batch)
        d = loss.derivatives_for([0, 0])
        eml.sgd_update(net["W1"], d, 0.1)
        eml.sgd_update(net["W2"], d, 0.1)
        arrival = loss.numbers_async(net["slot"])
        hist.clear()
        net["W1"].reset()
        net["W2"].reset()
        return arrival

    net["step"] = step
    return net


@pytest.mark.parametrize("dtype", DTYPES)
def test_recorded_step_replays_the_eager_step_bit_for_bit(dtype):
    """WengertList.record: the step captured as a CUDA graph and replayed N times must leave the same weights and
    losses as N eager steps (same kernels, same order; the SGD update is in place while recording)."""
    steps = 6
    eager = _small_net(dtype)
    losses_e = []
    for _ in range(1 + steps):
        eml.wait_arrival(eager["step"]())
        losses_e.append(float(eager["slot"].array[0]))
    rec = _small_net(dtype)
    eml.wait_arrival(rec["step"]())  # one eager step first: fills the buffer cache with every size the step needs
    losses_g = [float(rec["slot"].array[0])]
    graph = rec["hist"].record(rec["step"])
    assert len(rec["hist"]) == 48 * 32 + 32 * 6  # the tape is back where the step started
    for _ in range(steps):
        eml.wait_arrival(graph.launch())
        losses_g.append(float(rec["slot"].array[0]))
    assert losses_g == losses_e and losses_e[-1] < losses_e[0]
    for name in ("W1", "W2"):
        assert np.array_equal(rec[name].numbers().data, eager[name].numbers().data), name


def test_recorded_step_restrictions_are_errors():
    net = _small_net(np.float32, seed=2600)
    hist = net["hist"]
    # blocking call inside a recording
    eml.wait_arrival(net["step"]())
    with pytest.raises(eml.StateError, match="blocks on the device"):
        hist.record(lambda: net["W1"].numbers())
    # a step that does not return the tape to its starting state
    with pytest.raises(eml.StateError, match="must end in the tape state it started in"):
        hist.record(lambda: (net["X"] * net["W1"]))
    hist.clear()
    net["W1"].reset()
    net["W2"].reset()
    # the library still works eagerly afterwards
    eml.wait_arrival(net["step"]())
    assert np.isfinite(net["slot"].array[0])


# ------------------------------------------------------------------ fused all-gather epilogue on one GPU
@pytest.mark.parametrize("m,n,k,aligned", [(640, 512, 384, True), (300, 260, 132, True), (130, 70, 33, False)])
def test_gemm_scatter_epilogue_writes_every_destination(m, n, k, aligned):
    """eml_gemm_f64_scatter_dev stores every finished tile into the extra destinations too (on a multi-GPU box these
    are the peers' copies of C; here they are buffers of the same GPU).  Covers the TMA kernel's fused epilogue and
    the copy fallback for shapes that dispatch to a kernel without it, with and without `accumulate`."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(31)
    a = torch.rand((m, k), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    b = torch.rand((k, n), generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    vp = lambda t: C.c_void_p(t.data_ptr())
    for accumulate in (0, 1):
        c = torch.rand((m, n), generator=g, device="cuda", dtype=torch.float64)
        c0 = c.clone()
        extra = [torch.full((m, n), -3.0, device="cuda", dtype=torch.float64) for _ in range(3)]
        remote = (C.c_void_p * 3)(*[t.data_ptr() for t in extra])
        check(lib.eml_gemm_f64_scatter_dev(vp(a), vp(b), vp(c), m, n, k, k, n, n, accumulate, remote, 3, None))
        torch.cuda.synchronize()
        kernel = eml.last_kernel()
        assert (kernel == "dmma_f64_tma+peer_allgather") == aligned, kernel
        want = a @ b + (c0 if accumulate else 0)
        bound = 1e-12 * (a.abs() @ b.abs()) + 1e-15
        assert bool(((c - want).abs() <= bound).all())
        for t in extra:
            assert torch.equal(t, c)
    # eml_copy_async: the transport of the B panels (device to device here)
    dst = torch.empty_like(a)
    check(lib.eml_copy_async(vp(dst), vp(a), a.numel() * 8, None))
    torch.cuda.synchronize()
    assert torch.equal(dst, a)


def test_sgemm_non_finite_and_huge_values(orc):
    """inf, nan and near-FLT_MAX entries through the 3xTF32 path behave as in the reference's plain f32 products:
    an infinite entry gives +-inf (not the NaN its meeting the other operand's zero small part would give: the split
    passes flag non-finite inputs and a repair pass recomputes the NaNs as f32 fma chains), NaN stays NaN, and where
    the reference itself produces NaN (inf - inf, 0 * inf) so does the GPU.  Near-overflow values stay finite (the
    split truncates instead of rounding FLT_MAX up to infinity).  Host entry point and device-resident entry point,
    the packed kernels and the in-kernel-split kernel."""
    torch = pytest.importorskip("torch")
    m = n = k = 256
    a = np.zeros((m, k), np.float32)
    b = np.zeros((k, n), np.float32)
    a[np.arange(m), np.arange(m)] = 1.0            # identity: C = B, entry by entry
    b[:] = rng(2700).uniform(-1, 1, (k, n)).astype(np.float32)
    b[0, 0], b[1, 1], b[2, 2], b[3, 3] = np.inf, -np.inf, np.nan, np.finfo(np.float32).max
    b[4, 4] = -np.finfo(np.float32).max
    a[7, 9] = np.inf                               # row 7 of C: b[7, :] + inf * b[9, :]  -> +-inf by the sign of b[9, j]
    a[8, 0], a[8, 1] = 1.0, 1.0                    # C[8, 0] = b[8, 0] + inf + b[1, 0] = inf; C[8, 1] = ... - inf
    with np.errstate(all="ignore"):
        want = orc.matrix_mul(a, b, threads=4)     # the reference's own arithmetic on these operands
        got = gemm(a, b)
    assert eml.last_kernel() == "tf32x3_tcgen05"
    assert got[0, 0] == np.inf and got[1, 1] == -np.inf and np.isnan(got[2, 2])
    assert np.isfinite(got[3, 3]) and got[3, 3] >= np.float32(3.4e38) and got[4, 4] <= np.float32(-3.4e38)
    nonfinite = ~np.isfinite(want)
    assert nonfinite.sum() >= 256 + 4
    assert np.array_equal(np.isnan(got), np.isnan(want))                       # NaN exactly where the reference has NaN
    assert np.array_equal(got[nonfinite & ~np.isnan(want)], want[nonfinite & ~np.isnan(want)])   # +-inf with its sign
    clean = np.isfinite(want)
    clean[3, 3] = clean[4, 4] = False
    assert np.allclose(got[clean], want[clean], rtol=1e-5, atol=1e-6)
    # device-resident call, a larger product that takes the CTA-pair kernel, and a clean product afterwards (the
    # stream's flag has been cleared: nothing is recomputed, same bits as before any non-finite input was seen)
    r = rng(2701)
    a2, b2 = uniform(r, (512, 384), np.float32), uniform(r, (384, 640), np.float32)
    before = gemm(a2, b2)
    a3 = a2.copy()
    a3[5, 17], a3[300, 0] = -np.inf, np.inf
    with np.errstate(all="ignore"):
        want3 = orc.matrix_mul(a3, b2, threads=4)
        got3 = gemm(a3, b2)
    assert np.array_equal(np.isnan(got3), np.isnan(want3))
    inf3 = np.isinf(want3)
    assert inf3.sum() == 2 * 640 and np.array_equal(got3[inf3], want3[inf3])
    assert np.array_equal(gemm(a2, b2), before)
    # the batched contraction whose kernel splits in shared memory (config 3's path)
    x = torch.rand((4, 1024, 1024), device="cuda") - 0.5
    y = torch.rand((4, 1024, 1024), device="cuda") - 0.5
    x[2, 100, 7] = float("inf")
    out = torch.empty((4, 1024, 1024), device="cuda")
    s3 = lambda t: (C.c_int64 * 3)(*t.stride())
    vp = lambda t: C.c_void_p(t.data_ptr())
    check(lib.eml_contract_f32_dev(vp(x), vp(y), vp(out), 4, 1024, 1024, 1024, s3(x), s3(y), s3(out), 0, None))
    torch.cuda.synchronize()
    row = out[2, 100]
    assert bool(torch.isinf(row).all()) and bool((torch.sign(row) == torch.sign(y[2, 7])).all())
    assert bool(torch.isfinite(out[2, 99]).all()) and bool(torch.isfinite(out[3]).all())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("axis", [1, 0])
def test_gram_products_compute_the_upper_triangle_and_mirror_it(orc, dtype, axis):
    """C = G * G^T (A and B are the same buffer read through mirrored strides): the tensor-core kernels compute
    only the tiles on and above the diagonal and the upper triangle is mirrored.  On and above the diagonal the
    numbers are those of the ordinary product of two separate buffers, the result is exactly symmetric (as the
    reference's covariance is), and the covariance entry points (which go through it) still meet the bound."""
    torch = pytest.importorskip("torch")
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    f, s = 2304, 700  # 18 tiles of 128 / 9 pair tiles of 256: enough tiles that no split-K reformulation kicks in
    g = torch.Generator(device="cuda").manual_seed(77 + axis)
    G = torch.rand((s, f) if axis == 1 else (f, s), generator=g, device="cuda", dtype=tdt) - 0.3
    fn = lib.eml_contract_f64_dev if dtype == np.float64 else lib.eml_contract_f32_dev
    s3 = lambda *v: (C.c_int64 * 3)(*v)
    vp = lambda t: C.c_void_p(t.data_ptr())
    # A(i, k), B(k, j): feature-major reading of G on both sides
    sa, sb = ((0, 1, f), (0, f, 1)) if axis == 1 else ((0, s, 1), (0, 1, s))
    out = torch.full((f, f), float("nan"), device="cuda", dtype=tdt)
    before = eml.kernel_launches()
    check(fn(vp(G), vp(G), vp(out), 1, f, f, s, s3(*sa), s3(*sb), s3(0, f, 1), 0, None))
    launches = eml.kernel_launches() - before
    G2 = G.clone()
    full = torch.empty((f, f), device="cuda", dtype=tdt)
    check(fn(vp(G), vp(G2), vp(full), 1, f, f, s, s3(*sa), s3(*sb), s3(0, f, 1), 0, None))
    torch.cuda.synchronize()
    assert torch.equal(out, out.T)                                   # exactly symmetric
    iu = torch.triu_indices(f, f, device="cuda")
    assert torch.equal(out[iu[0], iu[1]], full[iu[0], iu[1]])        # same numbers on and above the diagonal
    assert launches == (2 if dtype == np.float64 else 3), launches   # (one pack +) GEMM + mirror
    # through the covariance entry point, against the oracle
    x = G.cpu().numpy()
    got = (eml.Matrix(x).covariance_column_features() if axis == 1 else eml.Matrix(x).covariance_row_features()).data
    assert np.array_equal(got, got.T)
    ref = orc.covariance(x[:, :48] if axis == 1 else x[:48], axis).astype(np.float64)
    x64 = x.astype(np.float64)
    xc = x64 - x64.mean(axis=0 if axis == 1 else 1, keepdims=True)
    scale = (np.abs(xc).T @ np.abs(xc) if axis == 1 else np.abs(xc) @ np.abs(xc).T) / s
    assert np.all(np.abs(got[:48, :48].astype(np.float64) - ref) <= TOL[np.dtype(dtype)] * scale[:48, :48] + 1e-30)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sweep_first_contribution_overwrites_like_zero_plus_x(orc, dtype):
    """The sweep does not clear derivative buffers whose first contribution comes from an elementwise rule: it
    writes 0 + dout * deriv instead.  Every observable must stay the reference's: the whole derivative vector
    equals the oracle's scalar tape bit for bit on exactly representable data, including +0 (not -0) where a
    zero derivative meets a negative rule, derivatives taken of an INTERMEDIATE (later nodes contribute zeros to
    it and must not wipe the seed), and a parent that is used twice by one rule and again by a later one."""
    x0 = np.array([[1.0, -2.0, 0.5], [4.0, 0.25, -8.0]], dtype)
    hist = eml.WengertList(dtype)
    X = eml.RecordTensor.variables(hist, eml.Tensor([("r", 2), ("c", 3)], x0))
    neg = -X                                   # constant derivative -1: no derivative array
    sq = neg.elementwise_multiply(neg)         # one rule, the same parent twice
    mix = (sq - neg) * 0.5 + 2.0               # x - y (1, -1), * c, + c
    out = mix / 4.0 - X                        # / c, and X used again by a later rule

    tape = orc.Tape(dtype)
    ox = tape.variables(x0)
    oneg = tape.unary(0, ox)
    osq = tape.binary(34, oneg, oneg, mode=3)
    omix = tape.unary(6, tape.unary(8, tape.binary(33, osq, oneg, mode=3), 0.5), 2.0)
    oout = tape.binary(33, tape.unary(9, omix, 4.0), ox, mode=3)
    assert len(hist) == len(tape) and np.array_equal(out.indexes(), oout[1])
    assert np.array_equal(out.numbers().data, oout[0])

    for target, otarget in ((out, oout), (mix, omix), (neg, oneg)):   # the output and two intermediates
        for (i, j) in ((0, 0), (1, 2)):
            got = np.asarray(target.derivatives_for([i, j]).to_vec(), dtype)
            want = np.asarray(tape.derivatives(otarget[1][i, j]), dtype)
            assert np.array_equal(got, want), (i, j)
            assert np.array_equal(np.signbit(got), np.signbit(want)), "a -0 where the reference holds +0"


@pytest.mark.parametrize("dtype", DTYPES)
def test_batched_host_contraction_pipelines_over_batch_chunks(dtype):
    """eml_contract_* from HOST buffers with many independent batches cuts the batch range into chunks whose upload,
    product and download overlap.  f64: same bits as the device-resident call on the same numbers (f32: the
    tensor-core kernel's deterministic split-K choice depends on how many tiles a launch holds, so a chunked and an
    unchunked call may round differently -- both within the bound); ragged last chunk (21 batches -> chunks of 3),
    padded batch / row strides, pageable (numpy) memory; elements of C's buffer that belong to no batch (the
    padding between batches) stay untouched."""
    torch = pytest.importorskip("torch")
    bt, m, k, n = 21, 384, 520, 640
    r = rng(61)
    lda, ldb = k + 8, n + 4                       # padded rows
    sa_b, sb_b, sc_b = m * lda + 64, k * ldb + 32, m * n + 128   # padded batches (C rows dense: m * n per batch)
    a = uniform(r, (bt * sa_b,), dtype)
    b = uniform(r, (bt * sb_b,), dtype)
    c = np.full(bt * sc_b, -7.0, dtype)
    sfx = "f32" if dtype == np.float32 else "f64"
    s3 = lambda *v: (C.c_int64 * 3)(*v)
    check(getattr(lib, f"eml_contract_{sfx}")(ptr(a), ptr(b), ptr(c), bt, m, n, k, s3(sa_b, lda, 1), s3(sb_b, ldb, 1),
                                             s3(sc_b, n, 1)))
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    tc = torch.full((bt * sc_b,), -7.0, device="cuda", dtype=ta.dtype)
    vp = lambda t: C.c_void_p(t.data_ptr())
    check(getattr(lib, f"eml_contract_{sfx}_dev")(vp(ta), vp(tb), vp(tc), bt, m, n, k, s3(sa_b, lda, 1), s3(sb_b, ldb, 1),
                                                 s3(sc_b, n, 1), 0, None))
    torch.cuda.synchronize()
    if dtype == np.float64:
        assert np.array_equal(c, tc.cpu().numpy())
    else:
        assert np.allclose(c, tc.cpu().numpy(), rtol=0, atol=1e-4)
    cv = c.reshape(bt, sc_b)
    assert np.all(cv[:, m * n:] == -7.0)          # the gaps between batches were not written
    for i in (0, 5, 20):
        av = a[i * sa_b:i * sa_b + m * lda].reshape(m, lda)[:, :k].astype(np.float64)
        bv = b[i * sb_b:i * sb_b + k * ldb].reshape(k, ldb)[:, :n].astype(np.float64)
        bound = TOL[np.dtype(dtype)] * (np.abs(av) @ np.abs(bv))
        assert np.all(np.abs(cv[i, :m * n].reshape(m, n).astype(np.float64) - av @ bv) <= bound)


@pytest.mark.parametrize("dtype", DTYPES)
def test_batched_host_contraction_with_batches_wider_than_a_staging_slot(dtype):
    """Pageable (numpy) operands whose single batch element exceeds the 16 MB page-locked staging slot
    (4 x 2100 x 2100 f32 = 17.6 MB per batch, f64 twice that): the staged copier cuts such a 'row' into slot-sized
    pieces.  The reference computes these shapes (Einsum2::to, src/tensors/einsum.rs:908-980), so must the drop-in."""
    bt, m = 4, 2100
    r = rng(62)
    a = uniform(r, (bt, m, m), dtype)
    b = uniform(r, (bt, m, m), dtype)
    c = np.empty((bt, m, m), dtype)
    sfx = "f32" if dtype == np.float32 else "f64"
    s3 = lambda *v: (C.c_int64 * 3)(*v)
    check(getattr(lib, f"eml_contract_{sfx}")(ptr(a), ptr(b), ptr(c), bt, m, m, m, s3(m * m, m, 1), s3(m * m, m, 1),
                                             s3(m * m, m, 1)))
    for i in (0, 3):
        rows = np.array([0, 1, 777, m - 1])
        av, bv = a[i][rows].astype(np.float64), b[i].astype(np.float64)
        bound = TOL[np.dtype(dtype)] * (np.abs(av) @ np.abs(bv))
        assert np.all(np.abs(c[i][rows].astype(np.float64) - av @ bv) <= bound)


def test_concurrent_pageable_host_calls_on_two_devices():
    """Two host threads, each driving its own GPU with pageable operands large enough for the pipelined path: the copy
    thread pool is shared by the whole process and must serve both without mixing their jobs."""
    import threading

    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = rng(63)
    a, b = uniform(r, (2304, 1024), np.float64), uniform(r, (1024, 1280), np.float64)
    want = gemm(a, b)
    errors = []

    def worker(dev):
        try:
            torch.cuda.set_device(dev)
            for _ in range(4):
                if not np.array_equal(gemm(a, b), want):
                    errors.append(f"device {dev}: mismatch")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(d,)) for d in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.set_device(0)
    assert not errors, errors
