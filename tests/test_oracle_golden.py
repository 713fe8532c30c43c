"""Pins the CPU oracle against every golden vector the reference's own tests hold for the hot path
(SURVEY.md section 8c).  CPU only."""
import numpy as np
import pytest

from scalar_record import Rec, matmul as rec_matmul

DTYPES = [np.float32, np.float64]


@pytest.mark.parametrize("dtype", DTYPES)
def test_matrix_mul_goldens(orc, golden, dtype):
    for g in golden["matmul"]:
        a, b = np.array(g["a"], dtype=dtype), np.array(g["b"], dtype=dtype)
        c = orc.matrix_mul(a, b)
        assert np.array_equal(c, np.array(g["expect"], dtype=dtype)), g["cite"]
        assert np.array_equal(orc.matrix_mul(a, b, threads=2), c)


def test_matrix_mul_panic(orc, golden):
    for g in golden["matmul_panic"]:
        with pytest.raises(orc.ReferencePanic) as e:
            orc.matrix_mul(np.zeros(g["a_shape"]), np.zeros(g["b_shape"]))
        assert str(e.value) == g["message"]


@pytest.mark.parametrize("dtype", DTYPES)
def test_tensor_mul_goldens(orc, golden, dtype):
    for g in golden["tensor_matmul"]:
        c, names = orc.tensor_mul(np.array(g["a"], dtype=dtype), g["a_names"], np.array(g["b"], dtype=dtype),
                                  g["b_names"])
        assert np.array_equal(c, np.array(g["expect"], dtype=dtype)), g["cite"]
        assert names == g["expect_names"]


def test_tensor_mul_panics(orc, golden):
    for g in golden["tensor_matmul_panic"]:
        with pytest.raises(orc.ReferencePanic) as e:
            orc.tensor_mul(np.zeros(g["a_shape"]), g["a_names"], np.zeros(g["b_shape"]), g["b_names"])
        assert str(e.value).startswith(g["message_prefix"]), str(e.value)


@pytest.mark.parametrize("dtype", DTYPES)
def test_einsum2_goldens(orc, golden, dtype):
    for g in golden["einsum2"]:
        out = orc.einsum2(np.array(g["a"], dtype=dtype), g["a_names"], np.array(g["b"], dtype=dtype),
                          g["b_names"], g["out_names"])
        assert np.array_equal(out, np.array(g["expect"], dtype=dtype)), g["cite"]


def _equiv(g, a, b, orc):
    e = g["equiv"]
    if e == "a @ b.T":
        return orc.matrix_mul(a, np.ascontiguousarray(b.T))
    if e == "a @ b":
        return orc.matrix_mul(a, b)
    if e == "a.T @ b":
        return orc.matrix_mul(np.ascontiguousarray(a.T), b)
    if e == "batched a @ b":
        return np.stack([orc.matrix_mul(a[i], b[i]) for i in range(a.shape[0])])
    if e == "dot":
        return np.array(orc.vector_dot(a, "a", b, "a"), dtype=a.dtype)
    if e == "a * b":
        return a * b
    if e == "outer":
        return np.outer(a, b)
    raise AssertionError(e)


@pytest.mark.parametrize("dtype", DTYPES)
def test_einsum2_equivalences(orc, golden, dtype):
    for g in golden["einsum2_equiv"]:
        a, b = np.array(g["a"], dtype=dtype), np.array(g["b"], dtype=dtype)
        out = orc.einsum2(a, g["a_names"], b, g["b_names"], g["out_names"])
        assert np.array_equal(out, _equiv(g, a, b, orc)), g["cite"]
    # ab,ab-> (tests/einsum.rs:100-109): the sum of the elementwise products
    a = np.array(golden["einsum2_equiv"][6]["a"], dtype=dtype)
    assert orc.einsum2(a, ["a", "b"], a, ["a", "b"], []) == (a * a).sum(dtype=dtype)


def test_einsum2_errors(orc):
    # src/tensors/einsum.rs:508-538: output name absent from every input / inconsistent lengths
    a, b = np.zeros((2, 3), np.float32), np.zeros((4, 5), np.float32)
    with pytest.raises(orc.InconsistentDimensionLengthError) as e:
        orc.einsum2(a, ["a", "b"], b, ["b", "c"], ["a", "c"])
    assert e.value.dimension == "b" and e.value.lengths == [3, 4]
    with pytest.raises(orc.InconsistentDimensionLengthError) as e:
        orc.einsum2(a, ["a", "b"], b, ["c", "d"], ["a", "z"])
    assert e.value.dimension == "z" and e.value.lengths == [None, None]
    with pytest.raises(orc.InconsistentDimensionLengthError) as e:
        orc.einsum2(a, ["a", "b"], b, ["a", "c"], ["a", "c"])
    assert e.value.dimension == "a" and e.value.lengths == [2, 4]


def test_scalar_product_order(orc):
    # reduce without a zero seed, strictly left to right, products rounded separately
    l = np.array([1e16, 1.0, -1e16, 1.0])
    r = np.ones(4)
    assert orc.scalar_product(l, r) == 1.0  # ((1e16 + 1) - 1e16) + 1 = 0 + 1
    assert orc.scalar_product(l[::-1].copy()[::-1], r) == 1.0
    x = np.float32(1.0) + np.float32(2.0) ** -12
    l32 = np.array([x, 1.0], np.float32)
    # product rounded to f32 before the add (no FMA): x*x rounds to 1 + 2^-11
    want = np.float32(np.float32(x * x) + np.float32(1.0))
    assert orc.scalar_product(l32, np.array([x, 1.0], np.float32)) == want
    # strided operands (column walk)
    m = np.arange(12, dtype=np.float64).reshape(3, 4)
    assert orc.scalar_product(m[:, 1], m[:, 2]) == 1 * 2 + 5 * 6 + 9 * 10


def test_record_matmul_matches_scalar_path(orc, golden):
    g = golden["record_matmul"][0]
    a, b = np.array(g["a"]), np.array(g["b"])
    t1, t2 = orc.Tape(np.float64), orc.Tape(np.float64)
    # container path
    ra, rb = t1.variables(a), t1.variables(b)
    c, ci = t1.matmul(ra, rb)
    # scalar Tensor<Record> path
    sa, sb = t2.variables(a), t2.variables(b)
    sc, sci = t2.scalar_record_matmul(sa, sb)
    assert np.array_equal(c, np.array(g["expect"]))
    assert np.array_equal(c, sc)
    assert len(t1) == len(t2) == 12 + 12 + 16 * (2 * 3 - 1)
    assert np.array_equal(ci, sci)
    # closed form of the output index: base + (i*N+j)(2K-1) + 2K-2
    base, (m, k), n = 24, a.shape, b.shape[1]
    want = base + (np.arange(m * n).reshape(m, n)) * (2 * k - 1) + 2 * k - 2
    assert np.array_equal(ci, want)
    for i in range(4):
        for j in range(4):
            d1, d2 = t1.derivatives(ci[i, j]), t2.derivatives(sci[i, j])
            assert np.array_equal(d1[ra[1]], d2[sa[1]]) and np.array_equal(d1[rb[1]], d2[sb[1]])
            da = np.zeros_like(a)
            da[i, :] = b[:, j]
            db = np.zeros_like(b)
            db[:, j] = a[i, :]
            assert np.array_equal(d1[ra[1]], da) and np.array_equal(d1[rb[1]], db)


def test_record_matmul_constant_operand_quirk(orc):
    """constants x variables: the container path records the constant side with Index 0
    (container_operations.rs:1291-1296), so the sweep adds sum(dC . B^T) into derivatives[0]; the scalar
    Record path treats the constant as a plain number (record_operations.rs:362-364)."""
    x = np.array([[1.0, 2.0], [3.0, 4.0]])
    w = np.array([[0.5, -1.0], [2.0, 0.25]])
    t1, t2 = orc.Tape(np.float64), orc.Tape(np.float64)
    rw, sw = t1.variables(w), t2.variables(w)
    c, ci = t1.matmul(orc.Tape.constants(x), rw)
    sc, sci = t2.scalar_record_matmul(orc.Tape.constants(x), sw, a_tracked=False)
    assert np.array_equal(c, sc)
    assert len(t1) == 4 + 4 * 3 and len(t2) == 4 + 4 * 3  # unary vs binary entries: same count
    d1, d2 = t1.derivatives(ci[0, 0]), t2.derivatives(sci[0, 0])
    # scalar path: dC00/dW = [[x00, 0], [x01, 0]]
    assert np.array_equal(d2[sw[1]], np.array([[1.0, 0.0], [2.0, 0.0]]))
    # container path: the same, plus dC00/dX summed into index 0 (= w[0,0]'s slot): w00 + w10
    assert np.array_equal(d1[rw[1]], np.array([[1.0 + (0.5 + 2.0), 0.0], [2.0, 0.0]]))


def test_container_unary_binary_rules(orc):
    t = orc.Tape(np.float64)
    x = t.variables(np.array([0.5, 2.0, 3.0]))
    y = t.variables(np.array([4.0, 0.25, 2.0]))
    z, zi = t.binary(35, x, y)  # Division
    assert np.array_equal(z, x[0] / y[0])
    d = t.derivatives(zi[1])
    assert d[x[1][1]] == 1.0 / 0.25 and d[y[1][1]] == -2.0 / (0.25 * 0.25)
    s, si = t.unary(12, (z, zi), scalar=1.0)  # div_swapped: 1 / z
    d = t.derivatives(si[0])
    assert d[zi[0]] == -1.0 / (z[0] * z[0])
    e, ei = t.unary(3, x)  # exp
    assert np.array_equal(e, np.exp(x[0]))
    assert t.derivatives(ei[2])[x[1][2]] == np.exp(3.0)
    # one tape entry per element
    assert len(t) == 6 + 3 + 3 + 3


def test_xor_network_replay(orc, golden):
    """tests/differentiation_nn.rs:74-120 replayed on the oracle tape with scalar Records."""
    g = golden["xor_nn"]
    f32 = np.float32
    tape = orc.Tape(f32)
    w1 = [[Rec.variable(x, tape) for x in row] for row in g["w1"]]
    w2 = [[Rec.variable(x, tape) for x in row] for row in g["w2"]]
    inputs = [[[Rec.constant(x, f32) for x in row]] for row in g["inputs"]]
    labels = [Rec.constant(x, f32) for x in g["labels"]]
    lr = f32(g["learning_rate"])

    def relu(r):
        return r if r.number > 0 else Rec.constant(0.0, f32)

    losses = []
    for _ in range(g["epochs"]):
        acc = Rec.constant(0.0, f32)
        for inp, correct in zip(inputs, labels):
            h = [[relu(v) for v in row] for row in rec_matmul(inp, w1)]
            out = rec_matmul(h, w2)[0][0]
            acc = acc + ((correct - out) * (correct - out))
        loss = acc / f32(len(inputs))
        d = tape.derivatives(loss.index)
        for w in (w1, w2):
            for row in w:
                for c in range(len(row)):
                    row[c] = row[c] - (d[row[c].index] * lr)
        tape.clear()
        for w in (w1, w2):
            for row in w:
                for r in row:
                    r.reset()
        losses.append(float(loss.number))
    assert losses[-1] < g["final_loss_below"], losses[-1]
    assert losses[0] > losses[-1]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_covariance_goldens(orc, golden, dtype):
    """Next row: covariance_column_features / covariance_row_features / Tensor::covariance."""
    for g in golden["covariance"]:
        x = np.array(g["x"], dtype)
        got = orc.covariance(x, g["feature_axis"])
        if g["exact"]:
            assert np.array_equal(got, np.array(g["expect"], dtype)), g["cite"]
        else:
            assert np.array_equal(np.round(got.astype(np.float64), g["decimals"]), np.array(g["expect"])), g["cite"]
        # row-features form of the same data
        assert np.array_equal(orc.covariance(np.ascontiguousarray(x.T), 1 - g["feature_axis"]), got)


def _einsum_n_expected(g, dtype):
    ops = [np.array(o, dtype) for o in g["operands"]]
    if "expect" in g:
        return ops, np.array(g["expect"], dtype)
    x0 = ops[0]
    if g["equiv"] == "sequential_sum(x0)":
        acc = dtype(0)
        for v in x0.reshape(-1):
            acc = dtype(acc + v)
        return ops, np.array(acc, dtype)
    return ops, np.ascontiguousarray(eval(g["equiv"], {"x0": x0})).astype(dtype)  # small integers: exact


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_einsum_n_goldens(orc, golden, dtype):
    """Next row: Einsum1 and Einsum3..6 against the literal expectations of tests/einsum.rs."""
    for g in golden["einsum_n"]:
        ops, want = _einsum_n_expected(g, dtype)
        got = orc.einsum_n(ops, g["names"], g["out_names"])
        assert got.shape == want.shape and np.array_equal(got, want), g["cite"]
    with pytest.raises(orc.InconsistentDimensionLengthError):
        orc.einsum_n([np.zeros((2, 3), dtype), np.zeros((4, 2), dtype), np.zeros((2, 2), dtype)],
                     [["a", "b"], ["b", "c"], ["c", "d"]], ["a", "d"])


# ---- next row: Cholesky and the multivariate Gaussian draw (callers of `*`) ----------------------
def test_cholesky_goldens(orc, golden):
    for g in golden["cholesky"]:
        dtype = np.float32 if g["exact"] else np.float64
        a = np.array(g["a"], dtype=dtype)
        low = orc.cholesky(a)
        expect = np.array(g["expect"], dtype=dtype)
        if g["exact"]:
            assert np.array_equal(low, expect), g["cite"]
            assert np.array_equal(orc.matrix_mul(low, np.ascontiguousarray(low.T)), a), g["cite"]  # L L^T == A exactly
        else:
            assert np.abs(low - expect).sum() < g["abs_sum_tolerance"], g["cite"]
            assert np.abs(orc.matrix_mul(low, np.ascontiguousarray(low.T)) - a).sum() < g["abs_sum_tolerance"]
    assert orc.cholesky(np.array([[1.0, 2.0], [2.0, 1.0]])) is None      # not positive definite
    assert orc.cholesky(np.zeros((2, 3))) is None                        # not square


def test_cholesky_off_diagonal_is_reciprocal_then_multiply(orc):
    """The reference computes L[i][j] = (a - sum) * (T::one() / L[j][j]) (src/linear_algebra.rs:1093-1096): a reciprocal
    and a multiply, TWO roundings -- not one division.  [[9, 5], [5, 10]]: L00 = 3 and L10 = 5 * fl(1/3), which rounds
    to 0x3FFAAAAAAAAAAAAA, while 5/3 rounds to 0x3FFAAAAAAAAAAAAB.  The expected bits are derived here with exact
    rationals (float(Fraction) is correctly rounded), independently of the oracle's arithmetic."""
    from fractions import Fraction
    import struct

    bits = lambda x: struct.unpack("<Q", struct.pack("<d", float(x)))[0]
    reciprocal = float(Fraction(1, 3))                       # fl(1 / 3)
    two_roundings = float(Fraction(5) * Fraction(reciprocal))  # fl(5 * fl(1/3))
    one_rounding = float(Fraction(5, 3))
    assert bits(two_roundings) == 0x3FFAAAAAAAAAAAAA and bits(one_rounding) == 0x3FFAAAAAAAAAAAAB
    low = orc.cholesky(np.array([[9.0, 5.0], [5.0, 10.0]]))
    assert low[0, 0] == 3.0 and bits(low[1, 0]) == bits(two_roundings)
    # third row: the sum is taken before the reciprocal multiply, L22's radicand uses the two-rounding entries
    a = np.array([[9.0, 5.0, 7.0], [5.0, 10.0, 2.0], [7.0, 2.0, 20.0]])
    low = orc.cholesky(a)
    l10, l20 = two_roundings, float(Fraction(7) * Fraction(reciprocal))
    l11 = float(np.sqrt(np.float64(float(Fraction(10) - Fraction(float(Fraction(l10) * Fraction(l10)))))))
    s = float(Fraction(0) + Fraction(float(Fraction(l20) * Fraction(l10))))
    l21 = float(Fraction(float(Fraction(2) - Fraction(s))) * Fraction(float(Fraction(1) / Fraction(l11))))
    assert (low[1, 0], low[2, 0], low[1, 1], low[2, 1]) == (l10, l20, l11, l21)


def test_multivariate_gaussian_draw_golden(orc, golden):
    """tests/distributions.rs:48-220: the reference's own assertions on its fixed random source."""
    g = golden["multivariate_gaussian"]
    mean = np.array(g["mean"])
    cov = np.array(g["covariance"])
    drawn = orc.multivariate_gaussian_draw(mean, cov, g["source"], g["max_samples"])
    assert drawn.shape == (25, 3)
    for f in range(3):
        feature_mean = drawn[:, f].sum() / g["max_samples"]
        sd = np.sqrt(cov[f, f])
        assert mean[f] - 0.5 * sd < feature_mean < mean[f] + 0.5 * sd, g["cite"]
    # N = 3 is odd: every sample consumes 4 numbers, so one number fewer than 100 is not enough
    assert orc.multivariate_gaussian_draw(mean, cov, g["source"][:99], g["max_samples"]) is None
    # a single Gaussian::draw keeps the surplus sample only when max_samples is even (:230-236)
    assert len(orc.gaussian_draw(1.0, 2.0, iter(g["source"]), 5, np.float64)) == 5
    assert orc.gaussian_draw(1.0, 2.0, iter(g["source"][:5]), 5, np.float64) is None
