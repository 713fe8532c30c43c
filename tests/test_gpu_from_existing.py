"""RecordTensor::from_existing / RecordMatrix::from_existing (reference container_record/mod.rs:219-224, :508): containers
made of arbitrary (number, Index) pairs -- ranges of other containers, rows of a constants matrix re-wrapped with the
weights' history (src/neural_networks.rs:105-120), duplicated elements.  Checked against the oracle's literal scalar tape,
whose operators take any index array: tape length and every Index exactly, derivatives within the path's tolerance
(sums over duplicated indexes run in a different -- fixed -- order than the reference's reverse sweep)."""
import numpy as np
import pytest

import easy_ml_b200 as eml

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]


def close(got, want, dtype):
    rtol = 2e-5 if dtype == np.float32 else 1e-12
    return np.allclose(got, want, rtol=rtol, atol=rtol * max(1.0, float(np.abs(want).max())))


def test_doc_example_range_of_variables():
    """mod.rs:205-217: a 2 x 2 variables tensor, range y 0..1 -> shape [("x", 2), ("y", 1)], same history, indexes 0 and 2."""
    hist = eml.WengertList(np.float64)
    x = eml.RecordTensor.variables(hist, eml.Tensor([("x", 2), ("y", 2)], np.array([[6.0, 9.0], [8.0, 12.0]])))
    fixed = x.range({"y": (0, 1)})
    assert fixed.shape() == [("x", 2), ("y", 1)]
    assert np.array_equal(fixed.indexes().ravel(), [0, 2]) and np.array_equal(fixed.numbers().data.ravel(), [6.0, 8.0])
    assert len(hist) == 4                                   # from_existing appends nothing
    y = fixed * 3.0                                         # tracked work on the view: entries 4, 5
    assert np.array_equal(y.indexes().ravel(), [4, 5]) and len(hist) == 6
    d = y.derivatives_for([1, 0])
    assert np.array_equal(d.at_tensor(x).data, [[0.0, 0.0], [3.0, 0.0]])
    assert np.array_equal(d.at_tensor(fixed).data.ravel(), [0.0, 3.0])   # Derivatives::at_tensor looks every Index up
    assert np.array_equal(d.to_vec(), [0.0, 0.0, 3.0, 0.0, 0.0, 1.0])


@pytest.mark.parametrize("dtype", DTYPES)
def test_rows_of_constants_rewrapped_with_the_weights_history(orc, dtype):
    """src/neural_networks.rs:105-120: every input row is its own RecordMatrix, from_existing(w1.history(), row range of
    the constants matrix): a container WITH a history whose every Index is 0.  Its would-be gradient is summed into
    derivatives[0] (= w1[0, 0]) by the sweep, as the reference's scalar tape does."""
    r = np.random.default_rng(11)
    n, din, dh = 5, 7, 4
    x = r.uniform(0, 1, (n, din)).astype(dtype)
    w1 = r.uniform(-0.5, 0.5, (din, dh)).astype(dtype)
    w2 = r.uniform(-0.5, 0.5, (dh, 1)).astype(dtype)
    hist = eml.WengertList(dtype)
    tape = orc.Tape(dtype)
    W1, ow1 = eml.RecordMatrix.variables(hist, eml.Matrix(w1)), tape.variables(w1)
    W2, ow2 = eml.RecordMatrix.variables(hist, eml.Matrix(w2)), tape.variables(w2)
    X = eml.RecordMatrix.constants(eml.Matrix(x), hist)
    ox = orc.Tape.constants(x, dtype)
    total, ototal = None, None
    for row in range(n):
        inp = X.range((row, 1), (0, din), history=hist)
        assert np.array_equal(inp.indexes(), np.zeros((1, din), np.int64))
        oinp = (ox[0][row:row + 1], ox[1][row:row + 1])
        out = ((inp * W1) * 2.0) * W2
        oout = tape.matmul(tape.unary(8, tape.matmul(oinp, ow1), 2.0), ow2)
        sq = out.elementwise_multiply(out)
        osq = tape.binary(34, oout, oout, mode=3)
        total = sq if total is None else total + sq
        ototal = osq if ototal is None else tape.binary(32, ototal, osq, mode=3)
    assert len(hist) == len(tape) and np.array_equal(total.indexes(), ototal[1])
    d = total.derivatives_for(0, 0)
    od = tape.derivatives(ototal[1][0, 0])
    g1, want1 = d.at_matrix(W1).data, od[ow1[1]]
    assert close(g1.ravel()[1:], want1.ravel()[1:], dtype)
    assert close(np.array(g1[0, 0]), np.array(want1[0, 0]), dtype)   # includes every row's constant-side sum
    assert close(d.at_matrix(W2).data, od[ow2[1]], dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_duplicated_and_permuted_indexes_scatter_deterministically(orc, dtype):
    """A container gathered from a variables tensor with repeats (a mask / resize that duplicates elements):
    derivatives[index] receives the sum over every element carrying that index."""
    r = np.random.default_rng(12)
    w = r.uniform(-1, 1, (3, 4)).astype(dtype)
    b = r.uniform(-1, 1, (4, 5)).astype(dtype)
    pick = np.array([[11, 0, 0, 7], [7, 7, 2, 11], [5, 5, 5, 5], [0, 1, 2, 3], [10, 9, 8, 0], [3, 3, 6, 6]])
    hist = eml.WengertList(dtype)
    tape = orc.Tape(dtype)
    W, ow = eml.RecordTensor.variables(hist, eml.Tensor([("r", 3), ("c", 4)], w)), tape.variables(w)
    B, ob = eml.RecordTensor.variables(hist, eml.Tensor([("k", 4), ("n", 5)], b)), tape.variables(b)
    numbers, indexes = w.ravel()[pick], ow[1].ravel()[pick]
    G = eml.RecordTensor.from_existing(hist, [("m", 6), ("k", 4)], numbers, indexes)
    og = (numbers.copy(), indexes.copy())
    assert np.array_equal(G.indexes(), indexes) and len(hist) == len(tape)
    y = (G * B).exp()
    oy = tape.unary(3, tape.matmul(og, ob))
    z = y.elementwise_multiply(G * B)
    oz = tape.binary(34, oy, tape.matmul(og, ob), mode=3)
    assert len(hist) == len(tape) and np.array_equal(z.indexes(), oz[1])
    runs = []
    for _ in range(2):
        d = z.derivatives_for([4, 2])
        od = tape.derivatives(oz[1][4, 2])
        assert close(d.at_tensor(W).data, od[ow[1]], dtype)
        assert close(d.at_tensor(B).data, od[ob[1]], dtype)
        assert close(d.at_tensor(G).data, od[indexes], dtype)        # each element reads derivatives[its index]
        runs.append(d.at_tensor(W).data.tobytes())
    assert runs[0] == runs[1]                                        # bit-identical run to run


def test_from_existing_without_history_is_constants_that_keep_their_indexes():
    hist = eml.WengertList(np.float64)
    w = eml.RecordTensor.variables(hist, eml.Tensor([("r", 2), ("c", 3)], np.arange(6.0).reshape(2, 3)))
    v = w.range({"c": (1, 3)}, history=None)
    assert np.array_equal(v.indexes(), [[1, 2], [4, 5]])
    y = v * 2.0                                            # history None: plain numbers, nothing recorded
    assert len(hist) == 6 and np.array_equal(y.indexes(), np.zeros((2, 2), np.int64))
    assert y.derivatives_for([0, 0]) is None


def test_from_existing_rejects_indexes_that_are_not_on_the_list():
    hist = eml.WengertList(np.float64)
    eml.RecordTensor.variables(hist, eml.Tensor([("r", 1), ("c", 2)], np.ones((1, 2))))
    with pytest.raises(eml.BadArgError):
        eml.RecordTensor.from_existing(hist, [("a", 1), ("b", 2)], np.ones((1, 2)), np.array([[0, 2]]))


# ------------------------------------------------------------------ scalar Records on the same list
@pytest.mark.parametrize("dtype", DTYPES)
def test_scalar_records_alone_are_bit_exact_with_the_reference_sweep(orc, dtype):
    """Record::variable, +, -, *, / between Records (src/differentiation/record_operations.rs) on the device-backed list:
    the sweep walks scalar entries in the reference's own order, so every derivative equals the oracle's literal tape
    bit for bit."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from scalar_record import Rec

    r = np.random.default_rng(21)
    vals = r.uniform(0.5, 2.0, 6).astype(dtype)
    hist = eml.WengertList(dtype)
    tape = orc.Tape(dtype)
    xs = [eml.Record.variable(v, hist) for v in vals]
    os_ = [Rec.variable(v, tape) for v in vals]

    def f(v):
        a = v[0] * v[1] + v[2] / v[3]
        b = (a - v[4]) * (a + 2.5) / (v[5] * v[5])
        return b * a - (a / b) + (v[0] - 3.0) * v[0]

    y, oy = f(xs), f(os_)
    assert y.index == oy.index and len(hist) == len(tape) and y.number == oy.number
    d, od = y.derivatives(), tape.derivatives(oy.index)
    assert np.array_equal(d.to_vec(), od)
    assert all(d.at(x) == od[o.index] for x, o in zip(xs, os_))


@pytest.mark.parametrize("dtype", DTYPES)
def test_network_example_mixes_container_products_and_scalar_records(orc, dtype):
    """mean_squared_loss_training, src/neural_networks.rs:95-131: each input row is re-wrapped (from_existing), goes
    through the layers' container products, its single output is taken out as a scalar Record, and the squared errors are
    folded with scalar Record arithmetic.  One list, one index space, one sweep."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from scalar_record import Rec

    r = np.random.default_rng(22)
    n, din, dh = 4, 3, 5
    x = r.uniform(0, 1, (n, din)).astype(dtype)
    labels = r.uniform(0, 1, (1, n)).astype(dtype)
    w1 = r.uniform(-0.5, 0.5, (din, dh)).astype(dtype)
    w2 = r.uniform(-0.5, 0.5, (dh, 1)).astype(dtype)
    hist = eml.WengertList(dtype)
    tape = orc.Tape(dtype)
    W1, ow1 = eml.RecordMatrix.variables(hist, eml.Matrix(w1)), tape.variables(w1)
    W2, ow2 = eml.RecordMatrix.variables(hist, eml.Matrix(w2)), tape.variables(w2)
    X, ox = eml.RecordMatrix.constants(eml.Matrix(x), hist), orc.Tape.constants(x, dtype)
    Y, oy = eml.RecordMatrix.constants(eml.Matrix(labels), hist), orc.Tape.constants(labels, dtype)
    acc, oacc = eml.Record.constant(0.0, dtype), Rec.constant(0.0, dtype)
    for row in range(n):
        inp = X.range((row, 1), (0, din), history=hist)
        out = ((inp * W1) * 0.5) * W2                                   # 1 x 1 container
        oinp = (ox[0][row:row + 1], ox[1][row:row + 1])
        oout = tape.matmul(tape.unary(8, tape.matmul(oinp, ow1), 0.5), ow2)
        output = out.get_as_record(0, 0)
        ooutput = Rec(oout[0][0, 0], tape, int(oout[1][0, 0]))
        correct = Y.get_as_record(0, row)                               # a constant: history None
        ocorrect = Rec(oy[0][0, row], None, 0)
        assert output.index == ooutput.index and correct.history is None
        acc = acc + (correct - output) * (correct - output)
        oacc = oacc + (ocorrect - ooutput) * (ocorrect - ooutput)
    loss, oloss = acc / float(n), oacc / float(n)
    assert loss.index == oloss.index and len(hist) == len(tape)
    rtol = 2e-5 if dtype == np.float32 else 1e-12
    assert abs(float(loss.number) - float(oloss.number)) <= rtol * abs(float(oloss.number))
    d, od = loss.derivatives(), tape.derivatives(oloss.index)
    g1, want1 = d.at_matrix(W1).data, od[ow1[1]]
    assert close(g1.ravel()[1:], want1.ravel()[1:], dtype) and close(np.array(g1[0, 0]), np.array(want1[0, 0]), dtype)
    assert close(d.at_matrix(W2).data, od[ow2[1]], dtype)
