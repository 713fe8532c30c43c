"""The reverse sweep's second stream (csrc/tape.cu, eml_tape::side) against the single-stream sweep, bit for bit.

Record::try_derivatives (reference src/differentiation.rs:623-648) walks the list once, sequentially.  On the device the
two products of a matmul node (dA = G B^T, dB = A^T G; record_scalar_product's two parents per product, reference
container_operations.rs:1291-1296) are independent, and so are the sums a constant operand sends to derivatives[0]:
the half nothing downstream waits for runs on a second stream.  That may not change a single bit: every derivative,
derivatives[0] (the parked sums are folded in sweep order, also when there are more of them than slots), the updated
weights and a recorded step's replays are compared with EML_SIDE=0; the constant-operand sums are also checked against
the oracle's literal scalar tape.
"""
import os

import numpy as np
import pytest

import easy_ml_b200 as eml

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]


def tape(dtype, side):
    old = os.environ.get("EML_SIDE")
    os.environ["EML_SIDE"] = "1" if side else "0"
    try:
        return eml.WengertList(dtype)  # the switch is read when the list is created
    finally:
        if old is None:
            del os.environ["EML_SIDE"]
        else:
            os.environ["EML_SIDE"] = old


def T(names, a):
    return eml.Tensor(list(zip(names, a.shape)), a)


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def sigmoid(z):
    return ((-z).exp() + 1.0).div_swapped(1.0)


def both(dtype, program):
    res = [program(tape(dtype, side)) for side in (True, False)]
    assert res[0].keys() == res[1].keys()
    for k in res[0]:
        assert same_bits(res[0][k], res[1][k]), f"{k}: the two-stream sweep differs from the single-stream sweep"
    return res[0]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,i,h,o", [(96, 40, 24, 10), (520, 300, 260, 12)])
def test_three_layer_network_with_a_shared_weight(dtype, n, i, h, o):
    """Inputs constant (their sums go to derivatives[0] on the second stream), every weight a leaf (its gradient product
    goes to the second stream), one weight used by three products (a later use has to wait for the earlier one's
    product on the second stream).  (520, 300, 260) takes the tensor-core kernels."""
    r = np.random.default_rng(11)
    x = r.uniform(0, 1, (n, i)).astype(dtype)
    lab = np.eye(o, dtype=dtype)[r.integers(0, o, n)]
    w1 = r.uniform(-0.3, 0.3, (i, h)).astype(dtype)
    wh = r.uniform(-0.3, 0.3, (h, h)).astype(dtype)
    w2 = r.uniform(-0.3, 0.3, (h, o)).astype(dtype)

    def program(hist):
        W1 = eml.RecordTensor.variables(hist, T("ih", w1))
        WH = eml.RecordTensor.variables(hist, T("hg", wh))
        W2 = eml.RecordTensor.variables(hist, T("ho", w2))
        X = eml.RecordTensor.constants(T("ni", x), hist)
        Y = eml.RecordTensor.constants(T("no", lab), hist)
        L = eml.RecordTensor.constants(T("un", np.ones((1, n), dtype)), hist)
        R = eml.RecordTensor.constants(T("ov", np.ones((o, 1), dtype)), hist)
        out = {}
        for step in range(2):
            a1 = sigmoid(X * W1)
            a2 = sigmoid(a1 * WH)
            a3 = sigmoid(a2 * WH)            # WH again: both products add into the same derivatives
            a4 = a3 * WH                     # ... and a third use
            mix = sigmoid(a4) - a2           # a2 has two consumers
            y = sigmoid(mix * W2)
            diff = y - Y
            loss = ((L * diff.elementwise_multiply(diff)) * R) / float(n)
            d = loss.derivatives_for([0, 0])
            out[f"g1_{step}"] = d.at_tensor(W1).data
            out[f"gh_{step}"] = d.at_tensor(WH).data
            out[f"g2_{step}"] = d.at_tensor(W2).data
            out[f"d0_{step}"] = np.array(d.at_index(0))
            out[f"loss_{step}"] = loss.numbers().data.copy()
            eml.sgd_update(W1, d, 0.5)
            eml.sgd_update(WH, d, 0.5)
            eml.sgd_update(W2, d, 0.5)
            hist.clear()
            W1.reset()
            WH.reset()
            W2.reset()
        out["w1"] = W1.numbers().data
        out["wh"] = WH.numbers().data
        out["w2"] = W2.numbers().data
        return out

    both(dtype, program)


@pytest.mark.parametrize("dtype", DTYPES)
def test_variables_times_computed_and_computed_times_variables(dtype):
    """The leaf may be the LEFT operand (W * H layouts) and both operands may be leaves."""
    r = np.random.default_rng(12)
    a = r.uniform(-1, 1, (48, 20)).astype(dtype)
    b = r.uniform(-1, 1, (20, 36)).astype(dtype)
    c = r.uniform(-1, 1, (36, 7)).astype(dtype)

    def program(hist):
        A = eml.RecordTensor.variables(hist, T("mk", a))
        B = eml.RecordTensor.variables(hist, T("kn", b))
        Cc = eml.RecordTensor.variables(hist, T("np", c))
        ab = A * B                      # both leaves
        h = sigmoid(ab)
        out = (h * Cc).elementwise_multiply(h * Cc)    # Cc: leaf on the right, twice
        d = out.derivatives_for([5, 3])
        return {"ga": d.at_tensor(A).data, "gb": d.at_tensor(B).data, "gc": d.at_tensor(Cc).data}

    both(dtype, program)

    wl = r.uniform(-1, 1, (30, 48)).astype(dtype)

    def program2(hist):
        WL = eml.RecordTensor.variables(hist, T("om", wl))
        A = eml.RecordTensor.variables(hist, T("mk", a))
        B = eml.RecordTensor.constants(T("kn", b), hist)
        hcomp = sigmoid(A * B)          # constant on the right: row-sum form of the parked sum
        y = WL * hcomp                  # leaf on the left
        d = y.derivatives_for([2, 2])
        return {"gwl": d.at_tensor(WL).data, "ga": d.at_tensor(A).data, "d0": np.array(d.at_index(0))}

    both(dtype, program2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("m", [3, 6])
def test_more_constant_operand_sums_than_slots(orc, dtype, m):
    """300 products with a constant operand in one sweep: more parked sums than the arena has slots (256 f32, 128 f64),
    so they are folded on the way.  m = 3 takes the thin kernels, m = 6 the column-sum kernel on the second stream.
    derivatives[0] is the in-order sum of all of them: compared with the oracle's scalar tape."""
    r = np.random.default_rng(13)
    k, n, count = 5, 4, 300
    xs = [r.uniform(-1, 1, (m, k)).astype(dtype) for _ in range(count)]
    w = r.uniform(-1, 1, (k, n)).astype(dtype)

    def program(hist):
        W = eml.RecordTensor.variables(hist, T("kn", w))
        acc = None
        for x in xs:
            c = eml.RecordTensor.constants(T("mk", x), hist) * W
            acc = c if acc is None else acc + c
        sq = acc.elementwise_multiply(acc)
        d = sq.derivatives_for([m - 1, 1])
        return {"gw": d.at_tensor(W).data, "d0": np.array(d.at_index(0))}

    got = both(dtype, program)

    t = orc.Tape(dtype)
    ow = t.variables(w)
    oacc = None
    for x in xs:
        oc = t.matmul(orc.Tape.constants(x, dtype), ow)
        oacc = oc if oacc is None else t.binary(32, oacc, oc, mode=3)
    osq = t.binary(34, oacc, oacc, mode=3)
    od = t.derivatives(osq[1][m - 1, 1])
    rtol = 2e-4 if dtype == np.float32 else 1e-11
    scale = float(np.abs(od).max())
    assert np.allclose(got["gw"], od[ow[1]], rtol=rtol, atol=rtol * scale)
    # derivatives[0] = W[0, 0]'s own derivative + every constant operand's sum
    assert np.allclose(float(got["d0"]), float(od[0]), rtol=rtol, atol=rtol * scale * count)


def test_recorded_step_with_the_second_stream_replays_bit_for_bit():
    """Fork and join are captured as parallel branches of the CUDA graph."""
    r = np.random.default_rng(14)
    n, i, h, o = 256, 64, 32, 10
    x = r.uniform(0, 1, (n, i)).astype(np.float32)
    lab = np.eye(o, dtype=np.float32)[r.integers(0, o, n)]
    w1 = r.uniform(-0.5, 0.5, (i, h)).astype(np.float32)
    w2 = r.uniform(-0.5, 0.5, (h, o)).astype(np.float32)

    def run(recorded, side):
        hist = tape(np.float32, side)
        W1 = eml.RecordTensor.variables(hist, T("ih", w1))
        W2 = eml.RecordTensor.variables(hist, T("ho", w2))
        X = eml.RecordTensor.constants(T("ni", x), hist)
        Y = eml.RecordTensor.constants(T("no", lab), hist)
        L = eml.RecordTensor.constants(T("un", np.ones((1, n), np.float32)), hist)
        R = eml.RecordTensor.constants(T("ov", np.ones((o, 1), np.float32)), hist)
        pinned = eml.PinnedBuffer((1,), np.float32)

        def body():
            out = sigmoid(sigmoid(X * W1) * W2)
            diff = out - Y
            loss = ((L * diff.elementwise_multiply(diff)) * R) / float(n)
            d = loss.derivatives_for([0, 0])
            eml.sgd_update(W1, d, 0.5)
            eml.sgd_update(W2, d, 0.5)
            arrival = loss.numbers_async(pinned)
            del out, diff, loss, d
            hist.clear()
            W1.reset()
            W2.reset()
            return arrival

        losses = []
        for _ in range(2):
            eml.wait_arrival(body())
            losses.append(float(pinned.array[0]))
        if recorded:
            step = hist.record(body)
            for _ in range(4):
                eml.wait_arrival(step.launch())
                losses.append(float(pinned.array[0]))
        else:
            for _ in range(4):
                eml.wait_arrival(body())
                losses.append(float(pinned.array[0]))
        return np.array(losses, np.float32), W1.numbers().data, W2.numbers().data

    base = run(False, False)
    for recorded, side in ((False, True), (True, True), (True, False)):
        got = run(recorded, side)
        for a, b in zip(base, got):
            assert same_bits(a, b), f"recorded={recorded} side={side}"


def tape_env(dtype, **env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return eml.WengertList(dtype)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,i,h,o", [
    (96, 40, 24, 10),        # FMA kernels, thin and short-contraction products
    (520, 300, 260, 12),     # tensor-core products with split-K (f32) / DMMA (f64), tall-skinny A^T B
    (1024, 2048, 2048, 300), # f32: the in-kernel-split GEMM and the CTA-pair kernel without split-K
])
def test_first_contributions_of_products_need_no_zero_fill(dtype, n, i, h, o):
    """vec![0; len] (reference src/differentiation.rs:627) is not materialised for buffers a product reaches first: the
    product is asked for 0 + sum (GemmOptions::zero_plus).  Compared bit for bit -- signs of zero included: X has an
    all-zero column, so a whole row of dW1 is a sum of signed zeros -- with accumulation into zero-filled buffers
    (EML_ZERO_PLUS=0)."""
    r = np.random.default_rng(17)
    x = r.uniform(-1, 1, (n, i)).astype(dtype)
    x[:, 3] = 0.0
    w1 = r.uniform(-0.05, 0.05, (i, h)).astype(dtype)
    w2 = r.uniform(-0.05, 0.05, (h, o)).astype(dtype)

    def program(hist):
        W1 = eml.RecordTensor.variables(hist, T("ih", w1))
        W2 = eml.RecordTensor.variables(hist, T("ho", w2))
        X = eml.RecordTensor.variables(hist, T("ni", x))
        y = sigmoid(X * W1) * W2
        out = y.elementwise_multiply(y) - y
        d = out.derivatives_for([n - 1, o // 2])
        return {"g1": d.at_tensor(W1).data, "g2": d.at_tensor(W2).data, "gx": d.at_tensor(X).data}

    a = program(tape_env(dtype, EML_ZERO_PLUS=1))
    b = program(tape_env(dtype, EML_ZERO_PLUS=0))
    for k in a:
        assert same_bits(a[k], b[k]), f"{k}: 0 + product differs from accumulating into zeros"
    assert not np.signbit(a["g1"][3]).any()  # sums of signed zeros added to +0 are +0


@pytest.mark.parametrize("recorded", [False, True])
def test_producers_write_the_small_part_planes_of_the_next_product(recorded):
    """f32: the weight update writes the TF32 small parts of the new weights (the next forward product reads them as
    its right operand in place), and the sweep kernel of an activation writes the small parts of the derivatives the
    backward product dW = X^T dZ reads -- two launches fewer per step, the same bits (EML_FUSE_PACKS=0), eagerly and
    as a recorded step."""
    r = np.random.default_rng(19)
    n, i, h, o = 1024, 256, 512, 10
    x = r.uniform(0, 1, (n, i)).astype(np.float32)
    lab = np.eye(o, dtype=np.float32)[r.integers(0, o, n)]
    w1 = r.uniform(-0.1, 0.1, (i, h)).astype(np.float32)
    w2 = r.uniform(-0.1, 0.1, (h, o)).astype(np.float32)

    def run(fuse_packs):
        hist = tape_env(np.float32, EML_FUSE_PACKS=1 if fuse_packs else 0)
        W1 = eml.RecordTensor.variables(hist, T("ih", w1))
        W2 = eml.RecordTensor.variables(hist, T("ho", w2))
        X = eml.RecordTensor.constants(T("ni", x), hist)
        Y = eml.RecordTensor.constants(T("no", lab), hist)
        L = eml.RecordTensor.constants(T("un", np.ones((1, n), np.float32)), hist)
        R = eml.RecordTensor.constants(T("ov", np.ones((o, 1), np.float32)), hist)
        pinned = eml.PinnedBuffer((1,), np.float32)

        def body():
            out = sigmoid(sigmoid(X * W1) * W2)
            diff = out - Y
            loss = ((L * diff.elementwise_multiply(diff)) * R) / float(n)
            d = loss.derivatives_for([0, 0])
            eml.sgd_update(W1, d, 0.5)
            eml.sgd_update(W2, d, 0.5)
            arrival = loss.numbers_async(pinned)
            del out, diff, loss, d
            hist.clear()
            W1.reset()
            W2.reset()
            return arrival

        losses = []
        for _ in range(3):
            eml.wait_arrival(body())
            losses.append(float(pinned.array[0]))
        before = eml.kernel_launches()
        if recorded:
            step = hist.record(body)
            per_step = eml.kernel_launches() - before
            for _ in range(3):
                eml.wait_arrival(step.launch())
                losses.append(float(pinned.array[0]))
        else:
            eml.wait_arrival(body())
            per_step = eml.kernel_launches() - before
            losses.append(float(pinned.array[0]))
            for _ in range(2):
                eml.wait_arrival(body())
                losses.append(float(pinned.array[0]))
        return np.array(losses, np.float32), W1.numbers().data, W2.numbers().data, per_step

    fused, plain = run(True), run(False)
    for a, b in zip(fused[:3], plain[:3]):
        assert same_bits(a, b)
    assert fused[3] == plain[3] - 2, f"{fused[3]} launches per step with producer-written planes, {plain[3]} without"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,i,h,o,hold", [
    (256, 64, 256, 10, False),   # the NN step's pattern in small: sigmoid(X W1) W2, row length 256
    (256, 64, 256, 10, True),    # ... with the hidden activations still held (their derivatives are stored as well)
    (192, 48, 512, 16, False),   # longest contraction the kernels form on the fly
    (130, 40, 128, 3, False),    # ragged row count
])
def test_short_products_of_the_sweep_ride_in_the_group_kernel(dtype, n, i, h, o, hold):
    """dH = dY W2^T of an output layer (contraction of <= 16 numbers) is owed to the last node of the activation's fused
    group and to nothing else: the group's backward kernel forms it on the fly instead of reading it from the arena, the
    product is not launched and dH is not stored (csrc/tape.cu ProductFeed, csrc/ew_fused.cu EW_INIT_PRODUCT).  Every
    derivative -- including the ones of the hidden activations, materialised on demand by Vec::from(Derivatives) -- must
    carry the bits of the sweep that launches the product (EML_FEED_PRODUCTS=0); one launch fewer."""
    r = np.random.default_rng(23)
    x = r.uniform(0, 1, (n, i)).astype(dtype)
    w1 = r.uniform(-0.3, 0.3, (i, h)).astype(dtype)
    w2 = r.uniform(-0.3, 0.3, (h, o)).astype(dtype)

    def program(hist):
        W1 = eml.RecordTensor.variables(hist, T("ih", w1))
        W2 = eml.RecordTensor.variables(hist, T("ho", w2))
        X = eml.RecordTensor.constants(T("ni", x), hist)
        hidden = sigmoid(X * W1)
        y = hidden * W2
        if not hold:
            del hidden
        out = y.elementwise_multiply(y) - y
        before = eml.kernel_launches()
        d = out.derivatives_for([n - 1, o // 2])
        launches = eml.kernel_launches() - before
        res = {"g1": d.at_tensor(W1).data, "g2": d.at_tensor(W2).data}
        if hold:
            res["gh"] = d.at_tensor(hidden).data
        res["all"] = d.to_vec()
        return res, launches

    (a, la), (b, lb) = program(tape_env(dtype, EML_FEED_PRODUCTS=1)), program(tape_env(dtype, EML_FEED_PRODUCTS=0))  # (opt-in)
    for k in a:
        assert same_bits(a[k], b[k]), f"{k}: the product formed inside the group kernel differs from the launched one"
    # (a block's stride must be a whole number of rows: 256 threads x one 16-byte vector)
    folded = (256 * (16 // np.dtype(dtype).itemsize)) % h == 0 and np.dtype(dtype).itemsize * h * o <= 32768
    assert la == lb - (1 if folded else 0), f"{la} launches with the product folded in, {lb} without"


@pytest.mark.parametrize("recorded", [False, True])
def test_split_k_weight_gradient_is_reduced_by_the_weight_update(recorded):
    """f32: dW1 = X^T dZ of a wide layer runs split-K on the CTA-pair kernel; when nothing else contributes to W1's
    derivatives the product leaves its partial sums to the reader (GemmOptions::defer_reduce): the weight update adds them
    in split order itself -- one launch fewer -- and at_tensor launches the reduction on demand.  Same bits as the
    product reducing them itself (the default, EML_DEFER_REDUCE=0: the deferred form measured slower), eagerly and as a
    recorded step."""
    r = np.random.default_rng(29)
    n, i, h, o = 2048, 300, 512, 10
    x = r.uniform(0, 1, (n, i)).astype(np.float32)
    lab = np.eye(o, dtype=np.float32)[r.integers(0, o, n)]
    w1 = r.uniform(-0.1, 0.1, (i, h)).astype(np.float32)
    w2 = r.uniform(-0.1, 0.1, (h, o)).astype(np.float32)

    def run(defer):
        hist = tape_env(np.float32, EML_DEFER_REDUCE=1 if defer else 0)
        W1 = eml.RecordTensor.variables(hist, T("ih", w1))
        W2 = eml.RecordTensor.variables(hist, T("ho", w2))
        X = eml.RecordTensor.constants(T("ni", x), hist)
        Y = eml.RecordTensor.constants(T("no", lab), hist)
        seen = {}

        def body(read_gradients=False):
            out = sigmoid(sigmoid(X * W1) * W2)
            diff = out - Y
            loss = diff.elementwise_multiply(diff)
            d = loss.derivatives_for([n // 2, 3])
            if read_gradients:  # readers other than the weight update: the reduction is launched on demand
                seen["g1"] = d.at_tensor(W1).data
                seen["all"] = d.at_tensor(W2).data
            eml.sgd_update(W1, d, 0.5)
            eml.sgd_update(W2, d, 0.5)
            del out, diff, loss, d
            hist.clear()
            W1.reset()
            W2.reset()

        body(read_gradients=True)
        before = eml.kernel_launches()
        if recorded:
            step = hist.record(lambda: body())
            per_step = eml.kernel_launches() - before
            for _ in range(3):
                step.launch()
        else:
            body()
            per_step = eml.kernel_launches() - before
            for _ in range(2):
                body()
        return seen["g1"], seen["all"], W1.numbers().data, W2.numbers().data, per_step

    a, b = run(True), run(False)
    for u, v in zip(a[:4], b[:4]):
        assert same_bits(u, v)
    assert a[4] == b[4] - 1, f"{a[4]} launches per step with the deferred reduction, {b[4]} without"
