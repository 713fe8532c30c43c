"""Fused elementwise groups of the device tape against the one-kernel-per-node path (EML_FUSE=0), bit for bit.

A run of container-level unary / binary operators (RecordTensor::unary / binary and their front-ends, reference
src/differentiation/container_record/container_operations.rs:1644-1804, swapped.rs:25-45, mod.rs:1224-1247) is evaluated
lazily as ONE kernel in the forward pass and ONE in the sweep.  Fusion may not change any observable of the reference:
tape length, every Index, every number, every derivative -- including Vec::from(Derivatives), which exposes the entries
of containers nobody holds any more (materialised on demand).  Both paths are built from the same separately rounded
rules, so the comparison is np.array_equal on the raw bytes; the unfused path itself is pinned to the oracle's literal
scalar tape by tests/test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200 import _ffi

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]


def tape(dtype, fuse):
    old = os.environ.get("EML_FUSE")
    os.environ["EML_FUSE"] = "1" if fuse else "0"
    try:
        return eml.WengertList(dtype)  # the switch is read when the list is created
    finally:
        if old is None:
            del os.environ["EML_FUSE"]
        else:
            os.environ["EML_FUSE"] = old


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def sigmoid(z):
    return ((-z).exp() + 1.0).div_swapped(1.0)


def names(shape):
    return [("r", shape[0]), ("c", shape[1])]


def run_both(dtype, program):
    """program(list, launches) -> dict of arrays; run on a fused and on an unfused tape."""
    out = []
    for fuse in (True, False):
        hist = tape(dtype, fuse)
        before = eml.kernel_launches()
        res = program(hist)
        res["launches"] = eml.kernel_launches() - before
        out.append(res)
    return out


def assert_same(fused, unfused, skip=("launches",)):
    assert fused.keys() == unfused.keys()
    for k in fused:
        if k in skip:
            continue
        assert same_bits(fused[k], unfused[k]), f"{k}: fused and unfused paths differ"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(64, 48), (7, 5), (1, 1), (33, 3)])
def test_sigmoid_chain_values_derivatives_indexes(dtype, shape):
    """The reference NN examples' activation, 1 / (1 + e^-x), as four container operators: one forward kernel, one
    sweep kernel; (7, 5) and (33, 3) take the scalar (non 16-byte-vector) variant of the kernels."""
    r = np.random.default_rng(1)
    w = r.uniform(-2, 2, shape).astype(dtype)

    def program(hist):
        W = eml.RecordTensor.variables(hist, eml.Tensor(names(shape), w))
        y = sigmoid(W)
        d = y.derivatives_for([shape[0] - 1, shape[1] // 2])
        return {"y": y.numbers().data, "idx": y.indexes(), "len": np.array(len(hist)), "dW": d.at_tensor(W).data,
                "dy": d.at_tensor(y).data, "vec": d.to_vec()}

    fused, unfused = run_both(dtype, program)
    assert_same(fused, unfused)
    assert fused["launches"] < unfused["launches"]
    # and the numbers are the sigmoid's: value and derivative y (1 - y) at the seeded element only
    y = 1.0 / (1.0 + np.exp(-w.astype(np.float64)))
    assert np.allclose(fused["y"], y, rtol=1e-5 if dtype == np.float32 else 1e-13)
    want = np.zeros(shape)
    e = (shape[0] - 1, shape[1] // 2)
    want[e] = (y * (1 - y))[e]
    assert np.allclose(fused["dW"], want, rtol=1e-4 if dtype == np.float32 else 1e-12)


@pytest.mark.parametrize("dtype", DTYPES)
def test_every_rule_in_one_group(dtype):
    """All 14 unary and 5 binary rules (src/differentiation/functions.rs) chained so that they form fused groups with
    several inputs, reuse of one container by two consumers (x * x: left parent first, then right, :641-644), a
    constant second operand, and held as well as dropped intermediates."""
    r = np.random.default_rng(2)
    shape = (24, 20)
    a = r.uniform(0.5, 1.5, shape).astype(dtype)
    b = r.uniform(0.5, 1.5, shape).astype(dtype)
    k = r.uniform(0.5, 1.5, shape).astype(dtype)

    def program(hist):
        A = eml.RecordTensor.variables(hist, eml.Tensor(names(shape), a))
        B = eml.RecordTensor.variables(hist, eml.Tensor(names(shape), b))
        K = eml.RecordTensor.constants(eml.Tensor(names(shape), k), hist)
        t1 = (A.sin() + B.cos()).sqrt()                      # three unary-ish rules, two tracked inputs
        t2 = (t1 * 1.5 - 0.25) / 3.0 + 0.125                 # constant-derivative rules only
        t3 = t2.pow(1.25).ln().exp()                         # pow_c, ln, exp
        t4 = t3.sub_swapped(4.0).div_swapped(2.0)            # c - x, c / x
        t5 = t4.elementwise_multiply(t4)                     # one container, both parents
        t6 = t5.elementwise_divide(A) - K                    # tracked / tracked, tracked - constant
        t7 = (t6 + B).elementwise_multiply(K)                # tracked + tracked, tracked * constant
        t8 = -t7
        d = t8.derivatives_for([3, 4])
        res = {"t8": t8.numbers().data, "t5": t5.numbers().data, "idx8": t8.indexes(), "idx2": t2.indexes(),
               "len": np.array(len(hist)), "dA": d.at_tensor(A).data, "dB": d.at_tensor(B).data,
               "dt5": d.at_tensor(t5).data, "dt2": d.at_tensor(t2).data, "vec": d.to_vec()}
        # derivative of an interior element: everything above it contributes 0 * derivative
        d2 = t2.derivatives_for([0, 0])
        res["d2A"] = d2.at_tensor(A).data
        res["vec2"] = d2.to_vec()
        return res

    fused, unfused = run_both(dtype, program)
    assert_same(fused, unfused)
    assert fused["launches"] < unfused["launches"]


@pytest.mark.parametrize("dtype", DTYPES)
def test_groups_between_matmuls_and_the_constant_operand_quirk(dtype):
    """A two-layer network written with container operators: groups are closed by the matrix products, a group's
    output is both stored (the next product reads it) and differentiated through; inputs are constants, so the
    reference's Index-0 behaviour (container_operations.rs:1291-1296) is part of the sweep."""
    r = np.random.default_rng(3)
    n, i, h, o = 96, 40, 24, 10
    x = r.uniform(0, 1, (n, i)).astype(dtype)
    lab = np.eye(o, dtype=dtype)[r.integers(0, o, n)]
    w1 = r.uniform(-0.5, 0.5, (i, h)).astype(dtype)
    w2 = r.uniform(-0.5, 0.5, (h, o)).astype(dtype)

    def program(hist):
        W1 = eml.RecordTensor.variables(hist, eml.Tensor([("i", i), ("h", h)], w1))
        W2 = eml.RecordTensor.variables(hist, eml.Tensor([("h", h), ("o", o)], w2))
        X = eml.RecordTensor.constants(eml.Tensor([("n", n), ("i", i)], x), hist)
        Y = eml.RecordTensor.constants(eml.Tensor([("n", n), ("o", o)], lab), hist)
        L = eml.RecordTensor.constants(eml.Tensor([("u", 1), ("n", n)], np.ones((1, n), dtype)), hist)
        R = eml.RecordTensor.constants(eml.Tensor([("o", o), ("v", 1)], np.ones((o, 1), dtype)), hist)
        losses = []
        for _ in range(3):
            out = sigmoid(sigmoid(X * W1) * W2)
            diff = out - Y
            loss = ((L * diff.elementwise_multiply(diff)) * R) / float(n)
            d = loss.derivatives_for([0, 0])
            g1, g2 = d.at_tensor(W1).data, d.at_tensor(W2).data
            losses.append(loss.numbers().data.copy())
            eml.sgd_update(W1, d, 0.5)
            eml.sgd_update(W2, d, 0.5)
            hist.clear()
            W1.reset()
            W2.reset()
        return {"losses": np.concatenate(losses), "g1": g1, "g2": g2, "w1": W1.numbers().data, "w2": W2.numbers().data}

    fused, unfused = run_both(dtype, program)
    assert_same(fused, unfused)
    assert fused["launches"] < unfused["launches"]


def test_group_limits_split_long_chains_without_changing_bits():
    """More operators than one program holds (EW_MAX_OPS / slot budget): the chain is cut into several groups."""
    r = np.random.default_rng(4)
    shape = (16, 16)
    a = r.uniform(0.5, 1.5, shape)

    def program(hist):
        A = eml.RecordTensor.variables(hist, eml.Tensor(names(shape), a))
        y = A
        held = []
        for step in range(40):
            y = (y * 1.01).sin().exp() if step % 3 == 0 else y.elementwise_multiply(A) + 0.5
            if step % 7 == 0:
                held.append(y)
        d = y.derivatives_for([5, 5])
        res = {"y": y.numbers().data, "dA": d.at_tensor(A).data, "vec": d.to_vec(), "len": np.array(len(hist))}
        for n, hcont in enumerate(held):
            res[f"h{n}"] = hcont.numbers().data
            res[f"dh{n}"] = d.at_tensor(hcont).data
        return res

    fused, unfused = run_both(np.float64, program)
    assert_same(fused, unfused)


def test_dropped_containers_are_never_stored_but_their_entries_exist():
    """A container released before anybody needed its numbers is not written by the forward kernel; its tape entries
    still count (len, indexes of later containers) and Derivatives::at(index) of such an entry is materialised on
    demand with the bits the unfused sweep stores."""
    r = np.random.default_rng(5)
    shape = (12, 8)
    a = r.uniform(-1, 1, shape).astype(np.float32)

    def program(hist):
        A = eml.RecordTensor.variables(hist, eml.Tensor(names(shape), a))
        y = sigmoid(A)                       # entries 96..479: neg, exp, +1 are dropped, 1/x is held
        idx = y.indexes()
        d = y.derivatives_for([2, 3])
        interior = np.array([float(d.at_index(96 + k * 96 + 19)) for k in range(4)])  # element (2, 3) of each node
        return {"idx": idx, "interior": interior, "len": np.array(len(hist))}

    fused, unfused = run_both(np.float32, program)
    assert_same(fused, unfused)
    assert fused["idx"][0, 0] == 96 * 4 and int(fused["len"]) == 96 * 5
    assert fused["interior"][3] == 1.0


def test_recorded_step_with_fused_groups_replays_bit_for_bit():
    """The CUDA-graph recording of a step captures the fused kernels (the flush points are inside the recording)."""
    r = np.random.default_rng(6)
    n, i, h, o = 128, 64, 32, 10
    x = r.uniform(0, 1, (n, i)).astype(np.float32)
    lab = np.eye(o, dtype=np.float32)[r.integers(0, o, n)]
    w1 = r.uniform(-0.5, 0.5, (i, h)).astype(np.float32)
    w2 = r.uniform(-0.5, 0.5, (h, o)).astype(np.float32)

    def run(recorded):
        hist = tape(np.float32, True)
        W1 = eml.RecordTensor.variables(hist, eml.Tensor([("i", i), ("h", h)], w1))
        W2 = eml.RecordTensor.variables(hist, eml.Tensor([("h", h), ("o", o)], w2))
        X = eml.RecordTensor.constants(eml.Tensor([("n", n), ("i", i)], x), hist)
        Y = eml.RecordTensor.constants(eml.Tensor([("n", n), ("o", o)], lab), hist)
        L = eml.RecordTensor.constants(eml.Tensor([("u", 1), ("n", n)], np.ones((1, n), np.float32)), hist)
        R = eml.RecordTensor.constants(eml.Tensor([("o", o), ("v", 1)], np.ones((o, 1), np.float32)), hist)
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
        return np.array(losses), W1.numbers().data, W2.numbers().data

    le, w1e, w2e = run(False)
    lr, w1r, w2r = run(True)
    assert same_bits(le, lr) and same_bits(w1e, w1r) and same_bits(w2e, w2r)
    assert le[-1] < le[0]


def _env_tape(dtype, **env):
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


@pytest.mark.parametrize("special", [False, True])
@pytest.mark.parametrize("n,i,h", [(1024, 200, 384), (520, 128, 260), (8192, 784, 512)])
def test_activation_chain_rides_in_the_gemm_epilogue(n, i, h, special):
    """sigmoid(X * W) at shapes the f32 CTA-pair kernel computes in one launch per tile (no split-K): the product is
    deferred, the four-operator group that follows it goes into the GEMM's epilogue (one launch fewer) and the numbers,
    the derivatives and every Index are those of the separate kernels (EML_EPILOGUE=0) and of one kernel per node
    (EML_FUSE=0), bit for bit -- also where the non-finite repair rewrites elements of the product (an infinite input)."""
    r = np.random.default_rng(21)
    x = r.uniform(-1, 1, (n, i)).astype(np.float32)
    w = r.uniform(-0.5, 0.5, (i, h)).astype(np.float32)
    if special:
        x[3, 5] = np.inf
        x[n - 1, 0] = -np.inf
        w[7, 9] = np.float32(3.0e38)

    def program(hist):
        before = eml.kernel_launches()
        W = eml.RecordTensor.variables(hist, eml.Tensor([("i", i), ("h", h)], w))
        X = eml.RecordTensor.constants(eml.Tensor([("n", n), ("i", i)], x), hist)
        z = X * W
        y = sigmoid(z)
        yv = y.numbers().data.copy()
        zv = z.numbers().data.copy()
        launches = eml.kernel_launches() - before
        d = y.derivatives_for([n // 2, h - 1])
        return {"y": yv, "z": zv, "gw": d.at_tensor(W).data, "idx": y.indexes(), "len": np.array(len(hist)),
                "launches": launches}

    fused = program(_env_tape(np.float32, EML_FUSE=1, EML_EPILOGUE=1))
    separate = program(_env_tape(np.float32, EML_FUSE=1, EML_EPILOGUE=0))
    per_node = program(_env_tape(np.float32, EML_FUSE=0))
    assert_same(fused, separate)
    assert_same(fused, per_node)
    assert fused["launches"] == separate["launches"] - 1, "the activation did not go into the GEMM epilogue"
