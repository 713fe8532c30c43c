"""Device-resident Matrix / Tensor (easy-ml_b200/device.py): operator chains whose
intermediates never leave HBM.  Elementwise results and data movement are bit-exact against numpy (one IEEE operation
per element, the reference's definition); products are the same kernels as the host entry points, so they agree bit
for bit with `Matrix.__mul__`, which the parity tests pin against the oracle."""
import ctypes as C
import itertools

import numpy as np
import pytest

import easy_ml_b200 as eml
from easy_ml_b200._ffi import check, lib

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

DTYPES = [np.float32, np.float64]


@pytest.mark.parametrize("dtype", DTYPES)
def test_matrix_elementwise_and_scalar_forms_are_bit_exact(dtype):
    r = np.random.default_rng(3)
    a = r.uniform(-4, 4, (37, 53)).astype(dtype)
    b = r.uniform(-4, 4, (37, 53)).astype(dtype)
    da, db = eml.DeviceMatrix.from_host(eml.Matrix(a)), eml.DeviceMatrix.from_host(eml.Matrix(b))
    assert np.array_equal((da + db).to_host().data, a + b)
    assert np.array_equal((da - db).to_host().data, a - b)
    assert np.array_equal((-da).to_host().data, -a)
    c = dtype(1.7)
    assert np.array_equal((da + 1.7).to_host().data, a + c)
    assert np.array_equal((da - 1.7).to_host().data, a - c)
    assert np.array_equal((da * 1.7).to_host().data, a * c)
    assert np.array_equal((da / 1.7).to_host().data, a / c)
    with pytest.raises(eml.Panic, match=r"Mismatched matrices, left is 37x53, right is 53x37, \+ is only defined for MxN \+ MxN"):
        da + db.transpose()
    with pytest.raises(eml.Panic, match=r"Mismatched matrices, left is 37x53, right is 53x37, - is only defined for MxN - MxN"):
        da - db.transpose()
    with pytest.raises(eml.Panic, match=r"Mismatched Matrices, left is 37x53, right is 37x53, \* is only defined for MxN \* NxL"):
        da * db


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(1, 1), (3, 2), (33, 65), (1000, 7), (5, 4097), (2048, 1536)])
def test_matrix_transpose(dtype, shape):
    r = np.random.default_rng(shape[0] * 31 + shape[1])
    a = r.standard_normal(shape).astype(dtype)
    t = eml.DeviceMatrix.from_host(eml.Matrix(a)).transpose()
    assert t.size() == (shape[1], shape[0])
    assert np.array_equal(t.to_host().data, a.T)


@pytest.mark.parametrize("dtype", DTYPES)
def test_tensor_reorder_and_transpose_all_permutations(dtype):
    names = ["a", "b", "c", "d"]
    lens = [3, 34, 5, 33]
    r = np.random.default_rng(8)
    x = r.standard_normal(lens).astype(dtype)
    t = eml.DeviceTensor.from_host(eml.Tensor(list(zip(names, lens)), x))
    kernels = set()
    for order in itertools.permutations(range(4)):
        want = np.ascontiguousarray(x.transpose(order))
        req = [names[d] for d in order]
        got = t.reorder(req)
        kernels.add(eml.last_kernel())
        assert got.shape() == [(names[d], lens[d]) for d in order]          # reorder: names move with the data
        assert np.array_equal(got.to_host().data, want)
        tr = t.transpose(req)
        assert tr.shape() == [(names[i], lens[d]) for i, d in enumerate(order)]  # transpose: names stay in place
        assert np.array_equal(tr.to_host().data, want)
        assert np.array_equal(eml.Tensor(list(zip(names, lens)), x).transpose(req).data, want)  # host mirror agrees
    assert kernels == {"reorder_tiled", "reorder_gather"}
    with pytest.raises(eml.Panic, match="must be the same set of dimension names in the tensor"):
        t.reorder(["a", "b", "c", "x"])


def test_reorder_entry_point_rank_limits_and_large_batch():
    x = torch.arange(2 * 70000 * 3, device="cuda", dtype=torch.float32).reshape(2, 70000, 3)
    out = torch.empty((3, 70000, 2), device="cuda", dtype=torch.float32)
    i64s = lambda *v: (C.c_int64 * len(v))(*v)
    # more than 65535 "rest" slices: the tiled kernel's grid-z loop
    check(lib.eml_reorder_f32_dev(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), 3, i64s(3, 70000, 2), i64s(1, 3, 210000), None))
    torch.cuda.synchronize()
    assert torch.equal(out, x.permute(2, 1, 0).contiguous())
    rc = lib.eml_reorder_f32_dev(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), 7, i64s(1, 1, 1, 1, 1, 1, 1), i64s(1, 1, 1, 1, 1, 1, 1), None)
    assert rc != 0 and "rank 7 outside 1..6" in lib.eml_last_error().decode()
    rc = lib.eml_reorder_f32_dev(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), 2, i64s(0, 4), i64s(1, 4), None)
    assert rc != 0 and "every length is >= 1" in lib.eml_last_error().decode()


@pytest.mark.parametrize("dtype", DTYPES)
def test_products_match_the_host_entry_point_bit_for_bit(dtype):
    r = np.random.default_rng(21)
    a = r.uniform(-1, 1, (300, 200)).astype(dtype)
    b = r.uniform(-1, 1, (200, 260)).astype(dtype)
    c = r.uniform(-1, 1, (260, 300)).astype(dtype)
    da, db, dc = (eml.DeviceMatrix.from_host(eml.Matrix(v)) for v in (a, b, c))
    chain = ((da * db) + dc.transpose()).transpose() * da          # four operators, nothing leaves the device
    host = eml.Matrix(np.ascontiguousarray(((eml.Matrix(a) * eml.Matrix(b)).data + c.T).T)) * eml.Matrix(a)
    assert np.array_equal(chain.to_host().data, host.data)
    # named rank-2 tensors: same product, same panics as Tensor.__mul__
    ta = eml.DeviceTensor.from_host(eml.Tensor([("r", 300), ("k", 200)], a))
    tb = eml.DeviceTensor.from_host(eml.Tensor([("k", 200), ("c", 260)], b))
    prod = ta * tb
    assert prod.shape() == [("r", 300), ("c", 260)]
    assert np.array_equal(prod.to_host().data, (eml.Matrix(a) * eml.Matrix(b)).data)
    with pytest.raises(eml.Panic, match="would create duplicate dimension names"):
        ta * eml.DeviceTensor.from_host(eml.Tensor([("k", 200), ("r", 260)], b))
    with pytest.raises(eml.Panic, match=r"\* is only defined for MxN \* NxL dimension lengths"):
        tb * ta.reorder(["k", "r"]).reorder(["r", "k"])
    with pytest.raises(eml.Panic, match="Dimensions of left and right tensors are not the same"):
        ta + tb


def test_householder_qr_with_products_chained_on_the_device():
    """qr_decomposition (src/linear_algebra.rs:1402-1543): `r = h * r; q = q * h` with r and q resident in HBM for
    the whole factorisation; same numbers as the host-operand version of tests/test_gpu_callers.py."""
    from test_gpu_callers import householder

    rows = cols = 96
    a = np.random.default_rng(5).uniform(0, 1, (rows, cols))
    r_dev, q_dev = eml.DeviceMatrix.from_host(eml.Matrix(a)), None
    r_host, q_host = eml.Matrix(a), None
    for c in range(cols - 1):
        column = r_host.data[c:, c].copy()
        h = np.eye(rows)
        h[c:, c:] = householder(column, np.float64)
        h_dev, h_host = eml.DeviceMatrix.from_host(eml.Matrix(h)), eml.Matrix(h)
        r_dev, r_host = h_dev * r_dev, h_host * r_host
        q_dev = h_dev if q_dev is None else q_dev * h_dev
        q_host = h_host if q_host is None else q_host * h_host
    q, r = q_dev.to_host(), r_dev.to_host()
    assert np.array_equal(q.data, q_host.data) and np.array_equal(r.data, r_host.data)
    assert np.all(np.abs((q_dev.transpose() * q_dev).to_host().data - np.eye(rows)) < 1e-5)
    assert np.all(np.abs((q_dev * r_dev).to_host().data - a) < 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_covariance_on_device_equals_host_entry_point(dtype):
    x = np.random.default_rng(9).uniform(-1, 1, (500, 40)).astype(dtype)
    d = eml.DeviceMatrix.from_host(eml.Matrix(x))
    assert np.array_equal(d.covariance_column_features().to_host().data, eml.Matrix(x).covariance_column_features().data)
    assert np.array_equal(d.covariance_row_features().to_host().data, eml.Matrix(x).covariance_row_features().data)
    t = eml.DeviceTensor.from_host(eml.Tensor([("samples", 500), ("features", 40)], x))
    cov = t.covariance("features")
    assert cov.shape() == [("i", 40), ("j", 40)]
    assert np.array_equal(cov.to_host().data, eml.Matrix(x).covariance_column_features().data)


def test_operator_chain_makes_no_allocator_call_in_steady_state():
    """Intermediates come from the library's stream-ordered block cache: repeating a chain reuses the same blocks."""
    a = eml.DeviceMatrix.from_host(eml.Matrix(np.random.default_rng(1).uniform(-1, 1, (256, 256)).astype(np.float32)))
    seen = []
    for _ in range(4):
        t = (a * a + a).transpose()
        seen.append(t._a.ptr.value)
        del t
    assert len(set(seen[1:])) == 1
