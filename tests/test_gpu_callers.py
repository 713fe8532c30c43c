"""Callers of `*` that reach the GPU path unchanged (SURVEY.md 8c fixtures downstream of the Mul impls, 8f row 3).

The reference's Householder QR (qr_decomposition, src/linear_algebra.rs:1402-1543) builds Q and R out of matrix
products `h * r`, `q * h` and the outer product `v_column * v_row` (:1378).  Here the same algorithm runs with
every product going through eml.Matrix.__mul__ (the drop-in GEMM), and the reference's own fuzz assertions
(tests/linear_algebra.rs:579-657: Q^T Q = I, R upper triangular, QR = A, all within 1e-5) are applied, at the
reference's tiny sizes and at sizes where the products take the tensor-core kernels.
"""
import numpy as np
import pytest

import easy_ml_b200 as eml

pytestmark = pytest.mark.gpu


def householder(x, dtype):
    """householder_matrix, src/linear_algebra.rs:1343-1380: I - 2 v v^T with the outer product on the GPU path."""
    length = np.sqrt(np.sum(x * x))
    a = length if x[0] > 0 else -length
    u = x.copy()
    u[0] = u[0] + a
    v = u / np.sqrt(np.sum(u * u))
    rows = x.shape[0]
    outer = eml.Matrix(v.reshape(rows, 1).astype(dtype)) * eml.Matrix(v.reshape(1, rows).astype(dtype))  # :1378
    return np.eye(rows, dtype=dtype) - outer.data * dtype(2)


def qr_decomposition(a):
    """qr_decomposition, src/linear_algebra.rs:1402-1543 (R = H_n ... H_1 A, Q = H_1 ... H_n)."""
    dtype = a.dtype.type
    rows, cols = a.shape
    r = eml.Matrix(a)
    q = None
    iterations = cols - 1 if rows == cols else cols
    for c in range(iterations):
        h_small = householder(r.data[c:, c].copy(), dtype)
        h = np.eye(rows, dtype=dtype)
        h[c:, c:] = h_small
        h = eml.Matrix(h)
        r = h * r                      # :1531
        q = h if q is None else q * h  # :1535
    return q, r


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-5), (np.float32, 1e-3)])
@pytest.mark.parametrize("shape", [(2, 2), (3, 2), (3, 3), (4, 4), (5, 3), (5, 5), (6, 6), (96, 64), (160, 160)])
def test_householder_qr_through_the_drop_in_products(dtype, tol, shape):
    r_ = np.random.default_rng(55 + shape[0] * 7 + shape[1])
    a = r_.uniform(0.0, 1.0, shape).astype(dtype)
    q, r = qr_decomposition(a)
    identity = (q.transpose() * q).data
    assert np.all(np.abs(identity - np.eye(shape[0], dtype=dtype)) < tol)       # Q is orthogonal
    assert np.all(np.abs(np.tril(r.data, -1)) < tol)                            # R is upper triangular
    assert np.all(np.abs((q * r).data - a) < tol)                               # QR = A
