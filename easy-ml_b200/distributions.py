"""Host mirror of the GEMM callers in easy_ml::distributions and linear_algebra.

The reference draws M samples of an N-dimensional Gaussian with M separate matrix-vector products
`mean + L * z_m` (draw_tensor_samples, src/distributions.rs:484-540).  Stacking the standard normal vectors as
the rows of Z turns the loop into ONE product  Z[M x N] * L^T[N x N]  on the drop-in GEMM: element (m, n) is
sum_k z[m,k] L[n,k], the same scalar product in the same k order as row n of L times z_m.  Everything else
(Cholesky factor of the N x N covariance, Box-Muller transform of the caller's uniform numbers, validation,
the way the random source is consumed) is the reference's host-side logic restated here; only the product runs
on the GPU, and there is no CPU fallback for it.
"""
import ctypes
import itertools

import numpy as np

from ._ffi import check, lib
from .host import Matrix, Panic, Tensor, _ptr, _sfx


def cholesky_decomposition(matrix):
    """linear_algebra::cholesky_decomposition, src/linear_algebra.rs:1008-1101 (Cholesky-Banachiewicz, row by
    row; `None` when the input is not square or not positive definite).  Host side: the factorisation is
    O(N^3 / 3) sequential work on the small covariance matrix and stays where the reference has it."""
    a = matrix.data if isinstance(matrix, Matrix) else np.asarray(matrix)
    n = a.shape[0]
    if a.ndim != 2 or a.shape[1] != n:
        return None
    dt = a.dtype.type
    low = np.zeros((n, n), dtype=a.dtype)
    for i in range(n):
        for j in range(i + 1):
            # sum_{k<j} L[i,k] L[j,k], added left to right starting from zero (:1075-1083)
            s = np.cumsum(low[i, :j] * low[j, :j], dtype=a.dtype)[-1] if j else dt(0)
            if i == j:
                entry_squared = a[i, j] - s
                if entry_squared <= 0:  # :1087: not positive definite
                    return None
                low[i, j] = np.sqrt(entry_squared)
            else:
                low[i, j] = (a[i, j] - s) * (dt(1) / low[j, j])  # :1093-1096: reciprocal, then multiply (two roundings)
    return Matrix(low) if isinstance(matrix, Matrix) else low


class _Source:
    """The caller's iterator of uniform numbers in [0, 1]; `take(n)` is n calls of `next()?`."""

    def __init__(self, source, dtype):
        self.dtype = dtype
        self.array = np.asarray(source, dtype=dtype).ravel() if isinstance(source, (np.ndarray, list, tuple)) else None
        self.pos = 0
        self.iterator = None if self.array is not None else iter(source)

    def take(self, n):
        if self.array is not None:
            got = self.array[self.pos:self.pos + n]
            self.pos += len(got)
        else:
            got = np.fromiter(itertools.islice(self.iterator, n), dtype=self.dtype)
        return got if len(got) == n else None


class Gaussian:
    """easy_ml::distributions::Gaussian, src/distributions.rs:131-250."""

    def __init__(self, mean, variance, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.mean = self.dtype.type(mean)
        self.variance = self.dtype.type(variance)

    def _draw_rows(self, source, rows, per_row):
        """`rows` consecutive calls of draw(source, per_row) (:205-237): each consumes ceil(per_row / 2) pairs and
        drops the surplus sample of its last pair.  Returns rows x per_row, or None when the source runs out."""
        t = self.dtype.type
        pairs = (per_row + 1) // 2
        numbers = source.take(rows * pairs * 2)
        if numbers is None:
            return None
        numbers = numbers.reshape(rows, pairs, 2)
        u, v = numbers[:, :, 0], numbers[:, :, 1]
        two = t(1) + t(1)
        minus_two = -two
        two_pi = two * t(np.pi)
        standard_deviation = np.sqrt(self.variance)
        radius = np.sqrt(minus_two * np.log(u))
        z = np.empty((rows, pairs, 2), dtype=self.dtype)
        z[:, :, 0] = (radius * np.cos(two_pi * v)) * standard_deviation + self.mean   # :222, :225
        z[:, :, 1] = (radius * np.sin(two_pi * v)) * standard_deviation + self.mean   # :223, :226
        return z.reshape(rows, pairs * 2)[:, :per_row]

    def draw(self, source, max_samples):
        source = source if isinstance(source, _Source) else _Source(source, self.dtype)
        drawn = self._draw_rows(source, 1, max_samples)
        return None if drawn is None else drawn[0].copy()


def _draw_samples(mean, covariance, source, max_samples):
    """draw_tensor_samples, src/distributions.rs:484-540, with the per-sample products batched into one GEMM."""
    dtype = covariance.dtype
    lower = cholesky_decomposition(covariance)
    if lower is None:
        return None
    features = mean.shape[0]
    z = Gaussian(0, 1, dtype)._draw_rows(_Source(source, dtype), max_samples, features)
    if z is None:
        return None
    z = np.ascontiguousarray(z)
    lower = np.ascontiguousarray(lower)
    product = np.empty((max_samples, features), dtype=dtype)
    s3 = lambda *s: (ctypes.c_int64 * 3)(*s)
    # product[m, n] = sum_k z[m, k] * lower[n, k]:  B[k][n] is element (n, k) of the row-major factor
    check(getattr(lib, f"eml_contract_{_sfx(dtype)}")(_ptr(z), _ptr(lower), _ptr(product), 1, max_samples, features,
                                                     features, s3(0, features, 1), s3(0, 1, features),
                                                     s3(0, features, 1)))
    return mean[None, :] + product  # `&column_vector_mean + (&lower_triangular * standard_normals)` (:528)


class MultivariateGaussian:
    """easy_ml::distributions::MultivariateGaussian, src/distributions.rs:262-369."""

    def __init__(self, mean, covariance):
        if mean.columns() != 1:
            raise Panic("Mean must be a column vector")
        if covariance.rows() != covariance.columns():
            raise Panic("Supplied 'covariance' matrix is not square")
        if mean.rows() != covariance.rows():
            raise Panic("Means must be same length as covariance matrix")
        if mean.data.dtype != covariance.data.dtype:
            raise TypeError("mean and covariance must have the same element type")
        self._mean, self._covariance = mean, covariance

    def mean(self):
        return self._mean

    def covariance(self):
        return self._covariance

    def draw(self, source, max_samples):
        """:346-368.  max_samples x N Matrix, or None (source exhausted / covariance not positive definite)."""
        drawn = _draw_samples(self._mean.data[:, 0], self._covariance.data, source, max_samples)
        return None if drawn is None else Matrix(drawn)


class MultivariateGaussianError(Exception):
    """MultivariateGaussianError, src/distributions.rs:382-416."""


class MultivariateGaussianTensor:
    """easy_ml::distributions::MultivariateGaussianTensor, src/distributions.rs:376-600."""

    def __init__(self, mean, covariance):
        shape = covariance.shape()
        if len(shape) != 2 or shape[0][1] != shape[1][1]:
            raise MultivariateGaussianError(f"Covariance matrix is not square: {covariance!r}")
        if len(mean.shape()) != 1 or mean.shape()[0][1] != shape[0][1]:
            raise MultivariateGaussianError(
                f"Mean vector has a different length {mean.shape()!r} to the covariance matrix size: {shape!r}")
        if mean.data.dtype != covariance.data.dtype:
            raise TypeError("mean and covariance must have the same element type")
        self._mean, self._covariance = mean, covariance

    def mean(self):
        return self._mean

    def covariance(self):
        return self._covariance

    def draw(self, source, max_samples, samples, features):
        """:560-581; `None` as well when the two dimension names are equal (:497-499)."""
        if samples == features:
            return None
        drawn = _draw_samples(self._mean.data, self._covariance.data, source, max_samples)
        if drawn is None:
            return None
        return Tensor([(samples, drawn.shape[0]), (features, drawn.shape[1])], drawn)
