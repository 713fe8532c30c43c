"""Device-resident Matrix / Tensor: the numbers stay in HBM between operations.

The host entry points (`Matrix.__mul__`, `Tensor.__mul__`) move their operands over PCIe for every product; a chain
such as the Householder QR's `r = h * r; q = q * h` (src/linear_algebra.rs:1531,1535) or `(a * b + c).transpose()`
pays that for every intermediate.  `DeviceMatrix` / `DeviceTensor` mirror the same operators with the same panics, but
own a device buffer: `*` is the device contraction (a transposed operand is a stride swap, not a copy), `+ - neg` and
the scalar forms are the HBM-bound elementwise kernels (src/matrices/operations.rs:41-163,1347-1552,
src/tensors/operations.rs:329-421,1811-1955), `transpose` / `reorder` the tiled reorder kernel
(src/tensors/mod.rs:1473-1600, src/matrices/mod.rs:1277-1330), `covariance*` the device covariance.  Elementwise
results are the reference's: one IEEE operation per element, so they are bit-exact.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from ._ffi import OPS, check, lib
from .host import Matrix, Panic, Tensor, _sfx, _shape_debug


class _Buffer:
    """Stream-ordered device memory (eml_alloc_async on the default stream, where every operator of this module is
    enqueued), handed back to the library's block cache with the last container that views it: a chain of operators
    makes no CUDA allocator call in steady state and never synchronises."""

    def __init__(self, nbytes):
        self.nbytes = max(int(nbytes), 1)
        self.ptr = C.c_void_p()
        check(lib.eml_alloc_async(C.byref(self.ptr), self.nbytes, None))

    def __del__(self):
        if getattr(self, "ptr", None) and self.ptr.value and lib is not None:
            lib.eml_free_async(self.ptr, self.nbytes, None)
            self.ptr = C.c_void_p()


def _i64s(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])


class _DeviceArray:
    """Contiguous row-major device array of `lens`, dtype f32 / f64."""

    def __init__(self, lens, dtype, buffer=None):
        self.lens = tuple(int(l) for l in lens)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.lens))
        self.buffer = buffer if buffer is not None else _Buffer(self.size * self.dtype.itemsize)

    @property
    def ptr(self):
        return self.buffer.ptr

    @classmethod
    def upload(cls, array):
        a = np.ascontiguousarray(array)
        out = cls(a.shape, a.dtype)
        check(lib.eml_copy_async(out.ptr, C.c_void_p(a.ctypes.data), a.nbytes, None))  # ordered on the same stream
        check(lib.eml_sync())  # `a` may be a temporary: it must outlive the copy
        return out

    def download(self):
        out = np.empty(self.lens, dtype=self.dtype)
        check(lib.eml_copy_async(C.c_void_p(out.ctypes.data), self.ptr, out.nbytes, None))
        check(lib.eml_sync())
        return out

    def _fn(self, name):
        return getattr(lib, f"eml_{name}_{_sfx(self.dtype)}_dev")

    def unary(self, op, c=0.0):
        out = _DeviceArray(self.lens, self.dtype)
        ct = C.c_float if self.dtype == np.float32 else C.c_double
        check(self._fn("unary")(OPS[op], ct(c), self.ptr, out.ptr, None, self.size, None))
        return out

    def binary(self, op, other):
        out = _DeviceArray(self.lens, self.dtype)
        check(self._fn("binary")(OPS[op], self.ptr, other.ptr, out.ptr, None, None, self.size, None))
        return out

    def reorder(self, order):
        """Array whose dimension d is this array's dimension order[d] (data moved, contiguous)."""
        strides = [int(np.prod(self.lens[d + 1:])) for d in range(len(self.lens))]
        out = _DeviceArray([self.lens[d] for d in order], self.dtype)
        check(self._fn("reorder")(self.ptr, out.ptr, len(order), _i64s(out.lens), _i64s([strides[d] for d in order]), None))
        return out

    def matmul(self, other):
        m, k = self.lens
        n = other.lens[1]
        out = _DeviceArray((m, n), self.dtype)
        check(self._fn("contract")(self.ptr, other.ptr, out.ptr, 1, m, n, k, _i64s((0, k, 1)), _i64s((0, n, 1)),
                                   _i64s((0, n, 1)), 0, None))
        return out

    def covariance(self, axis):
        f = self.lens[axis]
        out = _DeviceArray((f, f), self.dtype)
        check(self._fn("covariance")(self.ptr, self.lens[0], self.lens[1], axis, out.ptr, None))
        return out


def _scalar_ops(cls):
    """`+ c`, `- c`, `* c`, `/ c` (container on the left), as in src/matrices/operations.rs:1347-1552 and
    src/tensors/operations.rs:1811-1955."""
    for name, op in (("__add__", "ADD"), ("__sub__", "SUB"), ("__mul__", "MUL"), ("__truediv__", "DIV")):
        elementwise = getattr(cls, name, None)

        def method(self, other, _op=op, _elementwise=elementwise, _name=name):
            if isinstance(other, (int, float, np.floating, np.integer)):
                return self._like(self._a.unary(_op + "_C", float(other)))
            if _elementwise is None:
                return NotImplemented
            return _elementwise(self, other)

        setattr(cls, name, method)
    return cls


@_scalar_ops
class DeviceMatrix:
    """easy_ml::matrices::Matrix<T> with its numbers in HBM."""

    def __init__(self, array):
        self._a = array

    @staticmethod
    def from_host(matrix):
        return DeviceMatrix(_DeviceArray.upload(matrix.data if isinstance(matrix, Matrix) else Matrix(matrix).data))

    def to_host(self):
        return Matrix(self._a.download())

    def _like(self, array):
        return DeviceMatrix(array)

    def rows(self):
        return self._a.lens[0]

    def columns(self):
        return self._a.lens[1]

    def size(self):
        return self._a.lens

    def _same(self, other, symbol):
        if not isinstance(other, DeviceMatrix):
            raise TypeError("both operands must be DeviceMatrix")
        if self._a.dtype != other._a.dtype:
            raise TypeError("both matrices must have the same element type")
        if self.size() != other.size():  # src/matrices/operations.rs:56-63, :91-98
            raise Panic(f"Mismatched matrices, left is {self.rows()}x{self.columns()}, right is "
                        f"{other.rows()}x{other.columns()}, {symbol} is only defined for MxN {symbol} MxN")

    def __add__(self, other):
        self._same(other, "+")
        return DeviceMatrix(self._a.binary("ADD", other._a))

    def __sub__(self, other):
        self._same(other, "-")
        return DeviceMatrix(self._a.binary("SUB", other._a))

    def __neg__(self):
        return DeviceMatrix(self._a.unary("NEG"))

    def __mul__(self, other):
        # matrix_view_multiplication, src/matrices/operations.rs:165-196
        if not isinstance(other, DeviceMatrix):
            return NotImplemented
        if self._a.dtype != other._a.dtype:
            raise TypeError("both matrices must have the same element type")
        if self.columns() != other.rows():
            raise Panic(f"Mismatched Matrices, left is {self.rows()}x{self.columns()}, right is "
                        f"{other.rows()}x{other.columns()}, * is only defined for MxN * NxL")
        return DeviceMatrix(self._a.matmul(other._a))

    def mul_sliced(self, other, slices=0):
        """The same product as `*` through error-free int8 slices on the INT8 tensor cores (eml_gemm_f64_sliced_dev,
        f64 only): about twice the speed of the FP64 pipe for large matrices.  A guard proves the path's 1e-12
        contract for these operands first; when it cannot, the FP64 kernel runs.  `slices` = 0 uses as few slices as the
        guard can prove (7 or 8).  Returns (product, number of slices used or 0 for the FP64 kernel)."""
        if not isinstance(other, DeviceMatrix):
            raise TypeError("both operands must be DeviceMatrix")
        if self._a.dtype != np.float64 or other._a.dtype != np.float64:
            raise TypeError("mul_sliced is defined for f64 matrices")
        if self.columns() != other.rows():
            raise Panic(f"Mismatched Matrices, left is {self.rows()}x{self.columns()}, right is "
                        f"{other.rows()}x{other.columns()}, * is only defined for MxN * NxL")
        m, k, n = self.rows(), self.columns(), other.columns()
        out = _DeviceArray((m, n), np.float64)
        path = C.c_int(0)
        check(lib.eml_gemm_f64_sliced_dev(self._a.ptr, other._a.ptr, out.ptr, m, n, k, k, n, n, int(slices),
                                          C.byref(path), None))
        return DeviceMatrix(out), int(path.value)

    def transpose(self):
        """Matrix::transpose, src/matrices/mod.rs:1277-1300."""
        return DeviceMatrix(self._a.reorder((1, 0)))

    def covariance_column_features(self):
        return DeviceMatrix(self._a.covariance(1))

    def covariance_row_features(self):
        return DeviceMatrix(self._a.covariance(0))


@_scalar_ops
class DeviceTensor:
    """easy_ml::tensors::Tensor<T, D> with its numbers in HBM."""

    def __init__(self, shape, array):
        self._shape = [(str(n), int(l)) for n, l in shape]
        self._a = array

    @staticmethod
    def from_host(tensor):
        return DeviceTensor(tensor.shape(), _DeviceArray.upload(tensor.data))

    def to_host(self):
        return Tensor(self._shape, self._a.download())

    def _like(self, array):
        return DeviceTensor(self._shape, array)

    def shape(self):
        return list(self._shape)

    def names(self):
        return [n for n, _ in self._shape]

    def _same(self, other):
        if not isinstance(other, DeviceTensor):
            raise TypeError("both operands must be DeviceTensor")
        if self._a.dtype != other._a.dtype:
            raise TypeError("both tensors must have the same element type")
        if self._shape != other._shape:  # assert_same_dimensions, src/tensors/operations.rs:309-320
            raise Panic(f"Dimensions of left and right tensors are not the same: (left: {_shape_debug(self._shape)}, "
                        f"right: {_shape_debug(other._shape)})")

    def __add__(self, other):
        self._same(other)
        return self._like(self._a.binary("ADD", other._a))

    def __sub__(self, other):
        self._same(other)
        return self._like(self._a.binary("SUB", other._a))

    def __neg__(self):
        return self._like(self._a.unary("NEG"))

    def __mul__(self, other):
        # tensor_view_matrix_product, src/tensors/operations.rs:474-519
        if not isinstance(other, DeviceTensor):
            return NotImplemented
        if len(self._shape) != 2 or len(other._shape) != 2:
            raise TypeError("* is implemented for rank 2 tensors only")
        if self._a.dtype != other._a.dtype:
            raise TypeError("both tensors must have the same element type")
        ls, rs = self._shape, other._shape
        if ls[1][1] != rs[0][1]:
            raise Panic(f"Mismatched tensors, left is {_shape_debug(ls)}, right is {_shape_debug(rs)}, "
                        "* is only defined for MxN * NxL dimension lengths")
        if ls[0][0] == rs[1][0]:
            raise Panic(f"Matrix multiplication of tensors with shapes left {_shape_debug(ls)} and right "
                        f"{_shape_debug(rs)} would create duplicate dimension names as the shape "
                        f"{_shape_debug([ls[0], rs[1]])}. Rename one or both of the dimension names in the input "
                        "to prevent this. * is defined as MxN * NxL = MxL")
        return DeviceTensor([ls[0], rs[1]], self._a.matmul(other._a))

    def _order(self, names):
        names = list(names)
        if sorted(names) != sorted(self.names()):  # :1541-1547
            raise Panic(f"Dimension names provided {names!r} must be the same set of dimension names in the tensor: "
                        f"{_shape_debug(self._shape)}")
        return [self.names().index(n) for n in names]

    def reorder(self, names):
        """Tensor::reorder, src/tensors/mod.rs:1540-1553: data AND names move into the requested order."""
        order = self._order(names)
        return DeviceTensor([self._shape[d] for d in order], self._a.reorder(order))

    def transpose(self, names):
        """Tensor::transpose, :1473-1486: the data moves, the dimension names stay where they were."""
        moved = self.reorder(names)
        return DeviceTensor([(n, l) for (n, _), (_, l) in zip(self._shape, moved._shape)], moved._a)

    def covariance(self, feature_dimension):
        """linear_algebra::covariance, src/linear_algebra.rs:778-832."""
        if len(self._shape) != 2:
            raise TypeError("covariance is defined for rank 2 tensors")
        if feature_dimension not in self.names():
            raise Panic(f'Feature dimension "{feature_dimension}" is not present in the input tensor\'s shape: '
                        f"{_shape_debug(self._shape)}")
        out = self._a.covariance(self.names().index(feature_dimension))
        return DeviceTensor([("i", out.lens[0]), ("j", out.lens[1])], out)
