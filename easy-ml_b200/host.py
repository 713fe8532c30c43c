"""Host-side mirror of Easy ML's operator API for the dense-contraction path, over the C ABI.

This is the Python face used by the parity tests and the benchmark; the C++ mirror (host/easy_ml.hpp)
and the Rust shim follow the same rules.  Names, argument meaning and error behaviour
follow the reference (paths relative to the reference repo):

  Matrix * Matrix              src/matrices/operations.rs:165-196   -> eml_gemm_{f32,f64}
  Tensor<_,2> * Tensor<_,2>    src/tensors/operations.rs:474-519    -> eml_gemm_{f32,f64}
  Tensor<_,1>.scalar_product   src/tensors/mod.rs:1773              -> eml_dot_{f32,f64}
  Einsum::with_2(..).to(..)    src/tensors/einsum.rs:908-980        -> eml_contract_* / eml_einsum2_*
  RecordTensor / RecordMatrix  src/differentiation/container_record -> eml_tape_*

Every validation the reference performs happens here, BEFORE the FFI call, with the reference's panic
text (raised as `Panic`) or `Err` value (raised as InconsistentDimensionLengthError).  Numbers live in
numpy arrays of dtype float32 / float64 only -- other element types are not this path's business.
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import OPS, check, lib


class Panic(Exception):
    """The reference `panic!`s here; str(e) is the reference's message."""


class InconsistentDimensionLengthError(Exception):
    """Einsum `Err(InconsistentDimensionLengthError { lengths, dimension })`, src/tensors/einsum.rs:321."""

    def __init__(self, dimension, lengths):
        super().__init__(f"InconsistentDimensionLengthError(dimension={dimension!r}, lengths={lengths})")
        self.dimension, self.lengths = dimension, lengths


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"the cuda path handles f32/f64 only (got {dtype}); other numeric types stay on the generic path")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_float(data, dtype):
    a = np.array(data, dtype=dtype if dtype is not None else None)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def _shape_debug(shape):
    return "[" + ", ".join(f'("{n}", {l})' for n, l in shape) + "]"


def _gemm(a, b):
    """C = A * B for 2-D arrays whose last axis is contiguous (leading dimensions allowed)."""
    sfx = _sfx(a.dtype)
    if a.strides[1] != a.itemsize or a.strides[0] % a.itemsize or a.strides[0] < a.shape[1] * a.itemsize:
        a = np.ascontiguousarray(a)
    if b.strides[1] != b.itemsize or b.strides[0] % b.itemsize or b.strides[0] < b.shape[1] * b.itemsize:
        b = np.ascontiguousarray(b)
    m, k = a.shape
    n = b.shape[1]
    c = np.empty((m, n), dtype=a.dtype)
    check(getattr(lib, f"eml_gemm_{sfx}")(_ptr(a), _ptr(b), _ptr(c), m, n, k, a.strides[0] // a.itemsize,
                                          b.strides[0] // b.itemsize, n))
    return c


# ------------------------------------------------------------------------------------------------
class Matrix:
    """easy_ml::matrices::Matrix<T>, T in {f32, f64}: row-major, at least 1x1."""

    def __init__(self, rows, dtype=None):
        a = _as_float(rows, dtype)
        if a.ndim != 2 or a.size == 0:
            raise Panic("Matrices must be rectangular and have at least one element")
        self.data = a

    @staticmethod
    def from_flat_row_major(size, values, dtype=None):
        a = _as_float(values, dtype)
        if a.size != size[0] * size[1]:
            raise Panic("Inconsistent size")
        return Matrix(a.reshape(size))

    def rows(self):
        return self.data.shape[0]

    def columns(self):
        return self.data.shape[1]

    def size(self):
        return self.data.shape

    def transpose(self):
        return Matrix(np.ascontiguousarray(self.data.T))

    def get(self, row, column):
        return self.data[row, column]

    def __eq__(self, other):
        return isinstance(other, Matrix) and self.data.shape == other.data.shape and np.array_equal(self.data, other.data)

    def __mul__(self, other):
        # matrix_view_multiplication, src/matrices/operations.rs:165-196
        if not isinstance(other, Matrix):
            return NotImplemented
        if self.data.dtype != other.data.dtype:
            raise TypeError("both matrices must have the same element type")
        if self.columns() != other.rows():  # :176-183
            raise Panic(f"Mismatched Matrices, left is {self.rows()}x{self.columns()}, right is "
                        f"{other.rows()}x{other.columns()}, * is only defined for MxN * NxL")
        return Matrix(_gemm(self.data, other.data))

    def _covariance(self, axis):
        a = np.ascontiguousarray(self.data)
        f = a.shape[1] if axis == 1 else a.shape[0]
        out = np.empty((f, f), dtype=a.dtype)
        check(getattr(lib, f"eml_covariance_{_sfx(a.dtype)}")(_ptr(a), a.shape[0], a.shape[1], a.shape[1], axis, _ptr(out)))
        return Matrix(out)

    def covariance_column_features(self):
        """linear_algebra::covariance_column_features, src/linear_algebra.rs:615-639 (Matrix method mod.rs:1874)."""
        return self._covariance(1)

    def covariance_row_features(self):
        """linear_algebra::covariance_row_features, src/linear_algebra.rs:676-700."""
        return self._covariance(0)

    def __repr__(self):
        return f"Matrix({self.data.tolist()})"


# ------------------------------------------------------------------------------------------------
class Tensor:
    """easy_ml::tensors::Tensor<T, D>: named dimensions, row-major, every length >= 1."""

    def __init__(self, shape, data, dtype=None):
        shape = [(str(n), int(l)) for n, l in shape]
        names = [n for n, _ in shape]
        if len(set(names)) != len(names):  # src/tensors/mod.rs:96-123
            raise Panic(f"Dimension names must all be unique: {_shape_debug(shape)}")
        if any(l < 1 for _, l in shape):
            raise Panic(f"No dimension can have 0 elements: {_shape_debug(shape)}")
        a = _as_float(data, dtype)
        total = int(np.prod([l for _, l in shape])) if shape else 1
        if a.size != total:
            raise Panic(f"Length of dimensions must match size of data: {_shape_debug(shape)}, data: {a.size}")
        self._shape = shape
        self.data = a.reshape([l for _, l in shape])

    @staticmethod
    def from_fn(shape, f, dtype=np.float32):
        lens = [l for _, l in shape]
        a = np.empty(lens, dtype=dtype)
        for idx in np.ndindex(*lens):
            a[idx] = f(list(idx))
        return Tensor(shape, a)

    def shape(self):
        return list(self._shape)

    def names(self):
        return [n for n, _ in self._shape]

    def rename(self, names):
        return Tensor(list(zip(names, self.data.shape)), self.data)

    def transpose(self, names):
        """Tensor::transpose, src/tensors/mod.rs:1473-: moves the DATA into the order of `names` but keeps
        the dimension names of this tensor's shape in place (Easy ML does not swap names)."""
        order = [self.names().index(n) for n in names]
        moved = np.ascontiguousarray(self.data.transpose(order))
        return Tensor(list(zip(self.names(), moved.shape)), moved)

    def select(self, name, index):
        ax = self.names().index(name)
        return Tensor([s for i, s in enumerate(self._shape) if i != ax], np.take(self.data, index, axis=ax))

    def first(self):
        return self.data.reshape(-1)[0]

    def __eq__(self, other):
        return isinstance(other, Tensor) and self._shape == other._shape and np.array_equal(self.data, other.data)

    def __mul__(self, other):
        # tensor_view_matrix_product, src/tensors/operations.rs:474-519
        if not isinstance(other, Tensor):
            return NotImplemented
        if len(self._shape) != 2 or len(other._shape) != 2:
            raise TypeError("* is implemented for rank 2 tensors only")
        if self.data.dtype != other.data.dtype:
            raise TypeError("both tensors must have the same element type")
        ls, rs = self._shape, other._shape
        if ls[1][1] != rs[0][1]:  # :485-491
            raise Panic(f"Mismatched tensors, left is {_shape_debug(ls)}, right is {_shape_debug(rs)}, "
                        "* is only defined for MxN * NxL dimension lengths")
        if ls[0][0] == rs[1][0]:  # :492-501
            raise Panic(f"Matrix multiplication of tensors with shapes left {_shape_debug(ls)} and right "
                        f"{_shape_debug(rs)} would create duplicate dimension names as the shape "
                        f"{_shape_debug([ls[0], rs[1]])}. Rename one or both of the dimension names in the input "
                        "to prevent this. * is defined as MxN * NxL = MxL")
        return Tensor([ls[0], rs[1]], _gemm(self.data, other.data))

    def covariance(self, feature_dimension):
        """linear_algebra::covariance, src/linear_algebra.rs:778-832 (rank 2; output dimensions "i", "j")."""
        if len(self._shape) != 2:
            raise TypeError("covariance is defined for rank 2 tensors")
        names = self.names()
        if feature_dimension not in names:  # :797-802
            raise Panic(f'Feature dimension "{feature_dimension}" is not present in the input tensor\'s shape: '
                        f"{_shape_debug(self._shape)}")
        axis = names.index(feature_dimension)
        a = np.ascontiguousarray(self.data)
        f = a.shape[axis]
        out = np.empty((f, f), dtype=a.dtype)
        check(getattr(lib, f"eml_covariance_{_sfx(a.dtype)}")(_ptr(a), a.shape[0], a.shape[1], a.shape[1], axis, _ptr(out)))
        return Tensor([("i", f), ("j", f)], out)

    def scalar_product(self, other):
        # Tensor<T,1>::scalar_product, src/tensors/mod.rs:1773 -> operations.rs:1773-1808
        if len(self._shape) != 1 or len(other._shape) != 1:
            raise TypeError("scalar_product is defined for rank 1 tensors")
        if self._shape != other._shape:
            raise Panic(f"Dimensions of left and right tensors are not the same: (left: {_shape_debug(self._shape)}, "
                        f"right: {_shape_debug(other._shape)})")
        sfx = _sfx(self.data.dtype)
        out = np.zeros(1, dtype=self.data.dtype)
        check(getattr(lib, f"eml_dot_{sfx}")(_ptr(self.data), 1, _ptr(other.data), 1, self.data.size, _ptr(out)))
        return out[0]

    def __repr__(self):
        return f"Tensor({_shape_debug(self._shape)}, {self.data.tolist()})"


# ------------------------------------------------------------------------------------------------
def _length_of(dim, inputs):
    """src/tensors/einsum.rs:508-538"""
    lengths = []
    for shape in inputs:
        found = [l for n, l in shape if n == dim]
        lengths.append(found[0] if found else None)
    known = [l for l in lengths if l is not None]
    if not known or any(l != known[0] for l in known):
        raise InconsistentDimensionLengthError(dim, lengths)
    return known[0]


def _summation_dimensions(inputs, output):
    """src/tensors/einsum.rs:579-622"""
    unique = []
    for shape in inputs:
        for name, length in shape:
            if name in output:
                continue
            existing = [l for n, l in unique if n == name]
            if not existing:
                unique.append((name, length))
            elif existing[0] != length:
                lengths = []
                for s in inputs:
                    f = [l for n, l in s if n == name]
                    lengths.append(f[0] if f else None)
                raise InconsistentDimensionLengthError(name, lengths)
    return unique


def _strides(t):
    return {n: t.data.strides[i] // t.data.itemsize for i, (n, _) in enumerate(t._shape)}


def _merge(dims, lengths, stride_maps):
    """Collapse several named dimensions that play one GEMM role into a single (length, strides) pair.
    Possible only when, in every tensor that carries them, they nest contiguously in the given order."""
    dims = [d for d in dims if lengths[d] > 1] or dims[:1]
    if not dims:
        return 1, [0] * len(stride_maps)
    total = 1
    for d in dims:
        total *= lengths[d]
    out = []
    for sm in stride_maps:
        for a, b in zip(dims[:-1], dims[1:]):
            if sm[a] != sm[b] * lengths[b]:
                return None
        out.append(sm[dims[-1]])
    return total, out


class _Einsum2:
    def __init__(self, x, y):
        self.x, self.y = x, y

    def named(self, names_1, names_2):
        return _Einsum2(self.x.rename(names_1), self.y.rename(names_2))

    def to(self, output):
        """Einsum2::to, src/tensors/einsum.rs:908-980."""
        x, y = self.x, self.y
        if x.data.dtype != y.data.dtype:
            raise TypeError("both tensors must have the same element type")
        sfx = _sfx(x.data.dtype)
        output = list(output)
        inputs = [x._shape, y._shape]
        out_shape = [(d, _length_of(d, inputs)) for d in output]  # output_shape_for :563-572
        if len(set(output)) != len(output):
            raise Panic(f"Dimension names must all be unique: {_shape_debug(out_shape)}")  # Tensor::empty
        summation = _summation_dimensions(inputs, output)
        lengths = dict(out_shape)
        lengths.update(dict(summation))
        out = np.empty([l for _, l in out_shape], dtype=x.data.dtype)
        sx, sy = _strides(x), _strides(y)
        so = {n: out.strides[i] // out.itemsize for i, (n, _) in enumerate(out_shape)}
        in_x, in_y = set(x.names()), set(y.names())
        batch = [d for d in output if d in in_x and d in in_y]
        m_dims = [d for d in output if d in in_x and d not in in_y]
        n_dims = [d for d in output if d in in_y and d not in in_x]
        k_dims = [d for d, _ in summation if d in in_x and d in in_y]
        lone_sum = [d for d, _ in summation if not (d in in_x and d in in_y)]
        plan = None
        if k_dims and not lone_sum:
            zx = {d: sx.get(d, 0) for d in lengths}
            zy = {d: sy.get(d, 0) for d in lengths}
            b = _merge(batch, lengths, [zx, zy, so]) if batch else (1, [0, 0, 0])
            m = _merge(m_dims, lengths, [zx, so]) if m_dims else (1, [0, 0])
            n = _merge(n_dims, lengths, [zy, so]) if n_dims else (1, [0, 0])
            k = _merge(k_dims, lengths, [zx, zy])
            if None not in (b, m, n, k):
                plan = (b, m, n, k)
        if plan is not None:
            (B, (bx, by, bo)), (M, (mx, mo)), (N, (ny, no)), (K, (kx, ky)) = plan
            check(getattr(lib, f"eml_contract_{sfx}")(
                _ptr(x.data), _ptr(y.data), _ptr(out), B, M, N, K, _ffi.arr3((bx, mx, kx)), _ffi.arr3((by, ky, ny)),
                _ffi.arr3((bo, mo, no))))
        else:
            # not a (batched) GEMM: the strided loop nest, bit-exact with the reference
            ol = [l for _, l in out_shape]
            ao = [sx.get(d, 0) for d in output]
            bo_ = [sy.get(d, 0) for d in output]
            sl = [l for _, l in summation]
            a_s = [sx.get(d, 0) for d, _ in summation]
            b_s = [sy.get(d, 0) for d, _ in summation]
            check(getattr(lib, f"eml_einsum2_{sfx}")(
                _ptr(x.data), x.data.size, _ptr(y.data), y.data.size, _ptr(out), len(ol), _ffi.arrn(ol),
                _ffi.arrn(ao), _ffi.arrn(bo_), len(sl), _ffi.arrn(sl), _ffi.arrn(a_s), _ffi.arrn(b_s)))
        return Tensor(out_shape, out)


class _EinsumN:
    """Einsum1 / Einsum3..6 (src/tensors/einsum.rs:829-883, :1019-1666): the strided loop nest on the device, in
    the reference's order (bit-exact).  Two operands go through _Einsum2, which recognises (batched) GEMMs."""

    def __init__(self, tensors):
        self.tensors = list(tensors)

    def named(self, *names):
        if len(names) != len(self.tensors):
            raise TypeError("one list of names per tensor")
        return _EinsumN([t.rename(n) for t, n in zip(self.tensors, names)])

    def to(self, output):
        ts = self.tensors
        dtype = ts[0].data.dtype
        if any(t.data.dtype != dtype for t in ts):
            raise TypeError("all tensors must have the same element type")
        output = list(output)
        inputs = [t._shape for t in ts]
        out_shape = [(d, _length_of(d, inputs)) for d in output]  # output_shape_for :563-572
        if len(set(output)) != len(output):
            raise Panic(f"Dimension names must all be unique: {_shape_debug(out_shape)}")
        summation = _summation_dimensions(inputs, output)  # :579-622
        out = np.empty([l for _, l in out_shape], dtype=dtype)
        arrays = [np.ascontiguousarray(t.data) for t in ts]
        strides = [{n: a.strides[i] // a.itemsize for i, (n, _) in enumerate(t._shape)} for t, a in zip(ts, arrays)]
        n = len(ts)
        n_out, n_sum = len(out_shape), len(summation)
        os_ = [st.get(d, 0) for st in strides for d, _ in out_shape]
        ss_ = [st.get(d, 0) for st in strides for d, _ in summation]
        xs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
        elems = _ffi.arrn([a.size for a in arrays])
        check(getattr(lib, f"eml_einsum_n_{_sfx(dtype)}")(
            n, xs, elems, _ptr(out), n_out, _ffi.arrn([l for _, l in out_shape]), _ffi.arrn(os_), n_sum,
            _ffi.arrn([l for _, l in summation]), _ffi.arrn(ss_)))
        return Tensor(out_shape, out)


class Einsum:
    """easy_ml::tensors::einsum::Einsum (src/tensors/einsum.rs:98-313): with_1 .. with_6."""

    @staticmethod
    def with_1(x):
        return _EinsumN([x])

    @staticmethod
    def with_2(x, y):
        return _Einsum2(x, y)

    @staticmethod
    def with_3(a, b, c):
        return _EinsumN([a, b, c])

    @staticmethod
    def with_4(a, b, c, d):
        return _EinsumN([a, b, c, d])

    @staticmethod
    def with_5(a, b, c, d, e):
        return _EinsumN([a, b, c, d, e])

    @staticmethod
    def with_6(a, b, c, d, e, f):
        return _EinsumN([a, b, c, d, e, f])


# ------------------------------------------------------------------------------------------------
class WengertList:
    """easy_ml::differentiation::WengertList<T>, src/differentiation.rs:327-331, as a device tape."""

    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        _sfx(self.dtype)
        h = C.c_void_p()
        check(lib.eml_tape_create(_ffi.EML_F32 if self.dtype == np.float32 else _ffi.EML_F64, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib.eml_tape_destroy(self._h)
            self._h = None

    def __len__(self):
        n = C.c_int64()
        check(lib.eml_tape_len(self._h, C.byref(n)))
        return n.value

    def clear(self):
        check(lib.eml_tape_clear(self._h))

    def sync(self):
        check(lib.eml_tape_sync(self._h))

    def record(self, step):
        """Records the device work `step()` issues and returns it as a replayable CUDA graph (RecordedStep).
        No counterpart in the reference.  The step must have run eagerly once before (so every buffer size is in the
        cache), may not block on the device (use numbers_async), must release the containers it creates and must
        leave the tape as it found it (clear() + reset()); sgd_update works in place while recording."""
        check(lib.eml_tape_capture_begin(self._h))
        g = C.c_void_p()
        try:
            step()
        except BaseException:
            lib.eml_tape_capture_end(self._h, C.byref(g))
            if g:
                lib.eml_graph_destroy(g)
            raise
        check(lib.eml_tape_capture_end(self._h, C.byref(g)))
        return RecordedStep(g)


class RecordedStep:
    def __init__(self, handle):
        self._g = handle

    def __del__(self):
        if getattr(self, "_g", None):
            lib.eml_graph_destroy(self._g)
            self._g = None

    def launch(self):
        """Enqueues one replay; returns an arrival to pass to wait_arrival before reading what the step copied out."""
        a = C.c_void_p()
        check(lib.eml_graph_launch(self._g, C.byref(a)))
        return a


def wait_arrival(arrival):
    if arrival:
        check(lib.eml_arrival_wait(arrival))


class PinnedBuffer:
    """Page-locked host memory (eml_host_alloc) viewed as a numpy array: the target of asynchronous read-backs."""

    def __init__(self, shape, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        self._p = C.c_void_p()
        check(lib.eml_host_alloc(C.byref(self._p), max(n, 1) * self.dtype.itemsize))
        ctype = C.c_float if self.dtype == np.float32 else C.c_double
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(ctype)), shape=(max(n, 1),))[:n].reshape(shape)

    def __del__(self):
        if getattr(self, "_p", None):
            self.array = None
            lib.eml_host_free(self._p)
            self._p = None


class Record:
    """easy_ml::differentiation::Record<T> (src/differentiation.rs:452-476): a number, an optional list, a tape index.
    The operators follow src/differentiation/record_operations.rs (history None on a side -> append_unary with the
    other side's rule, both Some -> append_binary) with the rules of src/differentiation/functions.rs; entries go to the
    same device-backed list the containers use, so a Record taken from a container element
    (RecordContainer::get_as_record) and folded with scalar arithmetic differentiates through both."""

    __slots__ = ("number", "history", "index")

    def __init__(self, number, history=None, index=0):
        self.number, self.history, self.index = number, history, index

    @staticmethod
    def variable(x, history):  # :489-495: append_nullary
        i = C.c_int64()
        check(lib.eml_tape_append_nullary(history._h, C.byref(i)))
        return Record(history.dtype.type(x), history, i.value)

    @staticmethod
    def constant(x, dtype=np.float64):  # :479-485
        return Record(np.dtype(dtype).type(x), None, 0)

    def reset(self):  # :512-520
        if self.history is not None:
            i = C.c_int64()
            check(lib.eml_tape_append_nullary(self.history._h, C.byref(i)))
            self.index = i.value

    def _unary(self, number, derivative):
        i = C.c_int64()
        check(lib.eml_tape_append_unary(self.history._h, self.index, float(derivative), C.byref(i)))
        return Record(number, self.history, i.value)

    @staticmethod
    def _binary(x, y, number, dx, dy):
        _same_list(x.history, y.history)
        i = C.c_int64()
        check(lib.eml_tape_append_binary(x.history._h, x.index, float(dx), y.index, float(dy), C.byref(i)))
        return Record(number, x.history, i.value)

    def _coerce(self, o):
        return o if isinstance(o, Record) else Record(type(self.number)(o), None, 0)

    def _apply(self, o, z, dx, dy):
        if self.history is None and o.history is None:
            return Record(z)
        if o.history is None:
            return self._unary(z, dx)
        if self.history is None:
            return o._unary(z, dy)
        return Record._binary(self, o, z, dx, dy)

    def __add__(self, o):
        o = self._coerce(o)
        z = self.number + o.number
        one = type(z)(1)
        return self._apply(o, z, one, one)

    def __sub__(self, o):
        o = self._coerce(o)
        z = self.number - o.number
        one = type(z)(1)
        return self._apply(o, z, one, -one)

    def __mul__(self, o):
        o = self._coerce(o)
        return self._apply(o, self.number * o.number, o.number, self.number)

    def __truediv__(self, o):
        o = self._coerce(o)
        z = self.number / o.number
        one = type(z)(1)
        return self._apply(o, z, one / o.number, -self.number / (o.number * o.number))

    def derivatives(self):  # :606-648
        if self.history is None:
            raise Panic("Record has no WengertList to find derivatives from")
        d = C.c_void_p()
        check(lib.eml_tape_derivatives_at(self.history._h, self.index, C.byref(d)))
        return Derivatives(self.history, d)


def _same_list(a, b):
    if a is not None and b is not None and a is not b:
        raise Panic("Record containers must be using the same WengertList")


class _RecordContainer:
    """RecordContainer<'a, T, S, D> (container_record/mod.rs:66-79): device numbers + tape indexes."""

    def __init__(self, owner, handle, shape, history):
        self._owner, self._v, self._shape, self.history_list = owner, handle, shape, history

    def __del__(self):
        try:
            if self._owner._h:
                lib.eml_val_release(self._owner._h, self._v)
        except Exception:
            pass

    # -- constructors --
    @classmethod
    def _create(cls, owner, history, shape, data, variables):
        a = np.ascontiguousarray(np.asarray(data, dtype=owner.dtype))
        rows = int(np.prod(a.shape[:-1])) if a.ndim > 1 else 1
        cols = a.shape[-1] if a.ndim >= 1 else 1
        v = C.c_int64()
        f = lib.eml_tape_variables if variables else lib.eml_tape_constants
        check(f(owner._h, _ptr(a), rows, cols, C.byref(v)))
        return cls(owner, v.value, shape, history)

    @classmethod
    def _from_existing(cls, owner, history, shape, numbers, indexes):
        """RecordContainer::from_existing (container_record/mod.rs:219-224, :508): any (number, Index) pairs."""
        a = np.ascontiguousarray(np.asarray(numbers, dtype=owner.dtype))
        idx = np.ascontiguousarray(np.asarray(indexes, dtype=np.int64))
        if a.shape != idx.shape:
            raise Panic("from_existing: numbers and indexes must have the same shape")
        rows = int(np.prod(a.shape[:-1])) if a.ndim > 1 else 1
        cols = a.shape[-1] if a.ndim >= 1 else 1
        v = C.c_int64()
        check(lib.eml_tape_from_existing(owner._h, _ptr(a), _ptr(idx), rows, cols, 1 if history is not None else 0, C.byref(v)))
        return cls(owner, v.value, shape, history)

    def _range(self, row0, n_rows, col0, n_cols, shape, history):
        """from_existing over a MatrixRange / TensorRange of this container: numbers copied on the device, indexes follow."""
        v = C.c_int64()
        check(lib.eml_tape_range(self._owner._h, self._v, row0, n_rows, col0, n_cols, 1 if history is not None else 0, C.byref(v)))
        return type(self)(self._owner, v.value, shape, history)

    def history(self):
        return self.history_list

    def elements(self):
        return int(np.prod(self._lens()))

    def _rc(self):
        r, c, t = C.c_int64(), C.c_int64(), C.c_int()
        check(lib.eml_val_shape(self._owner._h, self._v, C.byref(r), C.byref(c), C.byref(t)))
        return r.value, c.value

    def _numbers(self):
        r, c = self._rc()
        out = np.empty(r * c, dtype=self._owner.dtype)
        check(lib.eml_val_numbers(self._owner._h, self._v, _ptr(out)))
        return out.reshape(self._lens())

    def numbers_async(self, pinned):
        """Enqueues the copy of the numbers into a PinnedBuffer; returns an arrival (None inside a recorded step,
        where the arrival comes from RecordedStep.launch)."""
        a = C.c_void_p()
        check(lib.eml_val_numbers_async(self._owner._h, self._v, pinned._p, C.byref(a)))
        return a if a else None

    def indexes(self):
        r, c = self._rc()
        out = np.empty(r * c, dtype=np.int64)
        check(lib.eml_val_indexes(self._owner._h, self._v, _ptr(out)))
        return out.reshape(self._lens())

    def get_as_record(self, *position):
        """RecordContainer::get_as_record (container_record/mod.rs:2569 / matrix :2511): element as a scalar Record."""
        number = self._numbers()[tuple(position)]
        return Record(number, self.history_list, int(self.indexes()[tuple(position)]))

    def reset(self):
        check(lib.eml_tape_reset(self._owner._h, self._v))

    def _new(self, handle, shape, history):
        return type(self)(self._owner, handle, shape, history)

    def _unary(self, op, c=0.0):
        v = C.c_int64()
        check(lib.eml_tape_unary(self._owner._h, OPS[op], float(c), self._v, C.byref(v)))
        return self._new(v.value, self._shape, self.history_list)

    def _binary(self, op, other):
        if self._lens() != other._lens() or not self._same_names(other):
            raise Panic("Record containers must have the same shape for a binary operation: "
                        f"(left: {self._shape_text()}, right: {other._shape_text()})")
        _same_list(self.history_list, other.history_list)
        v = C.c_int64()
        check(lib.eml_tape_binary(self._owner._h, OPS[op], self._v, other._v, C.byref(v)))
        return self._new(v.value, self._shape, self.history_list or other.history_list)

    # container (op) scalar, swapped forms, and the Real unary ops
    def __neg__(self):
        return self._unary("NEG")

    def exp(self):
        return self._unary("EXP")

    def ln(self):
        return self._unary("LN")

    def sin(self):
        return self._unary("SIN")

    def cos(self):
        return self._unary("COS")

    def sqrt(self):
        return self._unary("SQRT")

    def pow(self, c):
        return self._unary("POW_C", c)

    def sub_swapped(self, c):
        return self._unary("C_SUB", c)

    def div_swapped(self, c):
        return self._unary("C_DIV", c)

    def __add__(self, o):
        return self._binary("ADD", o) if isinstance(o, _RecordContainer) else self._unary("ADD_C", o)

    def __sub__(self, o):
        return self._binary("SUB", o) if isinstance(o, _RecordContainer) else self._unary("SUB_C", o)

    def __truediv__(self, o):
        if isinstance(o, _RecordContainer):
            raise TypeError("use elementwise_divide for container / container")
        return self._unary("DIV_C", o)

    def elementwise_multiply(self, o):
        return self._binary("MUL", o)

    def elementwise_divide(self, o):
        return self._binary("DIV", o)

    def _matmul(self, other, out_shape):
        _same_list(self.history_list, other.history_list)  # container_operations.rs:1331
        v = C.c_int64()
        check(lib.eml_tape_matmul(self._owner._h, self._v, other._v, C.byref(v)))
        return self._new(v.value, out_shape, self.history_list or other.history_list)

    def _derivatives(self, row, col):
        if self.history_list is None:
            return None
        d = C.c_void_p()
        check(lib.eml_tape_derivatives(self._owner._h, self._v, row, col, C.byref(d)))
        return Derivatives(self._owner, d)


class RecordTensor(_RecordContainer):
    """RecordTensor<'a, T, S, D>, container_record/mod.rs:89."""

    @staticmethod
    def variables(history, tensor):  # mod.rs:149-164
        return RecordTensor._create(history, history, tensor.shape(), tensor.data, True)

    @staticmethod
    def constants(tensor, owner):  # mod.rs:115-128 (owner = the list whose device holds the numbers)
        return RecordTensor._create(owner, None, tensor.shape(), tensor.data, False)

    @staticmethod
    def from_existing(history, shape, numbers, indexes, owner=None):
        """RecordTensor::from_existing(history, TensorView<(T, Index)>), mod.rs:219-224.  `history` None = constants;
        `owner` names the list whose device holds the numbers when history is None."""
        return RecordTensor._from_existing(history if history is not None else owner, history, list(shape), numbers, indexes)

    def range(self, ranges, history="same"):
        """RecordTensor::from_existing(history, TensorView::from(TensorRange::from(self, ranges))) for rank <= 2:
        `ranges` maps dimension names to (start, stop) (the doc example at mod.rs:205-217).  history: "same" keeps this
        container's, None makes constants, or a WengertList."""
        hist = self.history_list if history == "same" else history
        lens, names = self._lens(), [n for n, _ in self._shape]
        if len(lens) > 2:
            raise TypeError("range of a record tensor: rank 1 and 2 containers only")
        spans = [(0, l) for l in lens]
        for name, (a, b) in dict(ranges).items():
            if name not in names:
                raise Panic(f"Invalid dimension name {name!r} for tensor of shape {_shape_debug(self._shape)}")
            d = names.index(name)
            a, b = max(0, a), min(lens[d], b)
            if b <= a:
                raise Panic("range would leave no elements")
            spans[d] = (a, b)
        if len(lens) == 1:
            spans = [(0, 1)] + spans
        (r0, r1), (c0, c1) = spans
        new_shape = [(n, b - a) for n, (a, b) in zip(names, spans[-len(lens):])]
        return self._range(r0, r1 - r0, c0, c1 - c0, new_shape, hist)

    def shape(self):
        return list(self._shape)

    def _lens(self):
        return [l for _, l in self._shape]

    def _same_names(self, other):
        return self._shape == other._shape

    def _shape_text(self):
        return _shape_debug(self._shape)

    def numbers(self):
        return Tensor(self._shape, self._numbers())

    def __mul__(self, other):
        if not isinstance(other, RecordTensor):
            return self._unary("MUL_C", other)
        # record_tensor_matrix_multiply, container_operations.rs:1318-1381
        ls, rs = self._shape, other._shape
        if len(ls) != 2 or len(rs) != 2:
            raise TypeError("* between record tensors is matrix multiplication of rank 2 containers")
        if ls[1][1] != rs[0][1]:  # :1339-1345
            raise Panic(f"Mismatched record tensors, left is {_shape_debug(ls)}, right is {_shape_debug(rs)}, "
                        "* is only defined for MxN * NxL dimension lengths")
        if ls[0][0] == rs[1][0]:  # :1346-1355
            raise Panic(f"Matrix multiplication of record tensors with shapes left {_shape_debug(ls)} and right "
                        f"{_shape_debug(rs)} would create duplicate dimension names as the shape "
                        f"{_shape_debug([ls[0], rs[1]])}. Rename one or both of the dimension names in the input "
                        "to prevent this. * is defined as MxN * NxL = MxL")
        return self._matmul(other, [ls[0], rs[1]])

    def derivatives_for(self, indexes):  # mod.rs:1201-1213
        lens = self._lens()
        if len(indexes) != len(lens) or any(i < 0 or i >= l for i, l in zip(indexes, lens)):
            return None
        flat = int(np.ravel_multi_index(indexes, lens)) if lens else 0
        cols = lens[-1] if lens else 1
        return self._derivatives(flat // cols, flat % cols)


class RecordMatrix(_RecordContainer):
    """RecordMatrix<'a, T, S>, container_record/mod.rs:84."""

    @staticmethod
    def variables(history, matrix):  # mod.rs:426
        return RecordMatrix._create(history, history, matrix.size(), matrix.data, True)

    @staticmethod
    def constants(matrix, owner):  # mod.rs:392
        return RecordMatrix._create(owner, None, matrix.size(), matrix.data, False)

    @staticmethod
    def from_existing(history, numbers, indexes, owner=None):
        """RecordMatrix::from_existing(history, MatrixView<(T, Index)>), mod.rs:508."""
        a = np.asarray(numbers)
        return RecordMatrix._from_existing(history if history is not None else owner, history, tuple(a.shape), numbers, indexes)

    def range(self, rows, columns, history="same"):
        """RecordMatrix::from_existing(history, MatrixView::from(MatrixRange::from(self, IndexRange::new(start, length),
        IndexRange::new(start, length)))) -- how the reference's examples take one input row at a time
        (src/neural_networks.rs:105-120)."""
        hist = self.history_list if history == "same" else history
        (r0, nr), (c0, nc) = rows, columns
        r0, c0 = min(r0, self._shape[0]), min(c0, self._shape[1])
        nr, nc = min(nr, self._shape[0] - r0), min(nc, self._shape[1] - c0)
        if nr < 1 or nc < 1:
            raise Panic("range would leave no elements")
        return self._range(r0, nr, c0, nc, (nr, nc), hist)

    def _lens(self):
        return list(self._shape)

    def _same_names(self, other):
        return True

    def _shape_text(self):
        return f"({self._shape[0]}, {self._shape[1]})"

    def size(self):
        return tuple(self._shape)

    def numbers(self):
        return Matrix(self._numbers())

    def __mul__(self, other):
        if not isinstance(other, RecordMatrix):
            return self._unary("MUL_C", other)
        # record_matrix_matrix_multiply, container_operations.rs:1494-1531
        if self._shape[1] != other._shape[0]:  # :1506-1513
            raise Panic(f"Mismatched Matrices, left is {self._shape[0]}x{self._shape[1]}, right is "
                        f"{other._shape[0]}x{other._shape[1]}, * is only defined for MxN * NxL")
        return self._matmul(other, (self._shape[0], other._shape[1]))

    def derivatives_for(self, row, column):
        if not (0 <= row < self._shape[0] and 0 <= column < self._shape[1]):
            return None
        return self._derivatives(row, column)


class Derivatives:
    """easy_ml::differentiation::Derivatives<T>, src/differentiation.rs:364-414."""

    def __init__(self, owner, handle):
        self._owner, self._d = owner, handle

    def __del__(self):
        if getattr(self, "_d", None):
            lib.eml_derivs_free(self._d)
            self._d = None

    def __len__(self):
        n = C.c_int64()
        check(lib.eml_derivs_len(self._d, C.byref(n)))
        return n.value

    def at(self, record):  # Derivatives::at(&record), src/differentiation.rs:387-392
        return self.at_index(record.index)

    def at_index(self, index):  # Derivatives::at / Index<&Record>
        out = C.c_double()
        try:
            check(lib.eml_derivs_at(self._d, int(index), C.byref(out)))
        except _ffi.BadArgError as e:
            raise Panic(e.message)
        return self._owner.dtype.type(out.value)

    def _at(self, container):
        out = np.empty(container.elements(), dtype=self._owner.dtype)
        check(lib.eml_derivs_at_val(self._d, container._v, _ptr(out)))
        return out.reshape(container._lens())

    def at_tensor(self, record_tensor):  # container_record/mod.rs:1292
        return Tensor(record_tensor.shape(), self._at(record_tensor))

    def at_matrix(self, record_matrix):  # container_record/mod.rs:1328
        return Matrix(self._at(record_matrix))

    def to_vec(self):  # Vec::from(Derivatives), src/differentiation.rs:405-414
        out = np.empty(len(self), dtype=self._owner.dtype)
        check(lib.eml_derivs_to_host(self._d, _ptr(out)))
        return out


def sgd_update(weights, derivatives, learning_rate):
    """`w.map_mut(|x| x - derivatives[&x] * lr)` (src/neural_networks.rs:418-420) fused on the device."""
    check(lib.eml_tape_sgd_update(weights._owner._h, weights._v, derivatives._d, float(learning_rate)))
