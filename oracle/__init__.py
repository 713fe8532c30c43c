"""ctypes view of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (easy-ml_b200/) never does.
Parity pinned against the reference's golden vectors (tests/golden/); see oracle/oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

ORC_OK, ORC_ERR_PANIC, ORC_ERR_INCONSISTENT = 0, -1, -2


class ReferencePanic(Exception):
    """The reference would `panic!` here; str(e) is the reference's message."""


class InconsistentDimensionLengthError(Exception):
    """Einsum `Err(InconsistentDimensionLengthError { lengths, dimension })` (einsum.rs:321)."""

    def __init__(self, dimension, lengths):
        super().__init__(f"InconsistentDimensionLengthError(dimension={dimension!r}, lengths={lengths})")
        self.dimension, self.lengths = dimension, lengths


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.cpp", "oracle.h")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_scalar_product_f32.restype = C.c_float
        _lib.orc_scalar_product_f64.restype = C.c_double
        _lib.orc_tape_new.restype = C.c_void_p
        for f in ("orc_tape_len", "orc_tape_append_nullary", "orc_tape_append_nullary_repeating",
                  "orc_tape_append_unary", "orc_tape_append_binary"):
            getattr(_lib, f).restype = C.c_int64
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(f"oracle handles f32/f64 only, got {dtype}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _names(names):
    arr = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
    return arr


def _i64(v):
    return (C.c_int64 * max(len(v), 1))(*v)


def _check(rc):
    if rc == ORC_ERR_PANIC:
        raise ReferencePanic(lib().orc_last_error().decode())
    if rc != ORC_OK:
        raise RuntimeError(f"oracle rc={rc}")


def scalar_product(l, r):
    """A1 on two 1-D arrays (any stride)."""
    s, _ = _sfx(l.dtype)
    assert l.dtype == r.dtype and l.ndim == r.ndim == 1 and len(l) == len(r) and len(l) > 0
    isz = l.dtype.itemsize
    f = getattr(lib(), f"orc_scalar_product_{s}")
    return f(_p(l), C.c_int64(l.strides[0] // isz), _p(r), C.c_int64(r.strides[0] // isz), C.c_int64(len(l)))


def matrix_mul(a, b, rows=None, threads=1):
    """A2: Matrix * Matrix.  rows=(r0, r1) computes only that slab of C (others left zero)."""
    s, _ = _sfx(a.dtype)
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    c = np.zeros((a.shape[0], b.shape[1]), dtype=a.dtype)
    r0, r1 = rows if rows is not None else (0, a.shape[0])
    rc = getattr(lib(), f"orc_matrix_mul_{s}")(
        _p(a), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]), _p(b), C.c_int64(b.shape[0]), C.c_int64(b.shape[1]),
        _p(c), C.c_int64(r0), C.c_int64(r1), C.c_int(threads))
    _check(rc)
    return c


def matrix_mul_block(a, b, rows, cols, threads=1):
    """C[rows[0]:rows[1], cols[0]:cols[1]] of A * B (a bounded sample for CPU-baseline timing)."""
    s, _ = _sfx(a.dtype)
    assert a.flags.c_contiguous and b.flags.c_contiguous and a.shape[1] == b.shape[0]
    c = np.zeros((rows[1] - rows[0], cols[1] - cols[0]), dtype=a.dtype)
    getattr(lib(), f"orc_matrix_mul_block_{s}")(
        _p(a), C.c_int64(a.shape[1]), _p(b), C.c_int64(b.shape[1]), _p(c), C.c_int64(rows[0]), C.c_int64(rows[1]),
        C.c_int64(cols[0]), C.c_int64(cols[1]), C.c_int(threads))
    return c


def tensor_mul(a, a_names, b, b_names):
    """A3: rank-2 Tensor * Tensor.  Returns (c, out_names)."""
    s, _ = _sfx(a.dtype)
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    c = np.zeros((a.shape[0], b.shape[1] if a.shape[1] == b.shape[0] else 0), dtype=a.dtype)
    on = (C.c_char_p * 2)()
    rc = getattr(lib(), f"orc_tensor_mul_{s}")(
        _p(a), _names(a_names), _i64(a.shape), _p(b), _names(b_names), _i64(b.shape), _p(c), on)
    _check(rc)
    return c, [on[0].decode(), on[1].decode()]


def einsum2(a, a_names, b, b_names, out_names):
    """A4: Einsum::with_2(a, b).named(a_names, b_names).to(out_names).  Returns the output array."""
    s, _ = _sfx(a.dtype)
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    f = getattr(lib(), f"orc_einsum2_{s}")
    ol = (C.c_int64 * max(len(out_names), 1))()
    ed = C.c_char_p()
    el = (C.c_int64 * 2)(-1, -1)
    args = (_p(a), C.c_int(a.ndim), _names(a_names), _i64(a.shape), _p(b), C.c_int(b.ndim), _names(b_names),
            _i64(b.shape), C.c_int(len(out_names)), _names(out_names), ol)
    rc = f(*args, None, C.byref(ed), el)
    if rc == ORC_ERR_INCONSISTENT:
        raise InconsistentDimensionLengthError(ed.value.decode(), [None if x < 0 else int(x) for x in el])
    _check(rc)
    out = np.zeros([ol[i] for i in range(len(out_names))], dtype=a.dtype)
    rc = f(*args, _p(out), C.byref(ed), el)
    _check(rc)
    return out


def einsum_n(arrays, names, out_names):
    """Einsum::with_N(...).named(...).to(out_names) for N = 1..6 (src/tensors/einsum.rs:829-883, :1019-1666)."""
    n = len(arrays)
    s, _ = _sfx(arrays[0].dtype)
    arrays = [np.ascontiguousarray(a) for a in arrays]
    xs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    ranks = (C.c_int * n)(*[a.ndim for a in arrays])
    name_arrays = [_names(nm) for nm in names]
    len_arrays = [_i64(a.shape) for a in arrays]
    names_pp = (C.c_void_p * n)(*[C.cast(na, C.c_void_p).value for na in name_arrays])
    lens_pp = (C.c_void_p * n)(*[C.cast(la, C.c_void_p).value for la in len_arrays])
    f = getattr(lib(), f"orc_einsum_n_{s}")
    ol = (C.c_int64 * max(len(out_names), 1))()
    ed = C.c_char_p()
    el = (C.c_int64 * n)(*([-1] * n))
    on = _names(out_names)
    rc = f(C.c_int(n), xs, ranks, names_pp, lens_pp, C.c_int(len(out_names)), on, ol, None, C.byref(ed), el)
    if rc == ORC_ERR_INCONSISTENT:
        raise InconsistentDimensionLengthError(ed.value.decode(), [None if v < 0 else int(v) for v in el])
    _check(rc)
    out = np.zeros([ol[i] for i in range(len(out_names))], dtype=arrays[0].dtype)
    rc = f(C.c_int(n), xs, ranks, names_pp, lens_pp, C.c_int(len(out_names)), on, ol, _p(out), C.byref(ed), el)
    _check(rc)
    return out


def vector_dot(a, a_name, b, b_name):
    s, ct = _sfx(a.dtype)
    out = ct()
    rc = getattr(lib(), f"orc_vector_dot_{s}")(_p(np.ascontiguousarray(a)), a_name.encode(), C.c_int64(len(a)),
                                             _p(np.ascontiguousarray(b)), b_name.encode(), C.c_int64(len(b)),
                                             C.byref(out))
    _check(rc)
    return out.value


def covariance(x, feature_axis):
    """covariance_column_features (feature_axis=1) / covariance_row_features (0), linear_algebra.rs:615-700."""
    x = np.ascontiguousarray(x)
    sfx, _ = _sfx(x.dtype)
    f = x.shape[1] if feature_axis == 1 else x.shape[0]
    out = np.zeros((f, f), dtype=x.dtype)
    getattr(lib(), f"orc_covariance_{sfx}")(_p(x), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.c_int(feature_axis),
                                            _p(out))
    return out



def cholesky(a):
    """cholesky_decomposition_less_generic, src/linear_algebra.rs:1048-1101: literal triple loop, None when the
    input is not square or a diagonal entry would not be positive."""
    a = np.asarray(a)
    n = a.shape[0]
    if a.ndim != 2 or a.shape[1] != n:
        return None
    t = a.dtype.type
    low = np.zeros((n, n), dtype=a.dtype)
    for i in range(n):
        for j in range(i + 1):
            s = t(0)
            for k in range(j):
                s = s + low[i, k] * low[j, k]
            if i == j:
                entry_squared = a[i, j] - s
                if entry_squared <= t(0):
                    return None
                low[i, j] = np.sqrt(entry_squared)
            else:
                low[i, j] = (a[i, j] - s) * (t(1) / low[j, j])  # :1093-1096: reciprocal, then multiply (two roundings)
    return low


def gaussian_draw(mean, variance, source, max_samples, dtype):
    """Gaussian::draw, src/distributions.rs:205-237: Box-Muller over pairs pulled from `source` (a Python iterator),
    the surplus sample of the last pair popped; None when the source runs out."""
    t = np.dtype(dtype).type
    two = t(1) + t(1)
    minus_two = -two
    two_pi = two * t(np.pi)
    standard_deviation = np.sqrt(t(variance))
    samples = []
    while len(samples) < max_samples:
        try:
            u, v = t(next(source)), t(next(source))
        except StopIteration:
            return None
        z1 = np.sqrt(minus_two * np.log(u)) * np.cos(two_pi * v)
        z2 = np.sqrt(minus_two * np.log(u)) * np.sin(two_pi * v)
        samples.append(z1 * standard_deviation + t(mean))
        samples.append(z2 * standard_deviation + t(mean))
    if len(samples) > max_samples:
        samples.pop()
    return np.array(samples, dtype=dtype)


def multivariate_gaussian_draw(mean, covariance, source, max_samples):
    """draw_tensor_samples, src/distributions.rs:484-540: one `mean + L * z` matrix-vector product per sample row
    (the product through matrix_mul, i.e. the reference's scalar_product order)."""
    covariance = np.asarray(covariance)
    dtype = covariance.dtype
    mean = np.asarray(mean, dtype=dtype)
    low = cholesky(covariance)
    if low is None:
        return None
    n = mean.shape[0]
    source = iter(source)
    drawn = np.zeros((max_samples, n), dtype=dtype)
    for row in range(max_samples):
        z = gaussian_draw(0, 1, source, n, dtype)
        if z is None:
            return None
        drawn[row] = mean + matrix_mul(low, z.reshape(n, 1))[:, 0]
    return drawn


class Tape:
    """The literal scalar WengertList (src/differentiation.rs:327-352)."""

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)
        self.sfx, _ = _sfx(dtype)
        self._h = C.c_void_p(lib().orc_tape_new(self.dtype.itemsize))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_tape_free(self._h)
            self._h = None

    def __len__(self):
        return lib().orc_tape_len(self._h)

    def clear(self):
        lib().orc_tape_clear(self._h)

    def entry(self, i):
        lp, rp, ld, rd = C.c_int64(), C.c_int64(), C.c_double(), C.c_double()
        lib().orc_tape_get(self._h, C.c_int64(i), C.byref(lp), C.byref(rp), C.byref(ld), C.byref(rd))
        return lp.value, rp.value, ld.value, rd.value

    def derivatives(self, index):
        """Record::try_derivatives: full derivative vector (length = tape length)."""
        out = np.zeros(len(self), dtype=self.dtype)
        lib().orc_tape_derivatives(self._h, C.c_int64(index), _p(out))
        return out

    # -- containers as (numbers, indexes) pairs --
    def variables(self, x):
        """RecordTensor::variables (container_record/mod.rs:149-164)."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        start = lib().orc_tape_append_nullary_repeating(self._h, C.c_int64(x.size))
        return x.copy(), (start + np.arange(x.size, dtype=np.int64)).reshape(x.shape)

    @staticmethod
    def constants(x, dtype=None):
        """RecordTensor::constants (mod.rs:115-128): index 0 everywhere, no history."""
        x = np.ascontiguousarray(x, dtype=dtype)
        return x.copy(), np.zeros(x.shape, dtype=np.int64)

    def matmul(self, a, b, tracked=True):
        """record_tensor_matrix_multiply / record_matrix_matrix_multiply."""
        (an, ai), (bn, bi) = a, b
        m, k = an.shape
        n = bn.shape[1]
        c = np.zeros((m, n), dtype=self.dtype)
        ci = np.zeros((m, n), dtype=np.int64)
        getattr(lib(), f"orc_record_matmul_{self.sfx}")(
            self._h if tracked else None, _p(np.ascontiguousarray(an)), _p(np.ascontiguousarray(ai)),
            _p(np.ascontiguousarray(bn)), _p(np.ascontiguousarray(bi)), C.c_int64(m), C.c_int64(k), C.c_int64(n),
            _p(c), _p(ci))
        return c, ci

    def scalar_record_matmul(self, a, b, a_tracked=True, b_tracked=True):
        (an, ai), (bn, bi) = a, b
        m, k = an.shape
        n = bn.shape[1]
        c = np.zeros((m, n), dtype=self.dtype)
        ci = np.zeros((m, n), dtype=np.int64)
        getattr(lib(), f"orc_scalar_record_matmul_{self.sfx}")(
            self._h, _p(np.ascontiguousarray(an)), _p(np.ascontiguousarray(ai)), C.c_int(a_tracked),
            _p(np.ascontiguousarray(bn)), _p(np.ascontiguousarray(bi)), C.c_int(b_tracked), C.c_int64(m),
            C.c_int64(k), C.c_int64(n), _p(c), _p(ci))
        return c, ci

    def unary(self, op, x, scalar=0.0, tracked=True):
        xn, xi = x
        y = np.zeros(xn.shape, dtype=self.dtype)
        yi = np.zeros(xn.shape, dtype=np.int64)
        ct = C.c_float if self.sfx == "f32" else C.c_double
        getattr(lib(), f"orc_record_unary_{self.sfx}")(
            self._h if tracked else None, C.c_int(op), ct(scalar), _p(np.ascontiguousarray(xn)),
            _p(np.ascontiguousarray(xi)), C.c_int64(xn.size), _p(y), _p(yi))
        return y, yi

    def binary(self, op, x, y, mode=3):
        xn, xi = x
        yn, yi = y
        z = np.zeros(xn.shape, dtype=self.dtype)
        zi = np.zeros(xn.shape, dtype=np.int64)
        getattr(lib(), f"orc_record_binary_{self.sfx}")(
            self._h if mode else None, C.c_int(op), C.c_int(mode), _p(np.ascontiguousarray(xn)),
            _p(np.ascontiguousarray(xi)), _p(np.ascontiguousarray(yn)), _p(np.ascontiguousarray(yi)),
            C.c_int64(xn.size), _p(z), _p(zi))
        return z, zi
