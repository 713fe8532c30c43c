"""Scalar `Record<T>` arithmetic over the oracle's literal tape -- test helper.

Follows src/differentiation/record_operations.rs (Add :151-219, Sub, Mul :344-420, Div) with the
derivative rules of src/differentiation/functions.rs.  Used to replay the reference's scalar-Record
tests (tests/differentiation_nn.rs) against the oracle tape.
"""
import ctypes as C

import numpy as np

import oracle


class Rec:
    __slots__ = ("number", "tape", "index")

    def __init__(self, number, tape=None, index=0):
        self.number, self.tape, self.index = number, tape, index

    @staticmethod
    def variable(x, tape):
        i = oracle.lib().orc_tape_append_nullary(tape._h)
        return Rec(tape.dtype.type(x), tape, i)

    @staticmethod
    def constant(x, dtype):
        return Rec(np.dtype(dtype).type(x), None, 0)

    def reset(self):
        if self.tape is not None:
            self.index = oracle.lib().orc_tape_append_nullary(self.tape._h)

    def _un(self, number, d):
        i = oracle.lib().orc_tape_append_unary(self.tape._h, C.c_int64(self.index), C.c_double(float(d)))
        return Rec(number, self.tape, i)

    @staticmethod
    def _bin(x, y, number, dx, dy):
        i = oracle.lib().orc_tape_append_binary(x.tape._h, C.c_int64(x.index), C.c_double(float(dx)),
                                                C.c_int64(y.index), C.c_double(float(dy)))
        return Rec(number, x.tape, i)

    def _coerce(self, o):
        return o if isinstance(o, Rec) else Rec(type(self.number)(o), None, 0)

    def __add__(self, o):
        o = self._coerce(o)
        z = self.number + o.number
        one = type(z)(1)
        if self.tape is None and o.tape is None:
            return Rec(z)
        if o.tape is None:
            return self._un(z, one)
        if self.tape is None:
            return o._un(z, one)
        return Rec._bin(self, o, z, one, one)

    def __sub__(self, o):
        o = self._coerce(o)
        z = self.number - o.number
        one = type(z)(1)
        if self.tape is None and o.tape is None:
            return Rec(z)
        if o.tape is None:
            return self._un(z, one)
        if self.tape is None:
            return o._un(z, -one)  # sub_swapped: d/dy = -1
        return Rec._bin(self, o, z, one, -one)

    def __mul__(self, o):
        o = self._coerce(o)
        z = self.number * o.number
        if self.tape is None and o.tape is None:
            return Rec(z)
        if o.tape is None:
            return self._un(z, o.number)
        if self.tape is None:
            return o._un(z, self.number)
        return Rec._bin(self, o, z, o.number, self.number)

    def __truediv__(self, o):
        o = self._coerce(o)
        z = self.number / o.number
        one = type(z)(1)
        if self.tape is None and o.tape is None:
            return Rec(z)
        if o.tape is None:
            return self._un(z, one / o.number)
        if self.tape is None:
            return o._un(z, -self.number / (o.number * o.number))
        return Rec._bin(self, o, z, one / o.number, -self.number / (o.number * o.number))


def matmul(a, b):
    """Generic Matrix::mul with T = Record (src/matrices/operations.rs:165-196): reduce without a seed."""
    m, k, n = len(a), len(b), len(b[0])
    out = []
    for i in range(m):
        row = []
        for j in range(n):
            acc = a[i][0] * b[0][j]
            for kk in range(1, k):
                acc = acc + (a[i][kk] * b[kk][j])
            row.append(acc)
        out.append(row)
    return out
