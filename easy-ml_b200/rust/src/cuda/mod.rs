//! `cuda` feature: the f32/f64 dense-contraction hot path on a B200 through libeasyml_b200.
//!
//! This module is crate-private.  The public API (`Mul` for `Matrix`/`Tensor`/views, `Einsum`, `RecordTensor`,
//! `Record::derivatives`) is unchanged; its five private entry functions call the safe wrappers below when the
//! element type is `f32` or `f64` (see `patches/`).  There is no CPU fallback for those types under this feature:
//! a missing device is a panic carrying the library's message.
//!
//! Not compiled in the environment this was written in (no Rust toolchain); the call sequence is the one the
//! C++ mirror (`easy-ml_b200/host/easyml.hpp`) and the Python mirror (`easy-ml_b200/host.py`) exercise on a B200.
#![cfg(feature = "cuda")]
#![allow(dead_code)]

pub(crate) mod einsum;
pub(crate) mod ffi;
pub(crate) mod record;

use core::ffi::c_int;

/// Which implementation a numeric type takes.  `Mul` is implemented once for every `T: Numeric` through blanket
/// impls (src/numeric.rs:153-164) and stable Rust has no specialisation; `T` is not `'static` either
/// (`Record<'a, _>`), so `TypeId` is out, and `type_name`'s output is explicitly not guaranteed by std.  The robust
/// form is an associated const: `ZeroOne` -- a supertrait of `Numeric` that every numeric type implements by hand
/// (src/numeric.rs:179-243) -- gains `#[doc(hidden)] const CUDA_KIND: Kind = Kind::GENERIC`, and its `f32` / `f64`
/// impls override it (patches/numeric.rs).  `Kind` is public but lives in this private module, so no downstream
/// crate can name it: nobody outside the crate can claim to be `f32` (the "sealed" pattern), and the pointer casts
/// below are sound.
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub struct Kind(u8);

impl Kind {
    pub const GENERIC: Kind = Kind(0);
    pub const F32: Kind = Kind(1);
    pub const F64: Kind = Kind(2);
}

/// `T::CUDA_KIND`, resolved at compile time (the `if` below folds away for every other type).
#[inline(always)]
pub(crate) fn kind_of<T: crate::numeric::ZeroOne>() -> Kind {
    T::CUDA_KIND
}

/// f32 / f64 take the device path.
#[inline(always)]
pub(crate) fn on_device<T: crate::numeric::ZeroOne>() -> bool {
    T::CUDA_KIND != Kind::GENERIC
}

/// Non-zero code -> panic with the library's thread-local message.
#[track_caller]
pub(crate) fn check(code: c_int) {
    if code != ffi::EML_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::eml_last_error()) }.to_string_lossy();
        panic!("easy-ml cuda backend: {msg} (code {code})");
    }
}

/// Reinterprets a slice of `T` as `U` when `kind_of::<T>()` said they are the same type.
#[inline(always)]
unsafe fn cast<T, U>(s: &[T]) -> *const U {
    debug_assert_eq!(core::mem::size_of::<T>(), core::mem::size_of::<U>());
    s.as_ptr().cast()
}
#[inline(always)]
unsafe fn cast_mut<T, U>(s: &mut [T]) -> *mut U {
    debug_assert_eq!(core::mem::size_of::<T>(), core::mem::size_of::<U>());
    s.as_mut_ptr().cast()
}

/// C[m x n] = A[m x k] * B[k x n], row-major contiguous.  `T` must be f32 or f64 (checked by the caller with
/// `kind_of`).  Replaces the loop of matrix_view_multiplication (src/matrices/operations.rs:185-195) and of
/// tensor_view_matrix_product (src/tensors/operations.rs:506-518).
pub(crate) fn gemm<T: crate::numeric::ZeroOne>(a: &[T], b: &[T], c: &mut [T], m: usize, n: usize, k: usize) {
    assert!(a.len() == m * k && b.len() == k * n && c.len() == m * n);
    let (m, n, k) = (m as i64, n as i64, k as i64);
    unsafe {
        match kind_of::<T>() {
            Kind::F32 => check(ffi::eml_gemm_f32(cast(a), cast(b), cast_mut(c), m, n, k, k, n, n)),
            Kind::F64 => check(ffi::eml_gemm_f64(cast(a), cast(b), cast_mut(c), m, n, k, k, n, n)),
            _ => unreachable!("gemm is only dispatched for f32 / f64"),
        }
    }
}

/// Strided batched contraction C[b,m,n] = sum_k A[b,m,k] B[b,k,n]; strides in elements for (batch, row, column).
/// Replaces the loop nest of Einsum2::to (src/tensors/einsum.rs:931-975) for GEMM-shaped name patterns.
pub(crate) fn contract<T: crate::numeric::ZeroOne>(a: &[T], b: &[T], c: &mut [T], dims: [usize; 4], sa: [i64; 3], sb: [i64; 3], sc: [i64; 3]) {
    let [batch, m, n, k] = dims.map(|d| d as i64);
    unsafe {
        match kind_of::<T>() {
            Kind::F32 => check(ffi::eml_contract_f32(cast(a), cast(b), cast_mut(c), batch, m, n, k, sa.as_ptr(), sb.as_ptr(), sc.as_ptr())),
            Kind::F64 => check(ffi::eml_contract_f64(cast(a), cast(b), cast_mut(c), batch, m, n, k, sa.as_ptr(), sb.as_ptr(), sc.as_ptr())),
            _ => unreachable!(),
        }
    }
}

/// Generic N-operand einsum (1..=6 operands) in the reference's exact order: bit-exact (Einsum1, Einsum3..6, and
/// the two-operand patterns that are not GEMMs).  `out_strides` is `[n_ops][n_out]`, `sum_strides` `[n_ops][n_sum]`.
pub(crate) fn einsum_n<T: crate::numeric::ZeroOne>(xs: &[&[T]], out: &mut [T], out_len: &[i64], out_strides: &[i64], sum_len: &[i64], sum_strides: &[i64]) {
    let elems: Vec<i64> = xs.iter().map(|x| x.len() as i64).collect();
    unsafe {
        match kind_of::<T>() {
            Kind::F32 => {
                let ptrs: Vec<*const f32> = xs.iter().map(|x| cast(x)).collect();
                check(ffi::eml_einsum_n_f32(xs.len() as c_int, ptrs.as_ptr(), elems.as_ptr(), cast_mut(out), out_len.len() as c_int,
                                            out_len.as_ptr(), out_strides.as_ptr(), sum_len.len() as c_int, sum_len.as_ptr(), sum_strides.as_ptr()))
            }
            Kind::F64 => {
                let ptrs: Vec<*const f64> = xs.iter().map(|x| cast(x)).collect();
                check(ffi::eml_einsum_n_f64(xs.len() as c_int, ptrs.as_ptr(), elems.as_ptr(), cast_mut(out), out_len.len() as c_int,
                                            out_len.as_ptr(), out_strides.as_ptr(), sum_len.len() as c_int, sum_len.as_ptr(), sum_strides.as_ptr()))
            }
            _ => unreachable!(),
        }
    }
}

/// covariance_column_features / covariance_row_features / covariance (src/linear_algebra.rs:615-832).
pub(crate) fn covariance<T: crate::numeric::ZeroOne>(x: &[T], rows: usize, cols: usize, feature_axis: usize, out: &mut [T]) {
    unsafe {
        match kind_of::<T>() {
            Kind::F32 => check(ffi::eml_covariance_f32(cast(x), rows as i64, cols as i64, cols as i64, feature_axis as c_int, cast_mut(out))),
            Kind::F64 => check(ffi::eml_covariance_f64(cast(x), rows as i64, cols as i64, cols as i64, feature_axis as c_int, cast_mut(out))),
            _ => unreachable!(),
        }
    }
}

/// The device tape behind one `WengertList<f32 | f64>` (src/differentiation.rs:327-331).  One macro node per
/// container operation; tape length and every `Index` equal the reference's scalar tape (csrc/tape.cu).
pub(crate) struct DeviceTape {
    raw: *mut ffi::EmlTape,
}

impl DeviceTape {
    pub(crate) fn new<T: crate::numeric::ZeroOne>() -> Self {
        let mut raw = core::ptr::null_mut();
        let dtype = if kind_of::<T>() == Kind::F32 { ffi::EML_F32 } else { ffi::EML_F64 };
        check(unsafe { ffi::eml_tape_create(dtype, &mut raw) });
        DeviceTape { raw }
    }
    /// operations.len()
    pub(crate) fn len(&self) -> usize {
        let mut n = 0i64;
        check(unsafe { ffi::eml_tape_len(self.raw, &mut n) });
        n as usize
    }
    /// WengertList::clear (src/differentiation.rs:674)
    pub(crate) fn clear(&self) {
        check(unsafe { ffi::eml_tape_clear(self.raw) });
    }
    /// RecordTensor::variables / RecordMatrix::variables (container_record/mod.rs:149-164, :426)
    pub(crate) fn variables<T>(&self, numbers: &[T], rows: usize, cols: usize) -> ffi::EmlVal {
        let mut v = 0;
        check(unsafe { ffi::eml_tape_variables(self.raw, numbers.as_ptr().cast(), rows as i64, cols as i64, &mut v) });
        v
    }
    /// RecordTensor::constants (mod.rs:115-128)
    pub(crate) fn constants<T>(&self, numbers: &[T], rows: usize, cols: usize) -> ffi::EmlVal {
        let mut v = 0;
        check(unsafe { ffi::eml_tape_constants(self.raw, numbers.as_ptr().cast(), rows as i64, cols as i64, &mut v) });
        v
    }
    /// record_tensor_matrix_multiply / record_matrix_matrix_multiply (container_operations.rs:1318-1381, :1494-1531)
    pub(crate) fn matmul(&self, a: ffi::EmlVal, b: ffi::EmlVal) -> ffi::EmlVal {
        let mut c = 0;
        check(unsafe { ffi::eml_tape_matmul(self.raw, a, b, &mut c) });
        c
    }
    /// RecordTensor::unary front-ends (container_operations.rs:1644-1804, swapped.rs:25-45); `op` is an EML_OP_*
    pub(crate) fn unary(&self, op: c_int, scalar: f64, x: ffi::EmlVal) -> ffi::EmlVal {
        let mut y = 0;
        check(unsafe { ffi::eml_tape_unary(self.raw, op, scalar, x, &mut y) });
        y
    }
    /// RecordTensor::binary front-ends (container_record/mod.rs:962-1043, :1224, :1247)
    pub(crate) fn binary(&self, op: c_int, x: ffi::EmlVal, y: ffi::EmlVal) -> ffi::EmlVal {
        let mut z = 0;
        check(unsafe { ffi::eml_tape_binary(self.raw, op, x, y, &mut z) });
        z
    }
    /// RecordTensor::reset (mod.rs:349-365)
    pub(crate) fn reset(&self, v: ffi::EmlVal) {
        check(unsafe { ffi::eml_tape_reset(self.raw, v) });
    }
    /// The `T` half and the `Index` half of the `(T, Index)` view (mod.rs:2511, :2569)
    pub(crate) fn numbers<T: Clone + Default>(&self, v: ffi::EmlVal, n: usize) -> Vec<T> {
        let mut out = vec![T::default(); n];
        check(unsafe { ffi::eml_val_numbers(self.raw, v, out.as_mut_ptr().cast()) });
        out
    }
    pub(crate) fn indexes(&self, v: ffi::EmlVal, n: usize) -> Vec<usize> {
        let mut out = vec![0i64; n];
        check(unsafe { ffi::eml_val_indexes(self.raw, v, out.as_mut_ptr()) });
        out.into_iter().map(|i| i as usize).collect()
    }
    /// Record::try_derivatives (src/differentiation.rs:623-648) for element (row, col) of a container
    pub(crate) fn derivatives(&self, v: ffi::EmlVal, row: usize, col: usize) -> DeviceDerivatives {
        let mut d = core::ptr::null_mut();
        check(unsafe { ffi::eml_tape_derivatives(self.raw, v, row as i64, col as i64, &mut d) });
        DeviceDerivatives { raw: d }
    }
    pub(crate) fn release(&self, v: ffi::EmlVal) {
        check(unsafe { ffi::eml_val_release(self.raw, v) });
    }
    /// RecordTensor::from_existing / RecordMatrix::from_existing (mod.rs:219-224, :508): any (number, Index) pairs
    pub(crate) fn from_existing<T>(&self, numbers: &[T], indexes: &[usize], rows: usize, cols: usize, has_history: bool) -> ffi::EmlVal {
        assert!(numbers.len() == rows * cols && indexes.len() == rows * cols);
        let idx: Vec<i64> = indexes.iter().map(|&i| i as i64).collect();
        let mut v = 0;
        check(unsafe {
            ffi::eml_tape_from_existing(self.raw, numbers.as_ptr().cast(), idx.as_ptr(), rows as i64, cols as i64, has_history as c_int, &mut v)
        });
        v
    }
    /// WengertList::append_nullary (src/differentiation.rs:811-823)
    pub(crate) fn append_nullary(&self) -> usize {
        let mut i = 0i64;
        check(unsafe { ffi::eml_tape_append_nullary(self.raw, &mut i) });
        i as usize
    }
    /// WengertList::append_unary (:828-846)
    pub(crate) fn append_unary(&self, parent: usize, derivative: f64) -> usize {
        let mut i = 0i64;
        check(unsafe { ffi::eml_tape_append_unary(self.raw, parent as i64, derivative, &mut i) });
        i as usize
    }
    /// WengertList::append_binary (:855-878)
    pub(crate) fn append_binary(&self, lp: usize, ld: f64, rp: usize, rd: f64) -> usize {
        let mut i = 0i64;
        check(unsafe { ffi::eml_tape_append_binary(self.raw, lp as i64, ld, rp as i64, rd, &mut i) });
        i as usize
    }
    /// Record::try_derivatives (:623-648) from the Record at tape index `index`
    pub(crate) fn derivatives_at(&self, index: usize) -> DeviceDerivatives {
        let mut d = core::ptr::null_mut();
        check(unsafe { ffi::eml_tape_derivatives_at(self.raw, index as i64, &mut d) });
        DeviceDerivatives { raw: d }
    }
}

impl Drop for DeviceTape {
    fn drop(&mut self) {
        unsafe { ffi::eml_tape_destroy(self.raw) };
    }
}

/// `Derivatives<T>` (src/differentiation.rs:365-367) when the tape lives on the device.
pub(crate) struct DeviceDerivatives {
    raw: *mut ffi::EmlDerivs,
}

impl DeviceDerivatives {
    /// `derivatives[&record]` (src/differentiation.rs:387-403)
    pub(crate) fn at(&self, index: usize) -> f64 {
        let mut v = 0.0;
        check(unsafe { ffi::eml_derivs_at(self.raw, index as i64, &mut v) });
        v
    }
    /// Derivatives::at_tensor / at_matrix (container_record/mod.rs:1292, :1328)
    pub(crate) fn at_container<T: Clone + Default>(&self, v: ffi::EmlVal, n: usize) -> Vec<T> {
        let mut out = vec![T::default(); n];
        check(unsafe { ffi::eml_derivs_at_val(self.raw, v, out.as_mut_ptr().cast()) });
        out
    }
    /// `Vec::from(Derivatives)` (src/differentiation.rs:405-414)
    pub(crate) fn to_vec<T: Clone + Default>(&self) -> Vec<T> {
        let mut n = 0i64;
        check(unsafe { ffi::eml_derivs_len(self.raw, &mut n) });
        let mut out = vec![T::default(); n as usize];
        check(unsafe { ffi::eml_derivs_to_host(self.raw, out.as_mut_ptr().cast()) });
        out
    }
}

impl Drop for DeviceDerivatives {
    fn drop(&mut self) {
        unsafe { ffi::eml_derivs_free(self.raw) };
    }
}
