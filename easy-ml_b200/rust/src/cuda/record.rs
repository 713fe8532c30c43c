//! `WengertList<f32 | f64>`, `Record`, `RecordTensor` / `RecordMatrix` on the device-backed list.
//!
//! Under the `cuda` feature a `WengertList<T>` whose `T` is `f32` / `f64` owns a `DeviceTape` instead of a
//! `Vec<Operation<T>>`.  Everything that appends to the list goes to the device tape, so scalar Records and containers
//! keep sharing ONE index space exactly as in the reference:
//!
//!   WengertList::append_nullary / append_unary / append_binary   (src/differentiation.rs:811-878)  -> eml_tape_append_*
//!   WengertList::append_nullary_repeating (:708-727), RecordContainer::variables / reset            -> eml_tape_variables / reset
//!   record_tensor_matrix_multiply / record_matrix_matrix_multiply (container_operations.rs:1318-1531) -> matmul() below
//!   RecordContainer::unary / binary and the operator front-ends (mod.rs:845-1043)                  -> unary() / binary()
//!   Record::try_derivatives (:623-648), RecordContainer::derivatives_for                            -> eml_tape_derivatives_at
//!
//! The containers keep their `(T, Index)` storage on the host (the public API hands out views of it); an operation
//! uploads its operands with their indexes (`eml_tape_from_existing`: O(MK) next to the O(MNK) product), runs on the
//! device, and downloads the numbers and the closed-form indexes of the result.
#![cfg(feature = "cuda")]

use super::{ffi, DeviceDerivatives, DeviceTape};
use crate::differentiation::Index;

/// A container operand as the reference stores it: row-major `(number, Index)` pairs, `rows x cols`, and whether the
/// container has a history (`Some(&WengertList)`).
pub(crate) struct Pairs<'a, T> {
    pub pairs: &'a [(T, Index)],
    pub rows: usize,
    pub cols: usize,
    pub has_history: bool,
}

/// Uploads the operand; the handle lives until `release`.
fn upload<T: Clone>(tape: &DeviceTape, x: &Pairs<T>) -> ffi::EmlVal {
    let numbers: Vec<T> = x.pairs.iter().map(|(n, _)| n.clone()).collect();
    let indexes: Vec<usize> = x.pairs.iter().map(|(_, i)| *i).collect();
    tape.from_existing(&numbers, &indexes, x.rows, x.cols, x.has_history)
}

/// Downloads a result as `(number, Index)` pairs.
fn download<T: Clone + Default>(tape: &DeviceTape, v: ffi::EmlVal, elements: usize) -> Vec<(T, Index)> {
    let numbers: Vec<T> = tape.numbers(v, elements);
    let indexes = tape.indexes(v, elements);
    numbers.into_iter().zip(indexes).collect()
}

/// record_tensor_matrix_multiply / record_matrix_matrix_multiply after their asserts: C = A * B with the M*N*(2K-1)
/// tape entries of record_scalar_product (container_operations.rs:1263-1315) accounted for as one macro node.
pub(crate) fn matmul<T: Clone + Default>(tape: &DeviceTape, left: &Pairs<T>, right: &Pairs<T>) -> Vec<(T, Index)> {
    let (a, b) = (upload(tape, left), upload(tape, right));
    let c = tape.matmul(a, b);
    let out = download(tape, c, left.rows * right.cols);
    for v in [a, b, c] {
        tape.release(v);
    }
    out
}

/// RecordContainer::unary (mod.rs:845-861) for the closed set of rules of src/differentiation/functions.rs; `op` is an
/// `ffi::EML_OP_*`.  Closure-taking `unary` / `map` stay on the generic path (a closure cannot cross the C ABI).
pub(crate) fn unary<T: Clone + Default>(tape: &DeviceTape, op: i32, scalar: f64, x: &Pairs<T>) -> Vec<(T, Index)> {
    let a = upload(tape, x);
    let y = tape.unary(op, scalar, a);
    let out = download(tape, y, x.rows * x.cols);
    tape.release(a);
    tape.release(y);
    out
}

/// RecordContainer::binary (mod.rs:962-1043): the history selection (None, None) / (Some, None) / (None, Some) /
/// (Some, Some) follows from the operands' `has_history`.
pub(crate) fn binary<T: Clone + Default>(tape: &DeviceTape, op: i32, x: &Pairs<T>, y: &Pairs<T>) -> Vec<(T, Index)> {
    let (a, b) = (upload(tape, x), upload(tape, y));
    let z = tape.binary(op, a, b);
    let out = download(tape, z, x.rows * x.cols);
    for v in [a, b, z] {
        tape.release(v);
    }
    out
}

/// Record::try_derivatives (src/differentiation.rs:623-648): the whole derivative vector of the Record at `index`.
pub(crate) fn derivatives<T: Clone + Default>(tape: &DeviceTape, index: Index) -> Vec<T> {
    let d: DeviceDerivatives = tape.derivatives_at(index);
    d.to_vec()
}
