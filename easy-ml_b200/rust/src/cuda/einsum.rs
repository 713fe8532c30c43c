//! Einsum2::to (src/tensors/einsum.rs:908-980) on the device: classifies the dimension names of a two-operand Einstein
//! summation into GEMM roles and picks the entry point.
//!
//!   name in both inputs and the output      -> batch
//!   name in input 1 and the output only     -> M
//!   name in input 2 and the output only     -> N
//!   name in both inputs, not the output     -> K      (a summation dimension, src/tensors/einsum.rs:579-622)
//!   name summed on one side only            -> not a GEMM
//!
//! Several names may share one role when they nest contiguously, in the same order, in every tensor that carries them
//! (they merge into one strided dimension).  When every role merges the pattern is a (batched) GEMM and runs on the
//! tensor cores through `eml_contract_*` with the tensors' own element strides -- no transposes, no copies.  Everything
//! else (`ij,jk->ijk`, `ij,kl->il`, `ab,ab->`, a dimension summed on one side only ...) runs as the strided loop nest of
//! `eml_einsum2_*`, in the reference's own order, bit-exact.  The same classification is exercised on a B200 by
//! easy-ml_b200/host.py::_Einsum2.to and host/easyml.hpp::Einsum2::to against tests/einsum.rs.
#![cfg(feature = "cuda")]

use super::{check, ffi, kind_of, Kind};
use crate::numeric::ZeroOne;
use crate::tensors::Dimension;

/// One operand: its shape in storage order and the element stride of every dimension.
pub(crate) struct Operand<'a, T> {
    pub shape: &'a [(Dimension, usize)],
    pub strides: &'a [usize],
    pub data: &'a [T],
}

impl<T> Operand<'_, T> {
    fn stride_of(&self, name: Dimension) -> i64 {
        self.shape.iter().position(|(d, _)| *d == name).map(|i| self.strides[i] as i64).unwrap_or(0)
    }
    fn has(&self, name: Dimension) -> bool {
        self.shape.iter().any(|(d, _)| *d == name)
    }
}

/// A GEMM role after merging its names: total length and one stride per tensor that was asked about.
struct Role<const S: usize> {
    length: i64,
    strides: [i64; S],
}

/// Collapses `names` (each with its length) into one strided dimension.  `strides_of[t](name)` is the stride of `name`
/// in tensor `t`.  Possible only when in every tensor `stride(a) == stride(b) * length(b)` for neighbours a, b.
fn merge<const S: usize>(names: &[(Dimension, usize)], strides_of: [&dyn Fn(Dimension) -> i64; S]) -> Option<Role<S>> {
    if names.is_empty() {
        return Some(Role { length: 1, strides: [0; S] });
    }
    // dimensions of length 1 never constrain the nesting
    let mut kept: Vec<(Dimension, usize)> = names.iter().copied().filter(|(_, l)| *l > 1).collect();
    if kept.is_empty() {
        kept.push(names[0]);
    }
    let length = kept.iter().map(|(_, l)| *l as i64).product();
    let mut strides = [0i64; S];
    for (t, stride_of) in strides_of.iter().enumerate() {
        for pair in kept.windows(2) {
            let ((a, _), (b, lb)) = (pair[0], pair[1]);
            if stride_of(a) != stride_of(b) * lb as i64 {
                return None;
            }
        }
        strides[t] = stride_of(kept[kept.len() - 1].0);
    }
    Some(Role { length, strides })
}

/// Computes `out` (row-major over `output_shape`) = einsum of `x` and `y`.  `summation` is what
/// `summation_dimensions` returned (first appearance order), `output_shape` what `output_shape_for` returned: both
/// validations (and their `Err` values) stay in `Einsum2::to`.
pub(crate) fn einsum2<T: ZeroOne>(
    x: &Operand<T>,
    y: &Operand<T>,
    output_shape: &[(Dimension, usize)],
    summation: &[(Dimension, usize)],
    out: &mut [T],
) {
    // row-major strides of the output
    let mut out_strides = vec![1i64; output_shape.len()];
    for d in (0..output_shape.len().saturating_sub(1)).rev() {
        out_strides[d] = out_strides[d + 1] * output_shape[d + 1].1 as i64;
    }
    let so = |name: Dimension| -> i64 {
        output_shape.iter().position(|(d, _)| *d == name).map(|i| out_strides[i]).unwrap_or(0)
    };
    let sx = |name: Dimension| x.stride_of(name);
    let sy = |name: Dimension| y.stride_of(name);

    let batch: Vec<_> = output_shape.iter().copied().filter(|(d, _)| x.has(*d) && y.has(*d)).collect();
    let m_dims: Vec<_> = output_shape.iter().copied().filter(|(d, _)| x.has(*d) && !y.has(*d)).collect();
    let n_dims: Vec<_> = output_shape.iter().copied().filter(|(d, _)| y.has(*d) && !x.has(*d)).collect();
    let k_dims: Vec<_> = summation.iter().copied().filter(|(d, _)| x.has(*d) && y.has(*d)).collect();
    let lone_sum = summation.iter().any(|(d, _)| !(x.has(*d) && y.has(*d)));

    if !k_dims.is_empty() && !lone_sum {
        let b = merge(&batch, [&sx, &sy, &so]);
        let m = merge(&m_dims, [&sx, &so]);
        let n = merge(&n_dims, [&sy, &so]);
        let k = merge(&k_dims, [&sx, &sy]);
        if let (Some(b), Some(m), Some(n), Some(k)) = (b, m, n, k) {
            // C[b, m, n] = sum_k A[b, m, k] * B[b, k, n]; strides for (batch, row, column) of each tensor
            let sa = [b.strides[0], m.strides[0], k.strides[0]];
            let sb = [b.strides[1], k.strides[1], n.strides[0]];
            let sc = [b.strides[2], m.strides[1], n.strides[1]];
            super::contract(x.data, y.data, out, [b.length as usize, m.length as usize, n.length as usize, k.length as usize], sa, sb, sc);
            return;
        }
    }
    // not a (batched) GEMM: the strided loop nest, bit-exact with the reference
    let out_len: Vec<i64> = output_shape.iter().map(|(_, l)| *l as i64).collect();
    let a_out: Vec<i64> = output_shape.iter().map(|(d, _)| sx(*d)).collect();
    let b_out: Vec<i64> = output_shape.iter().map(|(d, _)| sy(*d)).collect();
    let sum_len: Vec<i64> = summation.iter().map(|(_, l)| *l as i64).collect();
    let a_sum: Vec<i64> = summation.iter().map(|(d, _)| sx(*d)).collect();
    let b_sum: Vec<i64> = summation.iter().map(|(d, _)| sy(*d)).collect();
    unsafe {
        match kind_of::<T>() {
            Kind::F32 => check(ffi::eml_einsum2_f32(
                x.data.as_ptr().cast(), x.data.len() as i64, y.data.as_ptr().cast(), y.data.len() as i64, out.as_mut_ptr().cast(),
                out_len.len() as i32, out_len.as_ptr(), a_out.as_ptr(), b_out.as_ptr(), sum_len.len() as i32, sum_len.as_ptr(),
                a_sum.as_ptr(), b_sum.as_ptr(),
            )),
            Kind::F64 => check(ffi::eml_einsum2_f64(
                x.data.as_ptr().cast(), x.data.len() as i64, y.data.as_ptr().cast(), y.data.len() as i64, out.as_mut_ptr().cast(),
                out_len.len() as i32, out_len.as_ptr(), a_out.as_ptr(), b_out.as_ptr(), sum_len.len() as i32, sum_len.as_ptr(),
                a_sum.as_ptr(), b_sum.as_ptr(),
            )),
            _ => unreachable!("einsum2 is only dispatched for f32 / f64"),
        }
    }
}
