// The five private entry functions of the reference gain a `#[cfg(feature = "cuda")]` branch taken only for
// f32 / f64.  Everything the reference validates stays where it is, with the same messages, BEFORE the call.
// (Shown as the lines to insert; `...` marks unchanged reference code.)

// ---- 1. src/matrices/operations.rs:165-196  matrix_view_multiplication ----------------------------------------
fn matrix_view_multiplication<T, S1, S2>(left: &S1, right: &S2) -> Matrix<T> /* where ... unchanged */ {
    assert!(left.view_columns() == right.view_rows(), /* ... unchanged message :176-183 */);
    #[cfg(feature = "cuda")]
    if crate::cuda::kind_of::<T>() != crate::cuda::Kind::Generic {
        let (m, k, n) = (left.view_rows(), left.view_columns(), right.view_columns());
        // Matrix<T> is already a contiguous row-major Vec<T> (src/matrices/mod.rs:67-71); any other MatrixRef
        // (ranges, masks, reversed ...) is packed by iterating the view -- O(MK) next to O(MNK).
        let a: Vec<T> = left.row_major_reference_iter().cloned().collect();
        let b: Vec<T> = right.row_major_reference_iter().cloned().collect();
        let mut c = vec![T::zero(); m * n];
        crate::cuda::gemm(&a, &b, &mut c, m, n, k);
        return Matrix::from_flat_row_major((m, n), c);
    }
    /* ... the generic loop :185-195, unchanged, for every other T */
}

// ---- 2. src/tensors/operations.rs:474-519  tensor_view_matrix_product ------------------------------------------
fn tensor_view_matrix_product<T, S1, S2>(left: S1, right: S2) -> Tensor<T, 2> /* where ... */ {
    /* ... both panics :485-501 unchanged */
    #[cfg(feature = "cuda")]
    if crate::cuda::kind_of::<T>() != crate::cuda::Kind::Generic {
        let (m, k, n) = (left_shape[0].1, left_shape[1].1, right_shape[1].1);
        let a: Vec<T> = TensorReferenceIterator::from(&left).cloned().collect();
        let b: Vec<T> = TensorReferenceIterator::from(&right).cloned().collect();
        let mut c = vec![T::zero(); m * n];
        crate::cuda::gemm(&a, &b, &mut c, m, n, k);
        return Tensor::from([left_shape[0], right_shape[1]], c);
    }
    /* ... generic loop :506-518 */
}

// ---- 3. src/tensors/einsum.rs:908-980  Einsum2::to (and Einsum1 / Einsum3..6 likewise with einsum_n) ------------
pub fn to<const O: usize>(self, output: [Dimension; O]) -> Result<Tensor<T, O>, InconsistentDimensionLengthError<2>> {
    let output_shape = output_shape_for(input, &output)?;              // :924 unchanged (same Err values)
    let summation_dimensions = summation_dimensions(input, &output)?;  // :927 unchanged
    #[cfg(feature = "cuda")]
    if crate::cuda::kind_of::<T>() != crate::cuda::Kind::Generic {
        // classify every name: in both inputs and the output -> batch; input 1 + output -> M; input 2 + output -> N;
        // both inputs, not the output -> K.  If every role is a jointly contiguous group in each operand the pattern
        // is a (batched) GEMM: crate::cuda::contract(...) with the tensors' element strides (no transposes).
        // Otherwise (ij,jk->ijk ; ij,kl->il ; a dimension summed on one side only ...): crate::cuda::einsum_n(...),
        // the strided loop nest in this function's own order, bit-exact.
        // (easy-ml_b200/host.py::_Einsum2.to and host/easyml.hpp::Einsum2::to are this classification, tested
        //  against tests/einsum.rs.)
    }
    /* ... generic loop nest :931-975 */
}

// ---- 4. src/differentiation/container_record/container_operations.rs:1318-1381, :1494-1531 ----------------------
//      record_tensor_matrix_multiply / record_matrix_matrix_multiply
// Under `cuda`, WengertList<f32|f64> lazily owns a crate::cuda::DeviceTape and a RecordContainer created through
// it keeps an EmlVal handle beside its shape:
//   asserts :1331-1355 / :1506-1513 unchanged  ->  let c = tape.matmul(left.handle, right.handle);
// RecordContainer::{variables, constants, reset, unary, binary, elementwise_multiply, elementwise_divide} map to
// DeviceTape::{variables, constants, reset, unary(EML_OP_*), binary(EML_OP_*)}; the (T, Index) view is
// DeviceTape::{numbers, indexes} (indexes are closed-form and bit-exact).  Closures over Record (`map`, `map_mut`)
// download numbers + indexes and run on the CPU as today.

// ---- 5. src/differentiation.rs:623-648  Record::try_derivatives --------------------------------------------------
// record_tensor.derivatives_for(indexes) / an element's Record::derivatives():
//   let d = tape.derivatives(container.handle, row, col);       // "Record has no WengertList ..." stays in Rust
//   Derivatives::at(&record)  -> d.at(record.index)
//   Derivatives::at_tensor(x) -> d.at_container(x.handle, x.elements())
