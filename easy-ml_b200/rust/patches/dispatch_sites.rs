// The reference's private entry functions with their `#[cfg(feature = "cuda")]` branches, taken only when the element
// type is f32 / f64 (`T::CUDA_KIND`, patches/numeric.rs).  Everything the reference validates stays where it is, with
// the same messages, BEFORE the call.  Each function below is the reference's function in full: the lines between the
// `cuda` markers are the insertion, the rest is unchanged reference code (elided bodies are marked `unchanged`).

// ===== 1. src/matrices/operations.rs:165-196 =====================================================================
#[track_caller]
#[inline]
fn matrix_view_multiplication<T, S1, S2>(left: &S1, right: &S2) -> Matrix<T>
where
    T: Numeric,
    for<'a> &'a T: NumericRef<T>,
    S1: MatrixRef<T> + NoInteriorMutability,
    S2: MatrixRef<T> + NoInteriorMutability,
{
    use crate::tensors::operations::scalar_product;
    assert!(
        left.view_columns() == right.view_rows(),
        "Mismatched Matrices, left is {}x{}, right is {}x{}, * is only defined for MxN * NxL",
        left.view_rows(),
        left.view_columns(),
        right.view_rows(),
        right.view_columns()
    );
    // ---- cuda ----
    #[cfg(feature = "cuda")]
    if crate::cuda::on_device::<T>() {
        let (m, k, n) = (left.view_rows(), left.view_columns(), right.view_columns());
        // Matrix<T> is a contiguous row-major Vec<T> (src/matrices/mod.rs:67-71); any other MatrixRef (ranges, masks,
        // reversed, stacked ...) is packed by walking the view -- O(MK) next to O(MNK)
        let a: Vec<T> = RowMajorReferenceIterator::from(left).cloned().collect();
        let b: Vec<T> = RowMajorReferenceIterator::from(right).cloned().collect();
        let mut c = vec![T::zero(); m * n];
        crate::cuda::gemm(&a, &b, &mut c, m, n, k);
        return Matrix::from_flat_row_major((m, n), c);
    }
    // ---- /cuda ----
    let mut result = Matrix::empty(T::zero(), (left.view_rows(), right.view_columns()));
    for ((i, j), x) in result.row_major_reference_mut_iter().with_index() {
        let left = RowReferenceIterator::from(left, i);
        let right = ColumnReferenceIterator::from(right, j);
        *x = scalar_product::<T, _, _>(left, right);
    }
    result
}

// ===== 2. src/tensors/operations.rs:474-519 ======================================================================
#[track_caller]
#[inline]
fn tensor_view_matrix_product<T, S1, S2>(left: S1, right: S2) -> Tensor<T, 2>
where
    T: Numeric,
    for<'a> &'a T: NumericRef<T>,
    S1: TensorRef<T, 2>,
    S2: TensorRef<T, 2>,
{
    let left_shape = left.view_shape();
    let right_shape = right.view_shape();
    if left_shape[1].1 != right_shape[0].1 {
        panic!(
            "Mismatched tensors, left is {:?}, right is {:?}, * is only defined for MxN * NxL dimension lengths",
            left.view_shape(),
            right.view_shape()
        );
    }
    if left_shape[0].0 == right_shape[1].0 {
        panic!(
            "Matrix multiplication of tensors with shapes left {:?} and right {:?} would \
             create duplicate dimension names as the shape {:?}. Rename one or both of the \
             dimension names in the input to prevent this. * is defined as MxN * NxL = MxL",
            left_shape,
            right_shape,
            [left_shape[0], right_shape[1]]
        )
    }
    // ---- cuda ----
    #[cfg(feature = "cuda")]
    if crate::cuda::on_device::<T>() {
        let (m, k, n) = (left_shape[0].1, left_shape[1].1, right_shape[1].1);
        let a: Vec<T> = TensorReferenceIterator::from(&left).cloned().collect();
        let b: Vec<T> = TensorReferenceIterator::from(&right).cloned().collect();
        let mut c = vec![T::zero(); m * n];
        crate::cuda::gemm(&a, &b, &mut c, m, n, k);
        return Tensor::from([left_shape[0], right_shape[1]], c);
    }
    // ---- /cuda ----
    let mut tensor = Tensor::empty([left.view_shape()[0], right.view_shape()[1]], T::zero());
    for ([i, j], x) in tensor.iter_reference_mut().with_index() {
        let left = TensorIndex::from(&left, [(left.view_shape()[0].0, i)]);
        let right = TensorIndex::from(&right, [(right.view_shape()[1].0, j)]);
        *x = scalar_product::<T, _, _>(
            TensorReferenceIterator::from(&left),
            TensorReferenceIterator::from(&right),
        )
    }
    tensor
}

// ===== 3. src/tensors/einsum.rs:908-980  Einsum2::to =============================================================
impl<T, S1, S2, const D1: usize, const D2: usize> Einsum2<T, S1, S2, D1, D2> {
    pub fn to<const O: usize>(
        self,
        output: [Dimension; O],
    ) -> Result<Tensor<T, O>, InconsistentDimensionLengthError<2>>
    where
        T: Numeric,
        for<'a> &'a T: NumericRef<T>,
        S1: TensorRef<T, D1>,
        S2: TensorRef<T, D2>,
    {
        let input_1_shape_const = &self.tensor_1.shape();
        let input_1_shape: &[(Dimension, usize)] = input_1_shape_const;
        let input_2_shape_const = &self.tensor_2.shape();
        let input_2_shape: &[(Dimension, usize)] = input_2_shape_const;
        let input = &[input_1_shape, input_2_shape];

        let output_shape = output_shape_for(input, &output)?;
        let mut output_tensor = Tensor::empty(output_shape, T::zero());
        let summation_dimensions = summation_dimensions(input, &output)?;
        // ---- cuda ----
        #[cfg(feature = "cuda")]
        if crate::cuda::on_device::<T>() {
            // operands as contiguous row-major buffers over their own dimension order (a Tensor is one already:
            // src/tensors/mod.rs:244-250; other TensorRefs are packed by walking the view)
            let x: Vec<T> = self.tensor_1.iter_reference().cloned().collect();
            let y: Vec<T> = self.tensor_2.iter_reference().cloned().collect();
            let strides_1 = crate::tensors::compute_strides(input_1_shape_const);
            let strides_2 = crate::tensors::compute_strides(input_2_shape_const);
            let mut out = vec![T::zero(); crate::tensors::dimensions::elements(&output_shape)];
            crate::cuda::einsum::einsum2(
                &crate::cuda::einsum::Operand { shape: input_1_shape, strides: &strides_1, data: &x },
                &crate::cuda::einsum::Operand { shape: input_2_shape, strides: &strides_2, data: &y },
                &output_shape,
                &summation_dimensions,
                &mut out,
            );
            return Ok(Tensor::from(output_shape, out));
        }
        // ---- /cuda ----
        let tensor_1_indexing = self.tensor_1.index();
        let tensor_2_indexing = self.tensor_2.index();
        for (indexes, element) in output_tensor.index_mut().iter_reference_mut().with_index() {
            /* the generic loop nest :931-975, unchanged */
        }
        Ok(output_tensor)
    }
}
// Einsum1::to and Einsum3..6::to (:829-883, :1019-1666) take the same shape with crate::cuda::einsum_n(&[&x1, &x2, ...],
// &mut out, out_len, out_strides, sum_len, sum_strides): the strided loop nest on the device, bit-exact.

// ===== 4a. src/differentiation.rs:327-331, :696-768, :811-878  the list ==========================================
#[derive(Debug)]
pub struct WengertList<T> {
    operations: RefCell<Vec<Operation<T>>>,
    // ---- cuda ----
    /// f32 / f64 lists live on the device: created on first use, `operations` then stays empty
    #[cfg(feature = "cuda")]
    device: core::cell::OnceCell<crate::cuda::DeviceTape>,
    // ---- /cuda ----
}

impl<T: Primitive + crate::numeric::ZeroOne> WengertList<T> {
    // ---- cuda ----
    #[cfg(feature = "cuda")]
    pub(crate) fn device(&self) -> Option<&crate::cuda::DeviceTape> {
        if crate::cuda::on_device::<T>() {
            Some(self.device.get_or_init(crate::cuda::DeviceTape::new::<T>))
        } else {
            None
        }
    }
    // ---- /cuda ----

    fn append_nullary(&self) -> Index {
        #[cfg(feature = "cuda")]
        if let Some(tape) = self.device() {
            return tape.append_nullary();
        }
        use std::ops::DerefMut;
        let mut borrow = self.operations.borrow_mut();
        BorrowedWengertList::new(borrow.deref_mut()).append_nullary()
    }

    fn append_unary(&self, parent: Index, derivative: T) -> Index
    where
        T: Into<f64>,
    {
        #[cfg(feature = "cuda")]
        if let Some(tape) = self.device() {
            return tape.append_unary(parent, derivative.into());
        }
        use std::ops::DerefMut;
        let mut borrow = self.operations.borrow_mut();
        BorrowedWengertList::new(borrow.deref_mut()).append_unary(parent, derivative)
    }

    fn append_binary(&self, left_parent: Index, left_derivative: T, right_parent: Index, right_derivative: T) -> Index
    where
        T: Into<f64>,
    {
        #[cfg(feature = "cuda")]
        if let Some(tape) = self.device() {
            return tape.append_binary(left_parent, left_derivative.into(), right_parent, right_derivative.into());
        }
        use std::ops::DerefMut;
        let mut borrow = self.operations.borrow_mut();
        BorrowedWengertList::new(borrow.deref_mut()).append_binary(left_parent, left_derivative, right_parent, right_derivative)
    }

    /// WengertList::clear (:674)
    pub fn clear(&self) {
        #[cfg(feature = "cuda")]
        if let Some(tape) = self.device() {
            return tape.clear();
        }
        self.operations.borrow_mut().clear();
    }
}

// ===== 4b. src/differentiation/container_record/container_operations.rs:1318-1381 ================================
#[track_caller]
fn record_tensor_matrix_multiply<'a, T, S1, S2>(
    lhs: &RecordTensor<'a, T, S1, 2>,
    rhs: &RecordTensor<'a, T, S2, 2>,
) -> RecordTensor<'a, T, Tensor<(T, Index), 2>, 2>
where
    T: Numeric + Primitive,
    for<'t> &'t T: NumericRef<T>,
    S1: TensorRef<(T, Index), 2>,
    S2: TensorRef<(T, Index), 2>,
{
    use crate::tensors::indexing::TensorReferenceIterator;
    use crate::tensors::views::TensorIndex;

    assert!(
        are_same_list(lhs.history, rhs.history),
        "Record containers must be using the same WengertList"
    );
    let left_shape = lhs.view_shape();
    let right_shape = rhs.view_shape();
    if left_shape[1].1 != right_shape[0].1 {
        panic!(
            "Mismatched record tensors, left is {:?}, right is {:?}, * is only defined for MxN * NxL dimension lengths",
            lhs.view_shape(),
            rhs.view_shape()
        );
    }
    if left_shape[0].0 == right_shape[1].0 {
        panic!(
            "Matrix multiplication of record tensors with shapes left {:?} and right {:?} would \
             create duplicate dimension names as the shape {:?}. Rename one or both of the \
             dimension names in the input to prevent this. * is defined as MxN * NxL = MxL",
            left_shape,
            right_shape,
            [left_shape[0], right_shape[1]]
        )
    }
    let history = match (lhs.history, rhs.history) {
        (None, None) => None,
        (Some(history), _) => Some(history),
        (_, Some(history)) => Some(history),
    };
    // ---- cuda ----
    #[cfg(feature = "cuda")]
    if crate::cuda::on_device::<T>() {
        // the list that owns the device tape: the operands' history, else (two constants) a scratch list
        let scratch;
        let list = match history {
            Some(list) => list,
            None => {
                scratch = WengertList::new();
                &scratch
            }
        };
        let tape = list.device().expect("f32 / f64 lists are device-backed under the `cuda` feature");
        let a: Vec<(T, Index)> = TensorReferenceIterator::from(lhs).cloned().collect();
        let b: Vec<(T, Index)> = TensorReferenceIterator::from(rhs).cloned().collect();
        let c = crate::cuda::record::matmul(
            tape,
            &crate::cuda::record::Pairs { pairs: &a, rows: left_shape[0].1, cols: left_shape[1].1, has_history: lhs.history.is_some() },
            &crate::cuda::record::Pairs { pairs: &b, rows: right_shape[0].1, cols: right_shape[1].1, has_history: rhs.history.is_some() },
        );
        let tensor = Tensor::from([left_shape[0], right_shape[1]], c);
        return RecordTensor::from_existing(history, TensorView::from(tensor));
    }
    // ---- /cuda ----
    let mut tensor = Tensor::empty([lhs.view_shape()[0], rhs.view_shape()[1]], (T::zero(), 0));
    for ([i, j], x) in tensor.iter_reference_mut().with_index() {
        let left = TensorIndex::from(&lhs, [(lhs.view_shape()[0].0, i)]);
        let right = TensorIndex::from(&rhs, [(rhs.view_shape()[1].0, j)]);
        *x = record_scalar_product::<T, _, _>(
            TensorReferenceIterator::from(&left),
            TensorReferenceIterator::from(&right),
            history,
        )
    }
    RecordTensor::from_existing(history, TensorView::from(tensor))
}
// record_matrix_matrix_multiply (:1494-1531) is the same insertion after its asserts (:1500-1513), with
// RowMajorReferenceIterator::from(lhs) / (rhs), view_rows() / view_columns() for the extents, and
// RecordMatrix::from_existing(history, MatrixView::from(Matrix::from_flat_row_major((m, n), c))) for the result.
// The operator front-ends of RecordContainer::unary / binary with a rule of src/differentiation/functions.rs (Neg, Sin,
// Cos, Exp, Ln, Sqrt, +-*/ scalar, sub_swapped, div_swapped, Pow; +, -, elementwise_multiply, elementwise_divide:
// container_operations.rs:832-1261, :1644-1893, mod.rs:1224-1247) call crate::cuda::record::unary(tape, ffi::EML_OP_*, c,
// pairs) / binary(tape, ffi::EML_OP_*, x, y) the same way; closure-taking unary / binary / map stay generic.

// ===== 5. src/differentiation.rs:623-648  Record::try_derivatives ================================================
impl<'a, T: Numeric + Primitive> Record<'a, T>
where
    for<'t> &'t T: NumericRef<T>,
{
    pub fn try_derivatives(&self) -> Option<Derivatives<T>> {
        let history = self.history?;
        // ---- cuda ----
        #[cfg(feature = "cuda")]
        if let Some(tape) = history.device() {
            // the whole vector, as the reference returns it (Vec::from(Derivatives) exposes every entry); entries
            // inside a matrix product are materialised by the library
            return Some(Derivatives { derivatives: crate::cuda::record::derivatives::<T>(tape, self.index) });
        }
        // ---- /cuda ----
        let operations = history.operations.borrow();
        let mut derivatives = vec![T::zero(); operations.len()];
        derivatives[self.index] = T::one();
        for i in (0..operations.len()).rev() {
            let operation = operations[i].clone();
            let derivative = derivatives[i].clone();
            derivatives[operation.left_parent] = derivatives[operation.left_parent].clone()
                + derivative.clone() * operation.left_derivative;
            derivatives[operation.right_parent] = derivatives[operation.right_parent].clone()
                + derivative * operation.right_derivative;
        }
        Some(Derivatives { derivatives })
    }
}
// RecordContainer::derivatives_for (container_record/mod.rs:1201-1213) builds the element's Record and calls the above.
// For tapes too long to materialise (a tracked 8192 x 784 x 512 product stands for 6.7e9 entries) the device keeps the
// derivatives: crate::cuda::DeviceTape::derivatives(handle, row, col) + DeviceDerivatives::at / at_container read only
// what is asked for (Derivatives::at_tensor, container_record/mod.rs:1292).
