/*
 * easyml_b200.h -- C ABI of the B200-native dense-contraction path for Easy ML (f32 / f64).
 *
 * This is the drop-in boundary.  The reference (Skeletonxf/easy-ml) has no FFI:
 * its "interface" for this path is five private Rust functions behind `std::ops::Mul` / `Einsum` /
 * `RecordTensor`.  Behind a `cuda` Cargo feature each of them validates exactly as today and then calls
 * one entry point below (the binding a maintainer would add is shown in INTEGRATION.md).  Every entry
 * point cites the reference code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *  - plain pointers and sizes, no C++/torch types; all extents and strides are int64_t ELEMENTS.
 *  - every function returns int: 0 = EML_OK, negative = EML_ERR_*; nothing throws or aborts across the
 *    boundary.  eml_last_error() returns a thread-local, library-owned message valid until the next call
 *    on that thread.
 *  - no CPU fallback: with no CUDA device every compute entry point returns EML_ERR_NO_DEVICE.
 *  - deterministic: fixed tile->CTA map, fixed k order, no atomics, fixed-order split-K; results are
 *    bit-identical run to run on the same GPU model.
 *  - host entry points (no suffix) take HOST pointers, stage through pinned memory and block until the
 *    result is in the caller's buffer.  `_dev` entry points take DEVICE pointers plus a cudaStream_t
 *    (as void*; NULL = CUDA's default stream, exactly as in the CUDA runtime) and are asynchronous.
 *  - the device used is the calling thread's current CUDA device (cudaSetDevice / torch.cuda.set_device);
 *    state is cached per device; entry points are re-entrant across threads.
 */
#ifndef EASYML_B200_H
#define EASYML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EML_OK 0
#define EML_ERR_NO_DEVICE (-1)
#define EML_ERR_BAD_ARG (-2)
#define EML_ERR_OOM (-3)
#define EML_ERR_CUDA (-4)
#define EML_ERR_NCCL (-5)
#define EML_ERR_STATE (-6)

#define EML_F32 0
#define EML_F64 1

/* Elementwise container ops (A10).  Values and derivative rules are those of
 * src/differentiation/functions.rs; the unary list is the set of `RecordTensor::unary` callers in
 * src/differentiation/container_record/container_operations.rs:1644-1804 and
 * container_operations/swapped.rs:25-45 (c = the scalar operand). */
#define EML_OP_NEG 0    /* -x            d = -1                 functions.rs:143-161 */
#define EML_OP_SIN 1    /* sin x         d = cos x              :163-181 */
#define EML_OP_COS 2    /* cos x         d = -sin x             :183-201 */
#define EML_OP_EXP 3    /* e^x           d = e^x                :203-221 */
#define EML_OP_LN 4     /* ln x          d = 1/x                :223-241 */
#define EML_OP_SQRT 5   /* sqrt x        d = 1/((1+1) sqrt x)   :243-261 */
#define EML_OP_ADD_C 6  /* x + c         d = 1                  :18-41  (d/dx) */
#define EML_OP_SUB_C 7  /* x - c         d = 1                  :43-66  (d/dx) */
#define EML_OP_MUL_C 8  /* x * c         d = c                  :68-91  (d/dx) */
#define EML_OP_DIV_C 9  /* x / c         d = 1/c                :93-116 (d/dx) */
#define EML_OP_POW_C 10 /* x ^ c         d = c x^(c-1)          :118-141 (d/dx) */
#define EML_OP_C_SUB 11 /* c - x         d = -1                 sub_swapped (d/dy) */
#define EML_OP_C_DIV 12 /* c / x         d = -c/(x x)           div_swapped (d/dy) */
#define EML_OP_C_POW 13 /* c ^ x         d = c^x ln c           pow swapped (d/dy) */
#define EML_OP_ADD 32   /* x + y         dx = 1     dy = 1 */
#define EML_OP_SUB 33   /* x - y         dx = 1     dy = -1 */
#define EML_OP_MUL 34   /* x * y         dx = y     dy = x        (elementwise_multiply, mod.rs:1224) */
#define EML_OP_DIV 35   /* x / y         dx = 1/y   dy = -x/(y y) (elementwise_divide,   mod.rs:1247) */
#define EML_OP_POW 36   /* x ^ y         dx = y x^(y-1)  dy = x^y ln x */

/* ---------------------------------------------------------------------------------------------
 * Library lifetime and errors
 * ------------------------------------------------------------------------------------------- */
/* Counts CUDA devices.  Returns EML_ERR_NO_DEVICE (and *n_devices = 0) when there is none: the f32/f64
 * path has no CPU fallback, the Rust shim turns this into a panic. */
int eml_init(int* n_devices);
void eml_shutdown(void);
const char* eml_last_error(void);
int eml_version(void);
/* How many kernels this library has launched so far in this process (all threads, all devices). */
int64_t eml_kernel_launches(void);
/* Name of the kernel family the last GEMM/contract call on this thread dispatched to
 * ("dmma_f64", "tf32x3_tcgen05", "simt", "einsum_generic", ...).  Library-owned string. */
const char* eml_last_kernel(void);

/* ---------------------------------------------------------------------------------------------
 * Tier 1 -- host-pointer drop-ins
 * ------------------------------------------------------------------------------------------- */
/* C[MxN] = A[MxK] * B[KxN], row-major with leading dimensions (elements).
 * Replaces matrix_view_multiplication, src/matrices/operations.rs:165-196 (the 16 Matrix/MatrixView Mul
 * impls), and tensor_view_matrix_product, src/tensors/operations.rs:474-519 (the 16 rank-2
 * Tensor/TensorView Mul impls); per element it replaces scalar_product, src/tensors/operations.rs:429-446.
 * Shape/name validation (operations.rs:176-183, :485-501) stays in the caller. */
int eml_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldc);
int eml_gemm_f64(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldc);

/* The same product over the first n_gpus GPUs of this box, one call from one host thread (BASELINE config 5: f64
 * 32768 x 32768 on 2 / 4 / 8 B200).  Replaces, for large operands, matrix_view_multiplication
 * src/matrices/operations.rs:165-196: rows of C are independent (:187-194), so A and C are cut into contiguous row
 * slabs, one per GPU, and B is needed everywhere.  Every GPU uploads its own slab of A and 1 / n_gpus of B over its own
 * PCIe link; the pieces of B are exchanged GPU to GPU over NVLink (copy engines) in K panels, each panel's product
 * starting as soon as its pieces are in; finished row panels of C are downloaded while the next one is multiplied.  No
 * cross-GPU reduction: the result carries the bits of eml_gemm_* on one GPU.  Host pointers (pageable or page-locked),
 * blocking.  EML_ERR_BAD_ARG when n_gpus exceeds the number of CUDA devices. */
int eml_gemm_f64_sharded(int n_gpus, const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                         int64_t lda, int64_t ldb, int64_t ldc);
int eml_gemm_f32_sharded(int n_gpus, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                         int64_t lda, int64_t ldb, int64_t ldc);

/* Strided-batched contraction  C[b,m,n] = sum_k A[b,m,k] * B[b,k,n].
 * sA = element strides of A for (batch, m, k); sB for (batch, k, n); sC for (batch, m, n).
 * Replaces Einsum2::to, src/tensors/einsum.rs:908-980, for every name pattern that classifies into
 * batch / M / N / K roles (each role may be a product of several named dimensions only when they are
 * jointly contiguous); the classification itself (output_shape_for :563, summation_dimensions :579)
 * stays in the caller.  Buffers are host pointers spanning the strided extents. */
int eml_contract_f32(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                     const int64_t sA[3], const int64_t sB[3], const int64_t sC[3]);
int eml_contract_f64(const double* A, const double* B, double* C, int64_t batch, int64_t M, int64_t N,
                     int64_t K, const int64_t sA[3], const int64_t sB[3], const int64_t sC[3]);

/* General two-operand Einstein summation for the patterns that are NOT (batched) GEMMs
 * (e.g. ij,jk->ijk ; ij,kl->il ; ab,ab-> ; a,b->ab):
 *   out[o] = sum over s of A[o . a_out_stride + s . a_sum_stride] * B[o . b_out_stride + s . b_sum_stride]
 * out is contiguous row-major over out_len[0..n_out); a stride of 0 means "this operand does not carry
 * that dimension".  The sum is seeded with zero and runs in the reference's order (summation dimensions in
 * order of first appearance, last fastest; src/tensors/einsum.rs:932-972, src/tensors/indexing.rs:876-933)
 * with separately rounded multiply and add, so this entry point is BIT-EXACT with the reference.
 * n_out <= 6, n_sum <= 12 (two rank-6 inputs).  Host pointers; a_elems / b_elems = buffer lengths. */
int eml_einsum2_f32(const float* A, int64_t a_elems, const float* B, int64_t b_elems, float* out, int n_out,
                    const int64_t* out_len, const int64_t* a_out_stride, const int64_t* b_out_stride, int n_sum,
                    const int64_t* sum_len, const int64_t* a_sum_stride, const int64_t* b_sum_stride);
int eml_einsum2_f64(const double* A, int64_t a_elems, const double* B, int64_t b_elems, double* out, int n_out,
                    const int64_t* out_len, const int64_t* a_out_stride, const int64_t* b_out_stride, int n_sum,
                    const int64_t* sum_len, const int64_t* a_sum_stride, const int64_t* b_sum_stride);

/* N-operand Einstein summation, N = 1..6: Einsum1::to and Einsum3..6::to,
 * src/tensors/einsum.rs:829-883 and :1019-1666 (transposes, sums and traces of one tensor; chains of three to six):
 *   out[o] = sum over s of ((x1 * x2) * x3) ...   each operand read at  o . out_strides[i] + s . sum_strides[i]
 * out is contiguous row-major over out_len[0..n_out); out_strides is [n_ops][n_out] and sum_strides [n_ops][n_sum]
 * (row-major; stride 0 = the operand does not carry that dimension).  Zero-seeded sum, summation dimensions in
 * the order given (the reference's: first appearance, last fastest), left-associated product, separately
 * rounded multiplies and adds: BIT-EXACT with the reference.  Host pointers; elems[i] = length of x[i]. */
int eml_einsum_n_f32(int n_ops, const float* const* x, const int64_t* elems, float* out, int n_out,
                     const int64_t* out_len, const int64_t* out_strides, int n_sum, const int64_t* sum_len,
                     const int64_t* sum_strides);
int eml_einsum_n_f64(int n_ops, const double* const* x, const int64_t* elems, double* out, int n_out,
                     const int64_t* out_len, const int64_t* out_strides, int n_sum, const int64_t* sum_len,
                     const int64_t* sum_strides);

/* Dot product of two strided vectors (n >= 1).  Replaces Tensor<T,1>::scalar_product,
 * src/tensors/mod.rs:1773 -> src/tensors/operations.rs:1773-1808. */
int eml_dot_f32(const float* x, int64_t incx, const float* y, int64_t incy, int64_t n, float* out);
int eml_dot_f64(const double* x, int64_t incx, const double* y, int64_t incy, int64_t n, double* out);

/* Covariance of features (the first "next" row: a SYRK-shaped caller of the GEMM):
 *   cov[i][j] = (1/S) sum_s (x[s,i] - mean_i)(x[s,j] - mean_j),  mean_i = (sum_s x[s,i]) / S,  no Bessel correction.
 * Replaces covariance_column_features (src/linear_algebra.rs:615-639; feature_axis = 1, S = rows),
 * covariance_row_features (:676-700; feature_axis = 0, S = cols) and covariance (:778-832; the caller resolves the
 * feature dimension name to an axis and keeps the reference's panic for an unknown name).  The reference
 * recomputes both means for every (i, j) -- O(F^2 S) extra; here: column/row sums, one centring pass, one GEMM
 * (X_c^T X_c or X_c X_c^T through strides), one scaling pass.  X: rows x cols row-major, ld = ldx; out: F x F. */
int eml_covariance_f32(const float* X, int64_t rows, int64_t cols, int64_t ldx, int feature_axis, float* out);
int eml_covariance_f64(const double* X, int64_t rows, int64_t cols, int64_t ldx, int feature_axis, double* out);
/* device-resident: X dense (ld = cols) */
int eml_covariance_f32_dev(const float* X, int64_t rows, int64_t cols, int feature_axis, float* out, void* stream);
int eml_covariance_f64_dev(const double* X, int64_t rows, int64_t cols, int feature_axis, double* out, void* stream);

/* Bit-exact GEMM: same contract as eml_gemm_*, but every C[i,j] is the reference's exact operation
 * sequence (first product, then k-ascending separately rounded multiply and add -- scalar_product,
 * src/tensors/operations.rs:441-445).  Slow (no tensor cores, no FMA); exists so parity can be shown
 * bit-for-bit against the oracle at sizes the fast kernels are only checked within tolerance. */
int eml_gemm_exact_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc);
int eml_gemm_exact_f64(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                       int64_t lda, int64_t ldb, int64_t ldc);

/* ---------------------------------------------------------------------------------------------
 * Tier 2 -- device-resident variants (benchmarks, keeping an NN step on the GPU)
 * ------------------------------------------------------------------------------------------- */
int eml_alloc(void** dptr, size_t bytes);
int eml_free(void* dptr);
/* Stream-ordered device memory from the library's per-(device, stream, size) block cache (the one the tape's
 * containers and all scratch use): a block may be reused by work enqueued LATER on the SAME stream without any
 * synchronisation, and in steady state no CUDA allocator call is made.  For device-resident containers whose
 * intermediates come and go with every operator (easy-ml_b200/device.py).  `bytes` at free = `bytes` at alloc. */
int eml_alloc_async(void** dptr, size_t bytes, void* stream);
int eml_free_async(void* dptr, size_t bytes, void* stream);
int eml_upload(void* dst_dev, const void* src_host, size_t bytes);   /* blocking */
int eml_download(void* dst_host, const void* src_dev, size_t bytes); /* blocking */
int eml_sync(void);                                                  /* current device */
/* Page-locked host memory.  The host entry points accept any host pointer; with pinned buffers the large
 * GEMMs overlap their PCIe copies with the compute (three streams), which pageable memory cannot do.  A Rust
 * shim would back `Matrix<f32|f64>` storage created under the `cuda` feature with these. */
/* Measured FP64 tensor-pipe (DMMA) peak of the current device at its current clocks, in TFLOP/s: a
 * register-resident DMMA loop with no memory traffic (~0.1 s).  The roofline denominator of the f64 GEMM. */
int eml_measure_fp64_tensor_peak(double* tflops);
int eml_host_alloc(void** hptr, size_t bytes);
int eml_host_free(void* hptr);

/* accumulate != 0:  C += A*B (the fused "accumulate into the derivative" epilogue of the backward). */
int eml_contract_f32_dev(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N,
                         int64_t K, const int64_t sA[3], const int64_t sB[3], const int64_t sC[3],
                         int accumulate, void* stream);
int eml_contract_f64_dev(const double* A, const double* B, double* C, int64_t batch, int64_t M, int64_t N,
                         int64_t K, const int64_t sA[3], const int64_t sB[3], const int64_t sC[3],
                         int accumulate, void* stream);
int eml_gemm_f32_dev(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldb, int64_t ldc, void* stream);
int eml_gemm_f64_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                     int64_t ldb, int64_t ldc, void* stream);
int eml_gemm_exact_f32_dev(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                           int64_t lda, int64_t ldb, int64_t ldc, void* stream);
int eml_gemm_exact_f64_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                           int64_t lda, int64_t ldb, int64_t ldc, void* stream);

/* Row-sharded GEMM with the all-gather fused into the epilogue (BASELINE config 5).
 * C (+)= A*B exactly as eml_gemm_f64_dev, and every finished tile is ALSO stored -- same offsets, same ldc --
 * into n_remote other buffers, normally the same rows of C in the peer GPUs' memory (pointers obtained with
 * eml_ipc_open; stores travel over NVLink from inside the GEMM kernel, no separate collective).  The caller
 * orders the peers afterwards (any stream-ordered barrier, e.g. a 1-element NCCL all-reduce).
 * Replaces, for one row block, matrix_view_multiplication src/matrices/operations.rs:165-196. */
int eml_gemm_f64_scatter_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                             int64_t lda, int64_t ldb, int64_t ldc, int accumulate, double* const* c_remote,
                             int n_remote, void* stream);
/* CUDA IPC plumbing for the peer-mapped buffers (one process per GPU): export a buffer obtained from
 * eml_alloc as a 64-byte handle, open a peer's handle in this process (peer access is enabled on demand). */
int eml_ipc_export(void* dptr, void* handle64);
int eml_ipc_open(const void* handle64, void** dptr);
int eml_ipc_close(void* dptr);
/* Asynchronous copy between any two device-visible buffers (local, or a peer's opened with eml_ipc_open:
 * such copies run on the copy engines over NVLink and take no SM from a concurrent GEMM). */
int eml_copy_async(void* dst, const void* src, size_t bytes, void* stream);

/* Backward of C = A*B for reverse-mode autodiff:  dA += dC * B^T  and/or  dB += A^T * dC  (either
 * output may be NULL).  This is what the reverse sweep Record::try_derivatives,
 * src/differentiation.rs:623-648, computes over the M*N*(2K-1) tape entries that
 * record_scalar_product, container_operations.rs:1263-1315, appends for one matmul.  Row-major, dense. */
int eml_matmul_bwd_f32_dev(const float* dC, const float* A, const float* B, float* dA, float* dB, int64_t M,
                           int64_t N, int64_t K, void* stream);
int eml_matmul_bwd_f64_dev(const double* dC, const double* A, const double* B, double* dA, double* dB,
                           int64_t M, int64_t N, int64_t K, void* stream);

/* Elementwise forward:  y[i] = f(x[i], c) and dydx[i] = f'(x[i], c)  (dydx may be NULL) -- the value and
 * the tape derivative of RecordTensor::unary, container_record/mod.rs:642-668, in one pass. */
int eml_unary_f32_dev(int op, float c, const float* x, float* y, float* dydx, int64_t n, void* stream);
int eml_unary_f64_dev(int op, double c, const double* x, double* y, double* dydx, int64_t n, void* stream);
/* z = f(x,y), dzdx, dzdy (either may be NULL): RecordTensor::binary, mod.rs:672-768. */
int eml_binary_f32_dev(int op, const float* x, const float* y, float* z, float* dzdx, float* dzdy, int64_t n,
                       void* stream);
int eml_binary_f64_dev(int op, const double* x, const double* y, double* z, double* dzdx, double* dzdy,
                       int64_t n, void* stream);
/* Backward chain rule of one elementwise tape segment:  dparent[i] += dout[i] * deriv[i]
 * (src/differentiation.rs:641-644 restricted to the segment). */
int eml_chain_f32_dev(const float* dout, const float* deriv, float* dparent, int64_t n, void* stream);
int eml_chain_f64_dev(const double* dout, const double* deriv, double* dparent, int64_t n, void* stream);
/* w[i] -= lr * d[i]: the weight update `w.map_mut(|x| x - derivatives[&x] * lr)`,
 * src/neural_networks.rs:418-420, as one kernel. */
int eml_sgd_f32_dev(float* w, const float* d, float lr, int64_t n, void* stream);
int eml_sgd_f64_dev(double* w, const double* d, double lr, int64_t n, void* stream);
/* C = A*B (row-major f64, device pointers), the product of matrix_view_multiplication (src/matrices/operations.rs:165-196),
 * computed beyond the FP64 pipe: both operands are cut into `slices` (6..10; 0 = as few as the guard can prove, at
 * least 7, at most 8) signed 7-bit slices relative to
 * their row / column maxima (exact), the slice products run as int8 GEMMs on the INT8 tensor cores with exact int32
 * sums, and the diagonals are recombined in f64.  A guard pass first PROVES, for these operands, that the truncation
 * stays within the path's contract |C - ref| <= 1e-12 sum_k |a_ik b_kj| for every element; when it cannot (wide dynamic
 * range inside a row / column, Inf / NaN, extreme exponents, K too long for int32) the DMMA kernel runs instead.
 * *path (may be NULL): the number of slices the sliced kernel used, 0 = DMMA kernel.  Opt-in: eml_gemm_f64* never
 * take this route. */
int eml_gemm_f64_sliced_dev(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                            int64_t ldb, int64_t ldc, int slices, int* path, void* stream);
/* Tensor::reorder / Tensor::transpose (src/tensors/mod.rs:1473-1600, data movement of TensorAccess::iter) and
 * Matrix::transpose (src/matrices/mod.rs:1277-1330) on device buffers:
 *   out (row-major, lengths out_lens[rank]) [i_0 .. i_{rank-1}] = in[sum_d i_d * in_strides[d]]
 * rank 1..6; strides in elements, >= 0.  Coalesced on both sides (32 x 32 shared-memory tiles) whenever the input
 * has a unit-stride dimension.  `out` must not overlap `in`. */
int eml_reorder_f32_dev(const float* in, float* out, int rank, const int64_t* out_lens, const int64_t* in_strides,
                        void* stream);
int eml_reorder_f64_dev(const double* in, double* out, int rank, const int64_t* out_lens, const int64_t* in_strides,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tier 3 -- device-resident reverse-mode tape (RecordTensor / RecordMatrix / WengertList)
 *
 * One eml_tape per WengertList<T> (src/differentiation.rs:327-331).  Containers are rank-2 (rows x cols,
 * row-major); rank and dimension names are caller-side metadata.  The tape never materialises the scalar
 * entries: a matmul is ONE macro node, yet every observable of the reference is reproduced --
 *   - tape length (== WengertList.operations.len(), grows by M*N*(2K-1) per tracked matmul),
 *   - every element's Index (variables: start+i, src/differentiation.rs:708-727; matmul outputs:
 *     base + (i*N+j)*(2K-1) + 2K-2, K==1: base + i*N+j; unary/binary outputs: base+i),
 *   - derivative values of the reverse sweep (src/differentiation.rs:623-648), including the
 *     reference's constant-operand behaviour: a constant container carries Index 0
 *     (container_record/mod.rs:120-126) and record_scalar_product records it as a parent
 *     (container_operations.rs:1291-1296), so its derivative sum lands in derivatives[0].
 * ------------------------------------------------------------------------------------------- */
typedef struct eml_tape eml_tape;
typedef struct eml_derivs eml_derivs;
typedef int64_t eml_val; /* handle of a record container living on the tape's device */

int eml_tape_create(int dtype, eml_tape** out); /* WengertList::new */
int eml_tape_destroy(eml_tape* t);
int eml_tape_len(eml_tape* t, int64_t* len);    /* operations.len() */
/* WengertList::clear, src/differentiation.rs:674.  Values created before the clear stay usable as
 * numbers; tracked ones must be eml_tape_reset before further tracked use (as in the reference). */
int eml_tape_clear(eml_tape* t);
/* RecordTensor::variables / RecordMatrix::variables, container_record/mod.rs:149-164,:426: appends
 * rows*cols nullary entries, Index = start + i in row-major order.  host_numbers: rows*cols elements. */
int eml_tape_variables(eml_tape* t, const void* host_numbers, int64_t rows, int64_t cols, eml_val* out);
/* RecordTensor::constants, mod.rs:115-128: no history, every Index = 0. */
int eml_tape_constants(eml_tape* t, const void* host_numbers, int64_t rows, int64_t cols, eml_val* out);
/* same, adopting a device buffer the caller already filled (copied; async on the tape's stream) */
int eml_tape_constants_dev(eml_tape* t, const void* dev_numbers, int64_t rows, int64_t cols, eml_val* out);
int eml_tape_variables_dev(eml_tape* t, const void* dev_numbers, int64_t rows, int64_t cols, eml_val* out);
/* RecordTensor::from_existing / RecordMatrix::from_existing, container_record/mod.rs:219-224, :508: a container made of
 * ANY (number, Index) pairs -- rows*cols numbers and rows*cols tape indexes, row-major -- typically the elements of
 * a range / mask / resize of other containers; duplicates allowed.  has_history != 0 is `Some(&list)`: the container
 * takes part in sweeps and the derivative of element e is added to derivatives[indexes[e]] (elements sharing an index
 * are summed in ascending element order: deterministic); every index must then be the Index of a container element
 * currently on the list.  has_history == 0 is `None`: a constant whose elements still report their indexes.  Nothing
 * is appended to the list.  The reference's examples re-wrap every row of a constants matrix this way
 * (src/neural_networks.rs:105-120: all indexes 0 with the weights' history). */
int eml_tape_from_existing(eml_tape* t, const void* host_numbers, const int64_t* host_indexes, int64_t rows, int64_t cols,
                           int has_history, eml_val* out);
/* The same for a rectangular range of a container that already lives on the device -- from_existing over
 * MatrixRange / TensorRange (src/matrices/views/ranges.rs, src/tensors/views/ranges.rs): numbers are copied on the
 * device, indexes follow the source's elements. */
int eml_tape_range(eml_tape* t, eml_val src, int64_t row0, int64_t n_rows, int64_t col0, int64_t n_cols, int has_history,
                   eml_val* out);
/* The SCALAR Record operators on the same list: WengertList::append_nullary / append_unary / append_binary, reference
 * src/differentiation.rs:811-878 (what Record::variable, and +, -, *, / ... between Records, record_operations.rs, push).
 * *index = the new entry's tape index.  Scalar entries and containers share one index space, so Records taken from
 * container elements and folded with scalar arithmetic (src/neural_networks.rs:105-131) differentiate through both.
 * The sweep walks a run of scalar entries in the reference's own order: bit-exact for purely scalar tapes. */
int eml_tape_append_nullary(eml_tape* t, int64_t* index);
int eml_tape_append_unary(eml_tape* t, int64_t parent, double derivative, int64_t* index);
int eml_tape_append_binary(eml_tape* t, int64_t left_parent, double left_derivative, int64_t right_parent,
                           double right_derivative, int64_t* index);
/* Record::derivatives / try_derivatives (src/differentiation.rs:606-648) of the Record at tape index `index`
 * (a scalar Record's index, or a container element's). */
int eml_tape_derivatives_at(eml_tape* t, int64_t index, eml_derivs** out);
/* RecordTensor::reset, mod.rs:349-365: re-append rows*cols nullary entries and renumber start + i. */
int eml_tape_reset(eml_tape* t, eml_val v);
int eml_val_release(eml_tape* t, eml_val v); /* drop a container (its node stays while referenced) */
int eml_val_shape(eml_tape* t, eml_val v, int64_t* rows, int64_t* cols, int* tracked);
int eml_val_numbers(eml_tape* t, eml_val v, void* host_out);     /* rows*cols numbers, row-major */
int eml_val_indexes(eml_tape* t, eml_val v, int64_t* host_out);  /* rows*cols tape indexes */
/* Asynchronous read-back: enqueues the copy of v's numbers into host_out (page-locked memory from eml_host_alloc)
 * on the tape's stream and returns; the data is there after eml_arrival_wait(*arrival) (or eml_tape_sync; arrival
 * may be NULL).  A training loop reads the loss of step i once step i+1 is enqueued, so the host never stalls
 * the device. */
int eml_val_numbers_async(eml_tape* t, eml_val v, void* pinned_host_out, void** arrival);
/* Blocks until the copy that produced `arrival` has landed, then releases the handle. */
int eml_arrival_wait(void* arrival);
int eml_tape_sync(eml_tape* t);
/* Borrowed pointer to v's numbers on the device: valid until v is released, and READ-ONLY -- containers are immutable
 * (operators and the weight update produce fresh ones) and a constant's split / packed TF32 planes are cached under
 * its identity, so writing through this pointer would leave later products on stale planes.  New input data for a
 * recorded step goes through eml_tape_constants_dev. */
int eml_val_device_ptr(eml_tape* t, eml_val v, const void** dev_numbers);

/* record_tensor_matrix_multiply / record_matrix_matrix_multiply,
 * container_operations.rs:1318-1381,:1494-1531 (+ the 8 Mul impls): c = a * b. */
int eml_tape_matmul(eml_tape* t, eml_val a, eml_val b, eml_val* c);
/* RecordTensor::unary and its operator front-ends (Neg, Sin, Cos, Exp, Ln, Sqrt, container (op) scalar,
 * sub_swapped, div_swapped): container_record/mod.rs:845-861, container_operations.rs:1644-1804. */
int eml_tape_unary(eml_tape* t, int op, double c, eml_val x, eml_val* y);
/* RecordTensor::binary (elementwise +, -, elementwise_multiply, elementwise_divide), mod.rs:962-1043:
 * history selection (None,None)/(Some,None)/(None,Some)/(Some,Some) as in the reference. */
int eml_tape_binary(eml_tape* t, int op, eml_val x, eml_val y, eml_val* z);

/* Record::derivatives of element (row, col) of v, src/differentiation.rs:606-648 /
 * RecordTensor::derivatives_for, mod.rs:1201.  Returns EML_ERR_STATE if v has no history
 * ("Record has no WengertList to find derivatives from"). */
int eml_tape_derivatives(eml_tape* t, eml_val v, int64_t row, int64_t col, eml_derivs** out);
int eml_derivs_free(eml_derivs* d);
int eml_derivs_len(eml_derivs* d, int64_t* len); /* == tape length at the time of the sweep */
/* Derivatives::at / Index<&Record>, src/differentiation.rs:387-403: d[index] as a double. */
int eml_derivs_at(eml_derivs* d, int64_t index, double* out);
/* Derivatives::at_tensor / at_matrix, mod.rs:1292,:1328: rows*cols values for container v. */
int eml_derivs_at_val(eml_derivs* d, eml_val v, void* host_out);
int eml_derivs_at_val_dev(eml_derivs* d, eml_val v, void* dev_out);
/* Vec::from(Derivatives), src/differentiation.rs:405-414: the whole vector (tape-length elements).
 * Entries interior to a matmul macro node are materialised on demand (they equal dC[i,j]). */
int eml_derivs_to_host(eml_derivs* d, void* host_out);
/* `w.map_mut(|x| x - derivatives[&x] * lr)`, src/neural_networks.rs:418-420, fused on the device:
 * numbers updated in place and, like the reference, one unary entry per element is appended (the
 * container's indexes move to the new entries until the next clear + reset). */
int eml_tape_sgd_update(eml_tape* t, eml_val w, eml_derivs* d, double lr);

/* ---------------------------------------------------------------------------------------------
 * Recorded steps (CUDA graphs).  A training step issues ~50 small dependent launches; recorded once, the same
 * device work is replayed as ONE graph launch.  No counterpart in the reference (it has no device); the tape
 * semantics are unchanged, the restrictions below only apply between begin and end:
 *   - run the identical step eagerly at least once first (every buffer size must already be in the cache);
 *   - no blocking host<->device call (eml_val_numbers, eml_tape_variables/constants, eml_derivs_at*, ...): they
 *     return EML_ERR_STATE; read results with eml_val_numbers_async into page-locked memory;
 *   - eml_tape_sgd_update updates the weights IN PLACE (eagerly it produces a fresh container, so nodes recorded
 *     earlier keep the old numbers; a recorded step must leave every persistent container where it found it);
 *   - release every container created inside, and end in the tape state the step started in (the usual
 *     clear() + reset() of the variables), so that a replay needs no host bookkeeping.
 * eml_graph_launch enqueues one replay on the tape's stream; *arrival (optional) completes when it has finished.
 * ------------------------------------------------------------------------------------------- */
typedef struct eml_graph eml_graph;
int eml_tape_capture_begin(eml_tape* t);
int eml_tape_capture_end(eml_tape* t, eml_graph** out);
int eml_graph_launch(eml_graph* g, void** arrival);
int eml_graph_destroy(eml_graph* g);
/* Development aid (no counterpart in the reference): what each kernel of a recorded step costs on the step's critical
 * path.  Times `replays` replays of the graph, then `replays` more per kernel node with that node switched off, and
 * writes a NUL-terminated text report (first two lines: microseconds per replay with and without the weight updates; then
 * one line per kernel, "<change in microseconds>\t<kernel name>") into report[0..capacity).  A negative number is what
 * the step would gain if that kernel cost nothing; kernels that overlap others show less than their own duration.
 * The step computes garbage while a kernel is off (weight updates are kept off meanwhile): use it on a model you are
 * about to throw away. */
int eml_graph_profile(eml_graph* g, int replays, char* report, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* EASYML_B200_H */
