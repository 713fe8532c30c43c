/*
 * oracle.h -- CPU restatement of Easy ML's dense-contraction hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.
 *
 * PARITY PINNING: the reference is Rust and there is no Rust toolchain in this image,
 * so oracle/_ref (a build of the real reference) cannot exist.  The restatement is pinned
 * against every golden vector the reference's own tests hold for this path
 * (tests/golden/reference_vectors.json, replayed by tests/test_oracle_golden.py).
 * Large-shape parity has no reference fixture: "pinned only by restatement".
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Built with `g++ -O2 -ffp-contract=off` (no FMA contraction, no fast-math), so the
 * IEEE operation sequence is the one rustc emits for the reference loops.
 */
#ifndef EASYML_ORACLE_H
#define EASYML_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_ERR_PANIC (-1)        /* the reference would panic!(); text in orc_last_error() */
#define ORC_ERR_INCONSISTENT (-2) /* Einsum: Err(InconsistentDimensionLengthError) */

const char* orc_last_error(void);

/* A1: scalar_product, src/tensors/operations.rs:429-446 (strided iterators) */
float orc_scalar_product_f32(const float* l, int64_t ls, const float* r, int64_t rs, int64_t k);
double orc_scalar_product_f64(const double* l, int64_t ls, const double* r, int64_t rs, int64_t k);

/* A2: matrix_view_multiplication, src/matrices/operations.rs:165-196 (row-major operands).
 * rows [row0,row1) of C only (row0=0,row1=a_rows for the whole product); threads>1 splits
 * the rows over std::threads (rows are independent, results stay bit-identical). */
int orc_matrix_mul_f32(const float* a, int64_t a_rows, int64_t a_cols, const float* b, int64_t b_rows,
                       int64_t b_cols, float* c, int64_t row0, int64_t row1, int threads);
int orc_matrix_mul_f64(const double* a, int64_t a_rows, int64_t a_cols, const double* b, int64_t b_rows,
                       int64_t b_cols, double* c, int64_t row0, int64_t row1, int threads);

/* Bounded sample of A2 for CPU-baseline timing: c (dense (row1-row0) x (col1-col0)) = C[row0:row1, col0:col1],
 * every element by the same scalar_product walk (B keeps its full row-major width b_cols). */
void orc_matrix_mul_block_f32(const float* a, int64_t a_cols, const float* b, int64_t b_cols, float* c, int64_t row0,
                              int64_t row1, int64_t col0, int64_t col1, int threads);
void orc_matrix_mul_block_f64(const double* a, int64_t a_cols, const double* b, int64_t b_cols, double* c,
                              int64_t row0, int64_t row1, int64_t col0, int64_t col1, int threads);

/* A3: tensor_view_matrix_product, src/tensors/operations.rs:474-519.  Names are C strings.
 * out_names[2] receives pointers to the (borrowed) output dimension names. */
int orc_tensor_mul_f32(const float* a, const char* const* a_names, const int64_t* a_len, const float* b,
                       const char* const* b_names, const int64_t* b_len, float* c, const char** out_names);
int orc_tensor_mul_f64(const double* a, const char* const* a_names, const int64_t* a_len, const double* b,
                       const char* const* b_names, const int64_t* b_len, double* c, const char** out_names);

/* A4: Einsum2::to, src/tensors/einsum.rs:908-980 (+ :508-706).  Inputs are contiguous row-major
 * tensors of rank da/db.  out may be NULL to query out_len only.  On ORC_ERR_INCONSISTENT,
 * err_dim receives the offending dimension name and err_len[2] the per-input lengths (-1 = None). */
int orc_einsum2_f32(const float* a, int da, const char* const* a_names, const int64_t* a_len, const float* b,
                    int db, const char* const* b_names, const int64_t* b_len, int dout,
                    const char* const* out_names, int64_t* out_len, float* out, const char** err_dim,
                    int64_t* err_len);
int orc_einsum2_f64(const double* a, int da, const char* const* a_names, const int64_t* a_len, const double* b,
                    int db, const char* const* b_names, const int64_t* b_len, int dout,
                    const char* const* out_names, int64_t* out_len, double* out, const char** err_dim,
                    int64_t* err_len);

/* Next row: Einsum1::to and Einsum3..6::to, src/tensors/einsum.rs:829-883, :1019-1666.
 * n_ops contiguous row-major inputs; err_len has n_ops entries. */
int orc_einsum_n_f32(int n_ops, const float* const* x, const int* ranks, const char* const* const* names,
                     const int64_t* const* lens, int dout, const char* const* out_names, int64_t* out_len, float* out,
                     const char** err_dim, int64_t* err_len);
int orc_einsum_n_f64(int n_ops, const double* const* x, const int* ranks, const char* const* const* names,
                     const int64_t* const* lens, int dout, const char* const* out_names, int64_t* out_len, double* out,
                     const char** err_dim, int64_t* err_len);

/* A5: Tensor<T,1>::scalar_product, src/tensors/mod.rs:1773 -> operations.rs:1773-1808 */
int orc_vector_dot_f32(const float* a, const char* a_name, int64_t a_len, const float* b, const char* b_name,
                       int64_t b_len, float* out);
int orc_vector_dot_f64(const double* a, const char* a_name, int64_t a_len, const double* b,
                       const char* b_name, int64_t b_len, double* out);

/* Next row: covariance_column_features / covariance_row_features / covariance,
 * src/linear_algebra.rs:615-639, :676-700, :778-832.  x is rows x cols row-major; feature_axis 1 = columns are
 * features (samples = rows), 0 = rows are features.  out is F x F.  Literal restatement: sequential sums from
 * zero, mean = sum / S, sum of (x - mean_i)(y - mean_j) / S, no Bessel correction. */
void orc_covariance_f32(const float* x, int64_t rows, int64_t cols, int feature_axis, float* out);
void orc_covariance_f64(const double* x, int64_t rows, int64_t cols, int feature_axis, double* out);

/* A9: WengertList, src/differentiation.rs:327-352, 674, 696-768, 811-878 -- the literal scalar tape */
typedef struct orc_tape orc_tape; /* element type fixed at creation: 4 = f32, 8 = f64 */
orc_tape* orc_tape_new(int elem_size);
void orc_tape_free(orc_tape*);
int64_t orc_tape_len(const orc_tape*);
void orc_tape_clear(orc_tape*);                                   /* WengertList::clear :674 */
int64_t orc_tape_append_nullary(orc_tape*);                       /* :811 */
int64_t orc_tape_append_nullary_repeating(orc_tape*, int64_t n);  /* :708 */
int64_t orc_tape_append_unary(orc_tape*, int64_t parent, double d);
int64_t orc_tape_append_binary(orc_tape*, int64_t lp, double ld, int64_t rp, double rd);
/* read entry i back: parents and derivatives (as double) */
void orc_tape_get(const orc_tape*, int64_t i, int64_t* lp, int64_t* rp, double* ld, double* rd);

/* A8: Record::try_derivatives, src/differentiation.rs:623-648.  out has orc_tape_len elements of the
 * tape's element type. */
void orc_tape_derivatives(const orc_tape*, int64_t index, void* out);

/* A6+A7: record_scalar_product / record_tensor_matrix_multiply / record_matrix_matrix_multiply,
 * src/differentiation/container_record/container_operations.rs:1263-1381,1494-1531.
 * Containers are given struct-of-arrays: numbers (row-major) + tape indexes.  tape==NULL is the
 * (None, None) history case (index 0 everywhere).  Shape/name validation is the caller's (A2/A3). */
void orc_record_matmul_f32(orc_tape* tape, const float* a, const int64_t* a_idx, const float* b,
                           const int64_t* b_idx, int64_t m, int64_t k, int64_t n, float* c, int64_t* c_idx);
void orc_record_matmul_f64(orc_tape* tape, const double* a, const int64_t* a_idx, const double* b,
                           const int64_t* b_idx, int64_t m, int64_t k, int64_t n, double* c, int64_t* c_idx);
/* The scalar `Tensor<Record<T>>` path for the same product (generic Mul with T = Record:
 * src/differentiation/record_operations.rs:151-190,344-376): constants (history None, flagged by
 * *_tracked == 0) are treated as plain numbers -> append_unary.  Used to show the container quirk. */
void orc_scalar_record_matmul_f64(orc_tape* tape, const double* a, const int64_t* a_idx, int a_tracked,
                                  const double* b, const int64_t* b_idx, int b_tracked, int64_t m, int64_t k,
                                  int64_t n, double* c, int64_t* c_idx);
void orc_scalar_record_matmul_f32(orc_tape* tape, const float* a, const int64_t* a_idx, int a_tracked,
                                  const float* b, const int64_t* b_idx, int b_tracked, int64_t m, int64_t k,
                                  int64_t n, float* c, int64_t* c_idx);

/* A10: elementwise container ops.  Op codes are shared with include/easyml_b200.h (EML_OP_*).
 * unary: RecordTensor::unary, container_record/mod.rs:642-668,845-861 with the rules of
 * src/differentiation/functions.rs and container_operations.rs:1644-1804, swapped.rs:25-45.
 * binary: RecordTensor::binary, mod.rs:672-768,962-1043; mode 3 = both histories, 1 = x only, 2 = y only. */
void orc_record_unary_f32(orc_tape* tape, int op, float scalar, const float* x, const int64_t* x_idx,
                          int64_t n, float* y, int64_t* y_idx);
void orc_record_unary_f64(orc_tape* tape, int op, double scalar, const double* x, const int64_t* x_idx,
                          int64_t n, double* y, int64_t* y_idx);
void orc_record_binary_f32(orc_tape* tape, int op, int mode, const float* x, const int64_t* x_idx,
                           const float* y, const int64_t* y_idx, int64_t n, float* z, int64_t* z_idx);
void orc_record_binary_f64(orc_tape* tape, int op, int mode, const double* x, const int64_t* x_idx,
                           const double* y, const int64_t* y_idx, int64_t n, double* z, int64_t* z_idx);

#ifdef __cplusplus
}
#endif
#endif
