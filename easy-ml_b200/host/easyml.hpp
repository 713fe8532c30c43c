// easyml.hpp -- C++ mirror of Easy ML's operator API for the f32/f64 dense-contraction path, over the C ABI
// of include/easyml_b200.h.  Header-only, C++17.
//
// The reference is a Rust crate (no Rust toolchain in this image), so this is the compiled-language host side
// of the drop-in boundary: same names, same argument meaning, same validation and the same messages as the
// reference -- a Rust `panic!` becomes `easyml::Panic`, `Err(InconsistentDimensionLengthError)` becomes the
// exception of that name.  Every check runs BEFORE the FFI call, exactly where the Rust shim keeps them
// (INTEGRATION.md section 4).  There is no CPU fallback: without a B200 every product throws
// `easyml::BackendError` carrying the library's message.
//
// Citations are file:line of the reference (Skeletonxf/easy-ml), relative to its repository root.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <cmath>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "easyml_b200.h"

namespace easyml {

struct Panic : std::runtime_error {  // the reference would panic!() with this text
    using std::runtime_error::runtime_error;
};
struct BackendError : std::runtime_error {  // non-zero code from the C ABI (no device, CUDA error ...)
    int code;
    BackendError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
using Dimension = std::string;  // `&'static str` in the reference, src/tensors/mod.rs:57
using Shape = std::vector<std::pair<Dimension, int64_t>>;

// Err(InconsistentDimensionLengthError { lengths, dimension }), src/tensors/einsum.rs:321
struct InconsistentDimensionLengthError : std::runtime_error {
    Dimension dimension;
    std::vector<std::optional<int64_t>> lengths;
    InconsistentDimensionLengthError(Dimension d, std::vector<std::optional<int64_t>> l)
        : std::runtime_error("InconsistentDimensionLengthError: dimension " + d), dimension(std::move(d)),
          lengths(std::move(l)) {}
};

namespace detail {
inline void check(int code) {
    if (code != EML_OK) throw BackendError(code, eml_last_error());
}
template <class T>
constexpr bool is_f32 = std::is_same<T, float>::value;
template <class T>
constexpr void require_float() {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value,
                  "the B200 path serves f32 and f64; other numeric types stay on the reference's generic path");
}
// Debug formatting of [(Dimension, usize); D], e.g. [("r", 2), ("c", 3)]
inline std::string shape_debug(const Shape& s) {
    std::ostringstream o;
    o << "[";
    for (size_t i = 0; i < s.size(); i++) o << (i ? ", " : "") << "(\"" << s[i].first << "\", " << s[i].second << ")";
    o << "]";
    return o.str();
}
}  // namespace detail

// ------------------------------------------------------------------------------------------------ Matrix
// easy_ml::matrices::Matrix<T>: flat row-major storage, src/matrices/mod.rs:67-71
template <class T>
class Matrix {
  public:
    Matrix(int64_t rows, int64_t cols, std::vector<T> row_major)  // from_flat_row_major, mod.rs:213
        : rows_(rows), cols_(cols), data_(std::move(row_major)) {
        detail::require_float<T>();
        if (rows < 1 || cols < 1 || (int64_t)data_.size() != rows * cols)
            throw Panic("Inconsistent size, attempted to construct a " + std::to_string(rows) + "x" +
                        std::to_string(cols) + " matrix but provided with " + std::to_string(data_.size()) +
                        " elements.");
    }
    int64_t rows() const { return rows_; }
    int64_t columns() const { return cols_; }
    const T& get(int64_t r, int64_t c) const { return data_[r * cols_ + c]; }
    const std::vector<T>& data() const { return data_; }
    bool operator==(const Matrix& o) const { return rows_ == o.rows_ && cols_ == o.cols_ && data_ == o.data_; }

    // <&Matrix<T> as Mul<&Matrix<T>>>::mul -> matrix_view_multiplication, src/matrices/operations.rs:165-196
    Matrix operator*(const Matrix& right) const {
        if (cols_ != right.rows_) {  // assert, :176-183
            std::ostringstream m;
            m << "Mismatched Matrices, left is " << rows_ << "x" << cols_ << ", right is " << right.rows_ << "x"
              << right.cols_ << ", * is only defined for MxN * NxL";
            throw Panic(m.str());
        }
        std::vector<T> c((size_t)(rows_ * right.cols_));
        if constexpr (detail::is_f32<T>)
            detail::check(eml_gemm_f32(data_.data(), right.data_.data(), c.data(), rows_, right.cols_, cols_, cols_,
                                       right.cols_, right.cols_));
        else
            detail::check(eml_gemm_f64(data_.data(), right.data_.data(), c.data(), rows_, right.cols_, cols_, cols_,
                                       right.cols_, right.cols_));
        return Matrix(rows_, right.cols_, std::move(c));
    }

    // linear_algebra::covariance_column_features / covariance_row_features, src/linear_algebra.rs:615-639, :676-700
    Matrix covariance_column_features() const { return covariance(1); }
    Matrix covariance_row_features() const { return covariance(0); }

  private:
    Matrix covariance(int axis) const {
        const int64_t f = axis == 1 ? cols_ : rows_;
        std::vector<T> out((size_t)(f * f));
        if constexpr (detail::is_f32<T>)
            detail::check(eml_covariance_f32(data_.data(), rows_, cols_, cols_, axis, out.data()));
        else
            detail::check(eml_covariance_f64(data_.data(), rows_, cols_, cols_, axis, out.data()));
        return Matrix(f, f, std::move(out));
    }
    int64_t rows_, cols_;
    std::vector<T> data_;
};

// ------------------------------------------------------------------------------------------------ DeviceMatrix
// Matrix<T> with its numbers in HBM: the same operators and panics as Matrix<T>, but an expression such as
// ((a * b) + c).transpose() or the QR loop r = h * r; q = q * h (src/linear_algebra.rs:1531,1535) never leaves the
// device.  `*` is the device contraction, `+ - neg` and the scalar forms the HBM-bound elementwise kernels
// (src/matrices/operations.rs:41-163,1347-1552: one IEEE operation per element, bit-exact), transpose the tiled
// reorder kernel (src/matrices/mod.rs:1277-1300).  Buffers come from the library's stream-ordered block cache
// (eml_alloc_async on the default stream, where every operator is enqueued): no allocator call in steady state.
template <class T>
class DeviceMatrix {
  public:
    explicit DeviceMatrix(const Matrix<T>& m) : DeviceMatrix(m.rows(), m.columns()) {
        detail::check(eml_copy_async(ptr_, m.data().data(), bytes(), nullptr));
        detail::check(eml_sync());  // the source may be a temporary
    }
    DeviceMatrix(DeviceMatrix&& o) noexcept : rows_(o.rows_), cols_(o.cols_), ptr_(o.ptr_) { o.ptr_ = nullptr; }
    DeviceMatrix& operator=(DeviceMatrix&& o) noexcept {
        if (this != &o) {
            release();
            rows_ = o.rows_, cols_ = o.cols_, ptr_ = o.ptr_;
            o.ptr_ = nullptr;
        }
        return *this;
    }
    DeviceMatrix(const DeviceMatrix&) = delete;
    DeviceMatrix& operator=(const DeviceMatrix&) = delete;
    ~DeviceMatrix() { release(); }

    int64_t rows() const { return rows_; }
    int64_t columns() const { return cols_; }
    Matrix<T> to_host() const {
        std::vector<T> out((size_t)(rows_ * cols_));
        detail::check(eml_copy_async(out.data(), ptr_, bytes(), nullptr));
        detail::check(eml_sync());
        return Matrix<T>(rows_, cols_, std::move(out));
    }

    DeviceMatrix operator*(const DeviceMatrix& right) const {  // matrix_view_multiplication, operations.rs:165-196
        if (cols_ != right.rows_) {
            std::ostringstream m;
            m << "Mismatched Matrices, left is " << rows_ << "x" << cols_ << ", right is " << right.rows_ << "x"
              << right.cols_ << ", * is only defined for MxN * NxL";
            throw Panic(m.str());
        }
        DeviceMatrix out(rows_, right.cols_);
        const int64_t sa[3] = {0, cols_, 1}, sb[3] = {0, right.cols_, 1}, sc[3] = {0, right.cols_, 1};
        if constexpr (detail::is_f32<T>)
            detail::check(eml_contract_f32_dev(ptr_, right.ptr_, out.ptr_, 1, rows_, right.cols_, cols_, sa, sb, sc, 0, nullptr));
        else
            detail::check(eml_contract_f64_dev(ptr_, right.ptr_, out.ptr_, 1, rows_, right.cols_, cols_, sa, sb, sc, 0, nullptr));
        return out;
    }
    // The same product through error-free int8 slices on the INT8 tensor cores (f64 only; eml_gemm_f64_sliced_dev):
    // about twice the FP64 pipe for large matrices.  A guard proves the path's 1e-12 contract for these operands or the
    // FP64 kernel runs; *slices_used (optional) reports which (0 = FP64 kernel).  slices = 0: the fewest the guard proves.
    DeviceMatrix mul_sliced(const DeviceMatrix& right, int slices = 0, int* slices_used = nullptr) const {
        static_assert(std::is_same<T, double>::value, "mul_sliced is defined for f64 matrices");
        if (cols_ != right.rows_) {
            std::ostringstream m;
            m << "Mismatched Matrices, left is " << rows_ << "x" << cols_ << ", right is " << right.rows_ << "x"
              << right.cols_ << ", * is only defined for MxN * NxL";
            throw Panic(m.str());
        }
        DeviceMatrix out(rows_, right.cols_);
        int used = 0;
        detail::check(eml_gemm_f64_sliced_dev(ptr_, right.ptr_, out.ptr_, rows_, right.cols_, cols_, cols_, right.cols_,
                                              right.cols_, slices, &used, nullptr));
        if (slices_used) *slices_used = used;
        return out;
    }
    DeviceMatrix operator+(const DeviceMatrix& o) const { return binary(EML_OP_ADD, o, "+"); }  // :41-72
    DeviceMatrix operator-(const DeviceMatrix& o) const { return binary(EML_OP_SUB, o, "-"); }  // :76-107
    DeviceMatrix operator-() const { return unary(EML_OP_NEG, T(0)); }
    DeviceMatrix operator+(T c) const { return unary(EML_OP_ADD_C, c); }                        // :1347-1552
    DeviceMatrix operator-(T c) const { return unary(EML_OP_SUB_C, c); }
    DeviceMatrix operator*(T c) const { return unary(EML_OP_MUL_C, c); }
    DeviceMatrix operator/(T c) const { return unary(EML_OP_DIV_C, c); }
    DeviceMatrix transpose() const {  // Matrix::transpose, src/matrices/mod.rs:1277-1300
        DeviceMatrix out(cols_, rows_);
        const int64_t lens[2] = {cols_, rows_}, strides[2] = {1, cols_};
        if constexpr (detail::is_f32<T>)
            detail::check(eml_reorder_f32_dev(ptr_, out.ptr_, 2, lens, strides, nullptr));
        else
            detail::check(eml_reorder_f64_dev(ptr_, out.ptr_, 2, lens, strides, nullptr));
        return out;
    }

  private:
    DeviceMatrix(int64_t rows, int64_t cols) : rows_(rows), cols_(cols) {
        detail::require_float<T>();
        void* p = nullptr;
        detail::check(eml_alloc_async(&p, bytes(), nullptr));
        ptr_ = static_cast<T*>(p);
    }
    size_t bytes() const { return sizeof(T) * (size_t)(rows_ * cols_); }
    void release() {
        if (ptr_) eml_free_async(ptr_, bytes(), nullptr);
        ptr_ = nullptr;
    }
    DeviceMatrix unary(int op, T c) const {
        DeviceMatrix out(rows_, cols_);
        if constexpr (detail::is_f32<T>)
            detail::check(eml_unary_f32_dev(op, c, ptr_, out.ptr_, nullptr, rows_ * cols_, nullptr));
        else
            detail::check(eml_unary_f64_dev(op, c, ptr_, out.ptr_, nullptr, rows_ * cols_, nullptr));
        return out;
    }
    DeviceMatrix binary(int op, const DeviceMatrix& o, const char* symbol) const {
        if (rows_ != o.rows_ || cols_ != o.cols_) {
            std::ostringstream m;
            m << "Mismatched matrices, left is " << rows_ << "x" << cols_ << ", right is " << o.rows_ << "x" << o.cols_
              << ", " << symbol << " is only defined for MxN " << symbol << " MxN";
            throw Panic(m.str());
        }
        DeviceMatrix out(rows_, cols_);
        if constexpr (detail::is_f32<T>)
            detail::check(eml_binary_f32_dev(op, ptr_, o.ptr_, out.ptr_, nullptr, nullptr, rows_ * cols_, nullptr));
        else
            detail::check(eml_binary_f64_dev(op, ptr_, o.ptr_, out.ptr_, nullptr, nullptr, rows_ * cols_, nullptr));
        return out;
    }
    int64_t rows_, cols_;
    T* ptr_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ GEMM callers
// linear_algebra::cholesky_decomposition, src/linear_algebra.rs:1008-1101 (Cholesky-Banachiewicz, row by row).
// Host side, as in the reference: O(N^3 / 3) sequential work on the small covariance matrix.
template <class T>
std::optional<Matrix<T>> cholesky_decomposition(const Matrix<T>& a) {
    const int64_t n = a.rows();
    if (a.columns() != n) return std::nullopt;
    std::vector<T> low((size_t)(n * n), T(0));
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j <= i; j++) {
            T sum = T(0);
            for (int64_t k = 0; k < j; k++) sum = sum + low[i * n + k] * low[j * n + k];
            if (i == j) {
                const T entry_squared = a.get(i, j) - sum;
                if (entry_squared <= T(0)) return std::nullopt;  // :1087: not positive definite
                low[i * n + j] = std::sqrt(entry_squared);
            } else {
                // :1093-1096: `(a - sum) * (T::one() / L[j][j])` -- reciprocal first, then a multiply (two roundings)
                const T reciprocal = T(1) / low[j * n + j];
                low[i * n + j] = (a.get(i, j) - sum) * reciprocal;
            }
        }
    return Matrix<T>(n, n, std::move(low));
}

// easy_ml::distributions::Gaussian, src/distributions.rs:131-250.  `Source` is any callable returning
// std::optional<T> (the reference's `Iterator<Item = T>`: nullopt = exhausted).
template <class T>
struct Gaussian {
    T mean, variance;
    // draw, :205-237: Box-Muller over pairs; the surplus sample of the last pair is popped
    template <class Source>
    std::optional<std::vector<T>> draw(Source& source, size_t max_samples) const {
        const T two = T(1) + T(1), minus_two = -two, two_pi = two * T(3.14159265358979323846264338327950288L);
        const T standard_deviation = std::sqrt(variance);
        std::vector<T> samples;
        samples.reserve(max_samples + 1);
        while (samples.size() < max_samples) {
            const std::optional<T> u = source();
            if (!u) return std::nullopt;
            const std::optional<T> v = source();
            if (!v) return std::nullopt;
            const T z1 = std::sqrt(minus_two * std::log(*u)) * std::cos(two_pi * *v);
            const T z2 = std::sqrt(minus_two * std::log(*u)) * std::sin(two_pi * *v);
            samples.push_back(z1 * standard_deviation + mean);
            samples.push_back(z2 * standard_deviation + mean);
        }
        if (samples.size() > max_samples) samples.pop_back();
        return samples;
    }
};

// easy_ml::distributions::MultivariateGaussian, src/distributions.rs:262-369.  The reference's draw_tensor_samples
// (:484-540) runs one `mean + L * z` matrix-vector product per sample; here the standard-normal vectors are stacked
// as the rows of Z (the source is consumed in the same order) and ONE product Z * L^T goes through the drop-in
// contraction, L addressed with transposed strides: element (m, n) is the same scalar product in the same k order.
template <class T>
class MultivariateGaussian {
  public:
    MultivariateGaussian(Matrix<T> mean, Matrix<T> covariance) : mean_(std::move(mean)), cov_(std::move(covariance)) {
        if (mean_.columns() != 1) throw Panic("Mean must be a column vector");
        if (cov_.rows() != cov_.columns()) throw Panic("Supplied 'covariance' matrix is not square");
        if (mean_.rows() != cov_.rows()) throw Panic("Means must be same length as covariance matrix");
    }
    const Matrix<T>& mean() const { return mean_; }
    const Matrix<T>& covariance() const { return cov_; }

    template <class Source>
    std::optional<Matrix<T>> draw(Source& source, int64_t max_samples) const {
        const std::optional<Matrix<T>> lower = cholesky_decomposition(cov_);
        if (!lower) return std::nullopt;
        const int64_t n = mean_.rows();
        const Gaussian<T> normal{T(0), T(1)};
        std::vector<T> z((size_t)(max_samples * n));
        for (int64_t m = 0; m < max_samples; m++) {
            const std::optional<std::vector<T>> row = normal.draw(source, (size_t)n);
            if (!row) return std::nullopt;
            std::copy(row->begin(), row->end(), z.begin() + m * n);
        }
        std::vector<T> drawn((size_t)(max_samples * n));
        const int64_t sz[3] = {0, n, 1}, sl[3] = {0, 1, n}, sd[3] = {0, n, 1};
        if constexpr (detail::is_f32<T>)
            detail::check(eml_contract_f32(z.data(), lower->data().data(), drawn.data(), 1, max_samples, n, n, sz, sl, sd));
        else
            detail::check(eml_contract_f64(z.data(), lower->data().data(), drawn.data(), 1, max_samples, n, n, sz, sl, sd));
        for (int64_t m = 0; m < max_samples; m++)
            for (int64_t f = 0; f < n; f++) drawn[m * n + f] = mean_.get(f, 0) + drawn[m * n + f];  // :528
        return Matrix<T>(max_samples, n, std::move(drawn));
    }

  private:
    Matrix<T> mean_, cov_;
};

// ------------------------------------------------------------------------------------------------ Tensor
// easy_ml::tensors::Tensor<T, D>: row-major Vec<T> + [(Dimension, usize); D], src/tensors/mod.rs:244-250.
// The rank is a runtime property here (C++ has no need of the const generic); rank-specific operations check it.
template <class T>
class Tensor {
  public:
    Tensor(Shape shape, std::vector<T> data) : shape_(std::move(shape)), data_(std::move(data)) {
        detail::require_float<T>();
        int64_t n = 1;
        for (auto& d : shape_) {
            if (d.second < 1) throw Panic("Invalid shape: " + detail::shape_debug(shape_));  // mod.rs:96-123
            n *= d.second;
        }
        for (size_t i = 0; i < shape_.size(); i++)
            for (size_t j = i + 1; j < shape_.size(); j++)
                if (shape_[i].first == shape_[j].first)
                    throw Panic("Dimension names must all be unique: " + detail::shape_debug(shape_));
        if ((int64_t)data_.size() != n)
            throw Panic("Length of data does not match size of the shape: " + detail::shape_debug(shape_));
    }
    const Shape& shape() const { return shape_; }
    const std::vector<T>& data() const { return data_; }
    size_t rank() const { return shape_.size(); }
    std::vector<Dimension> names() const {
        std::vector<Dimension> n;
        for (auto& d : shape_) n.push_back(d.first);
        return n;
    }
    std::vector<int64_t> strides() const {  // compute_strides, mod.rs:563
        std::vector<int64_t> s(shape_.size(), 1);
        for (int i = (int)shape_.size() - 2; i >= 0; i--) s[i] = s[i + 1] * shape_[i + 1].second;
        return s;
    }
    bool operator==(const Tensor& o) const { return shape_ == o.shape_ && data_ == o.data_; }

    // Mul for rank 2 -> tensor_view_matrix_product, src/tensors/operations.rs:474-519
    Tensor operator*(const Tensor& right) const {
        if (rank() != 2 || right.rank() != 2) throw std::invalid_argument("* is defined for rank 2 tensors");
        if (shape_[1].second != right.shape_[0].second)  // :485-491
            throw Panic("Mismatched tensors, left is " + detail::shape_debug(shape_) + ", right is " +
                        detail::shape_debug(right.shape_) +
                        ", * is only defined for MxN * NxL dimension lengths");
        Shape out{shape_[0], right.shape_[1]};
        if (shape_[0].first == right.shape_[1].first)  // :492-501
            throw Panic("Matrix multiplication of tensors with shapes left " + detail::shape_debug(shape_) +
                        " and right " + detail::shape_debug(right.shape_) +
                        " would create duplicate dimension names as the shape " + detail::shape_debug(out) +
                        ". Rename one or both of the dimension names in the input to prevent this. * is defined as "
                        "MxN * NxL = MxL");
        const int64_t m = shape_[0].second, k = shape_[1].second, n = right.shape_[1].second;
        std::vector<T> c((size_t)(m * n));
        if constexpr (detail::is_f32<T>)
            detail::check(eml_gemm_f32(data_.data(), right.data_.data(), c.data(), m, n, k, k, n, n));
        else
            detail::check(eml_gemm_f64(data_.data(), right.data_.data(), c.data(), m, n, k, k, n, n));
        return Tensor(out, std::move(c));
    }

    // Tensor<T, 1>::scalar_product, src/tensors/mod.rs:1773 -> operations.rs:1773-1808
    T scalar_product(const Tensor& right) const {
        if (rank() != 1 || right.rank() != 1) throw std::invalid_argument("scalar_product is defined for rank 1");
        if (shape_ != right.shape_)
            throw Panic("Dimensions of left and right tensors are not the same: (left: " + detail::shape_debug(shape_) +
                        ", right: " + detail::shape_debug(right.shape_) + ")");
        T out{};
        if constexpr (detail::is_f32<T>)
            detail::check(eml_dot_f32(data_.data(), 1, right.data_.data(), 1, (int64_t)data_.size(), &out));
        else
            detail::check(eml_dot_f64(data_.data(), 1, right.data_.data(), 1, (int64_t)data_.size(), &out));
        return out;
    }

  private:
    Shape shape_;
    std::vector<T> data_;
};

// ------------------------------------------------------------------------------------------------ Einsum
// Einsum::with_2(&x, &y).to([...]) -- Einsum2::to, src/tensors/einsum.rs:908-980
template <class T>
class Einsum2 {
  public:
    Einsum2(const Tensor<T>& x, const Tensor<T>& y) : x_(x), y_(y) {}

    Tensor<T> to(const std::vector<Dimension>& output) const {
        const Shape &xs = x_.shape(), &ys = y_.shape();
        auto find = [](const Shape& s, const Dimension& d) -> std::optional<int64_t> {
            for (auto& e : s)
                if (e.first == d) return e.second;
            return std::nullopt;
        };
        // output_shape_for :563-572 with length_of :508-538
        Shape out_shape;
        for (auto& d : output) {
            std::vector<std::optional<int64_t>> lengths{find(xs, d), find(ys, d)};
            std::optional<int64_t> known;
            bool ok = true;
            for (auto& l : lengths)
                if (l) {
                    if (known && *known != *l) ok = false;
                    known = l;
                }
            if (!known || !ok) throw InconsistentDimensionLengthError(d, lengths);
            out_shape.emplace_back(d, *known);
        }
        for (size_t i = 0; i < output.size(); i++)
            for (size_t j = i + 1; j < output.size(); j++)
                if (output[i] == output[j])
                    throw Panic("Dimension names must all be unique: " + detail::shape_debug(out_shape));
        auto in_out = [&](const Dimension& d) { return std::find(output.begin(), output.end(), d) != output.end(); };
        // summation_dimensions :579-622: input names not in the output, in order of first appearance
        Shape summation;
        for (const Shape* s : {&xs, &ys})
            for (auto& e : *s) {
                if (in_out(e.first)) continue;
                auto seen = find(summation, e.first);
                if (!seen)
                    summation.push_back(e);
                else if (*seen != e.second)
                    throw InconsistentDimensionLengthError(e.first, {find(xs, e.first), find(ys, e.first)});
            }
        std::map<Dimension, int64_t> len, sx, sy, so;
        for (auto& e : out_shape) len[e.first] = e.second;
        for (auto& e : summation) len[e.first] = e.second;
        {
            auto st = x_.strides();
            for (size_t i = 0; i < xs.size(); i++) sx[xs[i].first] = st[i];
            st = y_.strides();
            for (size_t i = 0; i < ys.size(); i++) sy[ys[i].first] = st[i];
            int64_t acc = 1;
            for (int i = (int)out_shape.size() - 1; i >= 0; i--) {
                so[out_shape[i].first] = acc;
                acc *= out_shape[i].second;
            }
        }
        auto stride = [](const std::map<Dimension, int64_t>& m, const Dimension& d) {
            auto it = m.find(d);
            return it == m.end() ? int64_t(0) : it->second;
        };
        int64_t total = 1;
        for (auto& e : out_shape) total *= e.second;
        std::vector<T> out((size_t)total);

        // classify every name by where it appears: batch / M / N / K
        std::vector<Dimension> batch, md, nd, kd;
        bool lone_sum = false;
        for (auto& d : output) {
            bool ix = sx.count(d), iy = sy.count(d);
            (ix && iy ? batch : ix ? md : nd).push_back(d);
        }
        for (auto& e : summation) {
            if (sx.count(e.first) && sy.count(e.first))
                kd.push_back(e.first);
            else
                lone_sum = true;
        }
        struct Role {
            bool ok = true;
            int64_t length = 1;
            std::vector<int64_t> strides;
        };
        // several names playing one role merge only when they nest contiguously, in order, in every tensor
        auto merge = [&](std::vector<Dimension> dims, std::vector<const std::map<Dimension, int64_t>*> maps) {
            Role r;
            r.strides.assign(maps.size(), 0);
            if (dims.empty()) return r;
            std::vector<Dimension> kept;
            for (auto& d : dims)
                if (len[d] > 1) kept.push_back(d);
            if (kept.empty()) kept.push_back(dims[0]);
            for (auto& d : kept) r.length *= len[d];
            for (size_t t = 0; t < maps.size(); t++) {
                for (size_t i = 0; i + 1 < kept.size(); i++)
                    if (stride(*maps[t], kept[i]) != stride(*maps[t], kept[i + 1]) * len[kept[i + 1]]) r.ok = false;
                r.strides[t] = stride(*maps[t], kept.back());
            }
            return r;
        };
        bool gemm = !kd.empty() && !lone_sum;
        Role b, m, n, k;
        if (gemm) {
            b = merge(batch, {&sx, &sy, &so});
            m = merge(md, {&sx, &so});
            n = merge(nd, {&sy, &so});
            k = merge(kd, {&sx, &sy});
            gemm = b.ok && m.ok && n.ok && k.ok;
        }
        if (gemm) {
            const int64_t sA[3] = {b.strides[0], m.strides[0], k.strides[0]};
            const int64_t sB[3] = {b.strides[1], k.strides[1], n.strides[0]};
            const int64_t sC[3] = {b.strides[2], m.strides[1], n.strides[1]};
            if constexpr (detail::is_f32<T>)
                detail::check(eml_contract_f32(x_.data().data(), y_.data().data(), out.data(), b.length, m.length,
                                               n.length, k.length, sA, sB, sC));
            else
                detail::check(eml_contract_f64(x_.data().data(), y_.data().data(), out.data(), b.length, m.length,
                                               n.length, k.length, sA, sB, sC));
        } else {
            // not a (batched) GEMM: the strided loop nest in the reference's order, bit-exact
            std::vector<int64_t> ol, ao, bo, sl, as, bs;
            for (auto& e : out_shape) {
                ol.push_back(e.second);
                ao.push_back(stride(sx, e.first));
                bo.push_back(stride(sy, e.first));
            }
            for (auto& e : summation) {
                sl.push_back(e.second);
                as.push_back(stride(sx, e.first));
                bs.push_back(stride(sy, e.first));
            }
            if constexpr (detail::is_f32<T>)
                detail::check(eml_einsum2_f32(x_.data().data(), (int64_t)x_.data().size(), y_.data().data(),
                                              (int64_t)y_.data().size(), out.data(), (int)ol.size(), ol.data(),
                                              ao.data(), bo.data(), (int)sl.size(), sl.data(), as.data(), bs.data()));
            else
                detail::check(eml_einsum2_f64(x_.data().data(), (int64_t)x_.data().size(), y_.data().data(),
                                              (int64_t)y_.data().size(), out.data(), (int)ol.size(), ol.data(),
                                              ao.data(), bo.data(), (int)sl.size(), sl.data(), as.data(), bs.data()));
        }
        return Tensor<T>(out_shape, std::move(out));
    }

  private:
    const Tensor<T>& x_;
    const Tensor<T>& y_;
};
// Einsum1 / Einsum3..6 (src/tensors/einsum.rs:829-883, :1019-1666): the strided loop nest on the device in the
// reference's order (zero-seeded sum, left-associated product) -- bit-exact.
template <class T>
class EinsumN {
  public:
    explicit EinsumN(std::vector<const Tensor<T>*> xs) : xs_(std::move(xs)) {}
    Tensor<T> to(const std::vector<Dimension>& output) const {
        auto find = [](const Shape& s, const Dimension& d) -> std::optional<int64_t> {
            for (auto& e : s)
                if (e.first == d) return e.second;
            return std::nullopt;
        };
        auto lengths_of = [&](const Dimension& d) {
            std::vector<std::optional<int64_t>> l;
            for (auto* x : xs_) l.push_back(find(x->shape(), d));
            return l;
        };
        Shape out_shape, summation;
        for (auto& d : output) {  // output_shape_for :563-572
            auto lengths = lengths_of(d);
            std::optional<int64_t> known;
            bool ok = true;
            for (auto& l : lengths)
                if (l) {
                    if (known && *known != *l) ok = false;
                    if (!known) known = l;
                }
            if (!known || !ok) throw InconsistentDimensionLengthError(d, lengths);
            out_shape.emplace_back(d, *known);
        }
        for (size_t i = 0; i < output.size(); i++)
            for (size_t j = i + 1; j < output.size(); j++)
                if (output[i] == output[j])
                    throw Panic("Dimension names must all be unique: " + detail::shape_debug(out_shape));
        for (auto* x : xs_)  // summation_dimensions :579-622
            for (auto& e : x->shape()) {
                if (std::find(output.begin(), output.end(), e.first) != output.end()) continue;
                auto seen = find(summation, e.first);
                if (!seen)
                    summation.push_back(e);
                else if (*seen != e.second)
                    throw InconsistentDimensionLengthError(e.first, lengths_of(e.first));
            }
        const int n = (int)xs_.size(), n_out = (int)out_shape.size(), n_sum = (int)summation.size();
        std::vector<int64_t> ol, sl, os((size_t)n * n_out, 0), ss((size_t)n * n_sum, 0), elems;
        std::vector<const T*> ptrs;
        int64_t total = 1;
        for (auto& e : out_shape) {
            ol.push_back(e.second);
            total *= e.second;
        }
        for (auto& e : summation) sl.push_back(e.second);
        for (int i = 0; i < n; i++) {
            auto st = xs_[i]->strides();
            const Shape& sh = xs_[i]->shape();
            for (size_t k = 0; k < sh.size(); k++) {
                for (int d = 0; d < n_out; d++)
                    if (out_shape[d].first == sh[k].first) os[(size_t)i * n_out + d] = st[k];
                for (int d = 0; d < n_sum; d++)
                    if (summation[d].first == sh[k].first) ss[(size_t)i * n_sum + d] = st[k];
            }
            ptrs.push_back(xs_[i]->data().data());
            elems.push_back((int64_t)xs_[i]->data().size());
        }
        std::vector<T> out((size_t)total);
        if constexpr (detail::is_f32<T>)
            detail::check(eml_einsum_n_f32(n, ptrs.data(), elems.data(), out.data(), n_out, ol.data(), os.data(), n_sum,
                                           sl.data(), ss.data()));
        else
            detail::check(eml_einsum_n_f64(n, ptrs.data(), elems.data(), out.data(), n_out, ol.data(), os.data(), n_sum,
                                           sl.data(), ss.data()));
        return Tensor<T>(out_shape, std::move(out));
    }

  private:
    std::vector<const Tensor<T>*> xs_;
};

struct Einsum {  // src/tensors/einsum.rs:98-313
    template <class T>
    static EinsumN<T> with_1(const Tensor<T>& x) { return EinsumN<T>({&x}); }
    template <class T>
    static Einsum2<T> with_2(const Tensor<T>& x, const Tensor<T>& y) { return Einsum2<T>(x, y); }
    template <class T>
    static EinsumN<T> with_3(const Tensor<T>& a, const Tensor<T>& b, const Tensor<T>& c) { return EinsumN<T>({&a, &b, &c}); }
    template <class T>
    static EinsumN<T> with_4(const Tensor<T>& a, const Tensor<T>& b, const Tensor<T>& c, const Tensor<T>& d) {
        return EinsumN<T>({&a, &b, &c, &d});
    }
    template <class T>
    static EinsumN<T> with_5(const Tensor<T>& a, const Tensor<T>& b, const Tensor<T>& c, const Tensor<T>& d,
                             const Tensor<T>& e) {
        return EinsumN<T>({&a, &b, &c, &d, &e});
    }
    template <class T>
    static EinsumN<T> with_6(const Tensor<T>& a, const Tensor<T>& b, const Tensor<T>& c, const Tensor<T>& d,
                             const Tensor<T>& e, const Tensor<T>& f) {
        return EinsumN<T>({&a, &b, &c, &d, &e, &f});
    }
};

// ------------------------------------------------------------------------------------------------ autodiff
// A step recorded once (WengertList::record) and replayed as one CUDA graph launch.  No counterpart in the
// reference; see the restrictions at eml_tape_capture_begin in easyml_b200.h.
class RecordedStep {
  public:
    explicit RecordedStep(eml_graph* g) : g_(g) {}
    RecordedStep(RecordedStep&& o) noexcept : g_(o.g_) { o.g_ = nullptr; }
    RecordedStep(const RecordedStep&) = delete;
    ~RecordedStep() {
        if (g_) eml_graph_destroy(g_);
    }
    // enqueues one replay; wait on the returned arrival (eml_arrival_wait) before reading what the step copied out
    void* launch() const {
        void* arrival = nullptr;
        detail::check(eml_graph_launch(g_, &arrival));
        return arrival;
    }
    // development aid: critical-path cost of every kernel of the step (see eml_graph_profile); the model is garbage after
    std::string profile(int replays = 200) const {
        std::string report(1 << 16, '\0');
        detail::check(eml_graph_profile(g_, replays, report.data(), (int64_t)report.size()));
        report.resize(std::strlen(report.c_str()));
        return report;
    }

  private:
    eml_graph* g_;
};


// WengertList<T>, src/differentiation.rs:327-331 -- here a handle to the device tape (one macro node per
// container operation; tape length and every Index equal the reference's scalar tape, see csrc/tape.cu).
template <class T>
class WengertList {
  public:
    WengertList() {
        detail::require_float<T>();
        detail::check(eml_tape_create(detail::is_f32<T> ? EML_F32 : EML_F64, &t_));
    }
    ~WengertList() {
        if (t_) eml_tape_destroy(t_);
    }
    WengertList(const WengertList&) = delete;
    WengertList& operator=(const WengertList&) = delete;
    int64_t len() const {  // operations.len()
        int64_t n = 0;
        detail::check(eml_tape_len(t_, &n));
        return n;
    }
    void clear() { detail::check(eml_tape_clear(t_)); }  // differentiation.rs:674
    // records the device work `step` issues (it must have run eagerly once before, and must leave the tape as it
    // found it: clear() + reset()) and returns it as a replayable graph
    template <class F>
    RecordedStep record(F&& step) {
        detail::check(eml_tape_capture_begin(t_));
        try {
            step();
        } catch (...) {
            eml_graph* dead = nullptr;
            eml_tape_capture_end(t_, &dead);
            if (dead) eml_graph_destroy(dead);
            throw;
        }
        eml_graph* g = nullptr;
        detail::check(eml_tape_capture_end(t_, &g));
        return RecordedStep(g);
    }
    eml_tape* raw() const { return t_; }

  private:
    eml_tape* t_ = nullptr;
};

template <class T>
class RecordTensor;

// Derivatives<T>, src/differentiation.rs:365-367
template <class T>
class Derivatives {
  public:
    explicit Derivatives(eml_derivs* d) : d_(d, [](eml_derivs* p) { eml_derivs_free(p); }) {}
    T at(int64_t index) const {  // Index<&Record>, differentiation.rs:387-403
        double v = 0;
        detail::check(eml_derivs_at(d_.get(), index, &v));
        return (T)v;
    }
    Tensor<T> at_tensor(const RecordTensor<T>& input) const;  // container_record/mod.rs:1292
    int64_t len() const {
        int64_t n = 0;
        detail::check(eml_derivs_len(d_.get(), &n));
        return n;
    }
    eml_derivs* raw() const { return d_.get(); }

  private:
    std::shared_ptr<eml_derivs> d_;
};

// RecordTensor<'a, T, S, 2> / RecordMatrix, container_record/mod.rs:66-89 (rank 2: the matmul path)
template <class T>
class RecordTensor {
  public:
    // RecordTensor::variables, mod.rs:149-164
    static RecordTensor variables(const WengertList<T>& history, const Tensor<T>& x) {
        require_rank2(x);
        eml_val v;
        detail::check(eml_tape_variables(history.raw(), x.data().data(), x.shape()[0].second, x.shape()[1].second, &v));
        return RecordTensor(&history, v, x.shape(), true);
    }
    // RecordTensor::constants, mod.rs:115-128 (no history; `owner` only says which device tape holds the numbers)
    static RecordTensor constants(const Tensor<T>& x, const WengertList<T>& owner) {
        require_rank2(x);
        eml_val v;
        detail::check(eml_tape_constants(owner.raw(), x.data().data(), x.shape()[0].second, x.shape()[1].second, &v));
        return RecordTensor(&owner, v, x.shape(), false);
    }
    RecordTensor(RecordTensor&& o) noexcept : list_(o.list_), v_(o.v_), shape_(std::move(o.shape_)), tracked_(o.tracked_) {
        o.v_ = -1;
    }
    RecordTensor(const RecordTensor&) = delete;
    ~RecordTensor() {
        if (v_ >= 0) eml_val_release(list_->raw(), v_);
    }
    const Shape& shape() const { return shape_; }
    bool has_history() const { return tracked_; }
    eml_val raw() const { return v_; }
    const WengertList<T>* list() const { return list_; }
    void reset() { detail::check(eml_tape_reset(list_->raw(), v_)); }  // mod.rs:349-365

    Tensor<T> numbers() const {
        std::vector<T> out((size_t)(shape_[0].second * shape_[1].second));
        detail::check(eml_val_numbers(list_->raw(), v_, out.data()));
        return Tensor<T>(shape_, std::move(out));
    }
    // asynchronous read-back into page-locked memory; wait on the returned arrival (eml_arrival_wait) before reading
    void* numbers_async(T* pinned_out) const {
        void* arrival = nullptr;
        detail::check(eml_val_numbers_async(list_->raw(), v_, pinned_out, &arrival));
        return arrival;
    }
    std::vector<int64_t> indexes() const {  // the Index half of the (T, Index) view, mod.rs:2511
        std::vector<int64_t> out((size_t)(shape_[0].second * shape_[1].second));
        detail::check(eml_val_indexes(list_->raw(), v_, out.data()));
        return out;
    }

    // &RecordTensor * &RecordTensor -> record_tensor_matrix_multiply, container_operations.rs:1318-1381
    RecordTensor operator*(const RecordTensor& right) const {
        if (tracked_ && right.tracked_ && list_ != right.list_)  // assert same list, :1331
            throw Panic("Record containers must be using the same WengertList");
        if (shape_[1].second != right.shape_[0].second)  // :1339-1345
            throw Panic("Mismatched tensors, left is " + detail::shape_debug(shape_) + ", right is " +
                        detail::shape_debug(right.shape_) +
                        ", * is only defined for MxN * NxL dimension lengths");
        Shape out{shape_[0], right.shape_[1]};
        if (shape_[0].first == right.shape_[1].first)  // :1346-1355
            throw Panic("Matrix multiplication of tensors with shapes left " + detail::shape_debug(shape_) +
                        " and right " + detail::shape_debug(right.shape_) +
                        " would create duplicate dimension names as the shape " + detail::shape_debug(out) +
                        ". Rename one or both of the dimension names in the input to prevent this. * is defined as "
                        "MxN * NxL = MxL");
        eml_val c;
        detail::check(eml_tape_matmul(list_->raw(), v_, right.v_, &c));
        return RecordTensor(list_, c, out, tracked_ || right.tracked_);
    }
    // unary container operators, container_operations.rs:1644-1804 and swapped.rs:25-45
    RecordTensor operator-() const { return unary(EML_OP_NEG, 0); }
    RecordTensor exp() const { return unary(EML_OP_EXP, 0); }
    RecordTensor ln() const { return unary(EML_OP_LN, 0); }
    RecordTensor sin() const { return unary(EML_OP_SIN, 0); }
    RecordTensor cos() const { return unary(EML_OP_COS, 0); }
    RecordTensor sqrt() const { return unary(EML_OP_SQRT, 0); }
    RecordTensor pow(T c) const { return unary(EML_OP_POW_C, c); }
    RecordTensor operator+(T c) const { return unary(EML_OP_ADD_C, c); }
    RecordTensor operator-(T c) const { return unary(EML_OP_SUB_C, c); }
    RecordTensor operator*(T c) const { return unary(EML_OP_MUL_C, c); }
    RecordTensor operator/(T c) const { return unary(EML_OP_DIV_C, c); }
    RecordTensor sub_swapped(T c) const { return unary(EML_OP_C_SUB, c); }
    RecordTensor div_swapped(T c) const { return unary(EML_OP_C_DIV, c); }
    // binary container operators, container_record/mod.rs:962-1043, :1224, :1247
    RecordTensor operator+(const RecordTensor& o) const { return binary(EML_OP_ADD, o); }
    RecordTensor operator-(const RecordTensor& o) const { return binary(EML_OP_SUB, o); }
    RecordTensor elementwise_multiply(const RecordTensor& o) const { return binary(EML_OP_MUL, o); }
    RecordTensor elementwise_divide(const RecordTensor& o) const { return binary(EML_OP_DIV, o); }

    // derivatives_for, mod.rs:1201-1213 -> Record::try_derivatives, differentiation.rs:623-648
    Derivatives<T> derivatives_for(int64_t row, int64_t col) const {
        if (!tracked_) throw Panic("Record has no WengertList to find derivatives from");
        eml_derivs* d = nullptr;
        detail::check(eml_tape_derivatives(list_->raw(), v_, row, col, &d));
        return Derivatives<T>(d);
    }

  private:
    RecordTensor(const WengertList<T>* l, eml_val v, Shape s, bool tracked)
        : list_(l), v_(v), shape_(std::move(s)), tracked_(tracked) {}
    static void require_rank2(const Tensor<T>& x) {
        if (x.rank() != 2) throw std::invalid_argument("the device tape holds rank 2 containers");
    }
    RecordTensor unary(int op, double c) const {
        eml_val y;
        detail::check(eml_tape_unary(list_->raw(), op, c, v_, &y));
        return RecordTensor(list_, y, shape_, tracked_);
    }
    RecordTensor binary(int op, const RecordTensor& o) const {
        if (shape_ != o.shape_)  // mod.rs:973-980
            throw Panic("Record containers must have the same shape for a binary operation: (left: " +
                        detail::shape_debug(shape_) + ", right: " + detail::shape_debug(o.shape_) + ")");
        eml_val z;
        detail::check(eml_tape_binary(list_->raw(), op, v_, o.v_, &z));
        return RecordTensor(list_, z, shape_, tracked_ || o.tracked_);
    }
    const WengertList<T>* list_;
    eml_val v_;
    Shape shape_;
    bool tracked_;
    friend class Derivatives<T>;
};

template <class T>
Tensor<T> Derivatives<T>::at_tensor(const RecordTensor<T>& input) const {
    std::vector<T> out((size_t)(input.shape()[0].second * input.shape()[1].second));
    detail::check(eml_derivs_at_val(d_.get(), input.raw(), out.data()));
    return Tensor<T>(input.shape(), std::move(out));
}

// `w.map_mut(|x| x - derivatives[&x] * lr)`, src/neural_networks.rs:418-420, fused on the device
template <class T>
void sgd_update(RecordTensor<T>& w, const Derivatives<T>& d, T lr) {
    detail::check(eml_tape_sgd_update(w.list()->raw(), w.raw(), d.raw(), (double)lr));
}

}  // namespace easyml
