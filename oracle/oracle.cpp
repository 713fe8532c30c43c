// oracle.cpp -- CPU restatement of Easy ML's dense-contraction hot path (see oracle.h).
// TEST INFRASTRUCTURE ONLY: checker and CPU baseline, never the product path.
// Parity pinned against the reference's own golden vectors (tests/golden/); no reference build
// is possible here (Rust toolchain absent), so large shapes are "pinned only by restatement".
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC  (see oracle/Makefile)
#include "oracle.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_error;

int panic(const std::string& msg) {
    g_error = msg;
    return ORC_ERR_PANIC;
}

std::string shape_debug(const char* const* names, const int64_t* len, int d) {
    // Rust {:?} of [(Dimension, usize); D], e.g. [("r", 2), ("c", 3)]
    std::string s = "[";
    for (int i = 0; i < d; i++) {
        if (i) s += ", ";
        s += "(\"" + std::string(names[i]) + "\", " + std::to_string(len[i]) + ")";
    }
    return s + "]";
}

// ---- A1 -----------------------------------------------------------------------------------
// src/tensors/operations.rs:441-445: zip -> map(x*y) -> reduce(x+y).  reduce seeds with the
// FIRST product (no zero), then adds the remaining products left to right.  Each product and
// each sum is rounded separately (rustc never contracts to FMA); -ffp-contract=off keeps it so.
template <class T>
T scalar_product(const T* l, int64_t ls, const T* r, int64_t rs, int64_t k) {
    T acc = l[0] * r[0];
    for (int64_t i = 1; i < k; i++) {
        T p = l[i * ls] * r[i * rs];
        acc = acc + p;
    }
    return acc;
}

// ---- A2 -----------------------------------------------------------------------------------
// src/matrices/operations.rs:165-196.  Row-major result walk (:187); left row iterator has stride 1
// (iterators.rs:776-791), right column iterator has stride = columns (iterators.rs:692-707).
template <class T>
void matrix_mul_rows(const T* a, int64_t a_cols, const T* b, int64_t b_cols, T* c, int64_t row0, int64_t row1) {
    for (int64_t i = row0; i < row1; i++)
        for (int64_t j = 0; j < b_cols; j++)
            c[i * b_cols + j] = scalar_product<T>(a + i * a_cols, 1, b + j, b_cols, a_cols);
}

// A bounded SAMPLE of the same product for CPU-baseline timing: elements C[row0:row1, col0:col1] only, computed
// exactly as above (B is still the full-width row-major matrix, so the column walk keeps its stride).  Every
// element of C costs the same 2K flop; threads split the (row, column-chunk) work, results stay bit-identical.
template <class T>
void matrix_mul_block(const T* a, int64_t a_cols, const T* b, int64_t b_cols, T* c, int64_t row0, int64_t row1,
                      int64_t col0, int64_t col1, int threads) {
    const int64_t rows = row1 - row0, cols = col1 - col0;
    auto work = [=](int t) {
        // thread t takes every threads-th (row, 64-column chunk)
        const int64_t chunks = (cols + 63) / 64;
        for (int64_t w = t; w < rows * chunks; w += threads) {
            const int64_t i = row0 + w / chunks, j0 = col0 + (w % chunks) * 64;
            const int64_t j1 = j0 + 64 < col1 ? j0 + 64 : col1;
            for (int64_t j = j0; j < j1; j++)
                c[(i - row0) * cols + (j - col0)] = scalar_product<T>(a + i * a_cols, 1, b + j, b_cols, a_cols);
        }
    };
    if (threads <= 1) {
        work(0);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
}

template <class T>
int matrix_mul(const T* a, int64_t a_rows, int64_t a_cols, const T* b, int64_t b_rows, int64_t b_cols, T* c,
               int64_t row0, int64_t row1, int threads) {
    if (a_cols != b_rows) {  // :176-183
        char buf[256];
        snprintf(buf, sizeof buf,
                 "Mismatched Matrices, left is %lldx%lld, right is %lldx%lld, * is only defined for MxN * NxL",
                 (long long)a_rows, (long long)a_cols, (long long)b_rows, (long long)b_cols);
        return panic(buf);
    }
    if (threads <= 1 || row1 - row0 < threads) {
        matrix_mul_rows<T>(a, a_cols, b, b_cols, c, row0, row1);
        return ORC_OK;
    }
    std::vector<std::thread> pool;
    int64_t rows = row1 - row0;
    for (int t = 0; t < threads; t++) {
        int64_t r0 = row0 + rows * t / threads, r1 = row0 + rows * (t + 1) / threads;
        pool.emplace_back([=] { matrix_mul_rows<T>(a, a_cols, b, b_cols, c, r0, r1); });
    }
    for (auto& th : pool) th.join();
    return ORC_OK;
}

// ---- A3 -----------------------------------------------------------------------------------
// src/tensors/operations.rs:474-519
template <class T>
int tensor_mul(const T* a, const char* const* an, const int64_t* al, const T* b, const char* const* bn,
               const int64_t* bl, T* c, const char** out_names) {
    if (al[1] != bl[0]) {  // :485-491 (names of the contracted pair are ignored)
        return panic("Mismatched tensors, left is " + shape_debug(an, al, 2) + ", right is " +
                     shape_debug(bn, bl, 2) + ", * is only defined for MxN * NxL dimension lengths");
    }
    if (std::strcmp(an[0], bn[1]) == 0) {  // :492-501
        const char* on[2] = {an[0], bn[1]};
        int64_t ol[2] = {al[0], bl[1]};
        return panic("Matrix multiplication of tensors with shapes left " + shape_debug(an, al, 2) +
                     " and right " + shape_debug(bn, bl, 2) +
                     " would create duplicate dimension names as the shape " + shape_debug(on, ol, 2) +
                     ". Rename one or both of the dimension names in the input to prevent this. * is defined "
                     "as MxN * NxL = MxL");
    }
    // :506-517 -- output shape [left_shape[0], right_shape[1]], iterated last-dimension-fastest
    // (src/tensors/indexing.rs:735-768); TensorIndex picks row i of left / column j of right.
    for (int64_t i = 0; i < al[0]; i++)
        for (int64_t j = 0; j < bl[1]; j++) c[i * bl[1] + j] = scalar_product<T>(a + i * al[1], 1, b + j, bl[1], al[1]);
    if (out_names) {
        out_names[0] = an[0];
        out_names[1] = bn[1];
    }
    return ORC_OK;
}

// ---- A4 -----------------------------------------------------------------------------------
struct Dim {
    const char* name;
    int64_t len;
};

bool same(const char* a, const char* b) { return std::strcmp(a, b) == 0; }

// src/tensors/einsum.rs:508-538 (length_of): first input containing the name fixes the length; any
// other input containing it with a different length, or no input containing it, is an error.
int length_of(const char* dim, const std::vector<std::vector<Dim>>& in, int64_t* out, int64_t* err_len) {
    std::vector<int64_t> lens(in.size(), -1);
    for (size_t t = 0; t < in.size(); t++)
        for (auto& d : in[t])
            if (same(d.name, dim)) {
                lens[t] = d.len;
                break;  // .find()
            }
    int64_t first = -1;
    for (auto l : lens)
        if (l >= 0) {
            first = l;
            break;
        }
    bool bad = first < 0;
    for (auto l : lens)
        if (l >= 0 && l != first) bad = true;
    if (bad) {
        for (size_t t = 0; t < in.size(); t++) err_len[t] = lens[t];
        return ORC_ERR_INCONSISTENT;
    }
    *out = first;
    return ORC_OK;
}

// src/tensors/einsum.rs:579-622 (summation_dimensions): input names not requested in the output, in
// order of first appearance; repeated names must agree on length.
int summation_dimensions(const std::vector<std::vector<Dim>>& in, const std::vector<const char*>& out,
                         std::vector<Dim>* sum, const char** err_dim, int64_t* err_len) {
    for (auto& shape : in)
        for (auto& d : shape) {
            bool in_out = false;
            for (auto o : out)
                if (same(o, d.name)) in_out = true;
            if (in_out) continue;
            const Dim* existing = nullptr;
            for (auto& s : *sum)
                if (same(s.name, d.name)) {
                    existing = &s;
                    break;
                }
            if (!existing)
                sum->push_back(d);
            else if (existing->len != d.len) {
                for (size_t t = 0; t < in.size(); t++) {
                    err_len[t] = -1;
                    for (auto& e : in[t])
                        if (same(e.name, d.name)) {
                            err_len[t] = e.len;
                            break;
                        }
                }
                *err_dim = d.name;
                return ORC_ERR_INCONSISTENT;
            }
        }
    return ORC_OK;
}

// src/tensors/einsum.rs:666-706 (filter_outer_and_summation_indexes) folded with the row-major
// offset of src/tensors/mod.rs:569-586 (get_index_direct).  Note the reference lets a summation
// match OVERWRITE an outer match (both loops run); a name cannot be in both lists, so it is moot.
int64_t input_offset(const std::vector<Dim>& shape, const std::vector<int64_t>& strides,
                     const std::vector<Dim>& outer, const std::vector<int64_t>& outer_idx,
                     const std::vector<Dim>& sum, const std::vector<int64_t>& sum_idx) {
    int64_t off = 0;
    for (size_t d = 0; d < shape.size(); d++) {
        int64_t idx = 0;
        for (size_t o = 0; o < outer.size(); o++)
            if (same(outer[o].name, shape[d].name)) {
                idx = outer_idx[o];
                break;
            }
        for (size_t s = 0; s < sum.size(); s++)
            if (same(sum[s].name, shape[d].name)) {
                idx = sum_idx[s];
                break;
            }
        off += idx * strides[d];
    }
    return off;
}

std::vector<int64_t> row_major_strides(const std::vector<Dim>& shape) {  // src/tensors/mod.rs:563
    std::vector<int64_t> s(shape.size(), 1);
    for (int d = (int)shape.size() - 2; d >= 0; d--) s[d] = s[d + 1] * shape[d + 1].len;
    return s;
}

// odometer with the last dimension fastest: ShapeIterator / DynamicShapeIterator,
// src/tensors/indexing.rs:735-768,876-933.  Returns false when exhausted.
bool advance(std::vector<int64_t>& idx, const std::vector<Dim>& shape) {
    for (int d = (int)shape.size() - 1; d >= 0; d--) {
        if (++idx[d] < shape[d].len) return true;
        idx[d] = 0;
    }
    return false;
}

template <class T>
int einsum2(const T* a, int da, const char* const* an, const int64_t* al, const T* b, int db,
            const char* const* bn, const int64_t* bl, int dout, const char* const* on, int64_t* out_len, T* out,
            const char** err_dim, int64_t* err_len) {
    std::vector<std::vector<Dim>> in(2);
    for (int i = 0; i < da; i++) in[0].push_back({an[i], al[i]});
    for (int i = 0; i < db; i++) in[1].push_back({bn[i], bl[i]});
    std::vector<const char*> onames(on, on + dout);
    // :924 output_shape_for (:563-572)
    std::vector<Dim> oshape;
    for (int o = 0; o < dout; o++) {
        int64_t len = 0;
        int rc = length_of(on[o], in, &len, err_len);
        if (rc != ORC_OK) {
            if (err_dim) *err_dim = on[o];
            return rc;
        }
        oshape.push_back({on[o], len});
        if (out_len) out_len[o] = len;
    }
    // :927
    std::vector<Dim> sum;
    const char* ed = nullptr;
    int rc = summation_dimensions(in, onames, &sum, &ed, err_len);
    if (rc != ORC_OK) {
        if (err_dim) *err_dim = ed;
        return rc;
    }
    if (!out) return ORC_OK;
    int64_t total = 1;
    for (auto& d : oshape) total *= d.len;
    if (total == 0) return ORC_OK;
    bool sum_valid = true;  // is_starting_index_valid: every summation length > 0
    for (auto& s : sum)
        if (s.len == 0) sum_valid = false;
    auto sa = row_major_strides(in[0]), sb = row_major_strides(in[1]);
    // An input dimension that is neither in the output nor a summation dimension cannot exist
    // (every non-output name becomes a summation dimension), so the reference's "Expected to find
    // an index" panic (:693-703) is unreachable from Einsum2::to.
    std::vector<int64_t> oidx(dout, 0);
    int64_t flat = 0;
    do {  // :931 row-major walk of the output
        T acc = T(0);  // :932
        if (sum.empty()) {
            std::vector<int64_t> none;
            T p1 = a[input_offset(in[0], sa, oshape, oidx, sum, none)];
            T p2 = b[input_offset(in[1], sb, oshape, oidx, sum, none)];
            T p = p1 * p2;
            acc = acc + p;  // :945
        } else if (sum_valid) {
            std::vector<int64_t> sidx(sum.size(), 0);
            do {
                T p1 = a[input_offset(in[0], sa, oshape, oidx, sum, sidx)];
                T p2 = b[input_offset(in[1], sb, oshape, oidx, sum, sidx)];
                T p = p1 * p2;
                acc = acc + p;  // :968
            } while (advance(sidx, sum));
        }
        out[flat++] = acc;
    } while (dout > 0 && advance(oidx, oshape));
    return ORC_OK;
}

// Einsum1::to / Einsum3..6::to, src/tensors/einsum.rs:829-883 and :1019-1666: the same loop nest over n_ops
// inputs; the term is the left-associated product ((p1 * p2) * p3) ..., each multiply separately rounded.
template <class T>
int einsum_n(int n_ops, const T* const* x, const int* ranks, const char* const* const* names,
             const int64_t* const* lens, int dout, const char* const* on, int64_t* out_len, T* out,
             const char** err_dim, int64_t* err_len) {
    std::vector<std::vector<Dim>> in(n_ops);
    for (int i = 0; i < n_ops; i++)
        for (int d = 0; d < ranks[i]; d++) in[i].push_back({names[i][d], lens[i][d]});
    std::vector<const char*> onames(on, on + dout);
    std::vector<Dim> oshape;
    for (int o = 0; o < dout; o++) {
        int64_t len = 0;
        int rc = length_of(on[o], in, &len, err_len);
        if (rc != ORC_OK) {
            if (err_dim) *err_dim = on[o];
            return rc;
        }
        oshape.push_back({on[o], len});
        if (out_len) out_len[o] = len;
    }
    std::vector<Dim> sum;
    const char* ed = nullptr;
    int rc = summation_dimensions(in, onames, &sum, &ed, err_len);
    if (rc != ORC_OK) {
        if (err_dim) *err_dim = ed;
        return rc;
    }
    if (!out) return ORC_OK;
    int64_t total = 1;
    for (auto& d : oshape) total *= d.len;
    if (total == 0) return ORC_OK;
    bool sum_valid = true;
    for (auto& sd : sum)
        if (sd.len == 0) sum_valid = false;
    std::vector<std::vector<int64_t>> strides;
    for (int i = 0; i < n_ops; i++) strides.push_back(row_major_strides(in[i]));
    std::vector<int64_t> oidx(dout, 0);
    int64_t flat = 0;
    auto term = [&](const std::vector<int64_t>& sidx) {
        T p = x[0][input_offset(in[0], strides[0], oshape, oidx, sum, sidx)];
        for (int i = 1; i < n_ops; i++) {
            T q = x[i][input_offset(in[i], strides[i], oshape, oidx, sum, sidx)];
            p = p * q;
        }
        return p;
    };
    do {
        T acc = T(0);
        if (sum.empty()) {
            std::vector<int64_t> none;
            T p = term(none);
            acc = acc + p;
        } else if (sum_valid) {
            std::vector<int64_t> sidx(sum.size(), 0);
            do {
                T p = term(sidx);
                acc = acc + p;
            } while (advance(sidx, sum));
        }
        out[flat++] = acc;
    } while (dout > 0 && advance(oidx, oshape));
    return ORC_OK;
}

// ---- A5 -----------------------------------------------------------------------------------
template <class T>
int vector_dot(const T* a, const char* an, int64_t al, const T* b, const char* bn, int64_t bl, T* out) {
    // assert_same_dimensions (src/tensors/operations.rs:1783 via dimensions.rs) then :464-471
    if (!same(an, bn) || al != bl) {
        const char* n1[1] = {an};
        const char* n2[1] = {bn};
        return panic("Dimensions of left and right tensors are not the same: (left: " + shape_debug(n1, &al, 1) +
                     ", right: " + shape_debug(n2, &bl, 1) + ")");
    }
    *out = scalar_product<T>(a, 1, b, 1, al);
    return ORC_OK;
}

}  // namespace

// ---- A9: the literal scalar tape -----------------------------------------------------------
struct orc_tape {
    int elem;
    // src/differentiation.rs:347-352 Operation{left_parent,right_parent,left_derivative,right_derivative}
    std::vector<int64_t> lp, rp;
    std::vector<float> ld32, rd32;
    std::vector<double> ld64, rd64;
    int64_t len() const { return (int64_t)lp.size(); }
    int64_t push(int64_t l, double dl, int64_t r, double dr) {
        int64_t i = len();
        lp.push_back(l);
        rp.push_back(r);
        if (elem == 4) {
            ld32.push_back((float)dl);
            rd32.push_back((float)dr);
        } else {
            ld64.push_back(dl);
            rd64.push_back(dr);
        }
        return i;
    }
};

namespace {

template <class T>
struct Tape {
    orc_tape* t;
    int64_t nullary() {  // :811-826 both parents = own index, derivatives 0
        int64_t i = t->len();
        return t->push(i, 0.0, i, 0.0);
    }
    int64_t unary(int64_t p, T d) {  // :836-850 right parent = own index, right derivative 0
        int64_t i = t->len();
        return t->push(p, (double)d, i, 0.0);
    }
    int64_t binary(int64_t l, T dl, int64_t r, T dr) { return t->push(l, (double)dl, r, (double)dr); }  // :862-878
};

// record_scalar_product, container_operations.rs:1263-1315.  The map is lazy and reduce pulls one
// product at a time, so the tape order per output element is mul0, mul1, add, mul2, add, ...
template <class T>
void record_scalar_product(orc_tape* tape, const T* l, const int64_t* li, int64_t ls, const T* r,
                           const int64_t* ri, int64_t rs, int64_t k, T* out, int64_t* out_idx) {
    if (!tape) {  // :1277-1283
        *out = scalar_product<T>(l, ls, r, rs, k);
        *out_idx = 0;
        return;
    }
    Tape<T> h{tape};
    T acc = l[0] * r[0];                                 // Multiplication::function :78
    int64_t acc_i = h.binary(li[0], r[0], ri[0], l[0]);  // d/dx = y, d/dy = x  (functions.rs:82-90)
    for (int64_t i = 1; i < k; i++) {
        T x = l[i * ls], y = r[i * rs];
        T p = x * y;
        int64_t p_i = h.binary(li[i * ls], y, ri[i * rs], x);
        acc = acc + p;                                 // Addition::function :28
        acc_i = h.binary(acc_i, T(1), p_i, T(1));      // :1304-1309
    }
    *out = acc;
    *out_idx = acc_i;
}

template <class T>
void record_matmul(orc_tape* tape, const T* a, const int64_t* ai, const T* b, const int64_t* bi, int64_t m,
                   int64_t k, int64_t n, T* c, int64_t* ci) {
    for (int64_t i = 0; i < m; i++)  // :1368 / :1522 row-major walk of the result
        for (int64_t j = 0; j < n; j++)
            record_scalar_product<T>(tape, a + i * k, ai + i * k, 1, b + j, bi + j, n, k, &c[i * n + j],
                                     &ci[i * n + j]);
}

// generic Mul with T = Record (src/matrices/operations.rs:165-196 over Record<T>): x*y and x+y follow
// record_operations.rs:344-376 and :151-190; a constant side (history None) is a plain number.
template <class T>
void scalar_record_matmul(orc_tape* tape, const T* a, const int64_t* ai, int at, const T* b, const int64_t* bi,
                          int bt, int64_t m, int64_t k, int64_t n, T* c, int64_t* ci) {
    Tape<T> h{tape};
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = 0; j < n; j++) {
            T acc = T(0);
            int64_t acc_i = 0;
            bool acc_tracked = false;
            for (int64_t kk = 0; kk < k; kk++) {
                T x = a[i * k + kk], y = b[kk * n + j];
                T p = x * y;
                int64_t p_i = 0;
                bool p_tracked = at || bt;
                if (at && bt)
                    p_i = h.binary(ai[i * k + kk], y, bi[kk * n + j], x);
                else if (at)
                    p_i = h.unary(ai[i * k + kk], y);  // Mul<&T> for &Record: d/dx = y
                else if (bt)
                    p_i = h.unary(bi[kk * n + j], x);  // rhs * &self.number
                if (kk == 0) {
                    acc = p;
                    acc_i = p_i;
                    acc_tracked = p_tracked;
                } else {
                    acc = acc + p;
                    if (acc_tracked && p_tracked) acc_i = h.binary(acc_i, T(1), p_i, T(1));
                    acc_tracked = acc_tracked || p_tracked;
                }
            }
            c[i * n + j] = acc;
            ci[i * n + j] = acc_i;
        }
}

// derivative rules: src/differentiation/functions.rs.  Op codes = EML_OP_* in include/easyml_b200.h.
template <class T>
void unary_rule(int op, T c, T x, T* y, T* d) {
    switch (op) {
        case 0: *y = -x; *d = -T(1); break;                                  // Negation :153-160
        case 1: *y = std::sin(x); *d = std::cos(x); break;                   // Sine :173-180
        case 2: *y = std::cos(x); *d = -std::sin(x); break;                  // Cosine :193-200
        case 3: *y = std::exp(x); *d = std::exp(x); break;                   // Exponential :213-220
        case 4: *y = std::log(x); *d = T(1) / x; break;                      // NaturalLogarithm :233-240
        case 5: *y = std::sqrt(x); *d = T(1) / ((T(1) + T(1)) * std::sqrt(x)); break;  // SquareRoot :253-260
        case 6: *y = x + c; *d = T(1); break;                                // Addition d/dx :33
        case 7: *y = x - c; *d = T(1); break;                                // Subtraction d/dx :58
        case 8: *y = x * c; *d = c; break;                                   // Multiplication d/dx :83
        case 9: *y = x / c; *d = T(1) / c; break;                            // Division d/dx :108
        case 10: *y = std::pow(x, c); *d = c * std::pow(x, c - T(1)); break; // Power d/dx :133
        case 11: *y = c - x; *d = -T(1); break;                              // sub_swapped: Subtraction d/dy :63
        case 12: *y = c / x; *d = -c / (x * x); break;                       // div_swapped: Division d/dy :113
        case 13: *y = std::pow(c, x); *d = std::pow(c, x) * std::log(c); break;  // pow swapped: Power d/dy :138
        default: *y = x; *d = T(1);
    }
}

template <class T>
void binary_rule(int op, T x, T y, T* z, T* dx, T* dy) {
    switch (op) {
        case 32: *z = x + y; *dx = T(1); *dy = T(1); break;
        case 33: *z = x - y; *dx = T(1); *dy = -T(1); break;
        case 34: *z = x * y; *dx = y; *dy = x; break;
        case 35: *z = x / y; *dx = T(1) / y; *dy = -x / (y * y); break;
        case 36: *z = std::pow(x, y); *dx = y * std::pow(x, y - T(1)); *dy = std::pow(x, y) * std::log(x); break;
        default: *z = x; *dx = T(1); *dy = T(0);
    }
}

template <class T>
void record_unary(orc_tape* tape, int op, T c, const T* x, const int64_t* xi, int64_t n, T* y, int64_t* yi) {
    Tape<T> h{tape};
    for (int64_t i = 0; i < n; i++) {  // container_record/mod.rs:660-665
        T d;
        unary_rule<T>(op, c, x[i], &y[i], &d);
        yi[i] = tape ? h.unary(xi[i], d) : 0;  // history None -> constants (:852)
    }
}

template <class T>
void record_binary(orc_tape* tape, int op, int mode, const T* x, const int64_t* xi, const T* y, const int64_t* yi,
                   int64_t n, T* z, int64_t* zi) {
    Tape<T> h{tape};
    for (int64_t i = 0; i < n; i++) {
        T dx, dy;
        binary_rule<T>(op, x[i], y[i], &z[i], &dx, &dy);
        if (!tape || mode == 0)
            zi[i] = 0;  // (None, None) :984
        else if (mode == 3)
            zi[i] = h.binary(xi[i], dx, yi[i], dy);  // :693-699
        else if (mode == 1)
            zi[i] = h.unary(xi[i], dx);  // binary_x_history :726-731
        else
            zi[i] = h.unary(yi[i], dy);  // binary_y_history :759-764
    }
}

// Record::try_derivatives, src/differentiation.rs:623-648
template <class T>
void derivatives(const orc_tape* t, const std::vector<T>& ld, const std::vector<T>& rd, int64_t index, T* d) {
    int64_t n = t->len();
    for (int64_t i = 0; i < n; i++) d[i] = T(0);  // :627
    d[index] = T(1);                              // :630
    for (int64_t i = n - 1; i >= 0; i--) {        // :633
        T di = d[i];
        T l = di * ld[i];
        d[t->lp[i]] = d[t->lp[i]] + l;  // :641-642
        T r = di * rd[i];
        d[t->rp[i]] = d[t->rp[i]] + r;  // :643-644
    }
}

}  // namespace

// covariance_*_features, src/linear_algebra.rs:615-639 / :676-700 / :778-832.  The reference recomputes both
// feature means inside the closure of every (i, j); the values are the same each time, so they are computed
// once here -- the operation sequence per number is unchanged.
template <class T>
static void covariance(const T* x, int64_t rows, int64_t cols, int axis, T* out) {
    const int64_t F = axis == 1 ? cols : rows, S = axis == 1 ? rows : cols;
    auto at = [&](int64_t s, int64_t f) { return axis == 1 ? x[s * cols + f] : x[f * cols + s]; };
    const T samples = (T)S;  // T::from_usize
    std::vector<T> mean((size_t)F);
    for (int64_t f = 0; f < F; f++) {
        T sum = T(0);
        for (int64_t s = 0; s < S; s++) sum = sum + at(s, f);  // .sum::<T>()
        mean[f] = sum / samples;
    }
    for (int64_t i = 0; i < F; i++)
        for (int64_t j = 0; j < F; j++) {
            T sum = T(0);
            for (int64_t s = 0; s < S; s++) sum = sum + (at(s, i) - mean[i]) * (at(s, j) - mean[j]);
            out[i * F + j] = sum / samples;
        }
}

extern "C" {

const char* orc_last_error(void) { return g_error.c_str(); }

float orc_scalar_product_f32(const float* l, int64_t ls, const float* r, int64_t rs, int64_t k) {
    return scalar_product<float>(l, ls, r, rs, k);
}
double orc_scalar_product_f64(const double* l, int64_t ls, const double* r, int64_t rs, int64_t k) {
    return scalar_product<double>(l, ls, r, rs, k);
}

int orc_matrix_mul_f32(const float* a, int64_t ar, int64_t ac, const float* b, int64_t br, int64_t bc, float* c,
                       int64_t row0, int64_t row1, int threads) {
    return matrix_mul<float>(a, ar, ac, b, br, bc, c, row0, row1, threads);
}
void orc_matrix_mul_block_f32(const float* a, int64_t ac, const float* b, int64_t bc, float* c, int64_t row0,
                              int64_t row1, int64_t col0, int64_t col1, int threads) {
    matrix_mul_block<float>(a, ac, b, bc, c, row0, row1, col0, col1, threads);
}
void orc_matrix_mul_block_f64(const double* a, int64_t ac, const double* b, int64_t bc, double* c, int64_t row0,
                              int64_t row1, int64_t col0, int64_t col1, int threads) {
    matrix_mul_block<double>(a, ac, b, bc, c, row0, row1, col0, col1, threads);
}
int orc_matrix_mul_f64(const double* a, int64_t ar, int64_t ac, const double* b, int64_t br, int64_t bc,
                       double* c, int64_t row0, int64_t row1, int threads) {
    return matrix_mul<double>(a, ar, ac, b, br, bc, c, row0, row1, threads);
}

int orc_tensor_mul_f32(const float* a, const char* const* an, const int64_t* al, const float* b,
                       const char* const* bn, const int64_t* bl, float* c, const char** on) {
    return tensor_mul<float>(a, an, al, b, bn, bl, c, on);
}
int orc_tensor_mul_f64(const double* a, const char* const* an, const int64_t* al, const double* b,
                       const char* const* bn, const int64_t* bl, double* c, const char** on) {
    return tensor_mul<double>(a, an, al, b, bn, bl, c, on);
}

int orc_einsum_n_f32(int n_ops, const float* const* x, const int* ranks, const char* const* const* names,
                     const int64_t* const* lens, int dout, const char* const* on, int64_t* out_len, float* out,
                     const char** err_dim, int64_t* err_len) {
    return einsum_n<float>(n_ops, x, ranks, names, lens, dout, on, out_len, out, err_dim, err_len);
}
int orc_einsum_n_f64(int n_ops, const double* const* x, const int* ranks, const char* const* const* names,
                     const int64_t* const* lens, int dout, const char* const* on, int64_t* out_len, double* out,
                     const char** err_dim, int64_t* err_len) {
    return einsum_n<double>(n_ops, x, ranks, names, lens, dout, on, out_len, out, err_dim, err_len);
}
int orc_einsum2_f32(const float* a, int da, const char* const* an, const int64_t* al, const float* b, int db,
                    const char* const* bn, const int64_t* bl, int dout, const char* const* on, int64_t* ol,
                    float* out, const char** ed, int64_t* el) {
    return einsum2<float>(a, da, an, al, b, db, bn, bl, dout, on, ol, out, ed, el);
}
int orc_einsum2_f64(const double* a, int da, const char* const* an, const int64_t* al, const double* b, int db,
                    const char* const* bn, const int64_t* bl, int dout, const char* const* on, int64_t* ol,
                    double* out, const char** ed, int64_t* el) {
    return einsum2<double>(a, da, an, al, b, db, bn, bl, dout, on, ol, out, ed, el);
}

int orc_vector_dot_f32(const float* a, const char* an, int64_t al, const float* b, const char* bn, int64_t bl,
                       float* out) {
    return vector_dot<float>(a, an, al, b, bn, bl, out);
}
int orc_vector_dot_f64(const double* a, const char* an, int64_t al, const double* b, const char* bn, int64_t bl,
                       double* out) {
    return vector_dot<double>(a, an, al, b, bn, bl, out);
}

void orc_covariance_f32(const float* x, int64_t rows, int64_t cols, int axis, float* out) {
    covariance<float>(x, rows, cols, axis, out);
}
void orc_covariance_f64(const double* x, int64_t rows, int64_t cols, int axis, double* out) {
    covariance<double>(x, rows, cols, axis, out);
}

orc_tape* orc_tape_new(int elem_size) {
    auto* t = new orc_tape();
    t->elem = elem_size == 4 ? 4 : 8;
    return t;
}
void orc_tape_free(orc_tape* t) { delete t; }
int64_t orc_tape_len(const orc_tape* t) { return t->len(); }
void orc_tape_clear(orc_tape* t) {
    t->lp.clear();
    t->rp.clear();
    t->ld32.clear();
    t->rd32.clear();
    t->ld64.clear();
    t->rd64.clear();
}
int64_t orc_tape_append_nullary(orc_tape* t) {
    int64_t i = t->len();
    return t->push(i, 0.0, i, 0.0);
}
int64_t orc_tape_append_nullary_repeating(orc_tape* t, int64_t n) {  // src/differentiation.rs:708-727
    int64_t start = t->len();
    for (int64_t i = 0; i < n; i++) t->push(start + i, 0.0, start + i, 0.0);
    return start;
}
int64_t orc_tape_append_unary(orc_tape* t, int64_t parent, double d) {
    int64_t i = t->len();
    return t->push(parent, d, i, 0.0);
}
int64_t orc_tape_append_binary(orc_tape* t, int64_t lp, double ld, int64_t rp, double rd) {
    return t->push(lp, ld, rp, rd);
}
void orc_tape_get(const orc_tape* t, int64_t i, int64_t* lp, int64_t* rp, double* ld, double* rd) {
    *lp = t->lp[i];
    *rp = t->rp[i];
    *ld = t->elem == 4 ? (double)t->ld32[i] : t->ld64[i];
    *rd = t->elem == 4 ? (double)t->rd32[i] : t->rd64[i];
}
void orc_tape_derivatives(const orc_tape* t, int64_t index, void* out) {
    if (t->elem == 4)
        derivatives<float>(t, t->ld32, t->rd32, index, (float*)out);
    else
        derivatives<double>(t, t->ld64, t->rd64, index, (double*)out);
}

void orc_record_matmul_f32(orc_tape* tape, const float* a, const int64_t* ai, const float* b, const int64_t* bi,
                           int64_t m, int64_t k, int64_t n, float* c, int64_t* ci) {
    record_matmul<float>(tape, a, ai, b, bi, m, k, n, c, ci);
}
void orc_record_matmul_f64(orc_tape* tape, const double* a, const int64_t* ai, const double* b,
                           const int64_t* bi, int64_t m, int64_t k, int64_t n, double* c, int64_t* ci) {
    record_matmul<double>(tape, a, ai, b, bi, m, k, n, c, ci);
}
void orc_scalar_record_matmul_f64(orc_tape* tape, const double* a, const int64_t* ai, int at, const double* b,
                                  const int64_t* bi, int bt, int64_t m, int64_t k, int64_t n, double* c,
                                  int64_t* ci) {
    scalar_record_matmul<double>(tape, a, ai, at, b, bi, bt, m, k, n, c, ci);
}
void orc_scalar_record_matmul_f32(orc_tape* tape, const float* a, const int64_t* ai, int at, const float* b,
                                  const int64_t* bi, int bt, int64_t m, int64_t k, int64_t n, float* c,
                                  int64_t* ci) {
    scalar_record_matmul<float>(tape, a, ai, at, b, bi, bt, m, k, n, c, ci);
}

void orc_record_unary_f32(orc_tape* tape, int op, float c, const float* x, const int64_t* xi, int64_t n,
                          float* y, int64_t* yi) {
    record_unary<float>(tape, op, c, x, xi, n, y, yi);
}
void orc_record_unary_f64(orc_tape* tape, int op, double c, const double* x, const int64_t* xi, int64_t n,
                          double* y, int64_t* yi) {
    record_unary<double>(tape, op, c, x, xi, n, y, yi);
}
void orc_record_binary_f32(orc_tape* tape, int op, int mode, const float* x, const int64_t* xi, const float* y,
                           const int64_t* yi, int64_t n, float* z, int64_t* zi) {
    record_binary<float>(tape, op, mode, x, xi, y, yi, n, z, zi);
}
void orc_record_binary_f64(orc_tape* tape, int op, int mode, const double* x, const int64_t* xi,
                           const double* y, const int64_t* yi, int64_t n, double* z, int64_t* zi) {
    record_binary<double>(tape, op, mode, x, xi, y, yi, n, z, zi);
}

}  // extern "C"
