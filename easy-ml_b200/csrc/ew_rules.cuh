// ew_rules.cuh -- the elementwise value + derivative rules of the record containers (reference
// src/differentiation/functions.rs) and the 16-byte vector helpers, shared by the one-op kernels (elementwise.cu)
// and the fused-group interpreter (ew_fused.cu).  Both translation units are compiled -fmad=false, so a rule
// evaluated inside a fused program yields the same bits as the one-op kernel.
#pragma once
#include <cstdint>

#include "easyml_b200.h"

namespace eml {
namespace {

// 16-byte vectors: float4 / double2
template <class T>
struct Vec;
template <>
struct Vec<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <>
struct Vec<double> {
    using type = double2;
    static constexpr int N = 2;
};

template <class T>
__device__ __forceinline__ void vload(const T* p, T (&v)[Vec<T>::N]) {
    typename Vec<T>::type t = *reinterpret_cast<const typename Vec<T>::type*>(p);
    const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int i = 0; i < Vec<T>::N; i++) v[i] = e[i];
}
template <class T>
__device__ __forceinline__ void vstore(T* p, const T (&v)[Vec<T>::N]) {
    typename Vec<T>::type t;
    T* e = reinterpret_cast<T*>(&t);
#pragma unroll
    for (int i = 0; i < Vec<T>::N; i++) e[i] = v[i];
    *reinterpret_cast<typename Vec<T>::type*>(p) = t;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- rules (reference src/differentiation/functions.rs) ----
template <class T, int OP>
__device__ __forceinline__ void unary_rule(T c, T x, T& y, T& d) {
    if (OP == EML_OP_NEG) { y = -x; d = -T(1); }
    else if (OP == EML_OP_SIN) { y = sin(x); d = cos(x); }
    else if (OP == EML_OP_COS) { y = cos(x); d = -sin(x); }
    else if (OP == EML_OP_EXP) { y = exp(x); d = y; }
    else if (OP == EML_OP_LN) { y = log(x); d = T(1) / x; }
    else if (OP == EML_OP_SQRT) { y = sqrt(x); d = T(1) / ((T(1) + T(1)) * y); }
    else if (OP == EML_OP_ADD_C) { y = x + c; d = T(1); }
    else if (OP == EML_OP_SUB_C) { y = x - c; d = T(1); }
    else if (OP == EML_OP_MUL_C) { y = x * c; d = c; }
    else if (OP == EML_OP_DIV_C) { y = x / c; d = T(1) / c; }
    else if (OP == EML_OP_POW_C) { y = pow(x, c); d = c * pow(x, c - T(1)); }
    else if (OP == EML_OP_C_SUB) { y = c - x; d = -T(1); }
    else if (OP == EML_OP_C_DIV) { y = c / x; d = -c / (x * x); }
    else if (OP == EML_OP_C_POW) { y = pow(c, x); d = y * log(c); }
}

template <class T, int OP>
__device__ __forceinline__ void binary_rule(T x, T y, T& z, T& dx, T& dy) {
    if (OP == EML_OP_ADD) { z = x + y; dx = T(1); dy = T(1); }
    else if (OP == EML_OP_SUB) { z = x - y; dx = T(1); dy = -T(1); }
    else if (OP == EML_OP_MUL) { z = x * y; dx = y; dy = x; }
    else if (OP == EML_OP_DIV) { z = x / y; dx = T(1) / y; dy = -x / (y * y); }
    else if (OP == EML_OP_POW) { z = pow(x, y); dx = y * pow(x, y - T(1)); dy = z * log(x); }
}

}  // namespace
}  // namespace eml
