// tf32_split.cuh -- the big/small TF32 split of the 3xTF32 f32 GEMM, shared by the pack pass (gemm_tf32x3.cu) and the
// in-kernel converters (gemm_tf32x3_fused.cu) so that both produce the same bits.
//
// big = x with its low 13 mantissa bits cleared -- which is exactly what tcgen05.mma.kind::tf32 does to a 32-bit operand
// word (scripts/probe/tf32_operand_probe.cu: 128 of 128 values truncated, none rounded).  So the RAW f32 tile can serve
// as the big operand: the in-kernel-split GEMM never writes a big tile, and the pack pass writing `big` explicitly gives
// the same products bit for bit.  small = rna(x - big): the difference is exact (|x - big| < 2^-10 |x|), rounded to TF32.
// Error of the three products kept (As Bb + Ab Bs + Ab Bb) against the exact a b: the dropped As Bs term, < 2^-20 |a b|,
// plus the rounding of the small parts, < 2^-21 |a b| each -- 1.9e-6 sum|a b| worst case (the contract is 1e-5).
#pragma once
#include <cstdint>

namespace eml {
namespace tf32 {

// round to nearest, ties away from zero, to a 10-bit mantissa (what cvt.rna.tf32.f32 computes for finite values):
// an integer add on the magnitude bits.  Only valid below the top binade (the carry must not reach the exponent's
// all-ones pattern) -- callers check.
__device__ __forceinline__ float rna_finite(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// x = big + small with both TF32-representable.  Non-finite and near-overflow values must not manufacture NaNs the
// reference would not produce: rounding FLT_MAX-sized values up to infinity is replaced by truncation, and an
// infinite x has a zero small part (inf - inf would be NaN).
__device__ __forceinline__ void split_special(float x, float& big, float& small) {
    // |x| >= 2^127, infinite or NaN
    if ((__float_as_uint(x) & 0x7f800000u) == 0x7f000000u) {
        // top binade: round both parts toward zero so that big + small never reaches 2^128
        big = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        small = __uint_as_float(__float_as_uint(x - big) & 0xffffe000u);
        return;
    }
    // inf: no small part (inf - inf would be NaN).  NaN: the canonical quiet NaN in both parts -- a payload living only in
    // the low 13 bits would otherwise be truncated away by the tensor core and turn the NaN into an infinity
    big = (x != x) ? __uint_as_float(0x7fc00000u) : x;
    small = (x != x) ? big : 0.0f;
}

__device__ __forceinline__ bool ordinary(float x) { return fabsf(x) < 0x1p127f; }  // false for NaN

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ void split(float x, float& big, float& small) {
    if (ordinary(x)) {
        big = trunc_tf32(x);
        small = rna_finite(x - big);  // |x - big| < 2^-10 |x|: exact difference, far from overflow
    } else {
        split_special(x, big, small);
    }
}

__device__ __forceinline__ bool ordinary4(float4 v) {
    return ordinary(v.x) & ordinary(v.y) & ordinary(v.z) & ordinary(v.w);
}
// the common case: no branch, four independent dependency chains
__device__ __forceinline__ void split4_ordinary(float4 v, float4& hi, float4& lo) {
    hi.x = trunc_tf32(v.x); lo.x = rna_finite(v.x - hi.x);
    hi.y = trunc_tf32(v.y); lo.y = rna_finite(v.y - hi.y);
    hi.z = trunc_tf32(v.z); lo.z = rna_finite(v.z - hi.z);
    hi.w = trunc_tf32(v.w); lo.w = rna_finite(v.w - hi.w);
}
// the rare case: a vector holding an infinity, a NaN or a value >= 2^127 (a real function call here costs the
// converters their register allocation: measured 30 % slower)
__device__ __forceinline__ void split4_any(float4 v, float4& hi, float4& lo) {
    split(v.x, hi.x, lo.x);
    split(v.y, hi.y, lo.y);
    split(v.z, hi.z, lo.z);
    split(v.w, hi.w, lo.w);
}
// four at once: one rarely-taken branch for the whole vector keeps the common path at ~6 instructions per element
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
    if (ordinary4(v))
        split4_ordinary(v, hi, lo);
    else
        split4_any(v, hi, lo);
}

}  // namespace tf32
}  // namespace eml
