// src/numeric.rs -- the compile-time switch of the `cuda` feature.
//
// `ZeroOne` is a supertrait of `Numeric` (src/numeric.rs:134-143) that every numeric type implements by hand (the
// zero_one_* macros :198-243, Wrapping / Saturating :179-196, Trace and Record in src/differentiation), so an
// associated const with a default reaches every `T: Numeric` without touching the blanket impl of `Numeric` itself.
// `crate::cuda::Kind` is `pub` inside the private module `crate::cuda`: downstream crates see the const in the docs as
// hidden and cannot name its type, hence cannot override it for their own types.

pub trait ZeroOne: Sized {
    fn zero() -> Self;
    fn one() -> Self;
    /// Implementation detail of the `cuda` feature: which device kernel family this type takes (none by default).
    #[doc(hidden)]
    #[cfg(feature = "cuda")]
    const CUDA_KIND: crate::cuda::Kind = crate::cuda::Kind::GENERIC;
}

macro_rules! zero_one_float {
    ($T:ty, $KIND:ident) => {
        impl ZeroOne for $T {
            #[inline]
            fn zero() -> $T {
                0.0
            }
            #[inline]
            fn one() -> $T {
                1.0
            }
            #[cfg(feature = "cuda")]
            const CUDA_KIND: crate::cuda::Kind = crate::cuda::Kind::$KIND;
        }
    };
}

zero_one_float!(f32, F32);
zero_one_float!(f64, F64);
