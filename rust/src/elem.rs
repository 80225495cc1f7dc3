//! Element types the device path is instantiated for.
use std::fmt::Debug;

use num_traits::{Num, NumCast};

/// f32, f64, i32 (the types the reference's tests exercise), i64, u32 and u64 (unsigned: arithmetic modulo 2^32 / 2^64,
/// what a release build of the reference computes).
pub trait NdiElem: Num + NumCast + PartialOrd + Copy + Debug + Send + 'static {
    /// `ndi_dtype` code of `include/ndi_b200.h`
    const DTYPE: i32;
}
impl NdiElem for f32 {
    const DTYPE: i32 = crate::ffi::NDI_F32;
}
impl NdiElem for f64 {
    const DTYPE: i32 = crate::ffi::NDI_F64;
}
impl NdiElem for i32 {
    const DTYPE: i32 = crate::ffi::NDI_I32;
}
impl NdiElem for i64 {
    const DTYPE: i32 = crate::ffi::NDI_I64;
}
impl NdiElem for u32 {
    const DTYPE: i32 = crate::ffi::NDI_U32;
}
impl NdiElem for u64 {
    const DTYPE: i32 = crate::ffi::NDI_U64;
}
