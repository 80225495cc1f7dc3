//! B200-native drop-in for the batched interpolation path of `ndarray-interp`.
//!
//! Same public surface as the reference crate (`Interp1D`/`Interp2D` builders, the
//! `Interp1DStrategy`/`Interp2DStrategy` traits, `Linear`, `CubicSpline`, `Bilinear`,
//! `interp_array`/`interp_into`, `BuilderError`/`InterpolateError`); every number is produced by
//! the CUDA kernels behind `include/ndi_b200.h`.  There is no CPU fallback.
//!
//! **Status: source only.**  The environment this was written in has no Rust toolchain, so this
//! crate has not been compiled or tested; the tested boundary is the C ABI, and the Python mirror
//! (`ndarray_interp_b200/`) carries the transcribed reference test-suite.
use thiserror::Error;

mod elem;
mod ffi;
pub mod interp1d;
pub mod interp2d;
pub mod vector_extensions;

pub use elem::NdiElem;

/// Errors during Interpolator creation
#[derive(Debug, Error)]
pub enum BuilderError {
    /// Insufficient data for the chosen interpolation strategy
    #[error("{0}")]
    NotEnoughData(String),
    /// A interpolation axis is not strict monotonic rising
    #[error("{0}")]
    Monotonic(String),
    #[error("{0}")]
    ShapeError(String),
    #[error("{0}")]
    ValueError(String),
}

/// Errors during Interpolation
#[derive(Debug, Error)]
pub enum InterpolateError {
    #[error("{0}")]
    OutOfBounds(String),
}

/// bind this process to one GPU (one process per GPU; call with LOCAL_RANK)
pub fn set_device(device: i32) {
    let st = unsafe { ffi::ndi_set_device(device) };
    assert!(st == ffi::NDI_OK, "ndi_set_device failed: {}", ffi::last_error());
}
