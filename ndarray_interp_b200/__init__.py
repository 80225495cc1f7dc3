"""ndarray_interp_b200 -- B200-native batched interpolation behind ndarray-interp's API.

Layout mirrors the reference crate (src/lib.rs:117-124):
    ndarray_interp_b200.interp1d           Interp1D, Interp1DBuilder, Linear, CubicSpline, ...
    ndarray_interp_b200.interp2d           Interp2D, Interp2DBuilder, Bilinear
    ndarray_interp_b200.vector_extensions  monotonic_prop, get_lower_index, Monotonic
    ndarray_interp_b200.BuilderError / InterpolateError

All arithmetic runs in hand-written CUDA for sm_100a behind the C ABI in include/ndi_b200.h
(csrc/).  There is no CPU fallback: without the built library or without a GPU, calls fail.
"""
from . import interp1d, interp2d, vector_extensions  # noqa: F401
from .errors import BuilderError, InterpolateError, Panic  # noqa: F401

__version__ = "0.1.0"
