"""BuilderError / InterpolateError with the reference's variants (src/lib.rs:127-146).

Rust enum variants are exposed as nested exception classes so that tests read like the
reference's: `pytest.raises(BuilderError.NotEnoughData)`.  Where the reference panics, the
mirror raises `Panic` carrying the reference's message.
"""


class BuilderError(Exception):
    """Errors during Interpolator creation"""


class _NotEnoughData(BuilderError):
    """Insufficient data for the chosen interpolation strategy"""


class _Monotonic(BuilderError):
    """A interpolation axis is not strict monotonic rising"""


class _ShapeError(BuilderError):
    pass


class _ValueError(BuilderError):
    pass


_NotEnoughData.__name__ = _NotEnoughData.__qualname__ = "BuilderError.NotEnoughData"
_Monotonic.__name__ = _Monotonic.__qualname__ = "BuilderError.Monotonic"
_ShapeError.__name__ = _ShapeError.__qualname__ = "BuilderError.ShapeError"
_ValueError.__name__ = _ValueError.__qualname__ = "BuilderError.ValueError"
BuilderError.NotEnoughData = _NotEnoughData
BuilderError.Monotonic = _Monotonic
BuilderError.ShapeError = _ShapeError
BuilderError.ValueError = _ValueError


class InterpolateError(Exception):
    """Errors during Interpolation"""


class _OutOfBounds(InterpolateError):
    pass


_OutOfBounds.__name__ = _OutOfBounds.__qualname__ = "InterpolateError.OutOfBounds"
InterpolateError.OutOfBounds = _OutOfBounds


class Panic(RuntimeError):
    """the reference panics here (assert!, unimplemented!, index out of bounds)"""


def rust_debug(v):
    """`{:?}` of a scalar: floats always show a fractional part, ints do not"""
    import numpy as np
    if isinstance(v, (float, np.floating)):
        f = float(v)
        if f != f:
            return "NaN"
        if f in (float("inf"), float("-inf")):
            return "inf" if f > 0 else "-inf"
        if isinstance(v, np.float32):
            r = np.format_float_positional(v, unique=True, trim="0")
        else:
            r = repr(f)
        if "e" in r or "E" in r:
            r = np.format_float_positional(v, unique=True, trim="0")
        return r if "." in r else r + ".0"
    return str(int(v))
