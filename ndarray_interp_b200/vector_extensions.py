"""Mirror of src/vector_extensions.rs: monotonic_prop and get_lower_index, computed on the device."""
import ctypes as C

import numpy as np

from . import _lib as L
from .errors import Panic


class Monotonic:
    """enum Monotonic { Rising{strict}, Falling{strict}, NotMonotonic } (vector_extensions.rs:24-29)"""

    def __init__(self, kind, strict=None):
        self.kind, self.strict = kind, strict

    @classmethod
    def Rising(cls, strict):
        return cls("Rising", bool(strict))

    @classmethod
    def Falling(cls, strict):
        return cls("Falling", bool(strict))

    def __eq__(self, other):
        return isinstance(other, Monotonic) and (self.kind, self.strict) == (other.kind, other.strict)

    def __hash__(self):
        return hash((self.kind, self.strict))

    def __repr__(self):
        return "NotMonotonic" if self.kind == "NotMonotonic" else f"{self.kind} {{ strict: {str(self.strict).lower()} }}"


Monotonic.NotMonotonic = Monotonic("NotMonotonic")
_FROM_CODE = [Monotonic.NotMonotonic, Monotonic.Rising(True), Monotonic.Rising(False), Monotonic.Falling(True),
              Monotonic.Falling(False)]


def monotonic_prop(x):
    """VectorExtensions::monotonic_prop (vector_extensions.rs:40-53).  `x` may be any 1-D view,
    including reversed / strided ones (vector_extensions.rs:380-384)."""
    x = np.asarray(x)
    if x.ndim != 1:
        raise TypeError("monotonic_prop needs a 1-D array")
    code = L.dtype_code(x.dtype)
    if len(x) <= 1:
        return Monotonic.NotMonotonic
    lib = L.require_device()
    if x.strides[0] % x.itemsize:
        x = np.ascontiguousarray(x)
    prop = C.c_int32(0)
    st = lib.ndi_monotonic_prop(code, C.c_void_p(x.ctypes.data), len(x), x.strides[0] // x.itemsize, C.byref(prop))
    L.check(st)
    return _FROM_CODE[prop.value]


def get_lower_index(grid, x):
    """VectorExtensions::get_lower_index (vector_extensions.rs:55-111) for a scalar or an array of
    queries.  NaN panics like the reference (vector_extensions.rs:83-84)."""
    grid = np.ascontiguousarray(grid)
    code = L.dtype_code(grid.dtype)
    if grid.ndim != 1 or len(grid) < 2:
        raise Panic("index out of bounds: get_lower_index needs a grid of at least 2 points")
    q = np.ascontiguousarray(x, dtype=grid.dtype)
    flat = q.reshape(-1)
    lib = L.require_device()
    idx = np.zeros(flat.shape, dtype=np.int64)
    bad = C.c_int64(-1)
    st = L.check(lib.ndi_lower_index(code, L.ptr(grid), len(grid), L.ptr(flat), flat.size, L.ptr(idx), C.byref(bad)))
    if st == L.NAN_QUERY:
        raise Panic("not implemented: failed to convert NaN to usize")
    return int(idx[0]) if q.ndim == 0 else idx.reshape(q.shape)
