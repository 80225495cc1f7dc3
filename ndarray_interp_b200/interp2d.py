"""Host-side mirror of src/interp2d (Interp2DBuilder, Interp2D, Bilinear) over numpy arrays.
Same names, validation order (lengths BEFORE monotonicity, unlike 1-D: interp2d/mod.rs:477-509)
and error behaviour as the reference; numeric results come from the CUDA path."""
import ctypes as C

import numpy as np

from . import _lib as L
from .errors import BuilderError, InterpolateError, Panic, rust_debug
from .vector_extensions import Monotonic, get_lower_index, monotonic_prop

__all__ = ["Interp2D", "Interp2DBuilder", "Interp2DStrategy", "Interp2DStrategyBuilder", "Bilinear"]


class Interp2DStrategyBuilder:
    """trait Interp2DStrategyBuilder (interp2d/strategies/mod.rs:14-44)"""
    MINIMUM_DATA_LENGHT = 2

    def build(self, x, y, data):
        raise NotImplementedError


class Interp2DStrategy:
    """trait Interp2DStrategy (interp2d/strategies/mod.rs:46-73)"""

    def interp_into(self, interpolator, target, x, y):
        raise NotImplementedError

    def interp_batch_into(self, interpolator, xs_flat, ys_flat, out_rows):
        """the reference's batch loop (interp2d/mod.rs:297-306)"""
        for i in range(xs_flat.shape[0]):
            self.interp_into(interpolator, out_rows[i, ...], xs_flat[i], ys_flat[i])

    def _bind(self, interpolator):
        pass


class Bilinear(Interp2DStrategyBuilder, Interp2DStrategy):
    """Bilinear strategy (interp2d/strategies/bilinear.rs)"""
    MINIMUM_DATA_LENGHT = 2

    def __init__(self):
        self._extrapolate = False

    @classmethod
    def new(cls):
        return cls()

    def extrapolate(self, yes):
        self._extrapolate = bool(yes)
        return self

    def build(self, x, y, data):                           # bilinear.rs:45-52
        return self

    def interp_batch_into(self, interpolator, xs_flat, ys_flat, out_rows):
        bad, axis = C.c_int64(-1), C.c_int32(-1)
        fn, h = ((L.load().ndi_interp2d_group_bilinear, interpolator._group.ptr) if getattr(interpolator, "_group", None)
                 else (L.load().ndi_interp2d_bilinear, interpolator._handle()))
        st = L.check(fn(h, L.ptr(xs_flat), L.ptr(ys_flat), xs_flat.size, int(self._extrapolate), L.ptr(out_rows),
                        C.byref(bad), C.byref(axis)))
        if st == L.OUT_OF_BOUNDS:                          # bilinear.rs:71-80: x is checked first
            name, v = ("x", xs_flat[bad.value]) if axis.value == 0 else ("y", ys_flat[bad.value])
            raise InterpolateError.OutOfBounds(f"{name} = {rust_debug(v)} is not in range")
        if st == L.NAN_QUERY:
            raise Panic("not implemented: failed to convert NaN to usize")

    def interp_into(self, interpolator, target, x, y):     # bilinear.rs:64-99
        dt = interpolator.data.dtype
        qx, qy = np.array([x], dtype=dt), np.array([y], dtype=dt)
        tgt = np.asarray(target)
        if tgt.flags.c_contiguous and tgt.dtype == dt:
            self.interp_batch_into(interpolator, qx, qy, tgt.reshape((1,) + tgt.shape))
        else:
            tmp = np.zeros((1,) + tgt.shape, dtype=dt)
            self.interp_batch_into(interpolator, qx, qy, tmp)
            tgt[...] = tmp[0]


class _Handle2D:
    def __init__(self, x, y, data, flags=0):
        lib = L.require_device()
        self.ptr = C.c_void_p()
        w = int(np.prod(data.shape[2:], dtype=np.int64))
        if L.is_dense(x) and L.is_dense(y) and L.is_dense(data):
            L.check(lib.ndi_interp2d_create(L.dtype_code(data.dtype), L.ptr(x), len(x), L.ptr(y), len(y), L.ptr(data), w,
                                            flags, C.byref(self.ptr)))
        else:                                              # views: uploaded as they lie, made dense on the device
            xk, xp, _, xs = L.view_args(x)
            yk, yp, _, ys = L.view_args(y)
            dk, dp, shape, strides = L.view_args(data)
            L.check(lib.ndi_interp2d_create_strided(L.dtype_code(data.dtype), xp, len(x), xs[0], yp, len(y), ys[0], dp,
                                                    dk.ndim, shape, strides, flags, C.byref(self.ptr)))

    def __del__(self):
        try:
            if self.ptr:
                L.load().ndi_interp2d_destroy(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


class Interp2D:
    """Two dimensional interpolator (interp2d/mod.rs:34-48)"""

    def __init__(self, x, y, data, strategy, _handle=None):
        self.x, self.y, self.data, self.strategy = x, y, data, strategy
        self._h = _handle
        if isinstance(strategy, Interp2DStrategy):
            strategy._bind(self)

    @staticmethod
    def builder(data):
        return Interp2DBuilder(data)

    @classmethod
    def new_unchecked(cls, x, y, data, strategy):          # interp2d/mod.rs:330-342
        x, y, data = _prepare(x, y, data)
        h = _Handle2D(x, y, data, L.ASSUME_VALID) if isinstance(strategy, Bilinear) else None
        return cls(x, y, data, strategy, h)

    def replicate(self, devices):
        """NOT in the reference: fan this interpolator out over several GPUs of this process (include/ndi_b200.h:
        ndi_interp2d_replicate); batch calls cut the queries into contiguous blocks, one per device."""
        from .interp1d import _Group
        if not isinstance(self.strategy, Bilinear):
            raise TypeError("only the built-in strategies run on the device")
        other = Interp2D.__new__(Interp2D)
        other.x, other.y, other.data, other.strategy, other._h = self.x, self.y, self.data, self.strategy, self._h
        other._group = _Group(L.load().ndi_interp2d_replicate, L.load().ndi_interp2d_group_destroy, self._handle(), devices, self)
        return other

    def _handle(self):
        if self._h is None:
            self._h = _Handle2D(self.x, self.y, self.data, L.ASSUME_VALID)
        return self._h.ptr

    def interp_scalar(self, x, y):                         # interp2d/mod.rs:107-113
        if self.data.ndim != 2:
            raise TypeError("interp_scalar needs 2-D data (Ix2)")
        buf = np.zeros((), dtype=self.data.dtype)
        dt = self.data.dtype.type
        self.strategy.interp_into(self, buf, dt(x), dt(y))
        return buf[()]

    def interp(self, x, y):                                # interp2d/mod.rs:132-146
        target = np.zeros(self.data.shape[2:], dtype=self.data.dtype)
        dt = self.data.dtype.type
        self.strategy.interp_into(self, target, dt(x), dt(y))
        return target

    def interp_into(self, x, y, buffer):                   # interp2d/mod.rs:160-167
        expect = list(self.data.shape[2:])
        if list(np.shape(buffer)) != expect:
            raise Panic(f"Zip: Producer dimension mismatch, expected: {expect}, got: {list(np.shape(buffer))}")
        dt = self.data.dtype.type
        self.strategy.interp_into(self, buffer, dt(x), dt(y))

    def interp_array(self, xs, ys):                        # interp2d/mod.rs:175-196
        xs, ys = np.asarray(xs), np.asarray(ys)
        if xs.shape != ys.shape:
            raise Panic("`xs.shape()` and `ys.shape()` do not match")
        zs = np.zeros(self._buffer_shape(xs.shape), dtype=self.data.dtype)
        self.interp_array_into(xs, ys, zs)
        return zs

    def interp_array_into(self, xs, ys, buffer):           # interp2d/mod.rs:215-285
        xs, ys = np.asarray(xs), np.asarray(ys)
        if xs.shape != ys.shape:
            raise Panic("`xs.shape()` and `ys.shape()` do not match")
        expect = self._buffer_shape(xs.shape)
        got = tuple(np.shape(buffer))
        if got != expect:
            from .interp1d import _panic_buffer_shape
            _panic_buffer_shape(xs.shape, self.data.shape[2:], expect, got)
        dt = self.data.dtype
        qx = np.ascontiguousarray(xs, dtype=dt).reshape(-1)
        qy = np.ascontiguousarray(ys, dtype=dt).reshape(-1)
        rows_shape = (qx.size,) + self.data.shape[2:]
        direct = isinstance(buffer, np.ndarray) and buffer.flags.c_contiguous and buffer.dtype == dt
        # dense stand-in for a strided buffer starts from the caller's values: rows at and after a failing query
        # come back untouched (interp2d/mod.rs:255-307)
        rows = buffer.reshape(rows_shape) if direct else np.array(buffer, dtype=dt).reshape(rows_shape)
        try:
            self.strategy.interp_batch_into(self, qx, qy, rows)
        finally:
            if not direct:
                buffer[...] = rows.reshape(expect)

    def _buffer_shape(self, query_shape):                  # interp2d/mod.rs:310-321
        return tuple(query_shape) + tuple(self.data.shape[2:])

    def index_point(self, x_idx, y_idx):                   # interp2d/mod.rs:348-364
        return self.x[x_idx], self.y[y_idx], self.data[x_idx, y_idx]

    def get_index_left_of(self, x, y):                     # interp2d/mod.rs:370-372
        return get_lower_index(self.x, x), get_lower_index(self.y, y)

    def is_in_x_range(self, x):
        return bool(self.x[0] <= x <= self.x[-1])

    def is_in_y_range(self, y):
        return bool(self.y[0] <= y <= self.y[-1])


def _prepare(x, y, data):
    data = np.asarray(data)                                # views stay views (_Handle2D)
    L.dtype_code(data.dtype)
    x, y = np.asarray(x), np.asarray(y)
    return (x if x.dtype == data.dtype else x.astype(data.dtype)), (y if y.dtype == data.dtype else y.astype(data.dtype)), data


class Interp2DBuilder:
    """Create and configure a Interp2D interpolator (interp2d/mod.rs:50-64, :382-519)"""

    def __init__(self, data):                              # Interp2DBuilder::new (:388-405)
        self._data = np.asarray(data)
        if self._data.ndim < 2:
            raise Panic("index out of bounds: Interp2DBuilder::new needs data with at least 2 dimensions")
        self._x = np.arange(self._data.shape[0]).astype(self._data.dtype)
        self._y = np.arange(self._data.shape[1]).astype(self._data.dtype)
        self._strategy = Bilinear.new()

    @classmethod
    def new(cls, data):
        return cls(data)

    def strategy(self, strategy):
        self._strategy = strategy
        return self

    def x(self, x):
        self._x = np.asarray(x)
        return self

    def y(self, y):
        self._y = np.asarray(y)
        return self

    def build(self):
        """Validate the input and create the configured Interp2D (interp2d/mod.rs:468-518)"""
        data, x, y, strat = self._data, np.asarray(self._x), np.asarray(self._y), self._strategy
        m = strat.MINIMUM_DATA_LENGHT
        if data.ndim < 2:
            raise BuilderError.ShapeError("data dimension needs to be at least 2")
        if data.shape[0] < m:
            raise BuilderError.NotEnoughData("The 0-dimension has not enough data for the chosen interpolation "
                                             f"strategy. Provided: {data.shape[0]}, Reqired: {m}")
        if data.shape[1] < m:
            raise BuilderError.NotEnoughData("The 1-dimension has not enough data for the chosen interpolation "
                                             f"strategy. Provided: {data.shape[1]}, Reqired: {m}")
        if len(x) != data.shape[0]:
            raise BuilderError.ShapeError(
                f"Lenghts of x-axis and data-0-axis need to match. Got x: {len(x)}, data-0: {data.shape[0]}")
        if len(y) != data.shape[1]:
            raise BuilderError.ShapeError(
                f"Lenghts of y-axis and data-1-axis need to match. Got y: {len(y)}, data-1: {data.shape[1]}")
        L.dtype_code(data.dtype)
        x, y = x.astype(data.dtype, copy=False), y.astype(data.dtype, copy=False)
        if monotonic_prop(x) != Monotonic.Rising(True):
            raise BuilderError.Monotonic("The x-axis needs to be strictly monotonic rising")
        if monotonic_prop(y) != Monotonic.Rising(True):
            raise BuilderError.Monotonic("The y-axis needs to be strictly monotonic rising")
        x, y, data = _prepare(x, y, data)
        finished = strat.build(x, y, data)
        h = _Handle2D(x, y, data, L.ASSUME_VALID) if isinstance(finished, Bilinear) else None
        return Interp2D(x, y, data, finished, h)
