"""Host-side mirror logic that needs no device: validation that precedes any upload, enum
mirrors, message formatting.  (The reference's builder tests that reach the monotonic check run
under -m gpu, because that check is a device kernel here.)"""
import numpy as np
import pytest

from ndarray_interp_b200 import BuilderError, InterpolateError, Panic
from ndarray_interp_b200.errors import rust_debug
from ndarray_interp_b200.interp1d import (BoundaryCondition, CubicSpline, Interp1DBuilder, Linear, RowBoundary,
                                          SingleBoundary, _panic_buffer_shape)
from ndarray_interp_b200.interp2d import Interp2D, Interp2DBuilder
from ndarray_interp_b200.vector_extensions import Monotonic, monotonic_prop


def test_not_enough_data_precedes_everything_1d():
    # tests/interp1d.rs:124-127 (i32 data) and tests/cubic_spline_strat.rs:30-35
    with pytest.raises(BuilderError.NotEnoughData):
        Interp1DBuilder.new(np.array([1], dtype=np.int32)).build()
    with pytest.raises(BuilderError.NotEnoughData):
        Interp1DBuilder(np.array([1.0, 2.0])).strategy(CubicSpline.new()).build()
    with pytest.raises(BuilderError.ShapeError):
        Interp1DBuilder(np.array(1.0)).build()


def test_builder_errors_2d_shape_checks():
    # tests/interp2d.rs:280-313: everything before the monotonic check is host logic
    i = np.int32
    with pytest.raises(BuilderError.NotEnoughData):
        Interp2D.builder(np.array([[1]], dtype=i)).build()
    with pytest.raises(BuilderError.NotEnoughData):
        Interp2D.builder(np.array([[1, 2]], dtype=i)).build()
    with pytest.raises(BuilderError.NotEnoughData):
        Interp2D.builder(np.array([[1], [2]], dtype=i)).build()
    d = np.array([[1, 2], [3, 4]], dtype=i)
    for bad in ([1], [1, 2, 3]):
        with pytest.raises(BuilderError.ShapeError):
            Interp2D.builder(d).x(np.array(bad, dtype=i)).build()
        with pytest.raises(BuilderError.ShapeError):
            Interp2D.builder(d).y(np.array(bad, dtype=i)).build()


def test_monotonic_prop_short_inputs_need_no_device():
    # vector_extensions.rs:41-43, :399-402
    assert monotonic_prop(np.array([1], dtype=np.int32)) == Monotonic.NotMonotonic
    assert monotonic_prop(np.array([], dtype=np.float64)) == Monotonic.NotMonotonic


def test_error_variants_are_distinct():
    assert issubclass(BuilderError.Monotonic, BuilderError)
    assert not issubclass(BuilderError.Monotonic, BuilderError.ShapeError)
    assert issubclass(InterpolateError.OutOfBounds, InterpolateError)
    assert BuilderError.NotEnoughData.__name__ == "BuilderError.NotEnoughData"


def test_rust_debug_formatting():
    assert rust_debug(1.0) == "1.0" and rust_debug(-0.1) == "-0.1" and rust_debug(np.int32(3)) == "3"
    assert rust_debug(float("nan")) == "NaN" and rust_debug(float("inf")) == "inf"
    assert rust_debug(np.float32(0.1)) == "0.1"


def test_buffer_shape_panic_messages():
    # pinned by src/interp1d/mod.rs:550-574
    with pytest.raises(Panic, match=r"expected: \[2\], got: \[1\]"):
        _panic_buffer_shape((2,), (4,), (2, 4), (1, 4))
    with pytest.raises(Panic, match=r"expected: \[4\], got: \[3\]"):
        _panic_buffer_shape((2,), (4,), (2, 4), (2, 3))


def test_individual_boundary_shape_error_is_host_logic():
    # tests/cubic_spline_strat.rs:413-439 ("Expected: [1, 2], got: [1, 3]" / "[2, 2]")
    y = np.array([[0.5, 1.0], [0.0, 1.5], [3.0, 0.5]])
    x = np.arange(3.0)
    b3 = BoundaryCondition.Individual([[RowBoundary.Natural, RowBoundary.Clamped, RowBoundary.NotAKnot]])
    with pytest.raises(BuilderError.ShapeError, match=r"Expected: \[1, 2\], got: \[1, 3\]"):
        CubicSpline.new().boundary(b3).build(x, y)
    b22 = BoundaryCondition.Individual([[RowBoundary.Natural, RowBoundary.NotAKnot],
                                        [RowBoundary.Natural, RowBoundary.NotAKnot]])
    with pytest.raises(BuilderError.ShapeError, match=r"Expected: \[1, 2\], got: \[2, 2\]"):
        CubicSpline.new().boundary(b22).build(x, y)


def test_boundary_enums():
    assert SingleBoundary.FirstDeriv(0.5) == SingleBoundary.FirstDeriv(0.5)
    assert RowBoundary.Mixed(SingleBoundary.NotAKnot, SingleBoundary.FirstDeriv(0.5)).kind == "Mixed"
    assert Linear.new().extrapolate(True)._extrapolate is True
    assert repr(Monotonic.Rising(True)) == "Rising { strict: true }"


def test_view_args_describe_an_ndarray_view_the_way_the_c_abi_takes_it():
    """pointer to the FIRST LOGICAL element, shape, strides in elements (ndarray's convention; numpy reports bytes)"""
    from ndarray_interp_b200 import _lib as L
    base = np.arange(4 * 6 * 5, dtype=np.float64).reshape(4, 6, 5)
    v = base[::-1, 1::2, ::-2]
    keep, ptr, shape, strides = L.view_args(v)
    assert keep is v and list(shape) == [4, 3, 3] and list(strides) == [-30, 10, -2]
    assert ptr.value == base.ctypes.data + base[3, 1, 4:].ctypes.data - base.ctypes.data     # element [3, 1, 4] of base
    assert not L.is_dense(v) and L.is_dense(base) and L.is_dense(base[1]) and not L.is_dense(base[:, 0])
    one = np.arange(10, dtype=np.int64)[::-3]
    keep, ptr, shape, strides = L.view_args(one)
    assert list(shape) == [4] and list(strides) == [-3] and ptr.value == one.ctypes.data
    # broadcast views have zero strides; a byte-misaligned view (structured field) is copied
    b = np.broadcast_to(np.arange(3.0)[:, None], (3, 4))
    assert list(L.view_args(b)[3]) == [1, 0]
    rec = np.zeros(5, dtype=[("pad", "u1"), ("v", "<f8")])["v"]
    keep, _, _, strides = L.view_args(rec)
    assert keep is not rec and list(strides) == [1]
