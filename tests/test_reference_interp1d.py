"""tests/interp1d.rs (13 tests) and the unit tests of src/interp1d/mod.rs:479-608, transcribed
test-for-test against the host mirror; every number comes from the CUDA path."""
import numpy as np
import pytest

import golden_util as G
from ndarray_interp_b200 import BuilderError, InterpolateError, Panic
from ndarray_interp_b200.interp1d import (Interp1D, Interp1DBuilder, Interp1DStrategy, Interp1DStrategyBuilder,
                                          Linear)

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def test_interp_y_only():
    interp = Interp1D.builder(np.array([1.5, 2.0, 3.0, 4.0, 5.0, 7.0, 7.0, 8.0, 9.0, 10.5])).build()
    assert interp.interp_scalar(0.0) == 1.5
    assert interp.interp_scalar(9.0) == 10.5
    assert interp.interp_scalar(4.5) == 6.0
    assert interp.interp_scalar(0.25) == 1.625
    assert interp.interp_scalar(8.75) == 10.125


def test_extrapolate_y_only():
    interp = Interp1D.builder(np.array([1.0, 2.0, 1.5])).strategy(Linear.new().extrapolate(True)).build()
    assert interp.interp_scalar(-1.0) == 0.0
    assert interp.interp_scalar(3.0) == 1.0


def test_interp_with_x_and_y():
    interp = (Interp1DBuilder.new(np.array([1.5, 2.0, 3.0, 4.0, 5.0, 7.0, 7.0, 8.0, 9.0, 10.5]))
              .x(np.array([-4.0, -3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0]))
              .strategy(Linear.new()).build())
    assert interp.interp_scalar(-4.0) == 1.5
    assert interp.interp_scalar(5.0) == 10.5
    assert interp.interp_scalar(0.5) == 6.0
    assert interp.interp_scalar(-3.75) == 1.625
    assert interp.interp_scalar(4.75) == 10.125


def test_interp_with_x_and_y_expspaced():
    interp = (Interp1DBuilder.new(np.array([1.0, 2.0, 3.0, 4.0, 5.0, 5.0, 4.0, 3.0, 2.0, 1.0]))
              .x(np.array([1.0, 2.0, 4.0, 8.0, 16.0, 32.0, 64.0, 128.0, 256.0, 512.0]))
              .strategy(Linear.new()).build())
    assert interp.interp_scalar(1.0) == 1.0
    assert interp.interp_scalar(512.0) == 1.0
    assert interp.interp_scalar(42.0) == 4.6875
    assert interp.interp_scalar(365.0) == 1.57421875


def test_extrapolate_with_x_and_y():
    interp = (Interp1DBuilder.new(np.array([1.0, 0.0, 1.5])).x(np.array([0.0, 1.0, 1.5]))
              .strategy(Linear.new().extrapolate(True)).build())
    assert interp.interp_scalar(-1.0) == 2.0
    assert interp.interp_scalar(2.0) == 3.0


def test_interp_array():
    interp = Interp1D.builder(np.array([1.0, 2.0, 3.0, 4.0, 5.0, 5.0, 4.0, 3.0, 2.0, 1.0])).build()
    x_query = np.array([[1.0, 2.0, 9.0], [4.0, 5.0, 7.5]])
    y_expect = np.array([[2.0, 3.0, 1.0], [5.0, 5.0, 2.5]])
    assert np.array_equal(interp.interp_array(x_query), y_expect)


def test_interp_y_only_out_of_bounds():
    interp = Interp1D.builder(np.array([1.0, 2.0, 3.0])).build()
    with pytest.raises(InterpolateError.OutOfBounds, match=r"x = -0\.1 is not in range"):
        interp.interp(-0.1)
    with pytest.raises(InterpolateError.OutOfBounds):
        interp.interp(9.0)


def test_interp_with_x_and_y_out_of_bounds():
    interp = (Interp1DBuilder.new(np.array([1.0, 2.0, 3.0])).x(np.array([-4.0, -3.0, 2.0]))
              .strategy(Linear.new()).build())
    with pytest.raises(InterpolateError.OutOfBounds):
        interp.interp(-4.1)
    with pytest.raises(InterpolateError.OutOfBounds):
        interp.interp(2.1)


def test_interp_builder_errors():
    i = np.int32
    with pytest.raises(BuilderError.NotEnoughData):
        Interp1DBuilder.new(np.array([1], dtype=i)).build()
    # monotonic is checked BEFORE the length match (interp1d/mod.rs:460 then :465)
    with pytest.raises(BuilderError.ShapeError):
        Interp1DBuilder.new(np.array([1, 2], dtype=i)).x(np.array([1, 2, 3], dtype=i)).build()
    with pytest.raises(BuilderError.Monotonic):
        Interp1DBuilder.new(np.array([1, 2, 3], dtype=i)).x(np.array([1, 2, 2], dtype=i)).build()
    with pytest.raises(BuilderError.Monotonic):     # both wrong: monotonic wins
        Interp1DBuilder.new(np.array([1, 2], dtype=i)).x(np.array([1, 2, 2], dtype=i)).build()


def test_interp_view_array():
    a = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0])
    interp = Interp1D.builder(a[::-1]).x(np.array([-4.0, -3.0, -2.0, -1.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0])).build()
    assert interp.interp_scalar(-4.0) == 10.0
    assert interp.interp_scalar(5.0) == 1.0
    assert interp.interp_scalar(0.0) == 6.0
    assert interp.interp_scalar(-3.5) == 9.5
    assert interp.interp_scalar(4.75) == 1.25


def test_interp_multi_fn():
    data = np.array([[0.1, 0.2, 0.3, 0.4, 0.5], [2.0, 2.0, 3.0, 4.0, 5.0], [10.0, 20.0, 30.0, 40.0, 50.0],
                     [20.0, 40.0, 60.0, 80.0, 100.0]])
    interp = Interp1DBuilder.new(data).x(np.array([1.0, 2.0, 3.0, 4.0])).build()
    res = interp.interp(1.5)
    assert np.abs(res - np.array([1.05, 1.1, 1.65, 2.2, 2.75])).max() <= EPS
    aa = interp.interp_array(np.array([[1.0, 1.5], [3.5, 4.0]]))
    assert np.abs(aa[1, 1, :] - np.array([20.0, 40.0, 60.0, 80.0, 100.0])).max() <= EPS
    expect = np.array([[[0.1, 0.2, 0.3, 0.4, 0.5], [1.05, 1.1, 1.65, 2.2, 2.75]],
                       [[15.0, 30.0, 45.0, 60.0, 75.0], [20.0, 40.0, 60.0, 80.0, 100.0]]])
    assert aa.shape == (2, 2, 5)
    assert np.abs(aa - expect).max() <= EPS


def test_interp_array_with_differnt_repr():
    interp = Interp1D.builder(np.array([1.0, 2.0, 3.0, 4.0, 5.0, 5.0, 4.0, 3.0, 2.0, 1.0])).build()
    x_query = np.array([[1.0, 2.0, 9.0], [4.0, 5.0, 7.5]])
    y_expect = np.array([[2.0, 3.0, 1.0], [5.0, 5.0, 2.5]])
    assert np.array_equal(interp.interp_array(x_query.view()), y_expect)
    assert np.array_equal(interp.interp_array(np.asfortranarray(x_query)), y_expect)


@pytest.mark.parametrize("case", G.load("linear"), ids=lambda c: c["name"])
def test_linear_golden_vectors(case):
    dt = G.DT[case["dtype"]]
    b = Interp1DBuilder.new(np.array(case["data"], dtype=dt))
    if case["x"] is not None:
        b = b.x(np.array(case["x"], dtype=dt))
    interp = b.strategy(Linear.new().extrapolate(case["extrapolate"])).build()
    out = interp.interp_array(np.array(case["query"], dtype=dt))
    exp = np.array(case["expect"], dtype=dt)
    assert out.shape == exp.shape
    if case["tol"]["abs"] == 0.0:
        assert np.array_equal(out, exp)
    else:
        assert np.abs(out - exp).max() <= case["tol"]["abs"]


# ---- src/interp1d/mod.rs:479-608 -------------------------------------------------------------------
def _rand(shape, seed=64):
    return np.random.default_rng(seed).uniform(0.0, 1.0, size=shape)


@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 6, 7])
def test_interp1d_nd(dim):                                   # test_dim! (:508-537)
    interp = Interp1D.builder(_rand((4,) * dim)).build()
    res = interp.interp(2.2)
    assert res.ndim == dim - 1
    buf = np.zeros(res.shape)
    interp.interp_into(2.2, buf)
    assert np.abs(buf - res).max(initial=0.0) <= EPS
    query = np.array([[0.5, 1.0], [1.5, 2.0]])
    res = interp.interp_array(query)
    assert res.ndim == dim - 1 + query.ndim
    buf = np.zeros(res.shape)
    interp.interp_array_into(query, buf)
    assert np.abs(buf - res).max() <= EPS


def test_interp1d_1d_scalar():
    r = Interp1D.builder(_rand(4)).build().interp_scalar(2.2)
    assert isinstance(r, np.float64)


def test_interp1d_2d_into_too_small():
    with pytest.raises(Panic, match=r"expected: \[4\], got: \[3\]"):
        Interp1D.builder(_rand((4, 4))).build().interp_into(2.2, np.zeros(3))


def test_interp1d_2d_into_too_big():
    with pytest.raises(Panic, match=r"expected: \[4\], got: \[5\]"):
        Interp1D.builder(_rand((4, 4))).build().interp_into(2.2, np.zeros(5))


@pytest.mark.parametrize("shape,msg", [((1, 4), r"expected: \[2\], got: \[1\]"), ((2, 3), None), ((3, 4), None),
                                       ((2, 5), None)])
def test_interp1d_2d_array_into_wrong_buffer(shape, msg):
    interp = Interp1D.builder(_rand((4, 4))).build()
    with pytest.raises(Panic, match=msg):
        interp.interp_array_into(np.array([2.2, 2.4]), np.zeros(shape))


# ---- examples/custom_strategy.rs ----------------------------------------------------------------------
class StepInterpolator(Interp1DStrategyBuilder, Interp1DStrategy):
    MINIMUM_DATA_LENGHT = 2

    def build(self, x, data):
        return self

    def interp_into(self, interpolator, target, x):
        idx = interpolator.get_index_left_of(x)
        x_left, data_left = interpolator.index_point(idx)
        x_right, data_right = interpolator.index_point(idx + 1)
        if (x_right - x_left) / 2.0 > (x - x_left):
            target[...] = data_left
        else:
            target[...] = data_right


def test_custom_strategy_example():
    data = np.array([2.0, 4.0, 5.0])
    query = G.linspace(-0.5, 2.5, 6)
    interp = Interp1D.builder(data).strategy(StepInterpolator()).build()
    result = interp.interp_array(query)
    assert np.abs(result - np.array([2.0, 2.0, 4.0, 4.0, 5.0, 5.0])).max() <= EPS


def test_rows_after_the_first_error_stay_untouched():
    """interp1d/mod.rs:336-340: the batch stops at the first Err; earlier rows are written"""
    interp = Interp1D.builder(np.array([[1.0, 10.0], [2.0, 20.0], [3.0, 30.0]])).build()
    buf = np.full((4, 2), -7.0)
    with pytest.raises(InterpolateError.OutOfBounds, match=r"x = 5\.0 is not in range"):
        interp.interp_array_into(np.array([0.5, 1.5, 5.0, 1.0]), buf)
    assert np.array_equal(buf, np.array([[1.5, 15.0], [2.5, 25.0], [-7.0, -7.0], [-7.0, -7.0]]))
