"""tests/interp2d.rs (8 tests) and src/interp2d/mod.rs:521-589, transcribed test-for-test."""
import numpy as np
import pytest

import golden_util as G
from ndarray_interp_b200 import BuilderError, InterpolateError, Panic
from ndarray_interp_b200.interp2d import Interp2D, Interp2DBuilder

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def data_i32():
    return np.array([[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]], dtype=np.int32)


def data_f64():
    return data_i32().astype(np.float64)


def test_cornerns_only_data_no_axis():
    interp = Interp2D.builder(data_i32()).build()
    assert interp.interp_scalar(0, 0) == 1
    assert interp.interp_scalar(2, 3) == 12
    assert interp.interp_scalar(2, 0) == 9
    assert interp.interp_scalar(0, 3) == 4
    assert interp.interp_scalar(0, 0).dtype == np.int32


def test_cornerns_only_x_axis():
    interp = Interp2D.builder(data_i32().view()).x(np.array([1, 2, 3], dtype=np.int32)).build()
    assert interp.interp_scalar(1, 0) == 1
    assert interp.interp_scalar(3, 3) == 12
    assert interp.interp_scalar(3, 0) == 9
    assert interp.interp_scalar(1, 3) == 4


def test_cornerns_only_y_axis():
    interp = Interp2D.builder(data_f64().view()).y(np.array([-3.0, -2.0, -1.0, 0.0])).build()
    assert interp.interp_scalar(0.0, -3.0) == 1.0
    assert interp.interp_scalar(2.0, 0.0) == 12.0
    assert interp.interp_scalar(2.0, -3.0) == 9.0
    assert interp.interp_scalar(0.0, 0.0) == 4.0


def test_extrapolate():
    interp = Interp2D.builder(data_i32()).build()
    with pytest.raises(InterpolateError.OutOfBounds, match="x = -1 is not in range"):
        interp.interp(-1, 1)
    with pytest.raises(InterpolateError.OutOfBounds, match="y = -1 is not in range"):
        interp.interp(1, -1)
    with pytest.raises(InterpolateError.OutOfBounds, match="x = 3 is not in range"):
        interp.interp(3, 1)
    with pytest.raises(InterpolateError.OutOfBounds, match="y = 4 is not in range"):
        interp.interp(1, 4)


def test_interpolate_array():
    case = [c for c in G.load("bilinear") if c["name"] == "interpolate_array"][0]
    data = G.linspace(0.0, 8.0, 9).reshape(3, 3)
    res = 11
    qx = np.repeat(G.linspace(1.0, 3.0, res), res).reshape(res, res)
    qy = np.tile(G.linspace(4.0, 6.0, res), res).reshape(res, res)
    interp = Interp2D.builder(data).x(np.array([1.0, 2.0, 3.0])).y(np.array([4.0, 5.0, 6.0])).build()
    out = interp.interp_array(qx, qy)
    expect = np.array(case["expect"]).reshape(res, res)
    assert np.abs(out - expect).max() <= EPS
    assert np.array_equal(out, expect)       # the exact-order kernel reproduces all 121 values bit for bit


def test_interp_nd_data():
    data = np.array([[[[1.0, 10.0], [-1.0, -10.0]], [[2.0, 20.0], [-2.0, -20.0]]],
                     [[[3.0, 30.0], [-3.0, -30.0]], [[5.0, 50.0], [-5.0, -50.0]]]])
    interp = Interp2DBuilder.new(data).build()
    res = interp.interp(0.0, 0.5)
    assert np.abs(res - np.array([[1.5, 15.0], [-1.5, -15.0]])).max() <= EPS
    res = interp.interp_array(np.array([0.0, 0.5]), np.array([0.5, 1.0]))
    expect = np.array([[[1.5, 15.0], [-1.5, -15.0]], [[3.5, 35.0], [-3.5, -35.0]]])
    assert np.abs(res - expect).max() <= EPS


def test_interp_array_with_unmatched_axis():
    data = G.linspace(0.0, 8.0, 9).reshape(3, 3)
    interp = Interp2D.builder(data).build()
    with pytest.raises(Panic, match=r"`xs.shape\(\)` and `ys.shape\(\)` do not match"):
        interp.interp_array(np.array([0.0, 1.0]), np.array([0.0, 1.0, 2.0]))


def test_builder_errors():
    i = np.int32
    d = np.array([[1, 2], [3, 4]], dtype=i)
    with pytest.raises(BuilderError.NotEnoughData):
        Interp2D.builder(np.array([[1]], dtype=i)).build()
    with pytest.raises(BuilderError.NotEnoughData):
        Interp2D.builder(np.array([[1, 2]], dtype=i)).build()
    with pytest.raises(BuilderError.NotEnoughData):
        Interp2D.builder(np.array([[1], [2]], dtype=i)).build()
    with pytest.raises(BuilderError.ShapeError):
        Interp2D.builder(d).x(np.array([1], dtype=i)).build()
    with pytest.raises(BuilderError.ShapeError):
        Interp2D.builder(d).x(np.array([1, 2, 3], dtype=i)).build()
    with pytest.raises(BuilderError.ShapeError):
        Interp2D.builder(d).y(np.array([1], dtype=i)).build()
    with pytest.raises(BuilderError.ShapeError):
        Interp2D.builder(d).y(np.array([1, 2, 3], dtype=i)).build()
    with pytest.raises(BuilderError.Monotonic, match="x-axis"):
        Interp2D.builder(d).x(np.array([2, 2], dtype=i)).build()
    with pytest.raises(BuilderError.Monotonic, match="y-axis"):
        Interp2D.builder(d).y(np.array([2, 2], dtype=i)).build()


@pytest.mark.parametrize("case", [c for c in G.load("bilinear") if c["name"] != "interpolate_array"],
                         ids=lambda c: c["name"])
def test_bilinear_golden_vectors(case):
    dt = G.DT[case["dtype"]]
    b = Interp2DBuilder.new(np.array(case["data"], dtype=dt))
    if case.get("x") is not None:
        b = b.x(np.array(case["x"], dtype=dt))
    if case.get("y") is not None:
        b = b.y(np.array(case["y"], dtype=dt))
    out = b.build().interp_array(np.array(case["qx"], dtype=dt), np.array(case["qy"], dtype=dt))
    exp = np.array(case["expect"], dtype=dt)
    assert out.shape == exp.shape
    if case["tol"]["abs"] == 0.0:
        assert np.array_equal(out, exp)
    else:
        assert np.abs(out - exp).max() <= case["tol"]["abs"]


# ---- src/interp2d/mod.rs:521-589 ----------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 3, 4, 5, 6, 7, 8])
def test_interp2d_nd(dim):
    data = np.random.default_rng(64).uniform(0.0, 1.0, size=(4,) * dim)
    interp = Interp2D.builder(data).build()
    res = interp.interp(2.2, 2.2)
    assert res.ndim == dim - 2
    buf = np.zeros(res.shape)
    interp.interp_into(2.2, 2.2, buf)
    assert np.abs(buf - res).max(initial=0.0) <= EPS
    q = np.array([[0.5, 1.0], [1.5, 2.0]])
    res = interp.interp_array(q, q)
    assert res.ndim == dim - 2 + q.ndim
    buf = np.zeros(res.shape)
    interp.interp_array_into(q, q, buf)
    assert np.abs(buf - res).max() <= EPS


def test_interp2d_2d_scalar():
    data = np.random.default_rng(64).uniform(0.0, 1.0, size=(4, 4))
    assert isinstance(Interp2D.builder(data).build().interp_scalar(2.2, 2.2), np.float64)
