"""tests/cubic_spline_strat.rs (18 tests) and the cubic_spline.rs doctests, transcribed
test-for-test.  Golden vectors come from tests/golden/reference_vectors.json."""
import numpy as np
import pytest

import golden_util as G
from ndarray_interp_b200 import BuilderError, InterpolateError
from ndarray_interp_b200.interp1d import (BoundaryCondition, CubicSpline, Interp1D, Interp1DBuilder, RowBoundary,
                                          SingleBoundary)

pytestmark = pytest.mark.gpu

_KIND = {"NotAKnot": BoundaryCondition.NotAKnot, "Natural": BoundaryCondition.Natural,
         "Clamped": BoundaryCondition.Clamped, "Periodic": BoundaryCondition.Periodic}


def _single(d):
    if d["kind"] == "FirstDeriv":
        return SingleBoundary.FirstDeriv(d["value"])
    if d["kind"] == "SecondDeriv":
        return SingleBoundary.SecondDeriv(d["value"])
    return getattr(SingleBoundary, d["kind"])


def _boundary(bc, data_shape):
    if bc["kind"] != "Individual":
        return _KIND[bc["kind"]]
    rows = []
    for r in bc["rows"]:
        rows.append(RowBoundary.Mixed(_single(r["left"]), _single(r["right"])) if r["kind"] == "Mixed"
                    else getattr(RowBoundary, r["kind"]))
    arr = np.empty((1,) + tuple(data_shape[1:]), dtype=object)
    arr.reshape(-1)[:] = rows
    return BoundaryCondition.Individual(arr)


def _assert_relative_eq(got, exp, eps, max_relative):
    got, exp = np.asarray(got, np.float64), np.asarray(exp, np.float64)
    diff = np.abs(got - exp)
    ok = (diff <= eps) | (diff <= np.maximum(np.abs(got), np.abs(exp)) * max_relative)
    assert ok.all(), f"max diff {diff.max()}"


@pytest.mark.parametrize("case", G.load("cubic"), ids=lambda c: c["name"])
def test_cubic_golden(case):
    """interp_natural, extrapolate_natural, extrapolate_not_a_knot (f32), not_a_knot_3_values,
    multidim_multi_bounds, extrapolate_clamped, extrapolate_deriv1/2, extrapolate_periodic*,
    and the Wikipedia doctest (cubic_spline.rs:62-82)"""
    dt = G.DT[case["dtype"]]
    data = np.array(case["data"], dtype=dt)
    b = Interp1D.builder(data)
    if case["x"] is not None:
        b = b.x(np.array(case["x"], dtype=dt))
    strat = CubicSpline.new().extrapolate(case["extrapolate"]).boundary(_boundary(case["bc"], data.shape))
    interp = b.strategy(strat).build()
    res = interp.interp_array(G.materialise(case["query"], dt))
    exp = np.array(case["expect"], dtype=dt).reshape(res.shape)
    _assert_relative_eq(res, exp, case["tol"]["abs"], case["tol"]["rel"])
    if case["name"] == "doctest_wikipedia":
        assert np.array_equal(res, exp)


def test_to_little_data():
    with pytest.raises(BuilderError.NotEnoughData):
        Interp1D.builder(np.array([1.0, 2.0])).strategy(CubicSpline.new()).build()


def test_enough_data():
    Interp1D.builder(np.array([1.0, 2.0, 1.0])).strategy(CubicSpline.new()).build()


def test_extrapolate_false():
    interp = Interp1D.builder(np.array([1.0, 2.0, 1.0])).strategy(CubicSpline.new()).build()
    with pytest.raises(InterpolateError.OutOfBounds):
        interp.interp(-0.5)
    with pytest.raises(InterpolateError.OutOfBounds):
        interp.interp(3.5)


def test_bounds_shape_error1():
    y = np.array([[0.5, 1.0], [0.0, 1.5], [3.0, 0.5]])
    bounds = BoundaryCondition.Individual([[RowBoundary.Natural, RowBoundary.Clamped, RowBoundary.NotAKnot]])
    with pytest.raises(BuilderError.ShapeError, match=r"Expected: \[1, 2\], got: \[1, 3\]"):
        Interp1DBuilder.new(y).strategy(CubicSpline.new().boundary(bounds)).build()


def test_bounds_shape_error2():
    y = np.array([[0.5, 1.0], [0.0, 1.5], [3.0, 0.5]])
    bounds = BoundaryCondition.Individual([[RowBoundary.Natural, RowBoundary.NotAKnot],
                                           [RowBoundary.Natural, RowBoundary.NotAKnot]])
    with pytest.raises(BuilderError.ShapeError, match=r"Expected: \[1, 2\], got: \[2, 2\]"):
        Interp1DBuilder.new(y).strategy(CubicSpline.new().boundary(bounds)).build()


def test_periodic_wrong_values():
    y = np.array([[0.5, 1.0], [0.0, 1.5], [0.5, 1.1]])
    msg = (r"First: \[0\.5, 1\.0\], shape=\[2\], strides=\[1\], layout=CFcf \(0xf\), const ndim=1, "
           r"last: \[0\.5, 1\.1\]")
    with pytest.raises(BuilderError.ValueError, match=msg):
        Interp1DBuilder.new(y).strategy(CubicSpline.new().boundary(BoundaryCondition.Periodic)).build()


def test_boundary_doc_example_builds():
    # cubic_spline.rs:123-151
    y = np.array([[0.5, 1.0], [0.0, 1.5], [3.0, 0.5]])
    x = np.array([-1.0, 0.0, 3.0])
    bounds = BoundaryCondition.Individual([[RowBoundary.Natural,
                                            RowBoundary.Mixed(SingleBoundary.NotAKnot, SingleBoundary.FirstDeriv(0.5))]])
    Interp1DBuilder.new(y).x(x).strategy(CubicSpline.new().boundary(bounds)).build()
