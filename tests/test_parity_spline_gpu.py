"""Cubic spline parity (K5 evaluation + K6 coefficient construction) against the CPU oracle,
through the C ABI.  The column-parallel Thomas keeps the reference's operation order, so the
coefficient arrays a, b and the evaluated values are compared bit for bit; the 1e-12 (f64) /
1e-5 (f32) relative bar of BASELINE.json is asserted alongside, measured against
max(|ref|, max|y| of the column) as SURVEY.md section 8(c) proposes (pure relative error is
undefined at zero crossings)."""
import ctypes as C

import numpy as np
import pytest

from ndarray_interp_b200 import BuilderError, InterpolateError, Panic, _lib as L
from ndarray_interp_b200.interp1d import (BoundaryCondition, CubicSpline, Interp1D, Interp1DBuilder, RowBoundary,
                                          SingleBoundary)
from oracle import oracle_py as O
from test_parity_gpu import make_queries, same

pytestmark = pytest.mark.gpu

SHARED = ["NotAKnot", "Natural", "Clamped", "Periodic"]
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def grid(rng, n, dt, uniform=False):
    g = np.arange(n) * 0.75 - 2.0 if uniform else np.cumsum(rng.uniform(0.5, 1.5, n))
    return g.astype(dt)


def build(g, y, bc_name, extrapolate=False):
    strat = CubicSpline.new().extrapolate(extrapolate).boundary(getattr(BoundaryCondition, bc_name))
    return Interp1DBuilder.new(y).x(g).strategy(strat).build()


def within_spec(got, ref, y):
    scale = np.maximum(np.abs(ref), np.abs(y).max(axis=0))
    return bool((np.abs(got.astype(np.float64) - ref) <= TOL[got.dtype] * scale).all())


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("bc", SHARED)
@pytest.mark.parametrize("n", [3, 4, 5, 6, 50, 1000])
@pytest.mark.parametrize("trailing", [(), (1,), (3,), (32,), (100,), (4, 5)], ids=lambda t: "w" + "x".join(map(str, t)))
def test_coefficients_match_oracle(dt, bc, n, trailing):
    rng = np.random.default_rng(n * 31 + len(trailing))
    g = grid(rng, n, dt)
    y = rng.normal(size=(n,) + trailing).astype(dt)
    if bc == "Periodic":
        y[-1] = y[0]
    st, a_ref, b_ref = O.spline_build(g, y, {"kind": bc})
    assert st == O.ST_OK
    interp = build(g, y, bc)
    a, b = interp.strategy.coefficients(interp)
    assert same(a, a_ref) and same(b, b_ref)


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("n", [3, 4, 7, 300])
def test_individual_boundaries_match_oracle(dt, n):
    rng = np.random.default_rng(n)
    w = 23
    g = grid(rng, n, dt)
    y = rng.normal(size=(n, w)).astype(dt)
    kinds = ["NotAKnot", "Natural", "Clamped", "FirstDeriv", "SecondDeriv"]
    rows, spec = [], []
    for c in range(w):
        if c % 4 == 0:
            k = kinds[c % 3]
            rows.append(getattr(RowBoundary, k)); spec.append({"kind": k})
        else:
            lk, rk = kinds[rng.integers(0, 5)], kinds[rng.integers(0, 5)]
            lv, rv = float(rng.normal()), float(rng.normal())

            def sb(k, v):
                return (SingleBoundary.FirstDeriv(v) if k == "FirstDeriv" else SingleBoundary.SecondDeriv(v)
                        if k == "SecondDeriv" else getattr(SingleBoundary, k))
            rows.append(RowBoundary.Mixed(sb(lk, lv), sb(rk, rv)))
            spec.append({"kind": "Mixed", "left": {"kind": lk, "value": lv}, "right": {"kind": rk, "value": rv}})
    st, a_ref, b_ref = O.spline_build(g, y, {"kind": "Individual", "rows": spec})
    assert st == O.ST_OK
    strat = CubicSpline.new().boundary(BoundaryCondition.Individual([rows]))
    interp = Interp1DBuilder.new(y).x(g).strategy(strat).build()
    a, b = interp.strategy.coefficients(interp)
    assert same(a, a_ref) and same(b, b_ref)


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("bc", SHARED)
@pytest.mark.parametrize("trailing", [(), (1,), (2,), (5,), (8,), (32,), (64,), (100,), (130,), (1024,)],
                         ids=lambda t: "w" + "x".join(map(str, t)))
def test_cubic_eval_matches_oracle(dt, bc, trailing):
    rng = np.random.default_rng(len(trailing) + int(np.prod(trailing, dtype=np.int64)))
    n = 64
    g = grid(rng, n, dt)
    y = rng.normal(size=(n,) + trailing).astype(dt)
    if bc == "Periodic":
        y[-1] = y[0]
    st, a_ref, b_ref = O.spline_build(g, y, {"kind": bc})
    for extrapolate in (False, True):
        interp = build(g, y, bc, extrapolate)
        mode = 0 if not extrapolate else (2 if bc == "Periodic" else 1)
        for nq in (1, 33, 4000):
            q = make_queries(rng, g, nq, dt, outside=extrapolate)
            if nq == 4000:
                q = np.sort(q)                    # sorted batch: exercises the row-reuse path
            st, ref, _ = O.interp1d_cubic(g, y, a_ref, b_ref, q, mode)
            assert st == O.ST_OK
            got = interp.interp_array(q)
            assert same(got, ref)
            assert within_spec(got, ref.astype(np.float64), y.reshape(n, -1).reshape(y.shape))


def test_cubic_periodic_wrap_and_nonfinite():
    rng = np.random.default_rng(4)
    g = np.cumsum(rng.uniform(0.5, 1.5, 40))
    y = rng.normal(size=(40, 3)); y[-1] = y[0]
    interp = build(g, y, "Periodic", True)
    st, a, b = O.spline_build(g, y, {"kind": "Periodic"})
    span = g[-1] - g[0]
    q = np.concatenate([rng.uniform(g[0] - 5 * span, g[-1] + 5 * span, 500), [g[0], g[-1], g[0] - span, g[-1] + span]])
    st, ref, _ = O.interp1d_cubic(g, y, a, b, q, 2)
    assert st == O.ST_OK
    assert same(interp.interp_array(q), ref)
    for bad in (np.inf, -np.inf, np.nan):          # rem_euclid(+-inf) is NaN -> the NaN panic
        qq = q.copy(); qq[77] = bad
        assert O.interp1d_cubic(g, y, a, b, qq, 2)[0] == O.ST_NAN_QUERY
        with pytest.raises(Panic, match="failed to convert NaN to usize"):
            interp.interp_array(qq)
    # Extrapolate::Yes: +-inf extrapolates (no panic), NaN panics
    ex = build(g, y, "Natural", True)
    st, a, b = O.spline_build(g, y, {"kind": "Natural"})
    qq = q.copy(); qq[5], qq[6] = np.inf, -np.inf
    st, ref, _ = O.interp1d_cubic(g, y, a, b, qq, 1)
    assert st == O.ST_OK and same(ex.interp_array(qq), ref)


def test_cubic_out_of_bounds_first_bad():
    rng = np.random.default_rng(6)
    g = np.cumsum(rng.uniform(0.5, 1.5, 100))
    y = rng.normal(size=(100, 16))
    interp = build(g, y, "NotAKnot", False)
    st, a, b = O.spline_build(g, y, {"kind": "NotAKnot"})
    for nq in (64, 50000):
        q = rng.uniform(g[0], g[-1], nq)
        q[nq // 3] = g[0] - 0.5
        q[nq // 3 + 7] = np.nan
        buf = np.full((nq, 16), 9.0)
        with pytest.raises(InterpolateError.OutOfBounds):
            interp.interp_array_into(q, buf)
        st, ref, bad = O.interp1d_cubic(g, y, a, b, q, 0, out=np.full((nq, 16), 9.0))
        assert (st, bad) == (O.ST_OUT_OF_BOUNDS, nq // 3)
        assert same(buf, ref)


def test_periodic_mismatch_reports_value_error():
    y = np.array([[0.5, 1.0], [0.0, 1.5], [0.2, 0.3], [0.5, 1.1]])
    with pytest.raises(BuilderError.ValueError, match="first and last value must be equal"):
        build(np.arange(4.0), y, "Periodic")
    with pytest.raises(BuilderError.ValueError, match=r"First: 0\.5, last: 0\.6"):
        build(np.arange(4.0), np.array([0.5, 0.1, 0.2, 0.6]), "Periodic")


def test_spline_build_large_column_count_and_long_columns():
    """both regimes north_star names: many columns (column-parallel Thomas fills the machine) and
    few long columns"""
    rng = np.random.default_rng(8)
    for n, w, dt in [(64, 40000, np.float32), (4096, 64, np.float64)]:
        g = grid(rng, n, dt)
        y = rng.normal(size=(n, w)).astype(dt)
        st, a_ref, b_ref = O.spline_build(g, y, {"kind": "Natural"})
        strat = CubicSpline.new().boundary(BoundaryCondition.Natural).solver("sequential")
        interp = Interp1DBuilder.new(y).x(g).strategy(strat).build()
        a, b = interp.strategy.coefficients(interp)
        assert same(a, a_ref) and same(b, b_ref)
        # an interpolant reproduces its knots: t = 0 gives exactly y[i]
        assert same(interp.interp_array(g[:-1]), y[:-1])


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
def test_spline_sign_of_zero_follows_the_oracle(dt):
    """a flat column with FirstDeriv(-0.0) on both sides: k, a and b are zeros whose SIGN the sweeps' hoisted
    division must keep (IEEE: -0 / b = -0 for b > 0); compared bytewise, not by value"""
    n, w = 50, 6
    g = np.arange(n, dtype=dt) * dt(0.5)
    y = np.zeros((n, w), dt); y[:, 3] = dt(2.5)
    rows = [RowBoundary.Mixed(SingleBoundary.FirstDeriv(-0.0), SingleBoundary.FirstDeriv(-0.0)) for _ in range(w)]
    spec = [{"kind": "Mixed", "left": {"kind": "FirstDeriv", "value": -0.0}, "right": {"kind": "FirstDeriv", "value": -0.0}}
            for _ in range(w)]
    st, a_ref, b_ref = O.spline_build(g, y, {"kind": "Individual", "rows": spec})
    assert st == O.ST_OK
    interp = Interp1DBuilder.new(y).x(g).strategy(CubicSpline.new().boundary(BoundaryCondition.Individual([rows]))).build()
    a, b = interp.strategy.coefficients(interp)
    assert a.tobytes() == a_ref.tobytes() and b.tobytes() == b_ref.tobytes()
