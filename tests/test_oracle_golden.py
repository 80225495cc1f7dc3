"""Pins the CPU oracle (oracle/ndi_oracle.cpp) against every golden vector the reference's own
tests hold for the hot path and against scipy as an independent check."""
import numpy as np
import pytest

import golden_util as G
from oracle import oracle_py as O


def _close(got, exp, tol):
    got = np.asarray(got, dtype=np.float64).ravel()
    exp = np.asarray(exp, dtype=np.float64).ravel()
    assert got.shape == exp.shape
    # approx::assert_relative_eq!(epsilon, max_relative): |a-b| <= eps OR |a-b| <= max(|a|,|b|)*rel
    diff = np.abs(got - exp)
    ok = (diff <= tol.get("abs", 0.0)) | (diff <= np.maximum(np.abs(got), np.abs(exp)) * tol.get("rel", 0.0))
    assert ok.all(), f"max diff {diff.max()} at {diff.argmax()}: got {got[diff.argmax()]} exp {exp[diff.argmax()]}"


@pytest.mark.parametrize("case", G.load("linear"), ids=lambda c: c["name"])
def test_linear_golden(case):
    dt = G.DT[case["dtype"]]
    data = np.array(case["data"], dtype=dt)
    x = np.array(case["x"], dtype=dt) if case["x"] is not None else G.default_axis(len(data), dt)
    q = np.array(case["query"], dtype=dt)
    st, out, _ = O.interp1d_linear(x, data, q, case["extrapolate"])
    assert st == O.ST_OK
    exp = np.array(case["expect"], dtype=dt)
    assert out.shape == exp.shape
    if case["tol"]["abs"] == 0.0:
        assert np.array_equal(out, exp)          # assert_eq! in the reference
    else:
        _close(out, exp, case["tol"])


@pytest.mark.parametrize("case", G.load("bilinear"), ids=lambda c: c["name"])
def test_bilinear_golden(case):
    dt = G.DT[case["dtype"]]
    data = G.materialise(case["data"], dt)
    x = np.array(case["x"], dtype=dt) if case.get("x") is not None else G.default_axis(data.shape[0], dt)
    y = np.array(case["y"], dtype=dt) if case.get("y") is not None else G.default_axis(data.shape[1], dt)
    qx, qy = G.materialise(case["qx"], dt), G.materialise(case["qy"], dt)
    st, out, _, _ = O.interp2d_bilinear(x, y, data, qx, qy, case.get("extrapolate", False))
    assert st == O.ST_OK
    exp = np.array(case["expect"], dtype=dt).reshape(out.shape)
    if case["name"] == "interpolate_array":
        # the restatement reproduces all 121 values bit-exactly
        assert np.array_equal(out, exp)
    elif case["tol"]["abs"] == 0.0:
        assert np.array_equal(out, exp)
    else:
        _close(out, exp, case["tol"])


@pytest.mark.parametrize("case", G.load("cubic"), ids=lambda c: c["name"])
def test_cubic_golden(case):
    dt = G.DT[case["dtype"]]
    data = np.array(case["data"], dtype=dt)
    x = np.array(case["x"], dtype=dt) if case["x"] is not None else G.default_axis(len(data), dt)
    q = G.materialise(case["query"], dt)
    st, out = O.cubic_interp(x, data, case["bc"], case["extrapolate"], q)
    assert st == O.ST_OK
    exp = np.array(case["expect"], dtype=dt).reshape(out.shape)
    if case["name"] == "doctest_wikipedia":
        assert np.array_equal(out, exp)          # bit-exact incl. -5.551115123125783e-17
    _close(out, exp, case["tol"])


def test_oob_pins():
    for case in G.load("oob1d"):
        data = np.array(case["data"])
        x = np.array(case["x"]) if case["x"] is not None else G.default_axis(len(data), np.float64)
        for i, qv in enumerate(case["queries"]):
            st, _, bad = O.interp1d_linear(x, data, np.array([0.0 + x[0], qv]), False)
            assert (st, bad) == (O.ST_OUT_OF_BOUNDS, 1)
    for case in G.load("oob1d_cubic"):
        data = np.array(case["data"])
        x = G.default_axis(len(data), np.float64)
        st, a, b = O.spline_build(x, data, {"kind": "NotAKnot"})
        assert st == O.ST_OK
        for qv in case["queries"]:
            st, _, bad = O.interp1d_cubic(x, data, a, b, np.array([qv]), 0)
            assert (st, bad) == (O.ST_OUT_OF_BOUNDS, 0)
    for case in G.load("oob2d"):
        data = np.array(case["data"], dtype=np.int32)
        x, y = G.default_axis(3, np.int32), G.default_axis(4, np.int32)
        for qx, qy, axis in case["queries"]:
            st, _, bad, ax = O.interp2d_bilinear(x, y, data, np.array([qx], np.int32), np.array([qy], np.int32), False)
            assert (st, bad, ax) == (O.ST_OUT_OF_BOUNDS, 0, axis)


@pytest.mark.parametrize("case", G.load("index"), ids=lambda c: c["name"])
def test_lower_index_golden(case):
    grid = G.index_grid(case["grid"])
    exp = np.array([p[0] for p in case["pairs"]], dtype=np.int64)
    q = np.array([G.index_query(p[1]) for p in case["pairs"]])
    st, idx, _ = O.lower_index(grid, q)
    assert st == O.ST_OK
    assert np.array_equal(idx, exp)


def test_lower_index_nan_panics():
    st, _, bad = O.lower_index(G.index_grid("linspace"), np.array([1.0, np.nan]))
    assert (st, bad) == (O.ST_NAN_QUERY, 1)


@pytest.mark.parametrize("case", G.load("monotonic"), ids=lambda c: c["name"])
def test_monotonic_golden(case):
    x = np.array(case["x"], dtype=G.DT[case["dtype"]])
    assert O.monotonic_prop(x, case["stride"]) == case["expect"]
    if case["stride"] == 1:
        assert O.monotonic_prop(x[::1], 1) == case["expect"]


def test_monotonic_nan_semantics():
    # vector_extensions.rs:136-170: an unordered pair is "else" in Init/NotStrict (-> Falling)
    # and "else" in the Likely states (-> NotMonotonic)
    nan = np.nan
    assert O.monotonic_prop(np.array([nan, 1.0])) == "FallingStrict"
    assert O.monotonic_prop(np.array([1.0, 1.0, nan])) == "Falling"
    assert O.monotonic_prop(np.array([1.0, 2.0, nan])) == "NotMonotonic"
    assert O.monotonic_prop(np.array([3.0, 2.0, nan])) == "NotMonotonic"


# ---- independent cross-check: scipy.interpolate.CubicSpline ---------------------------------------
@pytest.mark.parametrize("bc,scipy_bc", [("Natural", "natural"), ("NotAKnot", "not-a-knot"),
                                         ("Clamped", "clamped"), ("Periodic", "periodic")])
def test_cubic_vs_scipy(bc, scipy_bc):
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(7)
    n, w = 40, 3
    # NotAKnot: uniform grid only -- on a non-uniform grid the REFERENCE deviates from scipy
    # (see test_not_a_knot_right_boundary_quirk); the oracle follows the reference.
    x = np.arange(n) * 0.75 + 1.0 if bc == "NotAKnot" else np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.normal(size=(n, w))
    if bc == "Periodic":
        y[-1] = y[0]
    q = rng.uniform(x[0], x[-1], 500)
    st, out = O.cubic_interp(x, y, {"kind": bc}, False, q)
    assert st == O.ST_OK
    ref = CubicSpline(x, y, axis=0, bc_type=scipy_bc)(q)
    scale = np.abs(y).max()
    assert np.abs(out - ref).max() <= 1e-12 * scale


def test_cubic_deriv_bcs_vs_scipy():
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(8)
    n = 25
    x = np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.normal(size=n)
    q = rng.uniform(x[0], x[-1], 300)
    for lk, rk, sl, sr in [("FirstDeriv", "SecondDeriv", 1, 2), ("SecondDeriv", "FirstDeriv", 2, 1),
                           ("NotAKnot", "FirstDeriv", None, 1)]:
        bc = {"kind": "Individual", "rows": [{"kind": "Mixed", "left": {"kind": lk, "value": 0.3},
                                             "right": {"kind": rk, "value": -0.7}}]}
        st, out = O.cubic_interp(x, y, bc, False, q)
        assert st == O.ST_OK
        left = "not-a-knot" if sl is None else (sl, 0.3)
        ref = CubicSpline(x, y, bc_type=(left, (sr, -0.7)))(q)
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(y).max()


def test_not_a_knot_right_boundary_quirk():
    """cubic_spline.rs:635 sets the last diagonal entry to dx[n-2] (x[n-1]-x[n-2]) where the
    textbook/scipy system has dx[n-3].  Identical on uniform grids (all the reference's own
    NotAKnot tests), different otherwise.  Parity means following the reference, so the oracle
    (and the CUDA path) reproduce the quirk; this test documents it."""
    from scipy.interpolate import CubicSpline
    rng = np.random.default_rng(7)
    n = 40
    x = np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.normal(size=n)
    q = np.linspace(x[0], x[-1], 400)
    st, out = O.cubic_interp(x, y, {"kind": "NotAKnot"}, False, q)
    assert st == O.ST_OK
    ref = CubicSpline(x, y, bc_type="not-a-knot")(q)
    left = q < x[n // 2]
    assert np.abs(out - ref)[left].max() < 1e-6       # the left end is the textbook system
    assert np.abs(out - ref)[~left].max() > 1e-3      # the right end is not
    # still an interpolant: exact at the knots
    st, at_knots = O.cubic_interp(x, y, {"kind": "NotAKnot"}, False, x)
    assert np.abs(at_knots - y).max() < 1e-12


def test_periodic_mismatch_is_value_error():
    x = np.array([0.0, 1.0, 2.0, 3.0])
    st, _, _ = O.spline_build(x, np.array([[0.5, 1.0], [0.0, 1.5], [0.2, 0.1], [0.5, 1.1]]), {"kind": "Periodic"})
    assert st == O.ST_PERIODIC_MISMATCH
    st, _, _ = O.spline_build(x[:3], np.array([[0.5, 1.0], [0.0, 1.5], [0.5, 1.1]]), {"kind": "Periodic"})
    assert st == O.ST_PERIODIC_MISMATCH


def test_fma_would_break_parity():
    """the oracle must be built without FMA contraction.  Check that calc_frac
    equals the unfused numpy evaluation on a cancellation-heavy sample (numpy never fuses)."""
    rng = np.random.default_rng(1)
    n = 20000
    x1 = rng.uniform(0, 1, n).astype(np.float32); x2 = x1 + rng.uniform(0.01, 1, n).astype(np.float32)
    y1 = rng.uniform(-1, 1, n).astype(np.float32); y2 = rng.uniform(-1, 1, n).astype(np.float32)
    x = (x1 + (x2 - x1) * rng.uniform(0, 1, n).astype(np.float32)).astype(np.float32)
    ref = ((y2 - y1) / (x2 - x1)) * (x - x1) + y1
    got = np.array([O.calc_frac(float(a), float(b), float(c), float(d), float(e), np.float32)
                    for a, b, c, d, e in zip(x1, y1, x2, y2, x)], dtype=np.float32)
    assert np.array_equal(got, ref)


def test_row_split_study_solves_the_same_system_as_the_oracle():
    """scripts/pcr_accuracy_study.py (groundwork for a row-split spline build): its sequential
    solve must be the oracle's, bit for bit, and PCR + Thomas must stay far inside north_star's bars"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "pcr_accuracy_study", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "pcr_accuracy_study.py"))
    P = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(P)
    rng = np.random.default_rng(5)
    for dt, bar in ((np.float64, 1e-12), (np.float32, 1e-5)):
        x = np.cumsum(rng.uniform(0.5, 1.5, 300)).astype(dt)
        y = rng.normal(size=(300, 3)).astype(dt)
        low, mid, up, rhs = P.system_natural(x, y)
        a, b = P.coefficients(x, y, P.thomas(low, mid, up, rhs))
        st, ra, rb = O.spline_build(x, y, {"kind": "Natural"})
        assert st == O.ST_OK and np.array_equal(a, ra) and np.array_equal(b, rb)
        q = np.sort(rng.uniform(x[0], x[-1], 2000)).astype(dt)
        ref = P.evaluate(x, y, a, b, q)
        scale = np.maximum(np.abs(ref), np.abs(y).max(axis=0)[None, :])
        for levels in (1, 2, 4):
            pa, pb = P.coefficients(x, y, P.pcr_then_thomas(low, mid, up, rhs, levels))
            val = P.evaluate(x, y, pa, pb, q)
            assert float((np.abs(val - ref) / scale).max()) < bar / 10
            # the oracle's row-split variant (rowsplit_thomas: the operation-by-operation specification the
            # row-split build kernels are compared with) is this very computation
            st, oa, ob = O.spline_build(x, y, {"kind": "Natural"}, rowsplit_levels=levels)
            assert st == O.ST_OK and np.array_equal(oa, pa) and np.array_equal(ob, pb)
        # every boundary kind whose system is tridiagonal: the row-split solve stays inside the bars
        ind = {"kind": "Individual", "rows": [
            {"kind": "Mixed", "left": {"kind": "FirstDeriv", "value": 0.5}, "right": {"kind": "Natural"}},
            {"kind": "NotAKnot"},
            {"kind": "Mixed", "left": {"kind": "SecondDeriv", "value": -1.0}, "right": {"kind": "Clamped"}}]}
        yp = y.copy(); yp[-1] = yp[0]
        for bc in ({"kind": "NotAKnot"}, {"kind": "Clamped"}, ind, {"kind": "Periodic"}):
            if bc["kind"] == "Periodic":
                y = yp
            st, sa, sb = O.spline_build(x, y, bc)
            st2, ra2, rb2 = O.spline_build(x, y, bc, rowsplit_levels=3)
            assert st == O.ST_OK and st2 == O.ST_OK
            seq = P.evaluate(x, y, sa, sb, q)
            spl = P.evaluate(x, y, ra2, rb2, q)
            sc = np.maximum(np.abs(seq), np.abs(y).max(axis=0)[None, :])
            assert float((np.abs(spl - seq) / sc).max()) < bar / 10


def test_partition_specification_stays_inside_the_bars():
    """partition_thomas (the operation-by-operation specification of the partition build kernels,
    csrc/ndi_partition.cu) against the reference-order restatement: every boundary kind, direct solve, one, two and
    three split levels, tails of every length -- evaluated splines agree to a tenth of north_star's bars, and the
    f64 result agrees with scipy like the reference-order one does"""
    from scipy.interpolate import CubicSpline as SciSpline
    rng = np.random.default_rng(17)
    ind = {"kind": "Individual", "rows": [
        {"kind": "Mixed", "left": {"kind": "FirstDeriv", "value": 0.5}, "right": {"kind": "Natural"}},
        {"kind": "NotAKnot"},
        {"kind": "Mixed", "left": {"kind": "SecondDeriv", "value": -1.0}, "right": {"kind": "Clamped"}}]}
    for dt, bar in ((np.float64, 1e-12), (np.float32, 1e-5)):
        for n, block in ((10, 3), (97, 3), (300, 4), (1000, 5), (1025, 32), (4096, 32), (4127, 32)):
            x = np.cumsum(rng.uniform(0.5, 1.5, n)).astype(dt)
            y = rng.normal(size=(n, 3)).astype(dt)
            yp = y.copy(); yp[-1] = yp[0]
            q = np.sort(rng.uniform(x[0], x[-1], 1500)).astype(dt)
            for bc in ({"kind": "Natural"}, {"kind": "NotAKnot"}, {"kind": "Clamped"}, ind, {"kind": "Periodic"}):
                yy = yp if bc["kind"] == "Periodic" else y
                st, sa, sb = O.spline_build(x, yy, bc)
                st2, pa, pb = O.spline_build(x, yy, bc, partition_block=block)
                assert st == O.ST_OK and st2 == O.ST_OK
                _, seq, _ = O.interp1d_cubic(x, yy, sa, sb, q, 0)
                _, par, _ = O.interp1d_cubic(x, yy, pa, pb, q, 0)
                sc = np.maximum(np.abs(seq), np.abs(yy).max(axis=0)[None, :]).astype(np.float64)
                assert float((np.abs(par.astype(np.float64) - seq) / sc).max()) < bar / 10, (dt, n, block, bc["kind"])
                if dt is np.float64 and bc["kind"] == "Natural":
                    sci = SciSpline(x, yy, bc_type="natural")(q)
                    assert float((np.abs(par - sci) / sc).max()) < 1e-12


def _wrap(v, bits, signed):
    v &= (1 << bits) - 1
    return v - (1 << bits) if signed and v >> (bits - 1) else v


def _int_calc_frac(x1, y1, x2, y2, x, bits, signed):
    """Linear::calc_frac (linear.rs:29-36) as a release build of the reference evaluates it for an integer type:
    every operation modulo 2^bits, division truncating (signed) / flooring (unsigned), on Python's big integers"""
    num, den = _wrap(y2 - y1, bits, signed), _wrap(x2 - x1, bits, signed)
    if signed and num == -(1 << (bits - 1)) and den == -1:
        m = num
    else:
        m = abs(num) // abs(den) * (1 if (num < 0) == (den < 0) else -1) if signed else num // den
    return _wrap(_wrap(m * _wrap(x - x1, bits, signed), bits, signed) + y1, bits, signed)


@pytest.mark.parametrize("dt", [np.int32, np.int64, np.uint32, np.uint64], ids=["i32", "i64", "u32", "u64"])
def test_integer_arithmetic_of_the_oracle_is_the_wrapping_arithmetic_of_a_release_build(dt):
    """the oracle's integer Linear / Bilinear against big-integer arithmetic modulo 2^bits: large values whose
    products wrap, falling data (differences that are negative, i.e. wrap for the unsigned types), queries outside the
    grid on both sides"""
    bits, signed = np.dtype(dt).itemsize * 8, np.issubdtype(dt, np.signedinteger)
    rng = np.random.default_rng(bits + signed)
    top = (1 << (bits - 2))
    n, m, w = 40, 7, 3
    base = top if not signed else -top // 2
    g = (base + np.cumsum(rng.integers(1, 1000, n))).astype(dt)
    gy = (5 + np.cumsum(rng.integers(1, 9, m))).astype(dt)
    lo, hi = (0, 1 << (bits - 1)) if not signed else (-top, top)
    d1 = rng.integers(lo, hi, (n, w), dtype=np.int64 if bits == 32 else None).astype(dt) if bits == 32 else \
        (rng.integers(0, 1 << 62, (n, w)).astype(np.uint64) * 2).astype(dt)
    d2 = rng.integers(0, 1 << 20, (n, m, w)).astype(dt)
    q = (int(g[0]) - 30 + rng.integers(0, int(g[-1]) - int(g[0]) + 60, 500)).astype(dt)
    qy = (max(0, int(gy[0]) - 3) + rng.integers(0, int(gy[-1]) - int(gy[0]) + 6, 500)).astype(dt)
    st, out, _ = O.interp1d_linear(g, d1, q, True)
    assert st == O.ST_OK
    st, idx, _ = O.lower_index(g, q)
    for k in range(len(q)):
        i = int(idx[k])
        assert i == min(max(int(np.searchsorted(g, q[k], side="right")) - 1, 0), n - 2)
        for c in range(w):
            exp = _int_calc_frac(int(g[i]), int(d1[i, c]), int(g[i + 1]), int(d1[i + 1, c]), int(q[k]), bits, signed)
            assert int(out[k, c]) == exp, (k, c)
    st, out2, _, _ = O.interp2d_bilinear(g, gy, d2, q, qy, True)
    assert st == O.ST_OK
    st, jdx, _ = O.lower_index(gy, qy)
    for k in range(0, len(q), 7):
        i, j = int(idx[k]), int(jdx[k])
        for c in range(w):
            z1 = _int_calc_frac(int(g[i]), int(d2[i, j, c]), int(g[i + 1]), int(d2[i + 1, j, c]), int(q[k]), bits, signed)
            z2 = _int_calc_frac(int(g[i]), int(d2[i, j + 1, c]), int(g[i + 1]), int(d2[i + 1, j + 1, c]), int(q[k]), bits, signed)
            exp = _int_calc_frac(int(gy[j]), z1, int(gy[j + 1]), z2, int(qy[k]), bits, signed)
            assert int(out2[k, c]) == exp, (k, c)
