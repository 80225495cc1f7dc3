"""Row-split spline build (K6, second build mode: parallel cyclic reduction + Thomas, csrc/ndi_rowsplit.cu;
replaces the serial chains of CubicSpline::solve_for_k + thomas, cubic_spline.rs:409-721, on long columns).

Two bars, both asserted for every boundary kind:
  * bit for bit against the oracle's operation-by-operation specification of the same scheme
    (oracle/ndi_oracle.cpp: rowsplit_thomas, same number of reduction levels);
  * north_star's 1e-12 (f64) / 1e-5 (f32) against the oracle in the REFERENCE's elimination order, for the
    coefficients' effect on evaluated values, measured like tests/test_parity_spline_gpu.py (relative to
    max(|ref|, max|y| of the column)).
"""
import ctypes as C

import numpy as np
import pytest

from ndarray_interp_b200 import _lib as L
from ndarray_interp_b200.interp1d import (BoundaryCondition, CubicSpline, Interp1DBuilder, RowBoundary, SingleBoundary)
from oracle import oracle_py as O
from test_parity_gpu import same

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}
KINDS = ["NotAKnot", "Natural", "Clamped", "FirstDeriv", "SecondDeriv"]


def grid(rng, n, dt):
    return np.cumsum(rng.uniform(0.5, 1.5, n)).astype(dt)


def individual(rng, w):
    """random per-column boundaries: (RowBoundary objects for the mirror, the oracle's spec)"""
    rows, spec = [], []
    for c in range(w):
        if c % 5 == 0:
            k = KINDS[c % 3]
            rows.append(getattr(RowBoundary, k)); spec.append({"kind": k})
            continue
        lk, rk = KINDS[rng.integers(0, 5)], KINDS[rng.integers(0, 5)]
        lv, rv = float(rng.normal()), float(rng.normal())

        def sb(k, v):
            return (SingleBoundary.FirstDeriv(v) if k == "FirstDeriv" else SingleBoundary.SecondDeriv(v)
                    if k == "SecondDeriv" else getattr(SingleBoundary, k))
        rows.append(RowBoundary.Mixed(sb(lk, lv), sb(rk, rv)))
        spec.append({"kind": "Mixed", "left": {"kind": lk, "value": lv}, "right": {"kind": rk, "value": rv}})
    return rows, spec


def build(g, y, bc, solver, levels=0):
    strat = CubicSpline.new().boundary(bc).solver(solver, levels)
    return Interp1DBuilder.new(y).x(g).strategy(strat).build()


def eval_close(g, y, a, b, a_ref, b_ref, rng):
    """values of the two splines at random points agree inside north_star's bar"""
    q = np.sort(rng.uniform(g[0], g[-1], 4000)).astype(g.dtype)
    st, got, _ = O.interp1d_cubic(g, y, a, b, q, 0)
    st2, ref, _ = O.interp1d_cubic(g, y, a_ref, b_ref, q, 0)
    assert st == O.ST_OK and st2 == O.ST_OK
    ref64 = ref.astype(np.float64).reshape(len(q), -1)
    scale = np.maximum(np.abs(ref64), np.abs(y.reshape(len(g), -1)).max(axis=0)[None, :])
    err = float((np.abs(got.astype(np.float64).reshape(len(q), -1) - ref64) / scale).max())
    return err <= TOL[g.dtype], err


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("bc", ["NotAKnot", "Natural", "Clamped", "Periodic", "Individual"])
@pytest.mark.parametrize("n,w,levels", [(8, 3, 1), (9, 5, 2), (37, 33, 3), (100, 1, 4), (257, 40, 5), (1000, 7, 6),
                                        (1031, 64, 3), (4096, 65, 0), (5000, 2, 6)])
def test_rowsplit_matches_its_specification_and_the_reference_order(dt, bc, n, w, levels):
    rng = np.random.default_rng(n * 7 + w + levels)
    g = grid(rng, n, dt)
    y = rng.normal(size=(n, w)).astype(dt)
    if bc == "Periodic":
        y[-1] = y[0]
    if bc == "Individual":
        rows, spec = individual(rng, w)
        mirror_bc, oracle_bc = BoundaryCondition.Individual([rows]), {"kind": "Individual", "rows": spec}
    else:
        mirror_bc, oracle_bc = getattr(BoundaryCondition, bc), {"kind": bc}
    interp = build(g, y, mirror_bc, "rowsplit", levels)
    used = interp.strategy.rowsplit_levels(interp)
    sys_rows = n - 2 if bc == "Periodic" else n
    assert used >= 1 and (sys_rows >> used) >= 2
    if levels:
        assert used == min(levels, max(lv for lv in range(1, 7) if (sys_rows >> lv) >= 2))
    a, b = interp.strategy.coefficients(interp)
    st, a_spec, b_spec = O.spline_build(g, y, oracle_bc, rowsplit_levels=used)
    assert st == O.ST_OK
    assert same(a, a_spec) and same(b, b_spec)
    st, a_ref, b_ref = O.spline_build(g, y, oracle_bc)
    ok, err = eval_close(g, y, a, b, a_ref, b_ref, rng)
    assert ok, err
    # the same handle rebuilt in the reference's order gives the reference's bits again
    seq = build(g, y, mirror_bc, "sequential")
    assert seq.strategy.rowsplit_levels(seq) == 0
    sa, sb = seq.strategy.coefficients(seq)
    assert same(sa, a_ref) and same(sb, b_ref)


def test_auto_keeps_the_reference_order_on_short_systems_and_partitions_long_ones():
    rng = np.random.default_rng(3)
    for n, expect in [(100, 0), (1023, 0), (1024, -32), (4096, -32), (70000, -32)]:
        g = grid(rng, n, np.float64)
        y = rng.normal(size=(n, 4))
        interp = build(g, y, BoundaryCondition.Natural, "auto")
        used = interp.strategy.rowsplit_levels(interp)
        assert used == expect, (n, used)
        a, b = interp.strategy.coefficients(interp)
        st, a_ref, b_ref = O.spline_build_as(g, y, {"kind": "Natural"}, used)
        assert same(a, a_ref) and same(b, b_ref)


def test_an_unforced_rowsplit_request_picks_its_depth_by_the_system_length():
    rng = np.random.default_rng(4)
    for n, expect in [(2048, 3), (4096, 4), (70000, 6)]:
        g = grid(rng, n, np.float64)
        y = rng.normal(size=(n, 4))
        interp = build(g, y, BoundaryCondition.Natural, "rowsplit")
        assert interp.strategy.rowsplit_levels(interp) == expect


def test_rowsplit_periodic_mismatch_is_reported():
    from ndarray_interp_b200 import BuilderError
    g = np.arange(64.0)
    y = np.random.default_rng(1).normal(size=(64, 5)); y[-1] = y[0]; y[-1, 3] += 1.0
    with pytest.raises(BuilderError.ValueError, match="first and last value must be equal"):
        build(g, y, BoundaryCondition.Periodic, "rowsplit", 3)


def test_rowsplit_long_columns_f64_and_many_columns_f32():
    """the shape the mode exists for (few long columns) and a many-column shape; knots are reproduced exactly"""
    rng = np.random.default_rng(11)
    for n, w, dt in [(65536, 8, np.float64), (4096, 3000, np.float32)]:
        g = grid(rng, n, dt)
        y = rng.normal(size=(n, w)).astype(dt)
        interp = build(g, y, BoundaryCondition.NotAKnot, "rowsplit")
        used = interp.strategy.rowsplit_levels(interp)
        a, b = interp.strategy.coefficients(interp)
        st, a_spec, b_spec = O.spline_build(g, y, {"kind": "NotAKnot"}, rowsplit_levels=used)
        assert same(a, a_spec) and same(b, b_spec)
        st, a_ref, b_ref = O.spline_build(g, y, {"kind": "NotAKnot"})
        ok, err = eval_close(g, y, a, b, a_ref, b_ref, rng)
        assert ok, err
        assert same(interp.interp_array(g[:-1]), y[:-1])
