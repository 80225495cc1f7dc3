"""Partition spline build (K6, third build mode: blocks of rows solved in registers, the separator rows' system
recursively, csrc/ndi_partition.cu; replaces the serial chains of CubicSpline::solve_for_k + thomas,
cubic_spline.rs:409-721).

Two bars, both asserted for every boundary kind:
  * bit for bit against the oracle's operation-by-operation specification of the same scheme
    (oracle/ndi_oracle.cpp: partition_thomas, same block size);
  * north_star's 1e-12 (f64) / 1e-5 (f32) against the oracle in the REFERENCE's elimination order, for the
    coefficients' effect on evaluated values (measured like tests/test_rowsplit_gpu.py).
"""
import numpy as np
import pytest

from ndarray_interp_b200.interp1d import BoundaryCondition
from oracle import oracle_py as O
from test_parity_gpu import same
from test_rowsplit_gpu import build, eval_close, grid, individual

pytestmark = pytest.mark.gpu

# (rows, columns, requested block): direct solve only (8, 9), one split level, two and three split levels, tails of
# every kind (none, one row, a full block), the default block at the sizes the mode exists for
SHAPES = [(8, 3, 3), (9, 5, 3), (37, 33, 4), (100, 1, 3), (257, 40, 8), (1000, 7, 5), (1031, 64, 32), (4096, 65, 0),
          (5000, 2, 32), (1024, 31, 16), (129, 200, 32), (4100, 12, 64), (8191, 3, 48), (640, 70, 100)]


def block_used(requested):
    return min(max(requested, 3), 64) if requested else 32


def case(dt, bc, n, w, block):
    rng = np.random.default_rng(n * 7 + w + block)
    g = grid(rng, n, dt)
    y = rng.normal(size=(n, w)).astype(dt)
    if bc == "Periodic":
        y[-1] = y[0]
    if bc == "Individual":
        rows, spec = individual(rng, w)
        return rng, g, y, BoundaryCondition.Individual([rows]), {"kind": "Individual", "rows": spec}
    return rng, g, y, getattr(BoundaryCondition, bc), {"kind": bc}


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("bc", ["NotAKnot", "Natural", "Clamped", "Periodic", "Individual"])
@pytest.mark.parametrize("n,w,block", SHAPES)
def test_partition_matches_its_specification_and_the_reference_order(dt, bc, n, w, block):
    rng, g, y, mirror_bc, oracle_bc = case(dt, bc, n, w, block)
    interp = build(g, y, mirror_bc, "partition", block)
    used = interp.strategy.rowsplit_levels(interp)
    assert used == -block_used(block)
    a, b = interp.strategy.coefficients(interp)
    st, a_spec, b_spec = O.spline_build_as(g, y, oracle_bc, used)
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


def test_partition_periodic_mismatch_is_reported():
    from ndarray_interp_b200 import BuilderError
    g = np.arange(640.0)
    y = np.random.default_rng(1).normal(size=(640, 5)); y[-1] = y[0]; y[-1, 3] += 1.0
    with pytest.raises(BuilderError.ValueError, match="first and last value must be equal"):
        build(g, y, BoundaryCondition.Periodic, "partition", 8)


def test_partition_long_columns_f64_and_many_columns_f32():
    """few long columns (two split levels) and a many-column shape; knots are reproduced exactly"""
    rng = np.random.default_rng(11)
    for n, w, dt in [(65536, 8, np.float64), (4096, 3000, np.float32), (70001, 3, np.float64)]:
        g = grid(rng, n, dt)
        y = rng.normal(size=(n, w)).astype(dt)
        interp = build(g, y, BoundaryCondition.NotAKnot, "partition", 64 if n == 70001 else 0)
        used = interp.strategy.rowsplit_levels(interp)
        assert used == (-64 if n == 70001 else -32)
        a, b = interp.strategy.coefficients(interp)
        st, a_spec, b_spec = O.spline_build_as(g, y, {"kind": "NotAKnot"}, used)
        assert same(a, a_spec) and same(b, b_spec)
        st, a_ref, b_ref = O.spline_build(g, y, {"kind": "NotAKnot"})
        ok, err = eval_close(g, y, a, b, a_ref, b_ref, rng)
        assert ok, err
        assert same(interp.interp_array(g[:-1]), y[:-1])


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
def test_auto_keeps_the_reference_order_on_a_singular_not_a_knot_system(dt):
    """the reference's NotAKnot system takes x[n-1] - x[n-2] for its last diagonal entry (cubic_spline.rs:635); its last
    pivot vanishes when the last grid step is about 0.55 of the one before it, and the reference's result is then only
    reproduced by its own order of operations: AUTO keeps that order there (bit-identical coefficients), takes the
    partition build for the same grid under another boundary, and for NotAKnot on a grid without that property"""
    rng = np.random.default_rng(5)
    n, w = 2000, 6
    g = np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.normal(size=(n, w)).astype(dt)
    for last_over_prev, expect in ((0.5359, 0), (1.0, -32), (0.2, -32)):
        gg = g.copy()
        gg[-1] = gg[-2] + last_over_prev * (gg[-2] - gg[-3])
        gg = gg.astype(dt)
        interp = build(gg, y, BoundaryCondition.NotAKnot, "auto")
        info = interp.strategy.rowsplit_levels(interp)
        a, b = interp.strategy.coefficients(interp)
        st, a_ref, b_ref = O.spline_build_as(gg, y, {"kind": "NotAKnot"}, info)
        assert st == O.ST_OK and same(a, a_ref) and same(b, b_ref)
        if expect == 0:
            assert info == 0, (last_over_prev, info)
        nat = build(gg, y, BoundaryCondition.Natural, "auto")
        assert nat.strategy.rowsplit_levels(nat) == -32
        # per-column boundaries: one NotAKnot on the right is enough
        rows, spec = individual(rng, w)
        ind = build(gg, y, BoundaryCondition.Individual([rows]), "auto")
        has_right_nak = any(s.get("kind") == "NotAKnot" or (s.get("kind") == "Mixed" and s["right"]["kind"] == "NotAKnot") for s in spec)
        if expect == 0 and has_right_nak:
            assert ind.strategy.rowsplit_levels(ind) == 0
