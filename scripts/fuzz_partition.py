#!/usr/bin/env python3
"""Randomised sweep of the partition spline build against the oracle's specification (bit for bit) and against the
reference-order oracle (north_star's bars): rows 4 .. 20000, columns 1 .. 300, blocks 3 .. 64, every boundary kind, f32
and f64.  A longer-running companion of tests/test_partition_gpu.py.

    python scripts/fuzz_partition.py [cases] [seed] [auto]

With `auto` the tables have 1024 rows or more and are built with NDI_BUILD_AUTO: the partition build, except where a right
NotAKnot row meets a grid that makes the reference's system nearly singular (there: the reference's order, bit-identical);
a third of the NotAKnot grids are drawn INTO that region.  Under AUTO no case may exceed the bars.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ndarray_interp_b200.interp1d import BoundaryCondition  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from test_parity_gpu import same  # noqa: E402
from test_rowsplit_gpu import build, eval_close, individual  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
    auto = len(sys.argv) > 3 and sys.argv[3] == "auto"
    worst = {"float32": 0.0, "float64": 0.0}
    over = 0
    infos = {}
    for k in range(cases):
        dt = [np.float32, np.float64][k % 2]
        n = int(rng.choice([rng.integers(4, 40), rng.integers(40, 700), rng.integers(700, 5000), rng.integers(5000, 20000)]))
        w = int(rng.choice([1, 2, 3, rng.integers(4, 40), rng.integers(40, 300)]))
        block = int(rng.choice([0, 3, 4, 5, 7, 8, 16, 31, 32, 33, 48, 64]))
        bc = ["NotAKnot", "Natural", "Clamped", "Periodic", "Individual"][int(rng.integers(0, 5))]
        if auto:
            n, block = max(n, 1024 + n % 977), 0
        kind = int(rng.integers(0, 3))
        if kind == 0:
            g = np.cumsum(rng.uniform(0.5, 1.5, n))
        elif kind == 1:
            g = np.cumsum(np.exp(rng.uniform(-2, 2, n)))           # steps over two decades
        else:
            g = np.arange(n) * 0.25
        if auto and k % 3 == 0:                                   # the last step near 0.55 of the one before it
            g[-1] = g[-2] + rng.uniform(0.45, 0.65) * (g[-2] - g[-3])
        g = g.astype(dt)
        y = rng.normal(size=(n, w)).astype(dt)
        if bc == "Periodic":
            y[-1] = y[0]
        if bc == "Individual":
            rows, spec = individual(rng, w)
            mirror_bc, oracle_bc = BoundaryCondition.Individual([rows]), {"kind": "Individual", "rows": spec}
        else:
            mirror_bc, oracle_bc = getattr(BoundaryCondition, bc), {"kind": bc}
        interp = build(g, y, mirror_bc, "auto" if auto else "partition", block)
        used = interp.strategy.rowsplit_levels(interp)
        infos[used] = infos.get(used, 0) + 1
        a, b = interp.strategy.coefficients(interp)
        st, a_spec, b_spec = O.spline_build_as(g, y, oracle_bc, used)
        assert st == O.ST_OK
        assert same(a, a_spec) and same(b, b_spec), (k, dt.__name__, n, w, block, bc, kind)
        st, a_ref, b_ref = O.spline_build(g, y, oracle_bc)
        ok, err = eval_close(g, y, a, b, a_ref, b_ref, rng)
        worst[dt.__name__] = max(worst[dt.__name__], err)
        over += 0 if ok else 1
        if not ok:
            # is the reference order itself that far from an f64 solve of the same system?  (NotAKnot rows are not dominant)
            st, a64, b64 = O.spline_build(g.astype(np.float64), y.astype(np.float64), oracle_bc)
            _, e_ref = eval_close(g.astype(np.float64), y.astype(np.float64), a_ref.astype(np.float64), b_ref.astype(np.float64), a64, b64, rng)
            _, e_par = eval_close(g.astype(np.float64), y.astype(np.float64), a.astype(np.float64), b.astype(np.float64), a64, b64, rng)
            print("over the bar against the reference order:", k, dt.__name__, n, w, block, bc, kind, "%.3g" % err,
                  "| against an f64 solve: reference order %.3g, partition %.3g" % (e_ref, e_par), flush=True)
    print(f"{cases} cases bit-identical to the specification; worst distance to the reference order: {worst}; over the bar: {over}; "
          f"builds by ndi_interp1d_build_info: {infos}")
    if auto:
        assert over == 0


if __name__ == "__main__":
    main()
