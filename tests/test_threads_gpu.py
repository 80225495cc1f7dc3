"""One handle, many host threads at once: the reference's methods take &self and are called from rayon
workers (benches/bench_interp1d.rs:49-79).  Handles are immutable after build and every call's state
lives in a thread-local workspace, so concurrent calls must give the same bits as serial ones."""
import threading

import numpy as np
import pytest

from ndarray_interp_b200 import _lib as L
from ndarray_interp_b200.interp1d import BoundaryCondition, CubicSpline, Interp1DBuilder, Linear
from ndarray_interp_b200.interp2d import Bilinear, Interp2D
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def test_shared_handles_from_eight_threads():
    rng = np.random.default_rng(77)
    n, w = 500, 24
    g = np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.normal(size=(n, w))
    lin = Interp1DBuilder.new(y).x(g).strategy(Linear.new()).build()
    cub = Interp1DBuilder.new(y).x(g).strategy(CubicSpline.new().boundary(BoundaryCondition.Natural)).build()
    st, a, b = O.spline_build(g, y, {"kind": "Natural"})
    gx, gy = np.linspace(0, 1, 300).astype(np.float32), np.cumsum(rng.uniform(0.5, 1.5, 200)).astype(np.float32)
    z = rng.normal(size=(300, 200, 16)).astype(np.float32)
    bil = Interp2D.new_unchecked(gx, gy, z, Bilinear.new())
    L.check(L.load().ndi_interp2d_set_binning(bil._handle(), L.BIN_ON, 16))     # the binned path allocates per-call scratch
    failures = []

    def worker(tid):
        try:
            r = np.random.default_rng(1000 + tid)
            for it in range(12):
                nq = int(r.choice([1, 7, 300, 5000, 300_000]))                   # latency path, single launch, chunked pipeline
                q = r.uniform(g[0], g[-1], nq)
                if not np.array_equal(lin.interp_array(q), O.interp1d_linear(g, y, q, False)[1]):
                    failures.append((tid, it, "linear"))
                if not np.array_equal(cub.interp_array(q), O.interp1d_cubic(g, y, a, b, q, 0)[1]):
                    failures.append((tid, it, "cubic"))
                qx = r.uniform(0, 1, nq).astype(np.float32)
                qy = r.uniform(gy[0], gy[-1], nq).astype(np.float32).clip(gy[0], gy[-1])
                if not np.array_equal(bil.interp_array(qx, qy), O.interp2d_bilinear(gx, gy, z, qx, qy, False)[1]):
                    failures.append((tid, it, "bilinear"))
        except Exception as e:                                                   # noqa: BLE001
            failures.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures, failures[:5]
