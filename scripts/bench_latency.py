#!/usr/bin/env python3
"""Latency of the scalar / tiny-batch host calls (interp_scalar, interp): microseconds per C-ABI call."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndarray_interp_b200 import _lib as L  # noqa: E402
from ndarray_interp_b200.interp1d import BoundaryCondition, CubicSpline, Interp1DBuilder, Linear  # noqa: E402
from ndarray_interp_b200.interp2d import Interp2DBuilder  # noqa: E402


def time_call(fn, reps=2000):
    for _ in range(50):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    lib = L.require_device()
    rng = np.random.default_rng(0)
    out = []
    g = np.cumsum(rng.uniform(0.5, 1.5, 1000))
    for w in (1, 64):
        y = rng.normal(size=(1000, w)) if w > 1 else rng.normal(size=1000)
        lin = Interp1DBuilder.new(y).x(g).strategy(Linear.new()).build()
        cub = Interp1DBuilder.new(y).x(g).strategy(CubicSpline.new().boundary(BoundaryCondition.Natural)).build()
        for nq in (1, 16):
            q = rng.uniform(g[0], g[-1], nq)
            buf = np.zeros((nq, w))
            bad = C.c_int64(-1)
            us = time_call(lambda: lib.ndi_interp1d_linear(lin._handle(), L.ptr(q), nq, 0, L.ptr(buf), C.byref(bad)))
            out.append({"call": "ndi_interp1d_linear", "columns": w, "queries": nq, "us_per_call": round(us, 2)})
            us = time_call(lambda: lib.ndi_interp1d_cubic(cub._handle(), L.ptr(q), nq, 0, L.ptr(buf), C.byref(bad)))
            out.append({"call": "ndi_interp1d_cubic", "columns": w, "queries": nq, "us_per_call": round(us, 2)})
    z = rng.normal(size=(200, 100, 8)).astype(np.float32)
    bil = Interp2DBuilder.new(z).build()
    qx, qy = np.float32([17.3]), np.float32([42.9])
    buf = np.zeros((1, 8), np.float32)
    bad, ax = C.c_int64(-1), C.c_int32(-1)
    us = time_call(lambda: lib.ndi_interp2d_bilinear(bil._handle(), L.ptr(qx), L.ptr(qy), 1, 0, L.ptr(buf), C.byref(bad), C.byref(ax)))
    out.append({"call": "ndi_interp2d_bilinear", "columns": 8, "queries": 1, "us_per_call": round(us, 2)})
    lin1 = Interp1DBuilder.new(rng.normal(size=1000)).x(g).build()
    us = time_call(lambda: lin1.interp_scalar(123.4), reps=500)
    out.append({"call": "Interp1D.interp_scalar (Python mirror)", "columns": 1, "queries": 1, "us_per_call": round(us, 2)})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
