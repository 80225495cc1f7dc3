#!/usr/bin/env python3
"""per-call wall time of consecutive ndi_interp1d_spline_build calls (is the first boundary kind slow, or the first calls?)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndarray_interp_b200 import device as D  # noqa: E402

D.set_device(0)
for name, n, w, dt in [("c2", 4096, 1024, torch.float64), ("c5b-shard", 4096, 16384, torch.float32)]:
    for order in ([0, 1, 0, 1], [1, 0, 1, 0]):
        g = torch.cumsum(torch.rand(n, dtype=torch.float64, device="cuda") + 0.5, 0).to(dt)
        y = torch.randn(n, w, dtype=dt, device="cuda")
        ip = D.DeviceInterp1D(g, y)
        for code in order:
            ts = []
            for _ in range(6):
                t0 = time.perf_counter()
                st, _ = ip.spline_build(code)
                ts.append((time.perf_counter() - t0) * 1e3)
            print(name, "NotAKnot" if code == 0 else "Natural", " ".join(f"{t:.2f}" for t in ts), flush=True)
        del ip
