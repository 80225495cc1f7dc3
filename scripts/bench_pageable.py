#!/usr/bin/env python3
"""Host-array API with ordinary (pageable) numpy buffers vs pinned ones: GB/s of result rows delivered."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndarray_interp_b200.interp1d import Interp1D, Linear  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    n, w, nq = 65536, 16, 1 << 24
    g = np.cumsum(rng.uniform(0.5, 1.5, n)).astype(np.float32)
    y = rng.standard_normal((n, w), dtype=np.float32)
    q = rng.uniform(g[0], g[-1], nq).astype(np.float32)
    ip = Interp1D.new_unchecked(g, y, Linear.new().extrapolate(True))
    out_page = np.zeros((nq, w), np.float32)
    out_pin = torch.empty((nq, w), dtype=torch.float32, pin_memory=True).numpy()
    q_pin = torch.from_numpy(q).pin_memory().numpy()
    for name, qq, oo in (("pageable", q, out_page), ("pinned", q_pin, out_pin)):
        ip.interp_array_into(qq, oo)
        t0 = time.perf_counter()
        for _ in range(3):
            ip.interp_array_into(qq, oo)
        dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"buffers": name, "ms": round(dt * 1e3, 2), "result_GBps": round(oo.nbytes / dt / 1e9, 2)}))
    assert np.array_equal(out_page, out_pin)


if __name__ == "__main__":
    main()
