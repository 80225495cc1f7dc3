#!/usr/bin/env python3
"""K6 timing: CubicSpline coefficient construction for a few table shapes, in both build modes -- the reference's
elimination order ("sequential"), the row-split PCR + Thomas build at several depths and the partition build at several
block sizes (one JSON line each).

    python scripts/bench_spline_build.py [shape ...] [--levels 1,2,3,4,5,6] [--blocks 0,16] [--bc Natural,Periodic]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndarray_interp_b200 import _lib as L  # noqa: E402
from ndarray_interp_b200 import device as D  # noqa: E402

SHAPES = [("c2", 4096, 1024, torch.float64), ("c5b-shard", 4096, 16384, torch.float32), ("wide", 512, 262144, torch.float32),
          ("long", 65536, 64, torch.float64)]
BC = {"NotAKnot": 0, "Natural": 1, "Periodic": 3, "Individual": 4}


def main():
    D.set_device(0)
    args = sys.argv[1:]
    levels = [0]
    blocks = []
    bcs = list(BC)
    only = []
    i = 0
    while i < len(args):
        if args[i] == "--levels":
            levels = [int(v) for v in args[i + 1].split(",")]; i += 2
        elif args[i] == "--blocks":
            blocks = [int(v) for v in args[i + 1].split(",")]; i += 2
        elif args[i] == "--bc":
            bcs = args[i + 1].split(","); i += 2
        else:
            only.append(args[i]); i += 1
    modes = ([("sequential", L.BUILD_SEQUENTIAL, 0)] + [("rowsplit", L.BUILD_ROWSPLIT, lv) for lv in levels]
             + [("partition", L.BUILD_PARTITION, bl) for bl in blocks])
    for name, n, w, dt in SHAPES:
        if only and name not in only:
            continue
        g = torch.cumsum(torch.rand(n, dtype=torch.float64, device="cuda") + 0.5, 0).to(dt)
        y = torch.randn(n, w, dtype=dt, device="cuda")
        y[-1] = y[0]                                        # so that Periodic is admissible
        ip = D.DeviceInterp1D(g, y)
        rng = np.random.default_rng(0)
        ndt = np.float64 if dt == torch.float64 else np.float32
        ind = (rng.integers(0, 5, w).astype(np.int32), rng.normal(size=w).astype(ndt),
               rng.integers(0, 5, w).astype(np.int32), rng.normal(size=w).astype(ndt))
        for bc in bcs:
            code = BC[bc]
            extra = ind if bc == "Individual" else ()
            for mode_name, mode, lv in modes:
                ip.set_build_mode(mode, lv)
                for _ in range(6):                          # the first calls on a new handle grow the per-thread scratch
                    ip.spline_build(code, *extra)
                torch.cuda.synchronize()
                reps, t0 = 10, time.perf_counter()
                for _ in range(reps):
                    st, _ = ip.spline_build(code, *extra)
                    assert st == 0
                ms = (time.perf_counter() - t0) / reps * 1e3
                es = 8 if dt == torch.float64 else 4
                print(json.dumps({"shape": name, "rows": n, "columns": w, "dtype": str(dt), "boundary": bc, "mode": mode_name,
                                  "levels": ip.build_levels(), "ms": round(ms, 4),
                                  "algorithmic_GBps": round(es * (3 * n - 2) * w / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
