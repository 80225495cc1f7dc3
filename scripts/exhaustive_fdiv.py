#!/usr/bin/env python3
"""All 2^46 pairs of f32 mantissas: div_by(a, b, rcp_refined(b)) == __fdiv_rn(a, b) on this GPU?

    python scripts/exhaustive_fdiv.py [--chunk 8192] [--out profiles/r01/fdiv_exhaustive.txt]

The quotient's rounding depends on the two significands only (every operation of the sequence is
exact under scaling by powers of two while nothing leaves the normal range), so exponent 0 for both
operands covers the kernels' whole admitted range; a sample at the range ends is checked as well.
About two minutes on one B200."""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndarray_interp_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=8192)
    ap.add_argument("--limit", type=int, default=1 << 23, help="numerator mantissas to cover (default: all)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    lib = L.require_device()
    bad_total, t0 = 0, time.time()
    for s in range(0, a.limit, a.chunk):
        bad = C.c_uint64(0)
        L.check(lib.ndi_selftest_fdiv(s, min(a.chunk, a.limit - s), 0, 0, C.byref(bad)))
        bad_total += bad.value
    pairs = a.limit * (1 << 23)
    edge_bad = 0
    for ea, eb in [(-80, 40), (80, -40), (-80, -40), (80, 40)]:
        for s in range(0, 1 << 23, 1 << 17):
            bad = C.c_uint64(0)
            L.check(lib.ndi_selftest_fdiv(s, 16, ea, eb, C.byref(bad)))
            edge_bad += bad.value
    msg = (f"div_by vs __fdiv_rn: {pairs} mantissa pairs at exponent 0 ({a.limit} numerators x 2^23 divisors): "
           f"{bad_total} mismatches; range-end sample (4 exponent corners x 1024 numerators x 2^23 divisors): "
           f"{edge_bad} mismatches; {time.time() - t0:.1f} s")
    print(msg)
    if a.out:
        with open(a.out, "w") as f:
            f.write(msg + "\n")
    return 0 if bad_total == 0 and edge_bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
