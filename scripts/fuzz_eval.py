#!/usr/bin/env python3
"""Longer-running companion of tests/test_fuzz_gpu.py: the same randomised cases (shapes, row widths and batch sizes around
the kernels' internal tile sizes; direct, binned and swept bilinear; every element type; spline build + evaluation; per-column
boundaries), for many more seeds.  Every case is compared with the oracle bit for bit.

    python scripts/fuzz_eval.py [first_seed] [count]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_fuzz_gpu as F  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 44
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    for seed in range(first, first + count):
        F.test_fuzz_linear_and_bilinear(seed)
        F.test_fuzz_spline_build_and_eval(seed)
        if seed % 4 == 0:
            F.test_fuzz_individual_boundaries(seed)
    print(f"seeds {first} .. {first + count - 1}: linear + bilinear (direct, binned, swept), spline build + evaluation"
          f"{', per-column boundaries' if count >= 4 else ''}: all bit-identical to the oracle")


if __name__ == "__main__":
    main()
