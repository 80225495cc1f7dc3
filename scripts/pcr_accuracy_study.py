#!/usr/bin/env python3
"""Groundwork for the row-split spline build -- CPU only, numpy.

The shipped K6 solves the reference's tridiagonal system (cubic_spline.rs:409-674) with the reference's own
sequential Thomas order (:678-721), which is what makes the coefficients bit-identical -- and what makes a
build on few long columns chain-latency-bound.  This script measures what the alternative costs in accuracy:
k steps of parallel cyclic reduction (PCR) followed by Thomas on the 2^k interleaved systems, in f64 and f32,
against the sequential solve in the same precision and against an f64 (f32 case: f64) reference, on the
Natural boundary (SecondDeriv(0), :619-631, :656-668) and the grids of BASELINE.json's configurations.

Reported per case: the largest deviation of the evaluated spline, measured as SURVEY.md section 8(c) proposes,
|pcr - seq| / max(|seq|, max|y| of the column), next to north_star's bar (1e-12 for f64, 1e-5 for f32).
"""
import json
import sys

import numpy as np


def system_natural(x, y):
    """rows of A k = rhs as the reference forms them: interior :440-471, Natural rows :619-631 / :656-668"""
    dt = y.dtype.type
    n = len(x)
    dx = np.diff(x)
    low, mid, up = np.zeros(n, y.dtype), np.zeros(n, y.dtype), np.zeros(n, y.dtype)
    rhs = np.zeros_like(y)
    up[1:-1] = dx[:-1]                       # a_up = dx[n-1]
    mid[1:-1] = dt(2) * (dx[1:] + dx[:-1])
    low[1:-1] = dx[1:]                       # a_low = dx[n]
    dy = np.diff(y, axis=0)
    rhs[1:-1] = dt(3) * (dx[1:, None] * dy[:-1] / dx[:-1, None] + dx[:-1, None] * dy[1:] / dx[1:, None])
    # SecondDeriv(0) on both sides
    mid[0], up[0] = dt(2) * dx[0], dx[0]
    rhs[0] = dt(3) * dy[0]
    low[-1], mid[-1] = dx[-1], dt(2) * dx[-1]
    rhs[-1] = dt(3) * dy[-1]
    # the reference's thomas() names: a_low multiplies k[i-1], a_up multiplies k[i+1] (:690-720); in its interior
    # rows the coefficient of k[n-1] is dx[n] and of k[n+1] is dx[n-1]
    return low, mid, up, rhs


def thomas(low, mid, up, rhs):
    """thomas(), cubic_spline.rs:678-721, same operation order, every column at once"""
    n = len(mid)
    mid, rhs = mid.copy(), rhs.copy()
    for i in range(1, n):
        w = low[i] / mid[i - 1]
        mid[i] = mid[i] - w * up[i - 1]
        rhs[i] = rhs[i] - w * rhs[i - 1]
    k = np.zeros_like(rhs)
    k[-1] = rhs[-1] / mid[-1]
    for i in range(n - 2, -1, -1):
        k[i] = (rhs[i] - up[i] * k[i + 1]) / mid[i]
    return k


def pcr_then_thomas(low, mid, up, rhs, levels):
    n = len(mid)
    low, mid, up, rhs = low.copy(), mid.copy(), up.copy(), rhs.copy()
    s = 1
    for _ in range(levels):
        alpha, gamma = np.zeros_like(mid), np.zeros_like(mid)
        alpha[s:] = -low[s:] / mid[:-s]
        gamma[:-s] = -up[:-s] / mid[s:]
        nlow, nup, nmid, nrhs = np.zeros_like(low), np.zeros_like(up), mid.copy(), rhs.copy()
        nlow[s:] = alpha[s:] * low[:-s]
        nup[:-s] = gamma[:-s] * up[s:]
        nmid[s:] += alpha[s:] * up[:-s]
        nmid[:-s] += gamma[:-s] * low[s:]
        nrhs[s:] += alpha[s:, None] * rhs[:-s]
        nrhs[:-s] += gamma[:-s, None] * rhs[s:]
        low, mid, up, rhs = nlow, nmid, nup, nrhs
        s *= 2
    k = np.zeros_like(rhs)
    for j in range(min(s, n)):                # 2^levels independent systems: rows j, j+s, j+2s, ...
        sel = slice(j, n, s)
        k[sel] = thomas(low[sel], mid[sel], up[sel], rhs[sel])
    return k


def coefficients(x, y, k):
    dx = np.diff(x)[:, None]
    dy = np.diff(y, axis=0)
    return k[:-1] * dx - dy, dy - k[1:] * dx          # cubic_spline.rs:354-365


def evaluate(x, y, a, b, q):
    i = np.clip(np.searchsorted(x, q, side="right") - 1, 0, len(x) - 2)
    t = ((q - x[i]) / (x[i + 1] - x[i]))[:, None]
    one = y.dtype.type(1)
    return (one - t) * y[i] + t * y[i + 1] + t * (one - t) * (a[i] * (one - t) + b[i] * t)   # :825-827


def study(name, n, w, dt, levels_list, seed=0):
    rng = np.random.default_rng(seed)
    x64 = np.cumsum(rng.uniform(0.5, 1.5, n))
    y64 = rng.normal(size=(n, w))
    x, y = x64.astype(dt), y64.astype(dt)
    q = np.sort(rng.uniform(x[0], x[-1], 20000)).astype(dt)
    low, mid, up, rhs = system_natural(x, y)
    k_seq = thomas(low, mid, up, rhs)
    ref = evaluate(x, y, *coefficients(x, y, k_seq), q)
    scale = np.maximum(np.abs(ref), np.abs(y).max(axis=0)[None, :])
    # how far the sequential solve itself is from the exact answer in this precision (f32 only: f64 as truth)
    x_t, y_t = x.astype(np.float64), y.astype(np.float64)
    k_true = thomas(*system_natural(x_t, y_t))
    truth = evaluate(x_t, y_t, *coefficients(x_t, y_t, k_true), q.astype(np.float64))
    seq_err = float((np.abs(ref - truth) / scale).max())
    for lv in levels_list:
        k_pcr = pcr_then_thomas(low, mid, up, rhs, lv)
        val = evaluate(x, y, *coefficients(x, y, k_pcr), q)
        dev = float((np.abs(val - ref) / scale).max())
        err = float((np.abs(val - truth) / scale).max())
        print(json.dumps({"case": name, "rows": n, "columns": w, "dtype": np.dtype(dt).name, "pcr_levels": lv,
                          "chains": 2 ** lv, "chain_length": -(-n // 2 ** lv),
                          "max_dev_vs_sequential": dev, "max_err_vs_f64_truth": err,
                          "sequential_err_vs_f64_truth": seq_err,
                          "bar": 1e-12 if dt == np.float64 else 1e-5}), flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    study("c2 grid", 4096, 8, np.float64, [1, 3, 5])
    study("c2 grid", 4096, 8, np.float32, [1, 3, 5])
    if not quick:
        study("long grid", 65536, 4, np.float64, [3, 6, 8])
        study("long grid", 65536, 4, np.float32, [3, 6, 8])
