"""ctypes loader for the CPU oracle (oracle/ndi_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package ndarray_interp_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libndi_oracle.so")

ST_OK, ST_OUT_OF_BOUNDS, ST_NAN_QUERY, ST_PERIODIC_MISMATCH, ST_INVALID_ARGUMENT = 0, 1, 2, 3, 4
MONO_NAMES = ["NotMonotonic", "RisingStrict", "Rising", "FallingStrict", "Falling"]
BC = {"NotAKnot": 0, "Natural": 1, "Clamped": 2, "Periodic": 3, "Individual": 4}
SB = {"NotAKnot": 0, "Natural": 1, "Clamped": 2, "FirstDeriv": 3, "SecondDeriv": 4}

_SFX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.int32): "i32",
        np.dtype(np.int64): "i64", np.dtype(np.uint32): "u32", np.dtype(np.uint64): "u64"}
_CT = {"f32": C.c_float, "f64": C.c_double, "i32": C.c_int32, "i64": C.c_int64, "u32": C.c_uint32, "u64": C.c_uint64}


def build(force=False):
    """compile the oracle with the committed Makefile (g++ -O2 -ffp-contract=off)"""
    src = os.path.join(_HERE, "ndi_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libndi_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.ora_hardware_threads.restype = C.c_int32
        for s, ct in _CT.items():
            getattr(_lib, f"ora_calc_frac_{s}").restype = ct
            getattr(_lib, f"ora_calc_frac_{s}").argtypes = [ct] * 5
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _sfx(a):
    return _SFX[np.dtype(a.dtype)]


def hardware_threads():
    return int(lib().ora_hardware_threads())


def calc_frac(x1, y1, x2, y2, x, dtype=np.float64):
    s = _SFX[np.dtype(dtype)]
    return getattr(lib(), f"ora_calc_frac_{s}")(x1, y1, x2, y2, x)


def monotonic_prop(x, stride=1):
    """x: 1-D contiguous array; stride in elements (negative walks backwards from the end)"""
    x = np.ascontiguousarray(x)
    base = x.ctypes.data if stride > 0 else x.ctypes.data + (len(x) - 1) * x.itemsize
    n_eff = (len(x) + abs(stride) - 1) // abs(stride) if len(x) else 0
    r = getattr(lib(), f"ora_monotonic_prop_{_sfx(x)}")(C.c_void_p(base), C.c_int64(n_eff), C.c_int64(stride))
    return MONO_NAMES[r]


def lower_index(grid, q):
    grid = np.ascontiguousarray(grid)
    q = np.ascontiguousarray(q, dtype=grid.dtype).ravel()
    out = np.zeros(q.shape, dtype=np.int64)
    bad = C.c_int64(-1)
    st = getattr(lib(), f"ora_lower_index_{_sfx(grid)}")(_p(grid), C.c_int64(len(grid)), _p(q), C.c_int64(len(q)),
                                                         _p(out), C.byref(bad))
    return st, out, bad.value


def interp1d_linear(x, data, q, extrapolate, nthreads=0, out=None):
    x = np.ascontiguousarray(x)
    data = np.ascontiguousarray(data, dtype=x.dtype)
    q = np.ascontiguousarray(q, dtype=x.dtype)
    n = len(x)
    w = int(np.prod(data.shape[1:], dtype=np.int64))
    if out is None:
        out = np.zeros(q.shape + data.shape[1:], dtype=x.dtype)
    bad = C.c_int64(-1)
    args = [_p(x), C.c_int64(n), _p(data), C.c_int64(w), _p(q), C.c_int64(q.size), C.c_int32(int(extrapolate)),
            _p(out), C.byref(bad)]
    if nthreads:
        st = getattr(lib(), f"ora_interp1d_linear_mt_{_sfx(x)}")(*args, C.c_int32(nthreads))
    else:
        st = getattr(lib(), f"ora_interp1d_linear_{_sfx(x)}")(*args)
    return st, out, bad.value


def interp2d_bilinear(x, y, data, qx, qy, extrapolate, nthreads=0, out=None):
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y, dtype=x.dtype)
    data = np.ascontiguousarray(data, dtype=x.dtype)
    qx = np.ascontiguousarray(qx, dtype=x.dtype)
    qy = np.ascontiguousarray(qy, dtype=x.dtype)
    w = int(np.prod(data.shape[2:], dtype=np.int64))
    if out is None:
        out = np.zeros(qx.shape + data.shape[2:], dtype=x.dtype)
    bad, axis = C.c_int64(-1), C.c_int32(-1)
    args = [_p(x), C.c_int64(len(x)), _p(y), C.c_int64(len(y)), _p(data), C.c_int64(w), _p(qx), _p(qy),
            C.c_int64(qx.size), C.c_int32(int(extrapolate)), _p(out), C.byref(bad), C.byref(axis)]
    if nthreads:
        st = getattr(lib(), f"ora_interp2d_bilinear_mt_{_sfx(x)}")(*args, C.c_int32(nthreads))
    else:
        st = getattr(lib(), f"ora_interp2d_bilinear_{_sfx(x)}")(*args)
    return st, out, bad.value, axis.value


def bc_arrays(bc, w, dtype):
    """boundary spec (dict like the golden fixtures) -> (kind, left_kind, left_val, right_kind, right_val)"""
    kind = BC[bc["kind"]]
    if bc["kind"] != "Individual":
        return kind, None, None, None, None
    rows = bc["rows"]
    assert len(rows) == w, (len(rows), w)
    lk = np.zeros(w, np.int32); rk = np.zeros(w, np.int32)
    lv = np.zeros(w, dtype); rv = np.zeros(w, dtype)
    for i, r in enumerate(rows):
        if r["kind"] == "Mixed":
            lk[i], rk[i] = SB[r["left"]["kind"]], SB[r["right"]["kind"]]
            lv[i], rv[i] = r["left"].get("value", 0.0), r["right"].get("value", 0.0)
        else:
            lk[i] = rk[i] = SB[r["kind"]]
    return kind, lk, lv, rk, rv


def spline_build(x, data, bc, rowsplit_levels=0, partition_block=0):
    """rowsplit_levels > 0: the row-split variant of the solve (NOT the reference's order; the specification the
    row-split build kernels are compared with -- ndi_oracle.cpp, rowsplit_thomas); partition_block > 0: the
    partition variant with blocks of that many rows (ndi_oracle.cpp, partition_thomas)"""
    x = np.ascontiguousarray(x)
    data = np.ascontiguousarray(data, dtype=x.dtype)
    n = len(x)
    w = int(np.prod(data.shape[1:], dtype=np.int64))
    kind, lk, lv, rk, rv = bc_arrays(bc, w, x.dtype)
    a = np.zeros((n - 1,) + data.shape[1:], dtype=x.dtype)
    b = np.zeros_like(a)
    if partition_block:
        st = getattr(lib(), f"ora_spline_build_partition_{_sfx(x)}")(
            _p(x), C.c_int64(n), _p(data), C.c_int64(w), C.c_int32(kind), _p(lk), _p(lv), _p(rk), _p(rv),
            C.c_int32(partition_block), _p(a), _p(b))
        return st, a, b
    if rowsplit_levels:
        st = getattr(lib(), f"ora_spline_build_rowsplit_{_sfx(x)}")(
            _p(x), C.c_int64(n), _p(data), C.c_int64(w), C.c_int32(kind), _p(lk), _p(lv), _p(rk), _p(rv),
            C.c_int32(rowsplit_levels), _p(a), _p(b))
        return st, a, b
    st = getattr(lib(), f"ora_spline_build_{_sfx(x)}")(_p(x), C.c_int64(n), _p(data), C.c_int64(w), C.c_int32(kind),
                                                       _p(lk), _p(lv), _p(rk), _p(rv), _p(a), _p(b))
    return st, a, b


def spline_build_as(x, data, bc, build_info):
    """the specification matching ndi_interp1d_build_info's value: 0 reference order, L > 0 row-split, -m partition"""
    return spline_build(x, data, bc, rowsplit_levels=max(int(build_info), 0), partition_block=max(-int(build_info), 0))


def interp1d_cubic(x, data, a, b, q, extrap_mode, nthreads=0, out=None):
    x = np.ascontiguousarray(x)
    data = np.ascontiguousarray(data, dtype=x.dtype)
    q = np.ascontiguousarray(q, dtype=x.dtype)
    w = int(np.prod(data.shape[1:], dtype=np.int64))
    if out is None:
        out = np.zeros(q.shape + data.shape[1:], dtype=x.dtype)
    bad = C.c_int64(-1)
    args = [_p(x), C.c_int64(len(x)), _p(data), _p(a), _p(b), C.c_int64(w), _p(q), C.c_int64(q.size),
            C.c_int32(extrap_mode), _p(out), C.byref(bad)]
    if nthreads:
        st = getattr(lib(), f"ora_interp1d_cubic_mt_{_sfx(x)}")(*args, C.c_int32(nthreads))
    else:
        st = getattr(lib(), f"ora_interp1d_cubic_{_sfx(x)}")(*args)
    return st, out, bad.value


def cubic_interp(x, data, bc, extrapolate, q):
    """build + evaluate, returning (status, values) the way Interp1D+CubicSpline would"""
    st, a, b = spline_build(x, data, bc)
    if st != ST_OK:
        return st, None
    mode = 0 if not extrapolate else (2 if bc["kind"] == "Periodic" else 1)
    st, out, _ = interp1d_cubic(x, data, a, b, q, mode)
    return st, out
