"""Randomised differential parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Everything is compared BIT FOR BIT -- the kernels keep the reference's
operation order (no FMA, IEEE division), so the 4-ulp (linear / bilinear) and 1e-12 / 1e-5
relative (cubic, f64 / f32) bars of BASELINE.json are met with zero slack.  The tolerance the
spec allows is asserted as well, so a future kernel that trades exactness for speed has a bar
to be held to."""
import ctypes as C

import numpy as np
import pytest

from ndarray_interp_b200 import BuilderError, InterpolateError, Panic, _lib as L
from ndarray_interp_b200.interp1d import Interp1D, Linear
from ndarray_interp_b200.interp2d import Bilinear, Interp2D
from ndarray_interp_b200.vector_extensions import Monotonic, get_lower_index, monotonic_prop
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu

DTS = [np.float32, np.float64, np.int32, np.int64, np.uint32, np.uint64]


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return bool(np.array_equal(a.view(np.uint8), b.view(np.uint8))) or bool(
        np.array_equal(a, b, equal_nan=np.issubdtype(a.dtype, np.floating)))


def ulp_diff(a, b):
    """max distance in units of last place (floats)"""
    it = {4: np.int32, 8: np.int64}[a.itemsize]
    ai, bi = a.view(it).astype(np.int64), b.view(it).astype(np.int64)
    ai = np.where(ai < 0, np.iinfo(it).min - ai, ai)
    bi = np.where(bi < 0, np.iinfo(it).min - bi, bi)
    return int(np.abs(ai - bi).max(initial=0))


def make_grid(rng, n, dt, kind):
    if np.issubdtype(dt, np.integer):
        if kind == "uniform":
            return (np.arange(n) + 5).astype(dt)
        return np.cumsum(rng.integers(1, 9, n)).astype(dt)
    if kind == "uniform":
        return np.linspace(-3.0, 11.0, n).astype(dt)
    if kind == "exp":
        g = np.cumsum(np.exp(rng.uniform(-4, 2, n)))
        return g.astype(dt)
    g = np.cumsum(rng.uniform(0.5, 1.5, n))
    return g.astype(dt)


def make_queries(rng, g, nq, dt, outside):
    lo, hi = float(g[0]), float(g[-1])
    span = hi - lo
    if outside:
        q = rng.uniform(lo - 0.2 * span, hi + 0.2 * span, nq)
    else:
        q = rng.uniform(lo, hi, nq)
    if np.issubdtype(dt, np.integer):
        q = np.round(q)
        if not outside:
            q = np.clip(q, g[0], g[-1])
        if np.issubdtype(dt, np.unsignedinteger):
            q = np.maximum(q, 0.0)                    # below the grid, but not below zero
        return q.astype(dt)
    q = q.astype(dt)
    if not outside:
        q = np.clip(q, g[0], g[-1])
    # sprinkle exact knots and the two ends
    k = min(nq, 8)
    q[:k] = g[rng.integers(0, len(g), k)]
    if nq > 9:
        q[8], q[9] = g[0], g[-1]
    return q


def make_data(rng, shape, dt):
    if np.issubdtype(dt, np.unsignedinteger):
        return rng.integers(0, 2000, shape).astype(dt)   # falling data: y2 - y1 wraps, as in a release build of the reference
    if np.issubdtype(dt, np.integer):
        return rng.integers(-1000, 1000, shape).astype(dt)
    return rng.normal(size=shape).astype(dt)


# ---- K2 ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTS, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("kind", ["uniform", "random", "exp"])
@pytest.mark.parametrize("n", [2, 3, 10, 1000, 65536, 70001])
def test_lower_index_matches_oracle(dt, kind, n):
    if kind == "exp" and np.issubdtype(dt, np.integer):
        pytest.skip("float only")
    rng = np.random.default_rng(n * 7 + len(kind))
    g = make_grid(rng, n, dt, kind)
    if len(np.unique(g)) != n:
        pytest.skip("grid collapsed in this dtype")
    for nq in (1, 33, 40000):
        q = make_queries(rng, g, nq, dt, outside=True)
        if np.issubdtype(dt, np.floating) and nq > 20:
            q[10], q[11] = np.inf, -np.inf
        st, ref, _ = O.lower_index(g, q)
        assert st == O.ST_OK
        got = get_lower_index(g, q)
        assert np.array_equal(got, ref)          # indices are bit-exact by contract


def test_lower_index_nan_reports_first():
    g = np.linspace(0, 1, 50)
    q = np.random.default_rng(0).uniform(0, 1, 5000)
    q[[4000, 1234, 4999]] = np.nan
    lib = L.require_device()
    idx = np.zeros(q.shape, np.int64)
    bad = C.c_int64(-1)
    st = lib.ndi_lower_index(L.F64, L.ptr(g), len(g), L.ptr(q), q.size, L.ptr(idx), C.byref(bad))
    assert (st, bad.value) == (L.NAN_QUERY, 1234)
    with pytest.raises(Panic):
        get_lower_index(g, q)


# ---- K1 ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 3, 31, 257, 5000, 300001])
def test_monotonic_matches_oracle(n):
    rng = np.random.default_rng(n)
    names = {"NotMonotonic": Monotonic.NotMonotonic, "RisingStrict": Monotonic.Rising(True),
             "Rising": Monotonic.Rising(False), "FallingStrict": Monotonic.Falling(True),
             "Falling": Monotonic.Falling(False)}
    for dt in DTS:
        base = np.sort(rng.integers(0, 10 * n, n)) if np.issubdtype(dt, np.integer) else np.sort(rng.uniform(0, 1, n))
        variants = [base, base[::-1].copy(), np.unique(base) if len(np.unique(base)) > 1 else base]
        strict = np.arange(n) * 3
        variants += [strict, strict[::-1].copy(), np.zeros(n)]
        for pos in {0, n // 2, n - 2}:
            v = strict.copy()
            if 0 <= pos < n - 1:
                v[pos + 1] = v[pos]                   # one plateau
                variants.append(v)
                v2 = strict.copy(); v2[pos + 1] = v2[pos] - 1   # one dip
                variants.append(v2)
        if np.issubdtype(dt, np.floating):
            for pos in {0, n // 2, n - 1}:
                v = strict.astype(np.float64); v[pos] = np.nan
                variants.append(v)
        for v in variants:
            v = np.ascontiguousarray(v).astype(dt)
            assert monotonic_prop(v) == names[O.monotonic_prop(v)], (np.dtype(dt).name, n)
            assert monotonic_prop(v[::-1]) == names[O.monotonic_prop(v, -1)]


# ---- K3 ----------------------------------------------------------------------------------------------
WIDTHS = [(), (1,), (2,), (3,), (4,), (5,), (7,), (8,), (12,), (16,), (31,), (32,), (33,), (64,), (100,), (128,),
          (130,), (256,), (3, 5), (2, 2, 2), (1000,), (1024,)]


@pytest.mark.parametrize("dt", DTS, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("trailing", WIDTHS, ids=lambda t: "w" + "x".join(map(str, t)))
def test_linear_matches_oracle(dt, trailing):
    rng = np.random.default_rng(len(trailing) * 1000 + int(np.prod(trailing, dtype=np.int64)))
    n = 200
    g = make_grid(rng, n, dt, "random")
    data = make_data(rng, (n,) + trailing, dt)
    interp = Interp1D.new_unchecked(g, data, Linear.new().extrapolate(True))
    strict = Interp1D.new_unchecked(g, data, Linear.new())
    for nq in (1, 31, 33, 3000):
        q = make_queries(rng, g, nq, dt, outside=True)
        st, ref, _ = O.interp1d_linear(g, data, q, True)
        assert st == O.ST_OK
        got = interp.interp_array(q)
        assert same(got, ref)
        if np.issubdtype(dt, np.floating):
            assert ulp_diff(got, ref) <= 4           # the bar north_star states
        qi = make_queries(rng, g, nq, dt, outside=False)
        st, ref, _ = O.interp1d_linear(g, data, qi, False)
        assert st == O.ST_OK
        assert same(strict.interp_array(qi), ref)


@pytest.mark.parametrize("mode", [L.SEARCH_BINARY_GLOBAL, L.SEARCH_BINARY_SMEM, L.SEARCH_UNIFORM_GUESS, L.SEARCH_BUCKET_LUT, L.SEARCH_MERGE])
@pytest.mark.parametrize("kind", ["uniform", "random", "exp"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_linear_every_search_mode_gives_the_same_bits(mode, kind, dt):
    rng = np.random.default_rng(5)
    for n in (2, 17, 4096, 13000, 65536):
        g = make_grid(rng, n, dt, kind)
        if len(np.unique(g)) != n:
            continue
        data = make_data(rng, (n, 4), dt)
        interp = Interp1D.new_unchecked(g, data, Linear.new().extrapolate(True))
        L.check(L.load().ndi_interp1d_set_search_mode(interp._handle(), mode))
        q = make_queries(rng, g, 70000, dt, outside=True)
        st, ref, _ = O.interp1d_linear(g, data, q, True)
        assert same(interp.interp_array(q), ref)


def test_linear_query_dim_path():
    """BASELINE config 3 shape in miniature: 2-D query array, extrapolation, 5 % outside"""
    rng = np.random.default_rng(3)
    g = np.cumsum(np.exp(rng.uniform(-3, 1, 4096))).astype(np.float32)
    data = rng.normal(size=(4096, 16)).astype(np.float32)
    interp = Interp1D.new_unchecked(g, data, Linear.new().extrapolate(True))
    q = make_queries(rng, g, 256 * 128, np.float32, outside=True).reshape(256, 128)
    out = interp.interp_array(q)
    assert out.shape == (256, 128, 16)
    st, ref, _ = O.interp1d_linear(g, data, q, True)
    assert same(out, ref)


def test_linear_errors_first_bad_and_untouched_rows():
    rng = np.random.default_rng(11)
    g = np.cumsum(rng.uniform(0.5, 1.5, 300))
    data = rng.normal(size=(300, 64))
    interp = Interp1D.new_unchecked(g, data, Linear.new())
    for nq, bad_at in [(100, [17, 60]), (40000, [39999]), (40000, [12345, 777, 30000]), (40000, [0])]:
        q = rng.uniform(g[0], g[-1], nq)
        q[bad_at] = g[-1] + 1.0
        first = min(bad_at)
        buf = np.full((nq, 64), -7.0)
        with pytest.raises(InterpolateError.OutOfBounds):
            interp.interp_array_into(q, buf)
        st, ref, bad = O.interp1d_linear(g, data, q, False, out=np.full((nq, 64), -7.0))
        assert (st, bad) == (O.ST_OUT_OF_BOUNDS, first)
        assert same(buf, ref)                    # rows < first written, rows >= first untouched
    # NaN: OutOfBounds without extrapolation, the NaN panic with it
    q = rng.uniform(g[0], g[-1], 50); q[20] = np.nan
    with pytest.raises(InterpolateError.OutOfBounds, match="x = NaN"):
        interp.interp_array(q)
    ex = Interp1D.new_unchecked(g, data, Linear.new().extrapolate(True))
    with pytest.raises(Panic, match="failed to convert NaN to usize"):
        ex.interp_array(q)
    q = rng.uniform(g[0], g[-1], 40000); q[31000] = np.nan      # pre-pass path
    with pytest.raises(Panic, match="failed to convert NaN to usize"):
        ex.interp_array(q)


def test_linear_multi_chunk_pipeline():
    """> 64 MB of output: the two-stream chunked path"""
    rng = np.random.default_rng(13)
    g = np.cumsum(rng.uniform(0.5, 1.5, 512))
    data = rng.normal(size=(512, 1024))
    interp = Interp1D.new_unchecked(g, data, Linear.new())
    q = np.sort(rng.uniform(g[0], g[-1], 20011))
    st, ref, _ = O.interp1d_linear(g, data, q, False, nthreads=8)
    assert st == O.ST_OK
    assert same(interp.interp_array(q), ref)


# ---- K4 ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTS, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("trailing", [(), (1,), (2,), (3,), (8,), (16,), (32,), (33,), (128,), (200,), (2, 3)],
                         ids=lambda t: "w" + "x".join(map(str, t)))
def test_bilinear_matches_oracle(dt, trailing):
    rng = np.random.default_rng(77 + int(np.prod(trailing, dtype=np.int64)))
    n, m = 40, 57
    gx, gy = make_grid(rng, n, dt, "uniform"), make_grid(rng, m, dt, "random")
    data = make_data(rng, (n, m) + trailing, dt)
    ex = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True))
    strict = Interp2D.new_unchecked(gx, gy, data, Bilinear.new())
    for nq in (1, 33, 5000):
        qx, qy = make_queries(rng, gx, nq, dt, True), make_queries(rng, gy, nq, dt, True)
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
        assert st == O.ST_OK
        got = ex.interp_array(qx, qy)
        assert same(got, ref)
        if np.issubdtype(dt, np.floating):
            assert ulp_diff(got, ref) <= 4
        qx, qy = make_queries(rng, gx, nq, dt, False), make_queries(rng, gy, nq, dt, False)
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, False)
        assert same(strict.interp_array(qx, qy), ref)


def test_bilinear_error_axis_precedence():
    rng = np.random.default_rng(2)
    gx, gy = np.linspace(0, 1, 30), np.cumsum(rng.uniform(0.5, 1.5, 20))
    data = rng.normal(size=(30, 20, 8))
    interp = Interp2D.new_unchecked(gx, gy, data, Bilinear.new())
    for nq in (200, 70000):
        qx, qy = rng.uniform(0, 1, nq), rng.uniform(gy[0], gy[-1], nq)
        qy[nq // 2] = gy[-1] + 1                  # y fails first ...
        qx[nq // 2 + 5] = 2.0
        with pytest.raises(InterpolateError.OutOfBounds, match="y = "):
            interp.interp_array(qx, qy)
        qx[nq // 2] = -1.0                        # ... unless x of the same query fails too
        with pytest.raises(InterpolateError.OutOfBounds, match="x = -1.0"):
            interp.interp_array(qx, qy)
        buf = np.full((nq, 8), 3.0)
        with pytest.raises(InterpolateError.OutOfBounds):
            interp.interp_array_into(qx, qy, buf)
        st, ref, bad, ax = O.interp2d_bilinear(gx, gy, data, qx, qy, False, out=np.full((nq, 8), 3.0))
        assert (st, bad, ax) == (O.ST_OUT_OF_BOUNDS, nq // 2, 0)
        assert same(buf, ref)


@pytest.mark.parametrize("mode", [L.SEARCH_BINARY_GLOBAL, L.SEARCH_BINARY_SMEM, L.SEARCH_UNIFORM_GUESS, L.SEARCH_BUCKET_LUT, L.SEARCH_MERGE])
def test_bilinear_every_search_mode(mode):
    rng = np.random.default_rng(9)
    gx, gy = np.linspace(0, 1, 2048).astype(np.float32), np.cumsum(rng.uniform(0.5, 1.5, 777)).astype(np.float32)
    data = rng.normal(size=(2048, 777, 8)).astype(np.float32)
    interp = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True))
    L.check(L.load().ndi_interp2d_set_search_mode(interp._handle(), mode))
    qx, qy = make_queries(rng, gx, 50000, np.float32, True), make_queries(rng, gy, 50000, np.float32, True)
    st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
    assert same(interp.interp_array(qx, qy), ref)


# ---- K8: query binning by table band (ndi_bin.cu) -----------------------------------------------------
def _force_binning(interp, band_rows):
    L.check(L.load().ndi_interp2d_set_binning(interp._handle(), L.BIN_ON, band_rows))


@pytest.mark.parametrize("dt", DTS, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("trailing", [(), (2,), (8,), (32,), (33,), (200,)], ids=lambda t: "w" + "x".join(map(str, t)))
@pytest.mark.parametrize("band_rows", [1, 4, 64])
def test_bilinear_binned_matches_oracle(dt, trailing, band_rows):
    """binning changes the evaluation order only: same bits as the oracle, hence as the direct kernel"""
    rng = np.random.default_rng(501 + band_rows + int(np.prod(trailing, dtype=np.int64)))
    n, m = 300, 41                                  # band_rows = 1 -> more than 256 bands wanted: the plan must coarsen
    gx, gy = make_grid(rng, n, dt, "random"), make_grid(rng, m, dt, "uniform")
    data = make_data(rng, (n, m) + trailing, dt)
    ex = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True))
    strict = Interp2D.new_unchecked(gx, gy, data, Bilinear.new())
    _force_binning(ex, band_rows)
    _force_binning(strict, band_rows)
    for nq in (2, 33, 2048, 2049, 30011):
        qx, qy = make_queries(rng, gx, nq, dt, True), make_queries(rng, gy, nq, dt, True)
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
        assert st == O.ST_OK
        assert same(ex.interp_array(qx, qy), ref)
        qx, qy = make_queries(rng, gx, nq, dt, False), make_queries(rng, gy, nq, dt, False)
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, False)
        assert same(strict.interp_array(qx, qy), ref)


@pytest.mark.parametrize("dt,n", [(np.float32, 8000), (np.float64, 1000), (np.int64, 3000), (np.float32, 20000)],
                         ids=["f32-n8000", "f64-n1000", "i64-n3000", "f32-n20000-coarse"])
@pytest.mark.parametrize("mode", [L.SEARCH_BINARY_SMEM, L.SEARCH_BINARY_GLOBAL, L.SEARCH_MERGE])
def test_bilinear_binned_with_every_staged_search(dt, n, mode):
    """binning passes x every search strategy: the scatter kernel holds 30 - 47 KB of static shared memory, so an
    x-grid staged next to it needs the shared-memory opt-in although the dynamic part alone is below 48 KB"""
    rng = np.random.default_rng(n + mode)
    m, w, nq = 9, 4, 50_000
    gx, gy = make_grid(rng, n, dt, "random"), make_grid(rng, m, dt, "random")
    data = make_data(rng, (n, m, w), dt)
    ip = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True))
    L.check(L.load().ndi_interp2d_set_search_mode(ip._handle(), mode))
    _force_binning(ip, 64)
    qx, qy = make_queries(rng, gx, nq, dt, True), make_queries(rng, gy, nq, dt, True)
    st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
    assert st == O.ST_OK
    assert same(ip.interp_array(qx, qy), ref)


def test_bilinear_binned_errors_like_the_reference():
    rng = np.random.default_rng(23)
    gx, gy = np.linspace(0, 1, 500), np.cumsum(rng.uniform(0.5, 1.5, 20))
    data = rng.normal(size=(500, 20, 8))
    interp = Interp2D.new_unchecked(gx, gy, data, Bilinear.new())
    _force_binning(interp, 16)
    for nq in (200, 70000):
        qx, qy = rng.uniform(0, 1, nq), rng.uniform(gy[0], gy[-1], nq)
        qy[nq // 2] = gy[-1] + 1
        qx[nq // 2 + 5] = 2.0
        qx[nq - 1] = np.nan
        buf = np.full((nq, 8), 3.0)
        with pytest.raises(InterpolateError.OutOfBounds, match="y = "):
            interp.interp_array_into(qx, qy, buf)
        st, ref, bad, ax = O.interp2d_bilinear(gx, gy, data, qx, qy, False, out=np.full((nq, 8), 3.0))
        assert (st, bad, ax) == (O.ST_OUT_OF_BOUNDS, nq // 2, 1)
        assert same(buf, ref)


def test_bilinear_binned_device_api_error_word_and_rows():
    """_dev entry point: one call, first failure = minimum over the (unordered) binned lanes"""
    import torch
    from ndarray_interp_b200 import device as D
    rng = np.random.default_rng(29)
    n, m, w, nq = 700, 64, 8, 100_000
    gx, gy = np.cumsum(rng.uniform(0.5, 1.5, n)).astype(np.float32), np.linspace(-1, 1, m).astype(np.float32)
    data = rng.normal(size=(n, m, w)).astype(np.float32)
    qx = rng.uniform(gx[0], gx[-1], nq).astype(np.float32).clip(gx[0], gx[-1])
    qy = rng.uniform(-1, 1, nq).astype(np.float32).clip(-1, 1)
    bad = [91234, 5000, 5001, 77]
    qx[bad[0]] = gx[-1] + 1; qy[bad[1]] = 2.0; qx[bad[2]] = np.nan; qy[bad[3]] = -3.0
    st, ref, first, axis = O.interp2d_bilinear(gx, gy, data, qx, qy, False, out=np.zeros((nq, w), np.float32))
    assert (first, axis) == (77, 1)
    ip = D.DeviceInterp2D(torch.from_numpy(gx).cuda(), torch.from_numpy(gy).cuda(), torch.from_numpy(data).cuda())
    outs = []
    for mode, rows in [(L.BIN_OFF, 0), (L.BIN_ON, 32), (L.BIN_ON, 1), (L.BIN_SWEEP, 100), (L.BIN_SWEEP, 1)]:
        ip.set_binning(mode, rows)
        err = D.new_err_word()
        out = torch.zeros((nq, w), dtype=torch.float32, device="cuda")
        ip.bilinear(torch.from_numpy(qx).cuda(), torch.from_numpy(qy).cuda(), False, out=out, err=err)
        assert D.err_word_value(err) == 2 * 77 + 1
        outs.append(out.cpu().numpy())
    good = np.ones(nq, bool); good[bad] = False
    st, full, _, _ = O.interp2d_bilinear(gx, gy, data, qx[good], qy[good], False)
    for o in outs:
        assert same(o[good], full)                 # every passing row is written, failing rows are skipped
        assert not o[bad].any()


# ---- K9: band sweeps (ndi_sweep.cu) ---------------------------------------------------------------------
def _force_sweeps(interp, band_rows):
    L.check(L.load().ndi_interp2d_set_binning(interp._handle(), L.BIN_SWEEP, band_rows))


@pytest.mark.parametrize("dt", DTS, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("row_bytes", [16, 32, 64, 128, 48])
@pytest.mark.parametrize("band_rows,xkind", [(1, "random"), (7, "uniform"), (40, "random"), (1000, "uniform")])
def test_bilinear_sweeps_match_oracle(dt, row_bytes, band_rows, xkind):
    """sweeps change the evaluation order only (and, for 32-byte rows, which lane holds which operand): same bits
    as the oracle.  band_rows = 1 asks for more than 16 sweeps (the plan coarsens), 1000 gives a single sweep;
    48-byte rows are not a sweep shape and must fall through to the direct kernel."""
    rng = np.random.default_rng(901 + band_rows + row_bytes)
    n, m, w = 300, 41, row_bytes // np.dtype(dt).itemsize
    gx, gy = make_grid(rng, n, dt, xkind), make_grid(rng, m, dt, "random")
    data = make_data(rng, (n, m, w), dt)
    ex = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True))
    strict = Interp2D.new_unchecked(gx, gy, data, Bilinear.new())
    _force_sweeps(ex, band_rows)
    _force_sweeps(strict, band_rows)
    for nq in (1, 2, 33, 255, 256, 257, 2049, 30011):
        qx, qy = make_queries(rng, gx, nq, dt, True), make_queries(rng, gy, nq, dt, True)
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
        assert st == O.ST_OK
        assert same(ex.interp_array(qx, qy), ref)
        qx, qy = make_queries(rng, gx, nq, dt, False), make_queries(rng, gy, nq, dt, False)
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, False)
        assert same(strict.interp_array(qx, qy), ref)


@pytest.mark.parametrize("mode", [L.SEARCH_AUTO, L.SEARCH_BINARY_SMEM, L.SEARCH_BINARY_GLOBAL, L.SEARCH_UNIFORM_GUESS, L.SEARCH_BUCKET_LUT, L.SEARCH_MERGE])
def test_bilinear_sweeps_with_every_search_mode(mode):
    rng = np.random.default_rng(77 + mode)
    gx, gy = np.linspace(0, 1, 2048).astype(np.float32), np.cumsum(rng.uniform(0.5, 1.5, 777)).astype(np.float32)
    data = rng.normal(size=(2048, 777, 8)).astype(np.float32)
    interp = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True))
    L.check(L.load().ndi_interp2d_set_search_mode(interp._handle(), mode))
    _force_sweeps(interp, 512)
    qx, qy = make_queries(rng, gx, 50000, np.float32, True), make_queries(rng, gy, 50000, np.float32, True)
    st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
    assert same(interp.interp_array(qx, qy), ref)


def test_bilinear_sweeps_extreme_values_take_the_slow_division():
    """tables and queries outside the ranges of the hoisted-reciprocal division (tiny / huge values, far
    extrapolation, zero differences) through the pair form of the 32-byte rows"""
    rng = np.random.default_rng(5)
    n, m, w = 64, 33, 8
    gx, gy = np.cumsum(rng.uniform(0.5, 1.5, n)).astype(np.float32), np.linspace(0, 1, m).astype(np.float32)
    data = (rng.normal(size=(n, m, w)) * np.exp(rng.uniform(-80, 60, (n, m, w)))).astype(np.float32)
    data[rng.random((n, m, w)) < 0.2] = 0.0
    data[10:20] = data[9]                                   # zero first-stage numerators
    nq = 20000
    qx = rng.uniform(gx[0] - 1e6, gx[-1] + 1e6, nq).astype(np.float32)
    qy = rng.uniform(-3, 4, nq).astype(np.float32)
    qx[::3] = rng.uniform(gx[0], gx[-1], len(qx[::3])).astype(np.float32)
    ip = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True))
    with np.errstate(all="ignore"):
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
    direct = ip.interp_array(qx, qy)
    _force_sweeps(ip, 16)
    assert same(direct, ref)
    assert same(ip.interp_array(qx, qy), ref)


def test_bilinear_sweeps_errors_like_the_reference():
    rng = np.random.default_rng(24)
    gx, gy = np.linspace(0, 1, 500), np.cumsum(rng.uniform(0.5, 1.5, 20))
    data = rng.normal(size=(500, 20, 4))
    interp = Interp2D.new_unchecked(gx, gy, data, Bilinear.new())
    _force_sweeps(interp, 100)
    for nq in (200, 70000):
        qx, qy = rng.uniform(0, 1, nq), rng.uniform(gy[0], gy[-1], nq)
        qy[nq // 2] = gy[-1] + 1
        qx[nq // 2 + 5] = 2.0
        qx[nq - 1] = np.nan
        buf = np.full((nq, 4), 3.0)
        with pytest.raises(InterpolateError.OutOfBounds, match="y = "):
            interp.interp_array_into(qx, qy, buf)
        st, ref, bad, ax = O.interp2d_bilinear(gx, gy, data, qx, qy, False, out=np.full((nq, 4), 3.0))
        assert (st, bad, ax) == (O.ST_OUT_OF_BOUNDS, nq // 2, 1)
        assert same(buf, ref)


# ---- hoisted-reciprocal division (ndi_device.cuh: rcp_refined / div_by) ------------------------------
def test_fdiv_selftest_sample():
    """div_by(a, b, rcp_refined(b)) == __fdiv_rn(a, b): sampled numerator mantissas x ALL 2^23 divisor
    mantissas, at the ends and in the middle of the exponent range the kernels admit
    (b in [2^-40, 2^40], |a| in [2^-80, 2^80]).  The all-pairs run is scripts/exhaustive_fdiv.py."""
    lib = L.require_device()
    rng = np.random.default_rng(5)
    starts = [0, (1 << 23) - 64, 1 << 22] + [int(v) for v in rng.integers(0, (1 << 23) - 64, 13)]
    exps = [(0, 0), (-80, 40), (80, -40), (-80, -40), (80, 40), (3, -7)]
    for i, s in enumerate(starts):
        ea, eb = exps[i % len(exps)]
        bad = C.c_uint64(123)
        L.check(lib.ndi_selftest_fdiv(s, 64, ea, eb, C.byref(bad)))
        assert bad.value == 0, (s, ea, eb, bad.value)


@pytest.mark.parametrize("scale", [1e-30, 1e-17, 1.0, 1e9, 3e35], ids=lambda s: "x%g" % s)
def test_f32_tables_outside_the_fast_division_range(scale):
    """tables with tiny, huge, zero, denormal and infinite values: same bits as the oracle (IEEE division)"""
    rng = np.random.default_rng(41)
    n, m, w = 50, 37, 8
    g = np.cumsum(rng.uniform(0.5, 1.5, n)).astype(np.float32)
    gy = np.linspace(-2, 5, m).astype(np.float32)
    d1 = (rng.normal(size=(n, w)) * scale).astype(np.float32)
    d1[rng.integers(0, n, 40), rng.integers(0, w, 40)] = 0.0
    d1[5] = d1[6]                                     # equal neighbours: zero numerators
    d1[7, 0] = -0.0
    d1[9, 1] = np.float32(1e-42)                      # denormal
    d2 = (rng.normal(size=(n, m, w)) * scale).astype(np.float32)
    d2[rng.integers(0, n, 200), rng.integers(0, m, 200)] = 0.0
    d2[3] = d2[4]
    d2[11, 5, 2] = np.float32(-1e-41)
    with np.errstate(all="ignore"):
        for extra in (None, np.inf):
            if extra is not None:
                d1[20, 3] = extra; d2[20, 20, 3] = -extra
            q = make_queries(rng, g, 4000, np.float32, True)
            st, ref, _ = O.interp1d_linear(g, d1, q, True)
            assert same(Interp1D.new_unchecked(g, d1, Linear.new().extrapolate(True)).interp_array(q), ref)
            qx, qy = make_queries(rng, g, 4000, np.float32, True), make_queries(rng, gy, 4000, np.float32, True)
            st, ref, _, _ = O.interp2d_bilinear(g, gy, d2, qx, qy, True)
            assert same(Interp2D.new_unchecked(g, gy, d2, Bilinear.new().extrapolate(True)).interp_array(qx, qy), ref)


def test_f32_fast_division_edge_queries():
    """in-range tables (fast path armed) with the per-query and per-element exits: extreme
    extrapolation, grid steps outside [2^-40, 2^40], and second-stage numerators below 2^-80"""
    rng = np.random.default_rng(43)
    w = 8
    gx = np.array([0.0, 1e-13, 1.0, 2.0, 3.0, 1e13, 2e13], dtype=np.float32)      # steps 1e-13 and 1e13
    gy = np.array([0.0, 1.0, 2.0, 4.0], dtype=np.float32)
    data = rng.normal(size=(len(gx), len(gy), w)).astype(np.float32)
    data[0] = 0.0; data[1, :, ::2] = 0.0                                          # z1, z2 tiny next to the knot at 0
    qx = np.concatenate([np.float32(1e-30) * np.arange(1, 200, dtype=np.float32), rng.uniform(0, 3, 500).astype(np.float32),
                         np.float32([1e-14, 5e-14, 1e12, 1.5e13, -1e25, 1e30, -3e38, 3e38, 1e-45, 0.0])])
    qy = np.concatenate([rng.uniform(0, 4, len(qx) - 6), [1e-38, -1e30, 1e30, 3.9999, 0.0, 4.0]]).astype(np.float32)
    with np.errstate(all="ignore"):
        st, ref, _, _ = O.interp2d_bilinear(gx, gy, data, qx, qy, True)
        got = Interp2D.new_unchecked(gx, gy, data, Bilinear.new().extrapolate(True)).interp_array(qx, qy)
        assert same(got, ref)
        d1 = data[:, 1, :].copy()
        st, ref, _ = O.interp1d_linear(gx, d1, qx, True)
        assert same(Interp1D.new_unchecked(gx, d1, Linear.new().extrapolate(True)).interp_array(qx), ref)


def test_ddiv_selftest_sample():
    """Hoisted<double>::div == __ddiv_rn on 2^28 pseudo-random operand pairs (all exponents and signs)"""
    lib = L.require_device()
    for seed in (1, 2026):
        bad = C.c_uint64(123)
        L.check(lib.ndi_selftest_ddiv(seed, 1 << 27, C.byref(bad)))
        assert bad.value == 0, (seed, bad.value)


# ---- i64 (Linear / Bilinear are generic over Num) -----------------------
@pytest.mark.parametrize("mode", [L.SEARCH_AUTO, L.SEARCH_BINARY_GLOBAL, L.SEARCH_BINARY_SMEM, L.SEARCH_UNIFORM_GUESS,
                                  L.SEARCH_BUCKET_LUT, L.SEARCH_MERGE])
def test_i64_values_beyond_2_pow_53_and_wrapping(mode):
    """grid / data values that no double represents exactly (the bucket function rounds them, the search must
    not) and products that wrap like a Rust release build"""
    rng = np.random.default_rng(64)
    for n, w in ((2, 1), (50, 3), (5000, 8), (70001, 2)):
        g = (1 << 60) + np.cumsum(rng.integers(1, 9, n)).astype(np.int64)
        data = rng.integers(-(1 << 62), 1 << 62, (n, w)).astype(np.int64)
        q = rng.integers(int(g[0]) - 50, int(g[-1]) + 50, 20000).astype(np.int64)
        q[:4] = [g[0], g[-1], g[0] - 1, g[-1] + 1]
        st, ref_idx, _ = O.lower_index(g, q)
        assert st == O.ST_OK and np.array_equal(get_lower_index(g, q), ref_idx)
        interp = Interp1D.new_unchecked(g, data, Linear.new().extrapolate(True))
        L.check(L.load().ndi_interp1d_set_search_mode(interp._handle(), mode))
        st, ref, _ = O.interp1d_linear(g, data, q, True)
        assert st == O.ST_OK
        got = interp.interp_array(q)
        assert got.dtype == np.int64 and same(got, ref)
    gx = np.cumsum(rng.integers(1, 1000, 300)).astype(np.int64) - (1 << 40)
    gy = np.arange(40, dtype=np.int64) * 7 + (1 << 55)
    d2 = rng.integers(-(1 << 40), 1 << 40, (300, 40, 5)).astype(np.int64)
    qx = rng.integers(int(gx[0]), int(gx[-1]) + 1, 30000).astype(np.int64)
    qy = rng.integers(int(gy[0]), int(gy[-1]) + 1, 30000).astype(np.int64)
    ip = Interp2D.new_unchecked(gx, gy, d2, Bilinear.new())
    L.check(L.load().ndi_interp2d_set_search_mode(ip._handle(), mode))
    st, ref, _, _ = O.interp2d_bilinear(gx, gy, d2, qx, qy, False)
    assert st == O.ST_OK and same(ip.interp_array(qx, qy), ref)


# ---- u32 / u64 (the reference's generic bound admits them: linear.rs:29-36, T: Num) --------------------------------
@pytest.mark.parametrize("dt", [np.uint32, np.uint64], ids=["u32", "u64"])
@pytest.mark.parametrize("mode", [L.SEARCH_AUTO, L.SEARCH_BINARY_GLOBAL, L.SEARCH_BINARY_SMEM, L.SEARCH_UNIFORM_GUESS,
                                  L.SEARCH_BUCKET_LUT, L.SEARCH_MERGE])
def test_unsigned_values_in_the_upper_half_and_wrapping(dt, mode):
    """grid and data values with the top bit set (a signed comparison or division would get them wrong), falling data
    and queries left of the grid (differences that wrap, as in a release build of the reference), products that wrap"""
    bits = np.dtype(dt).itemsize * 8
    rng = np.random.default_rng(bits)
    top = 1 << (bits - 1)
    for n, w in ((2, 1), (50, 3), (5000, 8), (70001, 2)):
        g = (top - 40000 + np.cumsum(rng.integers(1, 9, n))).astype(dt)        # crosses 2^(bits-1)
        data = (rng.integers(0, 1 << 32, (n, w)).astype(np.uint64) << np.uint64(bits - 32)).astype(dt)
        q = (np.uint64(int(g[0]) - 50) + rng.integers(0, int(g[-1]) - int(g[0]) + 100, 20000).astype(np.uint64)).astype(dt)
        q[:4] = [g[0], g[-1], g[0] - dt(1), g[-1] + dt(1)]
        st, ref_idx, _ = O.lower_index(g, q)
        assert st == O.ST_OK and np.array_equal(get_lower_index(g, q), ref_idx)
        interp = Interp1D.new_unchecked(g, data, Linear.new().extrapolate(True))
        L.check(L.load().ndi_interp1d_set_search_mode(interp._handle(), mode))
        st, ref, _ = O.interp1d_linear(g, data, q, True)
        assert st == O.ST_OK
        got = interp.interp_array(q)
        assert got.dtype == np.dtype(dt) and same(got, ref)
    gx = (np.uint64(top) + np.cumsum(rng.integers(1, 1000, 300)).astype(np.uint64)).astype(dt)
    gy = (np.arange(40) * 7 + 3).astype(dt)
    d2 = rng.integers(0, 1 << 30, (300, 40, 5)).astype(dt)
    qx = (np.uint64(int(gx[0])) + rng.integers(0, int(gx[-1]) - int(gx[0]) + 1, 30000).astype(np.uint64)).astype(dt)
    qy = rng.integers(int(gy[0]), int(gy[-1]) + 1, 30000).astype(dt)
    ip = Interp2D.new_unchecked(gx, gy, d2, Bilinear.new())
    L.check(L.load().ndi_interp2d_set_search_mode(ip._handle(), mode))
    st, ref, _, _ = O.interp2d_bilinear(gx, gy, d2, qx, qy, False)
    assert st == O.ST_OK and same(ip.interp_array(qx, qy), ref)


@pytest.mark.parametrize("dt", [np.uint32, np.uint64], ids=["u32", "u64"])
def test_unsigned_builder_checks_and_errors(dt):
    """the same builder / OutOfBounds behaviour as i32 (tests/interp1d.rs:93-127); no splines on integers"""
    u = dt
    with pytest.raises(BuilderError.Monotonic):
        Interp1D.builder(np.array([1, 2, 3], u)).x(np.array([1, 2, 2], u)).build()
    interp = Interp1D.builder(np.array([10, 20, 40], u)).build()
    assert interp.interp_scalar(u(1)) == 20 and interp.interp_scalar(u(2)).dtype == np.dtype(dt)
    with pytest.raises(InterpolateError.OutOfBounds):
        interp.interp_scalar(u(3))
    big = 1 << (np.dtype(dt).itemsize * 8 - 1)
    assert monotonic_prop(np.array([big - 1, big, big], u)) == Monotonic.Rising(False)
    assert monotonic_prop(np.array([big + 5, big, 3], u)) == Monotonic.Falling(True)
    rising = Interp1D.builder(np.array([5, 9, 16], u)).x(np.array([big - 1, big, big + 7], u)).build()
    assert rising.interp_scalar(u(big + 7)) == 16 and rising.interp_scalar(u(big - 1)) == 5 and rising.interp_scalar(u(big + 3)) == 12


def test_i64_builder_checks_and_errors():
    """the same builder / OutOfBounds behaviour as i32 (tests/interp1d.rs:93-127)"""
    i = np.int64
    with pytest.raises(BuilderError.Monotonic):
        Interp1D.builder(np.array([1, 2, 3], i)).x(np.array([1, 2, 2], i)).build()
    interp = Interp1D.builder(np.array([10, 20, 40], i)).build()
    assert interp.interp_scalar(i(1)) == 20 and interp.interp_scalar(i(2)).dtype == np.int64
    with pytest.raises(InterpolateError.OutOfBounds):
        interp.interp_scalar(i(3))
    assert monotonic_prop(np.array([1 << 62, (1 << 62) + 1, (1 << 62) + 1], i)) == Monotonic.Rising(False)


# ---- strided views (tests/interp1d.rs:143-155 builds from negative strides) ---
def _views_1d(rng, dt):
    base = make_data(rng, (64, 12, 10), dt)
    g = make_grid(rng, 64, dt, "random")
    gpad = np.zeros(64 * 3, dt); gpad[::3] = g
    yield "negative axis 0 + reversed grid", g[::-1][::-1], base[::-1]
    yield "grid with stride 3", gpad[::3], base
    yield "trailing step", g, base[:, ::2, 1:9:3]
    yield "transposed trailing axes", g, base.transpose(0, 2, 1)
    yield "negative trailing", g, base[:, ::-1, ::-1]
    yield "sparse (host gather path)", g[::7][:9], base[::7, ::5, ::6][:9]
    yield "broadcast (zero stride)", g, np.broadcast_to(base[:, :1, :], base.shape)
    yield "1-d data view", g[::2], base[::2, 3, 4]
    yield "fortran order", g, np.asfortranarray(base)


@pytest.mark.parametrize("dt", DTS, ids=lambda d: np.dtype(d).name)
def test_strided_views_build_without_a_host_copy_1d(dt):
    rng = np.random.default_rng(77)
    for name, g, d in _views_1d(rng, dt):
        assert not (g.flags.c_contiguous and d.flags.c_contiguous), name
        q = make_queries(rng, np.ascontiguousarray(g), 3000, dt, outside=True)
        st, ref, _ = O.interp1d_linear(np.ascontiguousarray(g), np.ascontiguousarray(d), q, True)
        assert st == O.ST_OK
        ip = Interp1D.builder(d).x(g).strategy(Linear.new().extrapolate(True)).build()
        assert ip.data is d                                # the view itself is kept: no host-side copy
        assert same(ip.interp_array(q), ref), name
    # the monotonic check sees the view, not its memory order
    with pytest.raises(BuilderError.Monotonic):
        Interp1D.builder(np.zeros(10, dt)).x(np.arange(10).astype(dt)[::-1]).build()


def test_strided_views_2d_and_splines():
    rng = np.random.default_rng(78)
    base = rng.normal(size=(40, 30, 6))
    gx, gy = np.cumsum(rng.uniform(0.5, 1.5, 40)), np.cumsum(rng.uniform(0.5, 1.5, 30))
    for d, x, y in [(base[::-1, :, ::2], gx, gy), (base.transpose(1, 0, 2), gy, gx),
                    (base[::2, ::3], gx[::2], gy[::3]), (np.asfortranarray(base), gx[::-1][::-1], gy)]:
        qx = rng.uniform(x[0], x[-1], 5000); qy = rng.uniform(y[0], y[-1], 5000)
        st, ref, _, _ = O.interp2d_bilinear(np.ascontiguousarray(x), np.ascontiguousarray(y), np.ascontiguousarray(d), qx, qy, False)
        assert st == O.ST_OK
        assert same(Interp2D.builder(d).x(x).y(y).build().interp_array(qx, qy), ref)
    from ndarray_interp_b200.interp1d import CubicSpline
    d = base[:, ::-2, 1]
    q = rng.uniform(gx[0], gx[-1], 4000)
    got = Interp1D.builder(d).x(gx).strategy(CubicSpline.new()).build().interp_array(q)
    want = Interp1D.builder(np.ascontiguousarray(d)).x(gx).strategy(CubicSpline.new()).build().interp_array(q)
    assert same(got, want)


def test_strided_create_c_abi_rejects_bad_arguments_and_takes_device_views():
    import torch
    lib = L.require_device()
    h = C.c_void_p()
    x = np.arange(5.0); d = np.arange(10.0).reshape(5, 2)
    shape = (C.c_int64 * 2)(5, 2); strides = (C.c_int64 * 2)(2, 1)
    call = lambda n, ndim, flags: lib.ndi_interp1d_create_strided(L.F64, L.ptr(x), n, 1, L.ptr(d), ndim, shape, strides, flags, C.byref(h))
    assert call(4, 2, 0) == L.INVALID_ARGUMENT            # x length != shape[0]
    assert call(5, 9, 0) == L.INVALID_ARGUMENT            # too many dimensions
    assert call(5, 2, 4) == L.INVALID_ARGUMENT            # NDI_BORROW
    # device views: a transposed torch tensor is packed on the device
    t = torch.arange(24, dtype=torch.float64, device="cuda").reshape(4, 6).t()          # shape (6, 4), strides (1, 6)
    xs = torch.arange(6, dtype=torch.float64, device="cuda")
    shape = (C.c_int64 * 2)(6, 4); strides = (C.c_int64 * 2)(*t.stride())
    torch.cuda.synchronize()
    st = lib.ndi_interp1d_create_strided(L.F64, C.c_void_p(xs.data_ptr()), 6, 1, C.c_void_p(t.data_ptr()), 2, shape, strides,
                                         2, C.byref(h))
    assert st == L.OK
    q = np.array([0.5, 4.25]); out = np.zeros((2, 4)); bad = C.c_int64(-1)
    assert lib.ndi_interp1d_linear(h, L.ptr(q), 2, 0, L.ptr(out), C.byref(bad)) == L.OK
    tn = t.cpu().numpy()
    assert np.array_equal(out[0], tn[0] + 0.5 * (tn[1] - tn[0])) and np.array_equal(out[1], tn[4] + 0.25 * (tn[5] - tn[4]))
    lib.ndi_interp1d_destroy(h)


# ---- empty and 0-d query arrays (ndarray accepts both; the batch loops of interp1d/mod.rs:300-343 then run 0 / 1 times)
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_empty_and_zero_dimensional_queries(dt):
    from ndarray_interp_b200.interp1d import CubicSpline
    rng = np.random.default_rng(11)
    g = make_grid(rng, 30, dt, "random")
    d1 = make_data(rng, (30, 4, 3), dt)
    strategies = [Linear.new(), Linear.new().extrapolate(True)]
    if np.issubdtype(dt, np.floating):
        strategies.append(CubicSpline.new())
    for strat in strategies:
        ip = Interp1D.builder(d1).x(g).strategy(strat).build()
        for shape in [(0,), (0, 5), (3, 0, 2)]:
            out = ip.interp_array(np.zeros(shape, dt))
            assert out.shape == shape + (4, 3) and out.dtype == dt
        x0 = g[7]                                                  # a knot: the result is the data row itself
        out = ip.interp_array(np.array(x0))                        # Ix0 query -> data.shape[1:]
        assert out.shape == (4, 3) and same(out, d1[7])
        buf = np.full((0, 4, 3), 9, dt)
        ip.interp_array_into(np.zeros(0, dt), buf)                 # nothing to write, nothing to fail
    gy = make_grid(rng, 12, dt, "uniform")
    d2 = make_data(rng, (30, 12, 2), dt)
    ip2 = Interp2D.builder(d2).x(g).y(gy).build()
    assert ip2.interp_array(np.zeros((0, 4), dt), np.zeros((0, 4), dt)).shape == (0, 4, 2)
    assert same(ip2.interp_array(np.array(g[3]), np.array(gy[5])), d2[3, 5])
    # the C ABI itself: nq = 0 is a no-op with any pointers, nq < 0 is an argument error
    lib = L.require_device()
    bad = C.c_int64(-1)
    assert lib.ndi_interp1d_linear(ip._handle(), None, 0, 0, None, C.byref(bad)) == L.OK and bad.value == -1
    assert lib.ndi_interp1d_linear(ip._handle(), None, -1, 0, None, C.byref(bad)) == L.INVALID_ARGUMENT
