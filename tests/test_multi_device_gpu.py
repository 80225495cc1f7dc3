"""One interpolator on several GPUs of one process (include/ndi_b200.h: ndi_interp{1,2}d_replicate, the
`ndi_replicate` of SURVEY.md section 8(b)) and ndi_interp{1,2}d_clone_to_device: the fanned-out batch must equal the
single-device result bit for bit, report the batch-wide first failing query (also when it lies in a later block)
and leave rows at and after it untouched in EVERY block.  Needs two devices (skipped below)."""
import ctypes as C

import numpy as np
import pytest

from ndarray_interp_b200 import InterpolateError, Panic, _lib as L
from ndarray_interp_b200.interp1d import BoundaryCondition, CubicSpline, Interp1D, Interp1DBuilder, Linear
from ndarray_interp_b200.interp2d import Bilinear, Interp2D
from oracle import oracle_py as O
from test_parity_gpu import same

pytestmark = pytest.mark.gpu


def ndev():
    n = C.c_int32(0)
    L.load().ndi_device_count(C.byref(n))
    return n.value


needs2 = pytest.mark.skipif(ndev() < 2, reason="needs two GPUs")


def devices():
    return list(range(min(ndev(), 4)))


@needs2
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32], ids=["f32", "f64", "i32"])
def test_group_linear_equals_one_device_and_the_oracle(dt):
    rng = np.random.default_rng(3)
    n, w, nq = 3000, 16, 300_001
    g = np.cumsum(rng.integers(1, 9, n)).astype(dt) if np.issubdtype(dt, np.integer) else np.cumsum(rng.uniform(0.5, 1.5, n)).astype(dt)
    y = (rng.integers(-1000, 1000, (n, w)) if np.issubdtype(dt, np.integer) else rng.normal(size=(n, w))).astype(dt)
    q = rng.uniform(float(g[0]) - 5, float(g[-1]) + 5, nq).astype(dt)
    one = Interp1D.new_unchecked(g, y, Linear.new().extrapolate(True))
    many = one.replicate(devices())
    ref = O.interp1d_linear(g, y, q, True)[1]
    assert same(one.interp_array(q), ref) and same(many.interp_array(q), ref)
    assert same(many.interp_array(q[:100]), ref[:100])                 # small batch: first device only


@needs2
def test_group_first_failure_in_a_later_block_and_untouched_rows():
    rng = np.random.default_rng(4)
    n, w, nq = 500, 8, 200_000
    g = np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.normal(size=(n, w))
    strict = Interp1D.new_unchecked(g, y, Linear.new()).replicate(devices())
    nd = len(devices())
    for where in (nq // nd + 77, nq - 3, 5):                            # second block, last block, first block
        q = rng.uniform(g[0], g[-1], nq)
        q[where] = g[-1] + 1.0
        q[min(where + 1000, nq - 1)] = np.nan                           # a later failure must not be the one reported
        buf = np.full((nq, w), 7.0)
        with pytest.raises(InterpolateError.OutOfBounds):
            strict.interp_array_into(q, buf)
        st, ref, bad = O.interp1d_linear(g, y, q, False, out=np.full((nq, w), 7.0))
        assert (st, bad) == (O.ST_OUT_OF_BOUNDS, where)
        assert same(buf, ref)                                           # rows >= `where` untouched in every block


@needs2
def test_group_cubic_and_bilinear():
    rng = np.random.default_rng(5)
    n, w, nq = 2500, 32, 150_000
    g = np.cumsum(rng.uniform(0.5, 1.5, n)).astype(np.float32)
    y = rng.normal(size=(n, w)).astype(np.float32)
    strat = CubicSpline.new().boundary(BoundaryCondition.Natural).extrapolate(True)
    one = Interp1DBuilder.new(y).x(g).strategy(strat).build()
    many = one.replicate(devices())
    q = np.sort(rng.uniform(g[0] - 3, g[-1] + 3, nq)).astype(np.float32)
    assert same(many.interp_array(q), one.interp_array(q))
    qn = q.copy(); qn[nq - 10] = np.nan
    with pytest.raises(Panic, match="failed to convert NaN to usize"):
        many.interp_array(qn)

    gx, gy = np.linspace(0, 1, 300).astype(np.float32), np.cumsum(rng.uniform(0.5, 1.5, 200)).astype(np.float32)
    z = rng.normal(size=(300, 200, 8)).astype(np.float32)
    b1 = Interp2D.new_unchecked(gx, gy, z, Bilinear.new())
    bm = b1.replicate(devices())
    qx = rng.uniform(0, 1, nq).astype(np.float32).clip(0, 1)
    qy = rng.uniform(gy[0], gy[-1], nq).astype(np.float32).clip(gy[0], gy[-1])
    ref = O.interp2d_bilinear(gx, gy, z, qx, qy, False)[1]
    assert same(bm.interp_array(qx, qy), ref)
    qy[nq // 2 + 11] = gy[-1] + 1
    qx[nq // 2 + 12] = 2.0
    buf = np.full((nq, 8), 2.0, np.float32)
    with pytest.raises(InterpolateError.OutOfBounds, match="y = "):
        bm.interp_array_into(qx, qy, buf)
    st, ref, bad, ax = O.interp2d_bilinear(gx, gy, z, qx, qy, False, out=np.full((nq, 8), 2.0, np.float32))
    assert (bad, ax) == (nq // 2 + 11, 1) and same(buf, ref)


@needs2
def test_clone_to_device_gives_the_same_bits():
    import torch
    from ndarray_interp_b200 import device as D
    rng = np.random.default_rng(6)
    n, w, nq = 4096, 16, 100_000
    g = np.cumsum(np.exp(rng.uniform(-2, 2, n))).astype(np.float32)
    y = rng.normal(size=(n, w)).astype(np.float32)
    q = rng.uniform(g[0], g[-1], nq).astype(np.float32).clip(g[0], g[-1])
    ip0 = D.DeviceInterp1D(torch.from_numpy(g).cuda(0), torch.from_numpy(y).cuda(0))
    st, _ = ip0.spline_build(1)
    out0 = ip0.cubic(torch.from_numpy(q).cuda(0), 0)
    lin0 = ip0.linear(torch.from_numpy(q).cuda(0), False)
    clone = C.c_void_p()
    L.check(L.load().ndi_interp1d_clone_to_device(ip0.h, 1, C.byref(clone)))
    try:
        with torch.cuda.device(1):
            L.check(L.load().ndi_set_device(1))
            qd = torch.from_numpy(q).cuda(1)
            out1 = torch.empty((nq, w), dtype=torch.float32, device="cuda:1")
            err = torch.full((1,), -1, dtype=torch.int64, device="cuda:1")
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            L.check(L.load().ndi_interp1d_cubic_dev(clone, C.c_void_p(qd.data_ptr()), nq, 0, C.c_void_p(out1.data_ptr()),
                                                    C.c_void_p(err.data_ptr()), s))
            assert torch.equal(out1.cpu(), out0.cpu())
            L.check(L.load().ndi_interp1d_linear_dev(clone, C.c_void_p(qd.data_ptr()), nq, 0, C.c_void_p(out1.data_ptr()),
                                                     C.c_void_p(err.data_ptr()), s))
            assert torch.equal(out1.cpu(), lin0.cpu())
            assert int(err.item()) == -1
    finally:
        L.check(L.load().ndi_set_device(0))
        L.load().ndi_interp1d_destroy(clone)


def test_group_of_one_device_is_the_plain_call():
    rng = np.random.default_rng(8)
    g = np.cumsum(rng.uniform(0.5, 1.5, 100))
    y = rng.normal(size=(100, 3))
    q = rng.uniform(g[0], g[-1], 50_000)
    one = Interp1D.new_unchecked(g, y, Linear.new())
    assert same(one.replicate([0]).interp_array(q), one.interp_array(q))
    with pytest.raises(L.NdiLibraryError):
        one.replicate([0, 0])


def test_strided_buffer_keeps_rows_after_the_first_failure():
    """the advisor's case: a non-contiguous output buffer must come back untouched at and after the failing query"""
    rng = np.random.default_rng(9)
    g = np.cumsum(rng.uniform(0.5, 1.5, 50))
    y = rng.normal(size=(50, 4))
    q = rng.uniform(g[0], g[-1], 1000)
    q[400] = g[0] - 1
    backing = np.full((1000, 8), 5.0)
    view = backing[:, ::2]
    with pytest.raises(InterpolateError.OutOfBounds):
        Interp1D.new_unchecked(g, y, Linear.new()).interp_array_into(q, view)
    st, ref, bad = O.interp1d_linear(g, y, q, False, out=np.full((1000, 4), 5.0))
    assert bad == 400 and same(np.ascontiguousarray(view), ref)
    gy = np.arange(6.0)
    z = rng.normal(size=(50, 6, 4))
    qy = rng.uniform(0, 5, 1000)
    backing = np.full((1000, 8), 5.0)
    view = backing[:, ::2]
    with pytest.raises(InterpolateError.OutOfBounds):
        Interp2D.new_unchecked(g, gy, z, Bilinear.new()).interp_array_into(q, qy, view)
    st, ref, bad, ax = O.interp2d_bilinear(g, gy, z, q, qy, False, out=np.full((1000, 4), 5.0))
    assert bad == 400 and same(np.ascontiguousarray(view), ref)
