"""BASELINE.json's full sizes (configs C2, C3, C4) on the device-resident path.  The oracle cannot
produce 10^9 outputs in seconds, so parity is checked (a) bit for bit on a random sample of
queries and (b) through size-independent properties: an interpolant reproduces its knots exactly,
evaluation of a shuffled batch is the shuffled result, and every search strategy gives the same
bits."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from ndarray_interp_b200 import _lib as L  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    from ndarray_interp_b200 import device
    device.set_device(0)
    return device


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sample_rows(out, idx):
    return out[torch.from_numpy(idx).cuda()].cpu().numpy()


def test_c2_cubic_natural_f64_full_size(D):
    rng = np.random.default_rng(1234)
    n, w, nq = 4096, 1024, 1 << 20
    g = np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.standard_normal((n, w))
    q = np.sort(rng.uniform(g[0], g[-1], nq))
    ip = D.DeviceInterp1D(dev(g), dev(y))
    ip.set_build_mode(L.BUILD_SEQUENTIAL)                               # the reference's elimination order
    st, _ = ip.spline_build(1)
    assert st == 0 and ip.build_levels() == 0
    a, b = ip.coeffs_to_host()
    st, a_seq, b_seq = O.spline_build(g, y, {"kind": "Natural"})
    assert np.array_equal(a, a_seq) and np.array_equal(b, b_seq)        # K6 at the full C2 shape, reference order
    for mode, info in ((L.BUILD_ROWSPLIT, 4), (L.BUILD_AUTO, -32)):     # 4096 rows: AUTO takes the partition build
        ip.set_build_mode(mode)
        st, _ = ip.spline_build(1)
        assert st == 0 and ip.build_levels() == info
        a, b = ip.coeffs_to_host()
        st, a_ref, b_ref = O.spline_build_as(g, y, {"kind": "Natural"}, info)
        assert np.array_equal(a, a_ref) and np.array_equal(b, b_ref)    # bit for bit against its specification
        for got, ref in ((a, a_seq), (b, b_seq)):                       # and inside 1e-12 of the reference order
            assert float((np.abs(got - ref) / np.maximum(np.abs(ref), np.abs(y).max(axis=0)[None, :])).max()) <= 1e-12
    err = D.new_err_word()
    out = ip.cubic(dev(q), 0, err=err)
    assert D.err_word_value(err) == D.ERR_NONE
    idx = np.sort(rng.choice(nq, 2048, replace=False))
    st, ref, _ = O.interp1d_cubic(g, y, a_ref, b_ref, q[idx], 0, nthreads=8)
    assert np.array_equal(sample_rows(out, idx), ref)                    # (a) sampled bit-exact parity
    knots = ip.cubic(dev(g), 0)                                          # (b) knots reproduce the data
    assert torch.equal(knots[:-1], dev(y)[:-1])
    assert np.allclose(knots[-1].cpu().numpy(), y[-1], rtol=1e-12, atol=1e-12)
    perm = torch.randperm(1 << 16, device="cuda")                        # order independence
    sub = dev(q[: 1 << 16])
    assert torch.equal(ip.cubic(sub[perm].contiguous(), 0), ip.cubic(sub, 0)[perm])
    del out


def test_c3_linear_extrapolate_f32_full_size(D):
    rng = np.random.default_rng(1234)
    n, w = 65536, 16
    g = np.cumsum(np.exp(rng.uniform(-2.0, 2.0, n))).astype(np.float32)
    assert len(np.unique(g)) == n
    y = rng.standard_normal((n, w), dtype=np.float32)
    span = float(g[-1]) - float(g[0])
    q = (float(g[0]) + span * rng.uniform(-0.026, 1.026, (4096, 4096))).astype(np.float32)
    ip = D.DeviceInterp1D(dev(g), dev(y))
    outs = []
    for mode in (L.SEARCH_AUTO, L.SEARCH_BINARY_GLOBAL, L.SEARCH_BINARY_SMEM, L.SEARCH_UNIFORM_GUESS, L.SEARCH_BUCKET_LUT, L.SEARCH_MERGE):
        ip.set_search_mode(mode)
        err = D.new_err_word()
        out = ip.linear(dev(q), True, err=err)
        assert out.shape == (4096, 4096, 16)                             # query.shape ++ data.shape[1:]
        assert D.err_word_value(err) == D.ERR_NONE
        outs.append(out)
        if len(outs) == 2:
            assert torch.equal(outs[0], outs[1])
            outs.pop()
    flat = outs[0].view(-1, w)
    idx = np.sort(rng.choice(q.size, 4096, replace=False))
    st, ref, _ = O.interp1d_linear(g, y, q.reshape(-1)[idx], True)
    assert np.array_equal(sample_rows(flat, idx), ref)
    ip.set_search_mode(L.SEARCH_AUTO)
    knots = ip.linear(dev(g), True)                                      # knots reproduce the data exactly
    assert torch.equal(knots[:-1], dev(y)[:-1])                          # (x - x1 == 0: m*0 + y1)
    assert np.allclose(knots[-1].cpu().numpy(), y[-1], rtol=1e-5, atol=1e-5)   # last knot: x1 + dx*m, one rounding
    # without extrapolation the first outside query (row-major) is reported
    err = D.new_err_word()
    ip.linear(dev(q), False, out=outs[0], err=err)
    outside = np.flatnonzero((q.reshape(-1) < g[0]) | (q.reshape(-1) > g[-1]))
    assert D.err_word_value(err) == int(outside[0])


@pytest.mark.parametrize("extrapolate", [False, True])
def test_c4_bilinear_f32_full_size(D, extrapolate):
    rng = np.random.default_rng(1234)
    n = m = 2048
    w, nq = 8, 1 << 24
    gx = np.linspace(0.0, 1.0, n).astype(np.float32)
    gy = (np.cumsum(rng.uniform(0.5, 1.5, m)) / m).astype(np.float32)
    z = rng.standard_normal((n, m, w), dtype=np.float32)
    lo, hi = (-0.026, 1.026) if extrapolate else (0.0, 1.0)
    qx = np.clip((gx[0] + (gx[-1] - gx[0]) * rng.uniform(lo, hi, nq)).astype(np.float32), -9, 9)
    qy = (gy[0] + (gy[-1] - gy[0]) * rng.uniform(lo, hi, nq)).astype(np.float32)
    if not extrapolate:
        qx, qy = np.clip(qx, gx[0], gx[-1]), np.clip(qy, gy[0], gy[-1])
    ip = D.DeviceInterp2D(dev(gx), dev(gy), dev(z))
    err = D.new_err_word()
    out = ip.bilinear(dev(qx), dev(qy), extrapolate, err=err)
    assert D.err_word_value(err) == D.ERR_NONE
    idx = np.sort(rng.choice(nq, 8192, replace=False))
    st, ref, _, _ = O.interp2d_bilinear(gx, gy, z, qx[idx], qy[idx], extrapolate)
    assert st == 0
    assert np.array_equal(sample_rows(out, idx), ref)
    # grid nodes reproduce the data exactly
    ii, jj = rng.integers(0, n - 1, 100000), rng.integers(0, m - 1, 100000)   # not the last node: see C3
    nodes = ip.bilinear(dev(gx[ii]), dev(gy[jj]), extrapolate)
    assert torch.equal(nodes, dev(z)[torch.from_numpy(ii).cuda(), torch.from_numpy(jj).cuda()])
    if not extrapolate:
        # an outside x is reported before an outside y of the same query; word = 2*q + axis
        qx2, qy2 = qx.copy(), qy.copy()
        qy2[777], qx2[777], qy2[555] = 9.0, 9.0, -9.0
        err = D.new_err_word()
        ip.bilinear(dev(qx2), dev(qy2), False, out=out, err=err)
        assert D.err_word_value(err) == 2 * 555 + 1


def test_c5a_bilinear_f32_at_scale_binned_equals_direct(D):
    """C5a: 2.1 GB table, 2^25 random queries -- AUTO bins the batch by table band.  Binning changes the
    order of evaluation only: every output row equals the direct kernel's, and a sample equals the oracle."""
    rng = np.random.default_rng(55)
    n = m = 4096
    w, nq = 32, 1 << 25
    gx = torch.linspace(0.0, 1.0, n, dtype=torch.float32, device="cuda")
    gy_h = (np.cumsum(rng.uniform(0.5, 1.5, m)) / m).astype(np.float32)
    data = torch.randn((n, m, w), dtype=torch.float32, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(7)
    qx = torch.rand(nq, dtype=torch.float32, device="cuda", generator=gen)
    qy = (torch.rand(nq, dtype=torch.float32, device="cuda", generator=gen) * float(gy_h[-1] - gy_h[0]) + float(gy_h[0]))
    qy = qy.clamp(float(gy_h[0]), float(gy_h[-1]))
    ip = D.DeviceInterp2D(gx, dev(gy_h), data)
    launches0 = D.kernel_launch_count()
    err = D.new_err_word()
    out_auto = ip.bilinear(qx, qy, False, err=err)
    assert D.err_word_value(err) == D.ERR_NONE
    assert D.kernel_launch_count() - launches0 == 3                       # count pass, scatter pass, evaluation
    ip.set_binning(L.BIN_OFF)
    out_direct = ip.bilinear(qx, qy, False)
    assert torch.equal(out_auto, out_direct)
    idx = np.sort(rng.choice(nq, 4096, replace=False))
    rows = np.unique(np.concatenate([np.searchsorted(np.linspace(0, 1, n, dtype=np.float32), qx[idx].cpu().numpy(), "right") - 1]))
    # the oracle needs the table on the host: take the x-rows the sample touches (and their successors) only
    need = np.unique(np.clip(np.concatenate([rows, rows + 1]), 0, n - 1))
    sub = data[torch.from_numpy(need).cuda()].cpu().numpy()
    gx_h = gx.cpu().numpy()
    # evaluate the sample row by row on the two x-rows it needs (a 2-row table is a valid bilinear table)
    got = sample_rows(out_auto, idx)
    qxs, qys = qx[idx].cpu().numpy(), qy[idx].cpu().numpy()
    pos = {r: k for k, r in enumerate(need)}
    for k in range(0, len(idx), 64):                                      # 64 oracle calls: enough to pin the bits
        i = int(min(max(np.searchsorted(gx_h, qxs[k], "right") - 1, 0), n - 2))
        tbl = np.stack([sub[pos[i]], sub[pos[i + 1]]])
        st, ref, _, _ = O.interp2d_bilinear(gx_h[i:i + 2], gy_h, tbl, qxs[k:k + 1], qys[k:k + 1], False)
        assert st == O.ST_OK and np.array_equal(got[k:k + 1], ref)


def test_c5b_cubic_f32_at_scale_sorted(D):
    """C5b: 2^25 sorted queries on a (4096, 32) f32 spline: 8192 queries per interval, so nearly every
    32-query tile takes the one-interval path (rows gathered once per tile); sampled bit-exact parity
    and equality with the same queries in shuffled order (which take the per-round path)."""
    rng = np.random.default_rng(56)
    n, w, nq = 4096, 32, 1 << 25
    g = np.cumsum(rng.uniform(0.5, 1.5, n)).astype(np.float32)
    y = rng.standard_normal((n, w), dtype=np.float32)
    ip = D.DeviceInterp1D(dev(g), dev(y))
    st, _ = ip.spline_build(1)
    assert st == 0
    a, b = ip.coeffs_to_host()
    st, a_ref, b_ref = O.spline_build_as(g, y, {"kind": "Natural"}, ip.build_levels())
    assert ip.build_levels() == -32 and np.array_equal(a, a_ref) and np.array_equal(b, b_ref)
    q = torch.sort(torch.rand(nq, dtype=torch.float32, device="cuda") * float(g[-1] - g[0]) + float(g[0]))[0].clamp(float(g[0]), float(g[-1]))
    err = D.new_err_word()
    out = ip.cubic(q, 0, err=err)
    assert D.err_word_value(err) == D.ERR_NONE
    idx = np.sort(rng.choice(nq, 8192, replace=False))
    st, ref, _ = O.interp1d_cubic(g, y, a_ref, b_ref, q[torch.from_numpy(idx).cuda()].cpu().numpy(), 0)
    assert st == O.ST_OK and np.array_equal(sample_rows(out, idx), ref)
    perm = torch.randperm(1 << 22, device="cuda")
    sub = q[: 1 << 22].contiguous()
    assert torch.equal(ip.cubic(sub[perm].contiguous(), 0), out[: 1 << 22][perm])
