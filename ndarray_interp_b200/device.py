"""Device-resident, stream-ordered entry points (the `_dev` half of include/ndi_b200.h) for
callers that already hold their queries and results in HBM.

torch is used here for what it is good at -- device memory, streams, torch.distributed -- and
nothing else: tensors are passed to the C ABI as raw device pointers, and every kernel that runs
is one of ours.  One fused launch per call, no host synchronisation; the first failing query is
reported through a 64-bit device word (NDI_ERR_WORD_NONE when every query was fine).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_TORCH_DT = {torch.float32: L.F32, torch.float64: L.F64, torch.int32: L.I32, torch.int64: L.I64}
for _name, _code_ in (("uint32", L.U32), ("uint64", L.U64)):      # torch >= 2.3 has the (storage-only) unsigned types
    if hasattr(torch, _name):
        _TORCH_DT[getattr(torch, _name)] = _code_
ERR_NONE = L.ERR_WORD_NONE


def _code(t):
    try:
        return _TORCH_DT[t.dtype]
    except KeyError:
        raise TypeError(f"dtype {t.dtype} is not supported (f32, f64, i32, i64, u32, u64)") from None


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(stream):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def set_device(index):
    """one process per GPU: bind this process's library state to cuda:<index> (LOCAL_RANK)"""
    torch.cuda.set_device(index)
    L.check(L.require_device().ndi_set_device(int(index)))


def new_err_word(device=None):
    return torch.full((1,), -1, dtype=torch.int64, device=device or torch.cuda.current_device())


def err_word_value(err):
    """host read of the device error word (synchronises)"""
    return int(err.item()) & 0xFFFFFFFFFFFFFFFF


class DeviceInterp1D:
    """Interp1D whose grid, data and spline coefficients live in HBM (borrowed from torch tensors)."""

    def __init__(self, x, data, assume_valid=False):
        assert x.is_cuda and data.is_cuda and x.is_contiguous() and data.is_contiguous()
        assert x.dtype == data.dtype and x.dim() == 1 and data.shape[0] == x.shape[0]
        self.lib = L.require_device()
        self.x, self.data = x, data          # keep the tensors alive: the handle borrows them
        self.n = x.shape[0]
        self.trailing = tuple(data.shape[1:])
        self.w = int(np.prod(self.trailing, dtype=np.int64))
        self.a = self.b = None
        torch.cuda.current_stream().synchronize()      # tables must be complete before the grid check reads them
        self.h = C.c_void_p()
        flags = L.DEVICE_POINTERS | L.BORROW | (L.ASSUME_VALID if assume_valid else 0)
        st = L.check(self.lib.ndi_interp1d_create(_code(data), _p(x), self.n, _p(data), self.w, flags, C.byref(self.h)))
        if st == L.NOT_MONOTONIC:
            raise ValueError(L.last_error())

    def __del__(self):
        try:
            if self.h:
                self.lib.ndi_interp1d_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def set_search_mode(self, mode):
        L.check(self.lib.ndi_interp1d_set_search_mode(self.h, int(mode)))

    def set_build_mode(self, mode, levels=0):
        """how spline_build solves the tridiagonal system: L.BUILD_AUTO / BUILD_SEQUENTIAL (the reference's order,
        bit-identical coefficients) / BUILD_ROWSPLIT (`levels` steps of cyclic reduction, 0: library's choice) /
        BUILD_PARTITION (blocks of `levels` rows, 0: 32)"""
        L.check(self.lib.ndi_interp1d_set_build_mode(self.h, int(mode), int(levels)))

    def build_levels(self):
        """how the current coefficients were built: 0 the reference's order, L > 0 row-split levels, -m partition blocks"""
        lv = C.c_int32(-1)
        L.check(self.lib.ndi_interp1d_build_info(self.h, C.byref(lv)))
        return lv.value

    def spline_build(self, bc_kind=0, left_kind=None, left_val=None, right_kind=None, right_val=None):
        """CubicSpline::calc_coefficients on the device; returns the status (0 or PERIODIC_MISMATCH)"""
        torch.cuda.current_stream().synchronize()
        bad = C.c_int64(-1)
        st = L.check(self.lib.ndi_interp1d_spline_build(self.h, int(bc_kind), L.ptr(left_kind), L.ptr(left_val),
                                                        L.ptr(right_kind), L.ptr(right_val), C.byref(bad)))
        return st, bad.value

    def coeff_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        L.check(self.lib.ndi_interp1d_device_ptrs(self.h, None, None, C.byref(a), C.byref(b)))
        return a.value, b.value

    def coeffs_to_host(self):
        dt = {torch.float32: np.float32, torch.float64: np.float64}[self.data.dtype]
        a = np.zeros((self.n - 1,) + self.trailing, dtype=dt)
        b = np.zeros_like(a)
        L.check(self.lib.ndi_interp1d_spline_coeffs(self.h, L.ptr(a), L.ptr(b)))
        return a, b

    def _out(self, q, out):
        if out is None:
            out = torch.empty(tuple(q.shape) + self.trailing, dtype=self.data.dtype, device=q.device)
        assert out.is_contiguous() and out.numel() == q.numel() * self.w
        return out

    def linear(self, q, extrapolate=False, out=None, err=None, stream=None):
        assert q.is_cuda and q.is_contiguous() and q.dtype == self.data.dtype
        out = self._out(q, out)
        L.check(self.lib.ndi_interp1d_linear_dev(self.h, _p(q), q.numel(), int(bool(extrapolate)), _p(out), _p(err),
                                                 _stream(stream)))
        return out

    def cubic(self, q, extrap_mode=0, out=None, err=None, stream=None):
        assert q.is_cuda and q.is_contiguous() and q.dtype == self.data.dtype
        out = self._out(q, out)
        L.check(self.lib.ndi_interp1d_cubic_dev(self.h, _p(q), q.numel(), int(extrap_mode), _p(out), _p(err),
                                                _stream(stream)))
        return out


class DeviceInterp2D:
    """Interp2D + Bilinear with tables borrowed from torch tensors."""

    def __init__(self, x, y, data, assume_valid=False):
        assert x.is_cuda and y.is_cuda and data.is_cuda
        assert x.is_contiguous() and y.is_contiguous() and data.is_contiguous()
        assert data.shape[0] == x.shape[0] and data.shape[1] == y.shape[0]
        self.lib = L.require_device()
        self.x, self.y, self.data = x, y, data
        self.trailing = tuple(data.shape[2:])
        self.w = int(np.prod(self.trailing, dtype=np.int64))
        torch.cuda.current_stream().synchronize()
        self.h = C.c_void_p()
        flags = L.DEVICE_POINTERS | L.BORROW | (L.ASSUME_VALID if assume_valid else 0)
        st = L.check(self.lib.ndi_interp2d_create(_code(data), _p(x), x.shape[0], _p(y), y.shape[0], _p(data), self.w,
                                                  flags, C.byref(self.h)))
        if st == L.NOT_MONOTONIC:
            raise ValueError(L.last_error())

    def __del__(self):
        try:
            if self.h:
                self.lib.ndi_interp2d_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def set_search_mode(self, mode):
        L.check(self.lib.ndi_interp2d_set_search_mode(self.h, int(mode)))

    def set_binning(self, mode, band_rows=0):
        """locality binning of query batches by table band: L.BIN_AUTO / BIN_OFF / BIN_ON"""
        L.check(self.lib.ndi_interp2d_set_binning(self.h, int(mode), int(band_rows)))

    def bilinear(self, qx, qy, extrapolate=False, out=None, err=None, stream=None):
        assert qx.is_cuda and qy.is_cuda and qx.is_contiguous() and qy.is_contiguous() and qx.shape == qy.shape
        if out is None:
            out = torch.empty(tuple(qx.shape) + self.trailing, dtype=self.data.dtype, device=qx.device)
        L.check(self.lib.ndi_interp2d_bilinear_dev(self.h, _p(qx), _p(qy), qx.numel(), int(bool(extrapolate)), _p(out),
                                                   _p(err), _stream(stream)))
        return out


def lower_index(grid, q, search_mode=L.SEARCH_AUTO, err=None, stream=None):
    """get_lower_index for a device batch; returns int64 indices"""
    assert grid.is_cuda and q.is_cuda and grid.dtype == q.dtype and grid.is_contiguous() and q.is_contiguous()
    idx = torch.empty(q.shape, dtype=torch.int64, device=q.device)
    L.check(L.require_device().ndi_lower_index_dev(_code(grid), _p(grid), grid.numel(), _p(q), q.numel(), _p(idx),
                                                   _p(err), int(search_mode), _stream(stream)))
    return idx


def kernel_launch_count():
    return int(L.load().ndi_kernel_launch_count())
