// ndi_spline.cu -- K6: cubic-spline coefficient construction on the device.
//
// Replaces CubicSpline::calc_coefficients + solve_for_k + thomas
// (src/interp1d/strategies/cubic_spline.rs:310-368, :409-674, :678-721) and
// solve_for_k_individual (:370-403).
//
// Structure of the reference computation and how it is mapped here:
//   * the tridiagonal matrix depends only on x and on the boundary kinds, so for the five
//     whole-dataset boundary conditions it is SHARED by all trailing columns; only the right-hand
//     side is per column (:440-471).  Its Thomas forward elimination (multipliers w[i] and the
//     eliminated diagonal, :690-692) is therefore done once, by `spline_factor_kernel`, in exactly
//     the reference's serial order.  For Periodic the second solve (rhs2, :535-550) is the same for
//     every column too and is also done there.
//   * `spline_columns_kernel` is the column-parallel Thomas: one thread per trailing column, the
//     columns of a row are contiguous so every load and store is coalesced.  The forward sweep
//     builds the right-hand side on the fly from a sliding window of three data rows and parks the
//     swept values in the `b` output buffer; the backward sweep turns them into k and immediately
//     into a and b (:354-365), so k is never written to memory: the build reads data twice and
//     writes a and b once.
//   * BoundaryCondition::Individual gives every column its own first/last matrix row, hence its
//     own elimination; `spline_columns_individual_kernel` does the factorisation per column
//     (scratch: one eliminated diagonal per column).
//
// Every floating-point operation is performed in the reference's order with one rounding each
// (no FMA, true division), so the coefficients are bit-identical to the oracle's.
//
// Reference quirk kept on purpose: the last diagonal entry of the NotAKnot system is
// x[n-1]-x[n-2] (:635) where the textbook system has x[n-2]-x[n-3]; see
// tests/test_oracle_golden.py::test_not_a_knot_right_boundary_quirk.
#include "ndi_device.cuh"
#include "ndi_internal.h"

namespace ndi {

enum { SB_NAK = 0, SB_NATURAL = 1, SB_CLAMPED = 2, SB_FIRST = 3, SB_SECOND = 4 };
enum { BC_NAK = 0, BC_NATURAL = 1, BC_CLAMPED = 2, BC_PERIODIC = 3, BC_INDIVIDUAL = 4 };

template <class T>
struct Side { int kind; T val; };

// SingleBoundary::specialize (:287-296)
template <class T>
__host__ __device__ inline Side<T> specialize(Side<T> s) {
    if (s.kind == SB_NATURAL) return {SB_SECOND, (T)0};
    if (s.kind == SB_CLAMPED) return {SB_FIRST, (T)0};
    return s;
}

template <class T> struct A : Ar<T> {};
#define ADD A<T>::add
#define SUB A<T>::sub
#define MUL A<T>::mul
#define DIV A<T>::div

// matrix row i of the full (non-periodic) system: (:440-451) interior, (:584-590) the 3-point
// NotAKnot parabola system, (:599-669) boundary rows.
template <class T>
__device__ __forceinline__ void matrix_row(const T* __restrict__ x, int n, int i, int lk, int rk, bool nak3, T& up,
                                           T& mid, T& low) {
    const T two = (T)2, one = (T)1, zero = (T)0;
    if (i > 0 && i < n - 1) {
        const T dxn = SUB(x[i + 1], x[i]), dxn_1 = SUB(x[i], x[i - 1]);
        up = dxn_1; mid = MUL(two, ADD(dxn, dxn_1)); low = dxn;
    } else if (i == 0) {
        low = zero;
        const T dx0 = SUB(x[1], x[0]);
        if (nak3) { mid = one; up = one; }
        else if (lk == SB_NAK) { mid = SUB(x[2], x[1]); up = SUB(x[2], x[0]); }
        else if (lk == SB_FIRST) { mid = one; up = zero; }
        else { up = dx0; mid = MUL(two, dx0); }
    } else {
        up = zero;
        const T dx_1 = SUB(x[n - 1], x[n - 2]);
        if (nak3) { low = one; mid = one; }
        else if (rk == SB_NAK) { mid = dx_1; low = SUB(x[n - 1], x[n - 3]); }
        else if (rk == SB_FIRST) { mid = one; low = zero; }
        else { mid = MUL(two, dx_1); low = dx_1; }
    }
}
// row i of the condensed periodic system (:512-518), i in [0, n-3]
template <class T>
__device__ __forceinline__ void matrix_row_periodic(const T* __restrict__ x, int n, int i, T& up, T& mid, T& low) {
    const T two = (T)2;
    if (i == 0) {
        const T dx0 = SUB(x[1], x[0]), dx_1 = SUB(x[n - 1], x[n - 2]);
        mid = MUL(two, ADD(dx_1, dx0)); up = dx_1; low = (T)0;
    } else {
        const T dxn = SUB(x[i + 1], x[i]), dxn_1 = SUB(x[i], x[i - 1]);
        up = dxn_1; mid = MUL(two, ADD(dxn, dxn_1)); low = dxn;
    }
}

// right-hand sides ---------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T rhs_interior(T yl, T ym, T yr, T dxn, T dxn_1) {       // :468
    const T three = (T)3;
    return MUL(three, ADD(DIV(MUL(dxn, SUB(ym, yl)), dxn_1), DIV(MUL(dxn_1, SUB(yr, ym)), dxn)));
}
template <class T>
__device__ __forceinline__ T rhs_left(const T* __restrict__ x, Side<T> l, T y0, T y1, T y2) {
    const T two = (T)2, three = (T)3;
    const T dx0 = SUB(x[1], x[0]), dx1 = SUB(x[2], x[1]);
    if (l.kind == SB_NAK) {                                                           // :600-610
        const T d = SUB(x[2], x[0]);
        const T tmp1 = MUL(ADD(dx0, MUL(two, d)), dx1);
        return DIV(ADD(DIV(MUL(tmp1, SUB(y1, y0)), dx0), DIV(MUL(MUL(dx0, dx0), SUB(y2, y1)), dx1)), d);
    }
    if (l.kind == SB_FIRST) return l.val;                                             // :614-618
    return SUB(MUL(three, SUB(y1, y0)), DIV(MUL(l.val, MUL(dx0, dx0)), two));         // :629
}
template <class T>
__device__ __forceinline__ T rhs_right(const T* __restrict__ x, int n, Side<T> r, T y_1, T y_2, T y_3) {
    const T two = (T)2, three = (T)3;
    const T dx_1 = SUB(x[n - 1], x[n - 2]), dx_2 = SUB(x[n - 2], x[n - 3]);
    if (r.kind == SB_NAK) {                                                           // :635-647
        const T d = SUB(x[n - 1], x[n - 3]);
        const T tmp1 = MUL(ADD(MUL(two, d), dx_1), dx_2);
        return DIV(ADD(DIV(MUL(MUL(dx_1, dx_1), SUB(y_2, y_3)), dx_2), DIV(MUL(tmp1, SUB(y_1, y_2)), dx_1)), d);
    }
    if (r.kind == SB_FIRST) return r.val;                                             // :651-655
    return ADD(MUL(three, SUB(y_1, y_2)), DIV(MUL(r.val, MUL(dx_1, dx_1)), two));     // :666
}

// ---- shared-matrix factorisation (one thread; serial by definition) ---------------------------
// fac layout: up[n] | mid[n] (eliminated) | wl[n] | k2[n]
template <class T>
__global__ void spline_factor_kernel(const T* __restrict__ x, int n, int periodic, int lk, int rk, T* __restrict__ fac) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T* up = fac; T* mid = fac + n; T* wl = fac + 2 * (size_t)n; T* k2 = fac + 3 * (size_t)n;
    const bool nak3 = !periodic && n == 3 && lk == SB_NAK && rk == SB_NAK;
    const int len = periodic ? n - 2 : n;
    T u, m, l;
    if (periodic) matrix_row_periodic<T>(x, n, 0, u, m, l); else matrix_row<T>(x, n, 0, lk, rk, nak3, u, m, l);
    up[0] = u; mid[0] = m; wl[0] = (T)0;
    T m_prev = m, u_prev = u;
    for (int i = 1; i < len; ++i) {                                                   // thomas :690-692
        if (periodic) matrix_row_periodic<T>(x, n, i, u, m, l); else matrix_row<T>(x, n, i, lk, rk, nak3, u, m, l);
        const T w = DIV(l, m_prev);
        m = SUB(m, MUL(w, u_prev));
        up[i] = u; mid[i] = m; wl[i] = w;
        m_prev = m; u_prev = u;
    }
    if (periodic && n > 3) {                                                          // rhs2 solve :535-550
        const T dx0 = SUB(x[1], x[0]), dx_3 = SUB(x[n - 3], x[n - 4]);
        T prev = (T)0;
        for (int i = 0; i < len; ++i) {
            T r = (T)0;
            if (i == 0) r = -dx0;
            if (i == n - 3) r = -dx_3;
            if (i > 0) r = SUB(r, MUL(wl[i], prev));
            k2[i] = r; prev = r;
        }
        T kr = DIV(k2[len - 1], mid[len - 1]);
        k2[len - 1] = kr;
        for (int i = len - 2; i >= 0; --i) {
            kr = DIV(SUB(k2[i], MUL(up[i], kr)), mid[i]);
            k2[i] = kr;
        }
    }
}

// ---- column-parallel Thomas, shared matrix ------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(128) spline_columns_kernel(const T* __restrict__ x, int n, const T* __restrict__ y,
                                                             long long w, int periodic, Side<T> left, Side<T> right,
                                                             const T* __restrict__ fac, T* __restrict__ a,
                                                             T* __restrict__ b, unsigned long long* err) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= w) return;
    const T* up = fac; const T* mid = fac + n; const T* wl = fac + 2 * (size_t)n; const T* k2 = fac + 3 * (size_t)n;
    const T one = (T)1, two = (T)2, three = (T)3;
    auto Y = [&](int r) -> T { return __ldg(y + (long long)r * w + c); };
    auto Aat = [&](int r) -> T& { return a[(long long)r * w + c]; };
    auto Bat = [&](int r) -> T& { return b[(long long)r * w + c]; };
    const T dx0 = SUB(x[1], x[0]), dx1 = SUB(x[2], x[1]);
    const T dx_1 = SUB(x[n - 1], x[n - 2]), dx_2 = SUB(x[n - 2], x[n - 3]);

    if (periodic) {
        const T y0 = Y(0), yN = Y(n - 1);
        if (y0 != yN) { atomicMin(err, (unsigned long long)c); return; }            // :483, :501
        if (n == 3) {                                                                 // :480-496
            const T y1 = Y(1), y2 = yN;
            const T slope0 = DIV(SUB(y1, y0), dx0), slope1 = DIV(SUB(y2, y1), dx1);
            const T k = DIV(ADD(DIV(slope0, dx0), DIV(slope1, dx1)), ADD(DIV(one, dx0), DIV(one, dx1)));
            const T dy0 = SUB(y1, y0), dy1 = SUB(y2, y1);
            Aat(0) = SUB(MUL(k, dx0), dy0); Bat(0) = SUB(dy0, MUL(k, dx0));
            Aat(1) = SUB(MUL(k, dx1), dy1); Bat(1) = SUB(dy1, MUL(k, dx1));
            return;
        }
        const int len = n - 2;
        // forward sweep of rhs1 over the condensed system, parked in a[0..len)
        const T y1 = Y(1);
        const T slope0 = DIV(SUB(y1, y0), dx0);                                       // :521
        const T yn1 = yN, yn2 = Y(n - 2), yn3 = Y(n - 3);
        const T slope_1 = DIV(SUB(yn1, yn2), dx_1), slope_2 = DIV(SUB(yn2, yn3), dx_2);   // :526-527
        const T rhs_first = MUL(ADD(MUL(slope_1, dx0), MUL(slope0, dx_1)), three);    // :529-530
        const T rhs_last = MUL(ADD(MUL(slope_2, dx_1), MUL(slope_1, dx_2)), three);   // :531-532 (row n-2)
        T r_prev = rhs_first;
        Aat(0) = r_prev;
        T yl = y0, ym = y1;
        for (int i = 1; i < len; ++i) {
            const T yr = Y(i + 1);
            const T dxn = SUB(x[i + 1], x[i]), dxn_1 = SUB(x[i], x[i - 1]);
            const T r = SUB(rhs_interior<T>(yl, ym, yr, dxn, dxn_1), MUL(wl[i], r_prev));
            Aat(i) = r; r_prev = r; yl = ym; ym = yr;
        }
        // back substitution -> k1, parked in a[0..len)
        T kr = DIV(r_prev, mid[len - 1]);
        const T k1_last = kr;
        Aat(len - 1) = kr;
        for (int i = len - 2; i >= 0; --i) {
            kr = DIV(SUB(Aat(i), MUL(up[i], kr)), mid[i]);
            Aat(i) = kr;
        }
        const T k1_0 = kr;
        const T k_m1 = DIV(SUB(SUB(rhs_last, MUL(k1_0, dx_2)), MUL(k1_last, dx_1)),
                           ADD(ADD(MUL(k2[0], dx_2), MUL(k2[len - 1], dx_1)), MUL(two, ADD(dx_1, dx_2))));   // :552-557
        // k[i] = k1[i] + k_m1*k2[i] (i < n-2), k[n-2] = k_m1, k[n-1] = k[0]  (:559-563); then a, b (:354-365)
        const T k0 = ADD(k1_0, MUL(k_m1, k2[0]));
        T k_i = k0, y_i = y0;
        for (int i = 0; i <= n - 2; ++i) {
            T k_next;
            if (i + 1 < len) k_next = ADD(Aat(i + 1), MUL(k_m1, k2[i + 1]));
            else if (i + 1 == n - 2) k_next = k_m1;
            else k_next = k0;
            const T y_next = Y(i + 1);
            const T dx = SUB(x[i + 1], x[i]), dy = SUB(y_next, y_i);
            Aat(i) = SUB(MUL(k_i, dx), dy);
            Bat(i) = SUB(dy, MUL(k_next, dx));
            k_i = k_next; y_i = y_next;
        }
        return;
    }

    const Side<T> l = specialize(left), r = specialize(right);
    const bool nak3 = n == 3 && l.kind == SB_NAK && r.kind == SB_NAK;
    // forward sweep; swept right-hand side parked in b[0..n-1)
    T yl = Y(0), ym = Y(1), yr = Y(2);
    T r_prev;
    if (nak3) r_prev = MUL(DIV(SUB(ym, yl), dx0), two);                               // :592
    else r_prev = rhs_left<T>(x, l, yl, ym, yr);
    Bat(0) = r_prev;
#pragma unroll 4
    for (int i = 1; i < n - 1; ++i) {
        if (i > 1) yr = Y(i + 1);
        T rhs;
        if (nak3) {
            const T slope0 = DIV(SUB(ym, yl), dx0), slope1 = DIV(SUB(yr, ym), dx1);
            rhs = MUL(ADD(MUL(slope1, dx0), MUL(slope0, dx1)), three);                // :593-594
        } else {
            const T dxn = SUB(x[i + 1], x[i]), dxn_1 = SUB(x[i], x[i - 1]);
            rhs = rhs_interior<T>(yl, ym, yr, dxn, dxn_1);
        }
        const T rr = SUB(rhs, MUL(wl[i], r_prev));                                    // :698
        Bat(i) = rr; r_prev = rr;
        if (i < n - 2) { yl = ym; ym = yr; }
    }
    // now yl, ym, yr = y[n-3], y[n-2], y[n-1]
    T rhs_n;
    if (nak3) rhs_n = MUL(DIV(SUB(yr, ym), dx1), two);                                // :595
    else rhs_n = rhs_right<T>(x, n, r, yr, ym, yl);
    const T r_last = SUB(rhs_n, MUL(wl[n - 1], r_prev));
    // back substitution fused with a, b
    T k_right = DIV(r_last, mid[n - 1]);                                              // :704-708
    T y_right = yr;
#pragma unroll 4
    for (int i = n - 2; i >= 0; --i) {
        const T k = DIV(SUB(Bat(i), MUL(up[i], k_right)), mid[i]);                    // :716
        const T y_i = (i == n - 2) ? ym : Y(i);
        const T dx = SUB(x[i + 1], x[i]), dy = SUB(y_right, y_i);
        Aat(i) = SUB(MUL(k, dx), dy);                                                 // :362
        Bat(i) = SUB(dy, MUL(k_right, dx));                                           // :363
        k_right = k; y_right = y_i;
    }
}

// ---- column-parallel Thomas, per-column boundary conditions (Individual) ----------------------------
template <class T>
__global__ void __launch_bounds__(128) spline_columns_individual_kernel(
    const T* __restrict__ x, int n, const T* __restrict__ y, long long w, const int32_t* __restrict__ lks,
    const T* __restrict__ lvs, const int32_t* __restrict__ rks, const T* __restrict__ rvs, T* __restrict__ a,
    T* __restrict__ b, T* __restrict__ dia) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= w) return;
    const T two = (T)2, three = (T)3;
    auto Y = [&](int r) -> T { return __ldg(y + (long long)r * w + c); };
    auto Aat = [&](int r) -> T& { return a[(long long)r * w + c]; };
    auto Bat = [&](int r) -> T& { return b[(long long)r * w + c]; };
    auto Dat = [&](int r) -> T& { return dia[(long long)r * w + c]; };
    const Side<T> l = specialize(Side<T>{lks[c], lvs[c]}), r = specialize(Side<T>{rks[c], rvs[c]});
    const bool nak3 = n == 3 && l.kind == SB_NAK && r.kind == SB_NAK;
    const T dx0 = SUB(x[1], x[0]), dx1 = SUB(x[2], x[1]);

    T up_prev, m_prev, lo;
    matrix_row<T>(x, n, 0, l.kind, r.kind, nak3, up_prev, m_prev, lo);
    Dat(0) = m_prev;
    T yl = Y(0), ym = Y(1), yr = Y(2);
    T r_prev = nak3 ? MUL(DIV(SUB(ym, yl), dx0), two) : rhs_left<T>(x, l, yl, ym, yr);
    Bat(0) = r_prev;
    for (int i = 1; i < n - 1; ++i) {
        if (i > 1) yr = Y(i + 1);
        T u, m, lw;
        matrix_row<T>(x, n, i, l.kind, r.kind, nak3, u, m, lw);
        const T wgt = DIV(lw, m_prev);
        m = SUB(m, MUL(wgt, up_prev));
        Dat(i) = m;
        T rhs;
        if (nak3) {
            const T slope0 = DIV(SUB(ym, yl), dx0), slope1 = DIV(SUB(yr, ym), dx1);
            rhs = MUL(ADD(MUL(slope1, dx0), MUL(slope0, dx1)), three);
        } else {
            const T dxn = SUB(x[i + 1], x[i]), dxn_1 = SUB(x[i], x[i - 1]);
            rhs = rhs_interior<T>(yl, ym, yr, dxn, dxn_1);
        }
        const T rr = SUB(rhs, MUL(wgt, r_prev));
        Bat(i) = rr; r_prev = rr; m_prev = m; up_prev = u;
        if (i < n - 2) { yl = ym; ym = yr; }
    }
    T u, m, lw;
    matrix_row<T>(x, n, n - 1, l.kind, r.kind, nak3, u, m, lw);
    const T wgt = DIV(lw, m_prev);
    m = SUB(m, MUL(wgt, up_prev));
    const T rhs_n = nak3 ? MUL(DIV(SUB(yr, ym), dx1), two) : rhs_right<T>(x, n, r, yr, ym, yl);
    const T r_last = SUB(rhs_n, MUL(wgt, r_prev));
    T k_right = DIV(r_last, m);
    T y_right = yr;
    for (int i = n - 2; i >= 0; --i) {
        T ui, mi, li;
        matrix_row<T>(x, n, i, l.kind, r.kind, nak3, ui, mi, li);
        const T k = DIV(SUB(Bat(i), MUL(ui, k_right)), Dat(i));
        const T y_i = (i == n - 2) ? ym : Y(i);
        const T dx = SUB(x[i + 1], x[i]), dy = SUB(y_right, y_i);
        Aat(i) = SUB(MUL(k, dx), dy);
        Bat(i) = SUB(dy, MUL(k_right, dx));
        k_right = k; y_right = y_i;
    }
}

template <class T>
size_t spline_scratch_elems(int64_t n, int64_t w, int bc_kind) {
    return bc_kind == BC_INDIVIDUAL ? (size_t)n * (size_t)w : 4 * (size_t)n;
}

template <class T>
cudaError_t launch_spline_build(const T* x, int64_t n, const T* data, int64_t w, int bc_kind, const int32_t* lk,
                                const T* lv, const int32_t* rk, const T* rv, T* a, T* b, T* scratch,
                                unsigned long long* err, cudaStream_t st) {
    if (w <= 0) return cudaSuccess;
    // few columns: small blocks so that more SMs take part; many columns: 128-thread blocks
    const int block = (w <= 32ll * 2 * device_info().sm_count) ? 32 : 128;
    const int grid = (int)((w + block - 1) / block);
    if (bc_kind == BC_INDIVIDUAL) {
        spline_columns_individual_kernel<T><<<grid, block, 0, st>>>(x, (int)n, data, (long long)w, lk, lv, rk, rv, a, b, scratch);
        count_launch();
        return cudaGetLastError();
    }
    Side<T> l{SB_NAK, (T)0}, r{SB_NAK, (T)0};
    if (bc_kind == BC_NATURAL) l = r = Side<T>{SB_NATURAL, (T)0};
    if (bc_kind == BC_CLAMPED) l = r = Side<T>{SB_CLAMPED, (T)0};
    const int periodic = bc_kind == BC_PERIODIC;
    const Side<T> ls = specialize(l), rs = specialize(r);
    spline_factor_kernel<T><<<1, 32, 0, st>>>(x, (int)n, periodic, ls.kind, rs.kind, scratch);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    spline_columns_kernel<T><<<grid, block, 0, st>>>(x, (int)n, data, (long long)w, periodic, l, r, scratch, a, b, err);
    count_launch();
    return cudaGetLastError();
}

template cudaError_t launch_spline_build<float>(const float*, int64_t, const float*, int64_t, int, const int32_t*,
                                                const float*, const int32_t*, const float*, float*, float*, float*,
                                                unsigned long long*, cudaStream_t);
template cudaError_t launch_spline_build<double>(const double*, int64_t, const double*, int64_t, int, const int32_t*,
                                                 const double*, const int32_t*, const double*, double*, double*,
                                                 double*, unsigned long long*, cudaStream_t);
template size_t spline_scratch_elems<float>(int64_t, int64_t, int);
template size_t spline_scratch_elems<double>(int64_t, int64_t, int);

}  // namespace ndi
