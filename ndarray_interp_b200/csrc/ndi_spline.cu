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
//   * per column only two first-order recurrences are serial in the row index:
//         forward   r[i] = rhs[i] - w[i] * r[i-1]                 (:698)
//         backward  k[i] = (r[i] - up[i] * k[i+1]) / mid'[i]      (:716)
//     Everything else -- the right-hand side (:468, two divisions per element) and a, b from k
//     (:354-365) -- is independent per element.  The build is therefore three launches:
//         spline_rhs_kernel     all elements in parallel: rhs into the scratch matrix R (n x w)
//         spline_sweep_kernel   one thread per column: the two recurrences in place on R, rows streamed
//                               through shared memory with cp.async so that a step waits for arithmetic
//                               only; the division by mid'[i] uses the reciprocal formed once per row by
//                               the factor kernel (ndi_device.cuh, Hoisted) -- same quotient, 3 dependent
//                               operations instead of ~13
//         spline_ab_kernel      all elements in parallel: a, b from k and y
//     (first version: one fused kernel, one thread per column doing all of it serially: 4.3 ms for
//     4096 x 1024 f64 because 32 warps executed ~150 instructions per row each; profiles/r01.)
//     `spline_columns_kernel` is that fused version, kept for the 3-point special cases.
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
#include "ndi_spline.cuh"

namespace ndi {


template <class T>
__global__ void __launch_bounds__(kFacBlock) spline_factor_kernel(const T* __restrict__ x, int n, int periodic, int lk, int rk,
                                                                  T* __restrict__ fac, size_t fac_stride) {
    __shared__ T su[kFacTile], sm[kFacTile], sl[kFacTile];
    __shared__ T carry[2];
    if (gridDim.x > 1) { lk = ind_kind(blockIdx.x / 3); rk = ind_kind(blockIdx.x % 3); fac += blockIdx.x * fac_stride; }
    FacRow<T>* rows = reinterpret_cast<FacRow<T>*>(fac);
    T* k2 = fac + 4 * (size_t)n;
    const bool nak3 = !periodic && n == 3 && lk == SB_NAK && rk == SB_NAK;
    const int len = periodic ? n - 2 : n;
    for (int base = 0; base < len; base += kFacTile) {
        const int cnt = min(kFacTile, len - base);
        for (int j = threadIdx.x; j < cnt; j += kFacBlock) {
            T u, m, l;
            if (periodic) matrix_row_periodic<T>(x, n, base + j, u, m, l); else matrix_row<T>(x, n, base + j, lk, rk, nak3, u, m, l);
            su[j] = u; sm[j] = m; sl[j] = l;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            T m_prev = base ? carry[0] : (T)0, u_prev = base ? carry[1] : (T)0;
            for (int j = 0; j < cnt; ++j) {
                T m = sm[j];
                T w = (T)0;
                if (base + j > 0) {
                    w = DIV(sl[j], m_prev);
                    m = SUB(m, MUL(w, u_prev));
                }
                sm[j] = m; sl[j] = w;
                m_prev = m; u_prev = su[j];
            }
            carry[0] = m_prev; carry[1] = u_prev;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += kFacBlock) rows[base + j] = FacRow<T>{su[j], sm[j], sl[j], Hoisted<T>::rcp(sm[j])};
        __syncthreads();
    }
    if (periodic && n > 3 && threadIdx.x == 0) {                                      // rhs2 solve :535-550
        __threadfence_block();
        const T dx0 = SUB(x[1], x[0]), dx_3 = SUB(x[n - 3], x[n - 4]);
        T prev = (T)0;
        for (int i = 0; i < len; ++i) {
            T r = (T)0;
            if (i == 0) r = -dx0;
            if (i == n - 3) r = -dx_3;
            if (i > 0) r = SUB(r, MUL(rows[i].wl, prev));
            k2[i] = r; prev = r;
        }
        T kr = DIV(k2[len - 1], rows[len - 1].mid);
        k2[len - 1] = kr;
        for (int i = len - 2; i >= 0; --i) {
            kr = DIV(SUB(k2[i], MUL(rows[i].up, kr)), rows[i].mid);
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
    const FacRow<T>* rows = reinterpret_cast<const FacRow<T>*>(fac);
    const T* k2 = fac + 4 * (size_t)n;
    auto up = [&](int i) { return rows[i].up; };
    auto mid = [&](int i) { return rows[i].mid; };
    auto wl = [&](int i) { return rows[i].wl; };
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
            const T r = SUB(rhs_interior<T>(yl, ym, yr, dxn, dxn_1), MUL(wl(i), r_prev));
            Aat(i) = r; r_prev = r; yl = ym; ym = yr;
        }
        // back substitution -> k1, parked in a[0..len)
        T kr = DIV(r_prev, mid(len - 1));
        const T k1_last = kr;
        Aat(len - 1) = kr;
        for (int i = len - 2; i >= 0; --i) {
            kr = DIV(SUB(Aat(i), MUL(up(i), kr)), mid(i));
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
        const T rr = SUB(rhs, MUL(wl(i), r_prev));                                    // :698
        Bat(i) = rr; r_prev = rr;
        if (i < n - 2) { yl = ym; ym = yr; }
    }
    // now yl, ym, yr = y[n-3], y[n-2], y[n-1]
    T rhs_n;
    if (nak3) rhs_n = MUL(DIV(SUB(yr, ym), dx1), two);                                // :595
    else rhs_n = rhs_right<T>(x, n, r, yr, ym, yl);
    const T r_last = SUB(rhs_n, MUL(wl(n - 1), r_prev));
    // back substitution fused with a, b
    T k_right = DIV(r_last, mid(n - 1));                                              // :704-708
    T y_right = yr;
#pragma unroll 4
    for (int i = n - 2; i >= 0; --i) {
        const T k = DIV(SUB(Bat(i), MUL(up(i), k_right)), mid(i));                    // :716
        const T y_i = (i == n - 2) ? ym : Y(i);
        const T dx = SUB(x[i + 1], x[i]), dy = SUB(y_right, y_i);
        Aat(i) = SUB(MUL(k, dx), dy);                                                 // :362
        Bat(i) = SUB(dy, MUL(k_right, dx));                                           // :363
        k_right = k; y_right = y_i;
    }
}

// ---- the three-launch build ---------------------------------------------------------------------------
constexpr int kRows = 8, kRing = 4;       // sweep: kRing batches of kRows rows in flight per thread

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <class T>
__device__ __forceinline__ void cp_async_elem(unsigned smem_addr, const T* gmem_src) {
    if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gmem_src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gmem_src) : "memory");
}
template <class T>
__device__ __forceinline__ T lds_elem(unsigned smem_addr) {
    T v;
    // volatile keeps it between the cp.async wait before and the next prefetch after it (both volatile);
    // no memory clobber, so ordinary loads and stores may be scheduled across it
    if constexpr (sizeof(T) == 8) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_addr));
    else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_addr));
    return v;
}

// launch 1: right-hand sides.  Non-periodic: R[0] = left boundary row, R[1..n-2] = interior rows (:468),
// R[n-1] = right boundary row.  Periodic (n > 3): R[0] = condensed first row (:529-530), R[1..n-3] interior,
// R[n-2] = the last condensed equation's right-hand side (:531-532), kept for k_m1; columns whose first
// and last value differ are reported through err (:499-507).
template <class T, int V>
__global__ void __launch_bounds__(256) spline_rhs_kernel(const T* __restrict__ x, int n, const T* __restrict__ y, long long w,
                                                         int periodic, Side<T> left, Side<T> right, T* __restrict__ R,
                                                         unsigned long long* err, const int32_t* __restrict__ lks = nullptr,
                                                         const T* __restrict__ lvs = nullptr, const int32_t* __restrict__ rks = nullptr,
                                                         const T* __restrict__ rvs = nullptr, const int32_t* __restrict__ pos = nullptr) {
    const T three = (T)3;
    const int rows = periodic ? n - 1 : n;
    const long long wv = w / V, ntasks = row_task_count(wv, rows, blockDim.x);
    for (long long task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const RowTask t = row_task(task, wv, rows);
        if (!t.live) continue;
        const long long col0 = t.col * V;
        // window y[row-1 .. row+kRowGroup] shared by the rows of the group (clamped at the ends, where it is not used)
        T yv[kRowGroup + 2][V];
#pragma unroll
        for (int j = 0; j < kRowGroup + 2; ++j) ld_vec<T, V>(y + (long long)min(max(t.row - 1 + j, 0), n - 1) * w + col0, yv[j]);
        // grid steps of the intervals row-1 .. row+kRowGroup-1 and their reciprocals: the two divisions of an interior
        // row are divisions by grid steps, the same for all columns -- formed once per row group, every quotient is
        // the IEEE quotient (ndi_device.cuh, Hoisted) at 3 operations instead of ~25
        T dv[kRowGroup + 1], rv[kRowGroup + 1];
#pragma unroll
        for (int j = 0; j < kRowGroup + 1; ++j) {
            const int i0 = min(max(t.row - 1 + j, 0), n - 2);
            dv[j] = SUB(__ldg(x + i0 + 1), __ldg(x + i0));
            rv[j] = Hoisted<T>::rcp(dv[j]);
        }
#pragma unroll
        for (int j = 0; j < kRowGroup; ++j) {
            const int i = t.row + j;
            if (i >= rows) break;
            T out[V];
            const bool interior = periodic ? (i > 0 && i < n - 2) : (i > 0 && i < n - 1);
            if (interior) {
                const T dxn = dv[j + 1], dxn_1 = dv[j];                               // x[i+1] - x[i], x[i] - x[i-1]
#pragma unroll
                for (int c = 0; c < V; ++c)
                    out[c] = MUL(three, ADD(Hoisted<T>::div(MUL(dxn, SUB(yv[j + 1][c], yv[j][c])), dxn_1, rv[j]),
                                            Hoisted<T>::div(MUL(dxn_1, SUB(yv[j + 2][c], yv[j + 1][c])), dxn, rv[j + 1])));   // :468
            } else {
#pragma unroll 1
                for (int c = 0; c < V; ++c) {
                    const long long col = col0 + c;
                    const T* ycol = y + col;
                    auto Y = [&](int row) -> T { return __ldg(ycol + (long long)row * w); };
                    T v;
                    if (periodic) {
                        const T dx0 = SUB(x[1], x[0]), dx_1 = SUB(x[n - 1], x[n - 2]), dx_2 = SUB(x[n - 2], x[n - 3]);
                        if (i == 0) {
                            const T y0 = Y(0), yN = Y(n - 1);
                            if (y0 != yN) atomicMin(err, (unsigned long long)col);
                            const T slope0 = DIV(SUB(Y(1), y0), dx0);                 // :521
                            const T slope_1 = DIV(SUB(yN, Y(n - 2)), dx_1);           // :526
                            v = MUL(ADD(MUL(slope_1, dx0), MUL(slope0, dx_1)), three);   // :529-530
                        } else {
                            const T yn2 = Y(n - 2);
                            const T slope_1 = DIV(SUB(Y(n - 1), yn2), dx_1), slope_2 = DIV(SUB(yn2, Y(n - 3)), dx_2);   // :526-527
                            v = MUL(ADD(MUL(slope_2, dx_1), MUL(slope_1, dx_2)), three);   // :531-532
                        }
                    } else if (i == 0) {
                        // Individual: this column's own boundary
                        const Side<T> l = specialize(lks ? Side<T>{lks[col], lvs[col]} : left);
                        v = rhs_left<T>(x, l, Y(0), Y(1), Y(2));
                    } else {
                        const Side<T> r = specialize(rks ? Side<T>{rks[col], rvs[col]} : right);
                        v = rhs_right<T>(x, n, r, Y(n - 1), Y(n - 2), Y(n - 3));
                    }
                    out[c] = v;
                }
            }
            if (V == 1 && pos) R[(long long)i * w + pos[col0]] = out[0];              // Individual: columns grouped by matrix
            else st_vec<T, V>(R + (long long)i * w + col0, out);
        }
    }
}

// launch 2: the two recurrences over rows [0, len) of R, in place (thomas :690-720 with the matrix part
// already done by the factor kernel).  One thread per column.  A step must wait for arithmetic only,
// so everything it reads is staged in shared memory ahead of time with cp.async:
//   * the column's own rows go through a private ring (kRing batches of kRows rows in flight); a thread
//     reads back only what it copied itself, so cp.async.wait_group is all the synchronisation needed;
//   * the matrix rows (the same for every column) are staged per warp, 32 rows at a time, one row per
//     lane, double buffered; __syncwarp makes the other lanes' copies visible.
// Pointers advance by one row per step and all shared-memory offsets are compile-time constants.
template <class T>
__device__ __forceinline__ void cp_async_fac(unsigned smem_addr, const FacRow<T>* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(src) : "memory");
    if constexpr (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr + 16), "l"(reinterpret_cast<const char*>(src) + 16) : "memory");
}

template <class T, int BLOCK>
__global__ void __launch_bounds__(BLOCK) spline_sweep_kernel(int total_len, int nsys, long long w, long long ncols,
                                                             const T* __restrict__ fac, T* __restrict__ R, const SweepGroups groups) {
    extern __shared__ __align__(32) unsigned char ring_raw[];
    constexpr int kSuper = kRing * kRows;                     // rows per outer iteration
    static_assert(kSuper == 32, "one matrix row per lane and outer iteration");
    constexpr unsigned kSlot = BLOCK * sizeof(T);
    constexpr unsigned kFacBytes = sizeof(FacRow<T>);
    int block = blockIdx.x;
    if (groups.ngroups) {
        int g = 0;
        while (g + 1 < groups.ngroups && block >= groups.first_block[g + 1]) ++g;
        block -= groups.first_block[g];
        ncols = groups.count[g]; R += groups.col_off[g]; fac += g * groups.fac_stride;
    }
    // interleaved systems (row-split build): the blocks of a group are laid out system by system; system j is
    // rows j, j + nsys, ... of R, so a chain advances by nsys rows per step and reads every nsys-th matrix row
    const int bps = (int)((ncols + BLOCK - 1) / BLOCK);       // blocks per system
    const int sys = block / bps;
    block -= sys * bps;
    const int len = (total_len - sys + nsys - 1) / nsys;      // rows of this system
    const long long rs = (long long)nsys * w;                 // elements between consecutive rows of a chain
    R += (long long)sys * w;
    const long long c0 = (long long)block * BLOCK + threadIdx.x;
    const bool live = c0 < ncols;                             // dead lanes shadow the last column (they stage matrix rows too)
    const long long c = live ? c0 : ncols - 1;                // w: row stride of R, ncols: columns of this group
    const int lane = threadIdx.x & 31;
    const unsigned smem0 = (unsigned)__cvta_generic_to_shared(ring_raw);
    const unsigned ring = smem0 + threadIdx.x * (unsigned)sizeof(T);
    const unsigned facbuf = smem0 + kSuper * kSlot + (threadIdx.x >> 5) * (2 * kSuper * kFacBytes);   // this warp's [2][32] rows
    const FacRow<T>* rows = reinterpret_cast<const FacRow<T>*>(fac) + sys;   // row t of this system: rows[t * nsys]
    T* col = R + c;

    T k;                                                      // k[len-1] after the forward sweep, then the running k[i+1]
    // ---- forward: r[i] = rhs[i] - wl[i] * r[i-1], i = 1 .. len-1
    {
        const T* pf = col + rs;                                // next row to prefetch
        int pf_left = len - 1;
        auto prefetch = [&](int slot0) {
#pragma unroll
            for (int j = 0; j < kRows; ++j) {
                if (j < pf_left) cp_async_elem<T>(ring + (slot0 + j) * kSlot, pf);
                pf += rs;
            }
            pf_left -= kRows;
        };
        int fac_row = 1 + lane;                               // the matrix row this lane stages next
        auto stage_fac = [&](int buf) {
            if (fac_row < len) cp_async_fac<T>(facbuf + (buf * kSuper + lane) * kFacBytes, rows + (long long)fac_row * nsys);
            fac_row += kSuper;
        };
        stage_fac(0);
#pragma unroll
        for (int b = 0; b < kRing - 1; ++b) { prefetch(b * kRows); cp_async_commit(); }
        T prev = col[0];
        T* wp = col + rs;
        int buf = 0;
        for (int left = len - 1; left > 0; buf ^= 1) {
#pragma unroll
            for (int bb = 0; bb < kRing; ++bb) {
                if (bb == 0) stage_fac(buf ^ 1);
                prefetch(((bb + kRing - 1) % kRing) * kRows);
                cp_async_commit();
                cp_async_wait<kRing - 1>();
                if (bb == 0) __syncwarp();
                // all shared-memory reads of the batch first, then the chain: a step waits for its predecessor only
                T rv[kRows], wv[kRows];
#pragma unroll
                for (int j = 0; j < kRows; ++j) {
                    wv[j] = lds_elem<T>(facbuf + (buf * kSuper + bb * kRows + j) * kFacBytes + 2 * sizeof(T));
                    rv[j] = lds_elem<T>(ring + (bb * kRows + j) * kSlot);
                }
#pragma unroll
                for (int j = 0; j < kRows; ++j) {
                    if (j < left) {
                        prev = SUB(rv[j], MUL(wv[j], prev));                          // :698
                        if (live) *wp = prev;
                    }
                    wp += rs;
                }
                left -= kRows;
            }
            __syncwarp();                                     // everyone is done with buf before it is staged again
        }
        const FacRow<T> flast = ld_fac<T>(rows + (long long)(len - 1) * nsys);
        k = Hoisted<T>::div(prev, flast.mid, flast.rmid);                             // :704-708
        if (live) col[(long long)(len - 1) * rs] = k;
    }
    cp_async_wait<0>();
    __threadfence();                                          // the swept values are read back below
    __syncwarp();

    // ---- backward: k[i] = (r[i] - up[i] * k[i+1]) / mid[i], i = len-2 .. 0
    {
        const T* pf = col + (long long)(len - 2) * rs;
        int pf_left = len - 1;
        auto prefetch = [&](int slot0) {
#pragma unroll
            for (int j = 0; j < kRows; ++j) {
                if (j < pf_left) cp_async_elem<T>(ring + (slot0 + j) * kSlot, pf);
                pf -= rs;
            }
            pf_left -= kRows;
        };
        int fac_row = len - 2 - lane;
        auto stage_fac = [&](int buf) {
            if (fac_row >= 0) cp_async_fac<T>(facbuf + (buf * kSuper + lane) * kFacBytes, rows + (long long)fac_row * nsys);
            fac_row -= kSuper;
        };
        stage_fac(0);
#pragma unroll
        for (int b = 0; b < kRing - 1; ++b) { prefetch(b * kRows); cp_async_commit(); }
        T* wp = col + (long long)(len - 2) * rs;
        int buf = 0;
        for (int left = len - 1; left > 0; buf ^= 1) {
#pragma unroll
            for (int bb = 0; bb < kRing; ++bb) {
                if (bb == 0) stage_fac(buf ^ 1);
                prefetch(((bb + kRing - 1) % kRing) * kRows);
                cp_async_commit();
                cp_async_wait<kRing - 1>();
                if (bb == 0) __syncwarp();
                T rv[kRows], uv[kRows], mv[kRows], iv[kRows];
#pragma unroll
                for (int j = 0; j < kRows; ++j) {
                    const unsigned fa = facbuf + (buf * kSuper + bb * kRows + j) * kFacBytes;
                    uv[j] = lds_elem<T>(fa); mv[j] = lds_elem<T>(fa + sizeof(T)); iv[j] = lds_elem<T>(fa + 3 * sizeof(T));
                    rv[j] = lds_elem<T>(ring + (bb * kRows + j) * kSlot);
                }
#pragma unroll
                for (int j = 0; j < kRows; ++j) {
                    if (j < left) {
                        k = Hoisted<T>::div(SUB(rv[j], MUL(uv[j], k)), mv[j], iv[j]);  // :716
                        if (live) *wp = k;
                    }
                    wp -= rs;
                }
                left -= kRows;
            }
            __syncwarp();
        }
    }
}

// periodic only, between launch 2 and 3: k_m1 (:552-557) into R[n-2], k[n-1] = k[0] (:563) into R[n-1]
template <class T>
__global__ void __launch_bounds__(256) spline_periodic_close_kernel(const T* __restrict__ x, int n, long long w,
                                                                    const T* __restrict__ fac, T* __restrict__ R) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= w) return;
    const T* k2 = fac + 4 * (size_t)n;
    const int len = n - 2;
    const T two = (T)2;
    const T dx_1 = SUB(x[n - 1], x[n - 2]), dx_2 = SUB(x[n - 2], x[n - 3]);
    const T k1_0 = R[c], k1_last = R[(long long)(len - 1) * w + c], rhs_last = R[(long long)(n - 2) * w + c];
    const T k_m1 = DIV(SUB(SUB(rhs_last, MUL(k1_0, dx_2)), MUL(k1_last, dx_1)),
                       ADD(ADD(MUL(k2[0], dx_2), MUL(k2[len - 1], dx_1)), MUL(two, ADD(dx_1, dx_2))));
    R[(long long)(n - 2) * w + c] = k_m1;
    R[(long long)(n - 1) * w + c] = ADD(k1_0, MUL(k_m1, k2[0]));
}

// launch 3: a[i] = k[i] dx - dy, b[i] = dy - k[i+1] dx (:354-365).  Periodic: k[i] = k1[i] + k_m1 k2[i]
// for i < n-2 (:559-561), rows n-2 and n-1 of R already hold k.
template <class T, int V>
__global__ void __launch_bounds__(256) spline_ab_kernel(const T* __restrict__ x, int n, const T* __restrict__ y, long long w,
                                                        int periodic, const T* __restrict__ fac, const T* __restrict__ R,
                                                        T* __restrict__ a, T* __restrict__ b, const int32_t* __restrict__ pos = nullptr) {
    const T* k2 = fac + 4 * (size_t)n;
    const long long wv = w / V, ntasks = row_task_count(wv, n - 1, blockDim.x);
    for (long long task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const RowTask t = row_task(task, wv, n - 1);
        if (!t.live) continue;
        const long long col0 = t.col * V;
        const long long at0 = (long long)t.row * w + col0;
        const long long rcol = (V == 1 && pos) ? pos[col0] : col0;
        T kv[kRowGroup + 1][V], yv[kRowGroup + 1][V];
#pragma unroll
        for (int j = 0; j <= kRowGroup; ++j) {
            const int i = min(t.row + j, n - 1);
            ld_vec<T, V>(R + (long long)i * w + rcol, kv[j]);
            ld_vec<T, V>(y + (long long)i * w + col0, yv[j]);
        }
        if (periodic) {
            T k_m1[V];
            ld_vec<T, V>(R + (long long)(n - 2) * w + col0, k_m1);
#pragma unroll
            for (int j = 0; j <= kRowGroup; ++j)
                if (t.row + j < n - 2) {
                    const T k2v = k2[t.row + j];
#pragma unroll
                    for (int c = 0; c < V; ++c) kv[j][c] = ADD(kv[j][c], MUL(k_m1[c], k2v));
                }
        }
#pragma unroll
        for (int j = 0; j < kRowGroup; ++j) {
            const int i = t.row + j;
            if (i >= n - 1) break;
            const T dx = SUB(x[i + 1], x[i]);
            T av[V], bv[V];
#pragma unroll
            for (int c = 0; c < V; ++c) {
                const T dy = SUB(yv[j + 1][c], yv[j][c]);
                av[c] = SUB(MUL(kv[j][c], dx), dy);
                bv[c] = SUB(dy, MUL(kv[j + 1][c], dx));
            }
            st_vec<T, V>(a + at0 + (long long)j * w, av);
            st_vec<T, V>(b + at0 + (long long)j * w, bv);
        }
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

static size_t seq_fac_elems(int64_t n) { return (5 * (size_t)n + 3) & ~(size_t)3; }

template <class T>
size_t spline_scratch_elems(int64_t n, int64_t w, int bc_kind, int levels) {
    // factorisation(s) (FacRow[n] + k2[n] (+ the row-split coefficients), padded) + the matrix R;
    // Individual: nine factorisations (n < 4: a diagonal per column)
    const size_t fac = levels > 0 ? rowsplit_fac_elems(n, levels) : (levels < 0 ? partition_fac_elems(n, -levels) : seq_fac_elems(n));
    if (bc_kind == BC_INDIVIDUAL) return n < 4 ? (size_t)n * (size_t)w : 9 * fac + (size_t)n * (size_t)w;
    return fac + (size_t)n * (size_t)w;
}

template <class T>
cudaError_t launch_spline_sweep(int len, int nsys, long long w, const T* fac, size_t fac_stride, T* R, const int64_t* counts,
                                cudaStream_t st) {
    // few chains: small blocks so that more SMs take part; many: 128-thread blocks
    const int blk = ((long long)nsys * w <= 32ll * 2 * device_info().sm_count) ? 32 : 128;
    SweepGroups sg = {};
    long long nblocks = (long long)nsys * ((w + blk - 1) / blk);
    if (counts) {
        sg.ngroups = 9; sg.fac_stride = fac_stride;
        long long off = 0, first = 0;
        for (int g = 0; g < 9; ++g) {
            sg.first_block[g] = (int)first; sg.col_off[g] = off; sg.count[g] = counts[g];
            first += (long long)nsys * ((counts[g] + blk - 1) / blk); off += counts[g];
        }
        sg.first_block[9] = (int)first;
        nblocks = first;
    }
    if (nblocks <= 0) return cudaSuccess;
    const size_t smem = (size_t)kRing * kRows * blk * sizeof(T) + (size_t)(blk / 32) * 2 * kRing * kRows * sizeof(FacRow<T>);
    if (blk == 32) spline_sweep_kernel<T, 32><<<(unsigned)nblocks, 32, smem, st>>>(len, nsys, w, w, fac, R, sg);
    else spline_sweep_kernel<T, 128><<<(unsigned)nblocks, 128, smem, st>>>(len, nsys, w, w, fac, R, sg);
    count_launch();
    return cudaGetLastError();
}

int row_group_grid(long long wv, long long nrows) {
    const long long cap = (long long)device_info().sm_count * 8;
    const long long tasks = row_task_count(wv, nrows, 256);
    return (int)(tasks < cap ? tasks : cap);
}

template <class T>
cudaError_t launch_spline_ab(const T* x, int n, const T* y, long long w, int periodic, const T* fac, const T* R, T* a, T* b,
                             const int32_t* pos, cudaStream_t st) {
    constexpr int V = 16 / sizeof(T);
    if (!pos && vec_ok<T>(w, y, R, a, b)) spline_ab_kernel<T, V><<<row_group_grid(w / V, n - 1), 256, 0, st>>>(x, n, y, w, periodic, fac, R, a, b, nullptr);
    else spline_ab_kernel<T, 1><<<row_group_grid(w, n - 1), 256, 0, st>>>(x, n, y, w, periodic, fac, R, a, b, pos);
    count_launch();
    return cudaGetLastError();
}

template <class T>
static cudaError_t launch_spline_rhs_pos(const T* x, int n, const T* y, long long w, int periodic, Side<T> left, Side<T> right, T* R,
                                         unsigned long long* err, const int32_t* lks, const T* lvs, const int32_t* rks, const T* rvs,
                                         const int32_t* pos, cudaStream_t st) {
    constexpr int V = 16 / sizeof(T);
    const int rows = periodic ? n - 1 : n;
    if (!pos && vec_ok<T>(w, y, R, R, R)) spline_rhs_kernel<T, V><<<row_group_grid(w / V, rows), 256, 0, st>>>(x, n, y, w, periodic, left, right, R, err, lks, lvs, rks, rvs, nullptr);
    else spline_rhs_kernel<T, 1><<<row_group_grid(w, rows), 256, 0, st>>>(x, n, y, w, periodic, left, right, R, err, lks, lvs, rks, rvs, pos);
    count_launch();
    return cudaGetLastError();
}

template <class T>
cudaError_t launch_spline_rhs(const T* x, int n, const T* y, long long w, int periodic, Side<T> left, Side<T> right, T* R,
                              unsigned long long* err, const int32_t* lks, const T* lvs, const int32_t* rks, const T* rvs,
                              cudaStream_t st) {
    return launch_spline_rhs_pos<T>(x, n, y, w, periodic, left, right, R, err, lks, lvs, rks, rvs, nullptr, st);
}

template <class T>
cudaError_t launch_spline_periodic_close(const T* x, int n, long long w, const T* fac, T* R, cudaStream_t st) {
    spline_periodic_close_kernel<T><<<(int)((w + 255) / 256), 256, 0, st>>>(x, n, w, fac, R);
    count_launch();
    return cudaGetLastError();
}

// levels == 0: the reference's elimination order (coefficients bit-identical to the reference arithmetic);
// levels < 0: partition build with blocks of -levels rows (ndi_partition.cu);
// levels > 0: row-split build -- `levels` steps of parallel cyclic reduction, then 2^levels interleaved systems
// per column (ndi_rowsplit.cu); the caller has checked that the systems keep at least two rows.
template <class T>
cudaError_t launch_spline_build(const T* x, int64_t n, const T* data, int64_t w, int bc_kind, const int32_t* lk,
                                const T* lv, const int32_t* rk, const T* rv, const int32_t* pos, const int64_t* group_count,
                                int levels, T* a, T* b, T* scratch, unsigned long long* err, cudaStream_t st) {
    if (w <= 0) return cudaSuccess;
    if (levels < 0 && n >= 4)                                // partition build (ndi_partition.cu), blocks of -levels rows
        return launch_partition_build<T>(x, n, data, w, bc_kind, -levels, lk, lv, rk, rv, a, b, scratch, err, st);
    if (levels < 0) levels = 0;
    // few columns: small blocks so that more SMs take part; many columns: 128-thread blocks
    const int block = (w <= 32ll * 2 * device_info().sm_count) ? 32 : 128;
    const int grid = (int)((w + block - 1) / block);
    const size_t fac_elems = levels > 0 ? rowsplit_fac_elems(n, levels) : seq_fac_elems(n);
    const int nsys = 1 << levels;
    cudaError_t e;
    if (bc_kind == BC_INDIVIDUAL) {
        if (n < 4 || !pos || !group_count) {                 // the 3-point special cases: one factorisation per column
            spline_columns_individual_kernel<T><<<grid, block, 0, st>>>(x, (int)n, data, (long long)w, lk, lv, rk, rv, a, b, scratch);
            count_launch();
            return cudaGetLastError();
        }
        // columns grouped by (left kind, right kind): nine shared-matrix builds in the same three launches
        T* R = scratch + 9 * fac_elems;
        if (levels > 0) {
            if ((e = launch_rowsplit_front<T>(x, n, data, w, bc_kind, levels, lk, lv, rk, rv, pos, scratch, fac_elems, R, err, st)) != cudaSuccess) return e;
        } else {
            spline_factor_kernel<T><<<9, kFacBlock, 0, st>>>(x, (int)n, 0, 0, 0, scratch, fac_elems);
            count_launch();
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = launch_spline_rhs_pos<T>(x, (int)n, data, (long long)w, 0, Side<T>{SB_NAK, (T)0}, Side<T>{SB_NAK, (T)0}, R, err, lk, lv, rk, rv, pos, st)) != cudaSuccess) return e;
        }
        if ((e = launch_spline_sweep<T>((int)n, nsys, (long long)w, scratch, fac_elems, R, group_count, st)) != cudaSuccess) return e;
        return launch_spline_ab<T>(x, (int)n, data, (long long)w, 0, scratch, R, a, b, pos, st);
    }
    Side<T> l{SB_NAK, (T)0}, r{SB_NAK, (T)0};
    if (bc_kind == BC_NATURAL) l = r = Side<T>{SB_NATURAL, (T)0};
    if (bc_kind == BC_CLAMPED) l = r = Side<T>{SB_CLAMPED, (T)0};
    const int periodic = bc_kind == BC_PERIODIC;
    const Side<T> ls = specialize(l), rs = specialize(r);
    if (n >= 4) {
        T* R = scratch + fac_elems;
        if (levels > 0) {
            if ((e = launch_rowsplit_front<T>(x, n, data, w, bc_kind, levels, nullptr, nullptr, nullptr, nullptr, nullptr, scratch, 0, R, err, st)) != cudaSuccess) return e;
        } else {
            spline_factor_kernel<T><<<1, kFacBlock, 0, st>>>(x, (int)n, periodic, ls.kind, rs.kind, scratch, 0);
            count_launch();
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = launch_spline_rhs<T>(x, (int)n, data, (long long)w, periodic, l, r, R, err, nullptr, nullptr, nullptr, nullptr, st)) != cudaSuccess) return e;
        }
        if ((e = launch_spline_sweep<T>(periodic ? (int)n - 2 : (int)n, nsys, (long long)w, scratch, 0, R, nullptr, st)) != cudaSuccess) return e;
        if (periodic && (e = launch_spline_periodic_close<T>(x, (int)n, (long long)w, scratch, R, st)) != cudaSuccess) return e;
        return launch_spline_ab<T>(x, (int)n, data, (long long)w, periodic, scratch, R, a, b, nullptr, st);
    }
    spline_factor_kernel<T><<<1, kFacBlock, 0, st>>>(x, (int)n, periodic, ls.kind, rs.kind, scratch, 0);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    spline_columns_kernel<T><<<grid, block, 0, st>>>(x, (int)n, data, (long long)w, periodic, l, r, scratch, a, b, err);
    count_launch();
    return cudaGetLastError();
}

#define NDI_INST_SPLINE(T)                                                                                                   \
    template cudaError_t launch_spline_build<T>(const T*, int64_t, const T*, int64_t, int, const int32_t*, const T*,         \
                                                const int32_t*, const T*, const int32_t*, const int64_t*, int, T*, T*, T*,   \
                                                unsigned long long*, cudaStream_t);                                          \
    template cudaError_t launch_spline_sweep<T>(int, int, long long, const T*, size_t, T*, const int64_t*, cudaStream_t);    \
    template cudaError_t launch_spline_ab<T>(const T*, int, const T*, long long, int, const T*, const T*, T*, T*,            \
                                             const int32_t*, cudaStream_t);                                                  \
    template cudaError_t launch_spline_periodic_close<T>(const T*, int, long long, const T*, T*, cudaStream_t);              \
    template cudaError_t launch_spline_rhs<T>(const T*, int, const T*, long long, int, Side<T>, Side<T>, T*,                \
                                              unsigned long long*, const int32_t*, const T*, const int32_t*, const T*,        \
                                              cudaStream_t);                                                                  \
    template size_t spline_scratch_elems<T>(int64_t, int64_t, int, int);
NDI_INST_SPLINE(float)
NDI_INST_SPLINE(double)

}  // namespace ndi
