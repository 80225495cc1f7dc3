// ndi_partition.cu -- K6, third build mode: block partition of the rows ("substructuring").
//
// The reference solves its tridiagonal system with one serial Thomas sweep per column
// (cubic_spline.rs:678-721 under solve_for_k, :409-674).  The row-split mode (ndi_rowsplit.cu) shortens the chains by
// 2^L at the price of L elementwise passes with a halo; it is bound by the latency of those passes (profiles/r02).
// This mode cuts the rows into blocks of m - 1 rows separated by single rows (rows m-1, 2m-1, ...).  With the
// separators' unknowns known the blocks are independent, so
//   1. every (block, column) pair is ONE thread that keeps its right-hand sides in registers and runs the two Thomas
//      recurrences of the block on them (part_local_kernel): g = A_block^-1 rhs;
//   2. the block's response to its two separators (the "spikes" p = A_block^-1 low[first] e_first,
//      q = A_block^-1 up[last] e_last) depends on x only and is formed once per block (part_factor_kernel);
//   3. the separators' equations  -low[s] p[s-1] k[s-m] + (mid[s] - low[s] q[s-1] - up[s] p[s+1]) k[s]
//      - up[s] q[s+1] k[s+m] = rhs[s] - low[s] g[s-1] - up[s] g[s+1]  are again a tridiagonal system with a matrix
//      shared by all columns, m times shorter: the same three steps are applied to it until at most min(4 m, 128)
//      rows are left, which one lane per column solves directly out of shared memory (part_top_kernel);
//   4. k = g - p k_left - q k_right, elementwise, from the top level down (part_corr_kernel); on level 0 that
//      correction is part of the kernel that writes a and b (part_ab_kernel), so k itself is never stored.
// Every chain is m - 1 steps of one fused multiply-add (+ one multiplication by a reciprocal formed once per row); no
// pass needs a halo.  All matrix work runs on a second stream beside the right-hand-side kernel.  (Forming the
// right-hand sides inside the block solve was measured and dropped: 67 us against 30 + 15 us at the C2 shape, 192
// registers per thread; profiles/r02/partition_build.md.)
//
// The rounding differs from the reference's elimination order, so this is NOT bit-identical to the reference
// arithmetic; like the row-split mode it is held bit for bit to the checker's operation-by-operation specification
// of the same scheme (partition_thomas; tests/test_partition_gpu.py) and to north_star's 1e-12 (f64) / 1e-5 (f32)
// bars against the reference-order checker.
#include <map>

#include "ndi_spline.cuh"

namespace ndi {

#define FMA A<T>::fma

constexpr int kPartLevelsMax = 16;
constexpr int kPartTopMax = 128;                     // rows of the directly solved system: min(kPartTopMax, 4 m)

// Level l: a tridiagonal system of `len` rows; its row j lives in row (j + 1) * stride - 1 of the scratch matrix R
// (level 0: stride 1; the rows of level l + 1 are the separators of level l).  Per level, at fac + off:
//   FacRow[len] {up, eliminated mid, elimination weight, 1 / eliminated mid} of the block factorisations
//   low[len] | mid[len] | up[len]   the level's matrix        p[len] | q[len]   the spikes
// fac[0, n): the grid steps dx[i] = x[i+1] - x[i], fac[n, 2n): their reciprocals as Hoisted<T>::rcp forms them,
// fac[4n, 5n): k2 (periodic), where the close and a / b kernels of ndi_spline.cu look for it.
struct PartLevel { int len, stride; unsigned long long off; };
struct PartPlan { int nsplit, m; unsigned long long elems; PartLevel lv[kPartLevelsMax + 1]; };

int partition_block_for(int requested) {
    if (requested <= 0) return kPartBlockDefault;
    return requested < 3 ? 3 : (requested > kPartBlockMax ? kPartBlockMax : requested);
}
static int part_top_rows(int m) { return 4 * m < kPartTopMax ? 4 * m : kPartTopMax; }

static PartPlan part_plan(int64_t n, int64_t len, int m) {
    PartPlan p{};
    p.m = m;
    unsigned long long off = (5ull * (unsigned long long)n + 3) & ~3ull;
    long long cur = len, stride = 1;
    int l = 0;
    auto put = [&](int at) {
        p.lv[at] = PartLevel{(int)cur, (int)stride, off};
        off += (9ull * (unsigned long long)cur + 3) & ~3ull;
    };
    while (cur > part_top_rows(m) && l < kPartLevelsMax) { put(l); cur /= m; stride *= m; ++l; }
    put(l);
    p.nsplit = l;
    p.elems = off;
    return p;
}
size_t partition_fac_elems(int64_t n, int block) { return (size_t)part_plan(n, n, block).elems; }

template <class T>
struct PartArrays {
    FacRow<T>* fr; T *low, *mid, *up, *p, *q;
    __host__ __device__ PartArrays(T* facb, const PartLevel& L) {
        T* b = facb + L.off;
        fr = reinterpret_cast<FacRow<T>*>(b);
        low = b + 4 * (size_t)L.len; mid = low + L.len; up = mid + L.len; p = up + L.len; q = p + L.len;
    }
};

// matrix row j of level l: level 0 from x (solve_for_k :440-451, :599-669; periodic: the condensed system :512-518),
// level l >= 1 from the spikes of level l - 1 (step 3 above)
template <class T>
__device__ __forceinline__ void part_row(const T* __restrict__ x, int n, int periodic, int lk, int rk, const PartPlan& pl,
                                         T* facb, int l, int j, T& low, T& mid, T& up) {
    if (l == 0) {
        if (periodic) matrix_row_periodic<T>(x, n, j, up, mid, low); else matrix_row<T>(x, n, j, lk, rk, false, up, mid, low);
        return;
    }
    const PartLevel& pv = pl.lv[l - 1];
    const PartArrays<T> a(facb, pv);
    const int s = j * pl.m + pl.m - 1, lastL = s - 1, firstR = s + 1;
    const bool right = firstR < pv.len;
    const T ls = a.low[s], us = a.up[s];
    low = -MUL(ls, a.p[lastL]);
    T md = FMA(-ls, a.q[lastL], a.mid[s]);
    if (right) md = FMA(-us, a.p[firstR], md);
    mid = md;
    up = right ? -MUL(us, a.q[firstR]) : (T)0;
}
template <class T>
__device__ __forceinline__ T part_rhs2(const T* __restrict__ x, int n, int len, int j) {          // :535-538
    const T dx0 = SUB(x[1], x[0]), dx_3 = SUB(x[n - 3], x[n - 4]);
    return j == 0 ? -dx0 : (j == len - 1 ? -dx_3 : (T)0);
}
// what every level-0 row leaves behind besides its matrix row: rhs2 of the periodic system, the grid step and its reciprocal
template <class T>
__device__ __forceinline__ void part_row_extras(const T* __restrict__ x, int n, int periodic, int len, T* fac0, T* facb, int j) {
    if (periodic) facb[4 * (size_t)n + j] = part_rhs2<T>(x, n, len, j);
    if (blockIdx.y == 0 && j + 1 < n) {
        const T d = SUB(x[j + 1], x[j]);
        fac0[j] = d; fac0[(size_t)n + j] = Hoisted<T>::rcp(d);
    }
}

// One warp per block of a split level: the lanes form the block's matrix rows (and its separator's); lane 0 runs the
// Thomas elimination of the block (:690-692), a chain of divisions; the reciprocals of the eliminated diagonal are
// formed by all lanes; lanes 0 and 1 run the two spikes.
constexpr int kPfWarps = 4;
template <class T>
__global__ void __launch_bounds__(32 * kPfWarps) part_factor_kernel(const T* __restrict__ x, int n, int periodic, int lk, int rk,
                                                                    const PartPlan pl, int l, T* fac, size_t fac_stride) {
    __shared__ T s_lo[kPfWarps][kPartBlockMax], s_mi[kPfWarps][kPartBlockMax], s_uu[kPfWarps][kPartBlockMax],
        s_wl[kPfWarps][kPartBlockMax], s_rm[kPfWarps][kPartBlockMax];
    if (gridDim.y > 1) { lk = ind_kind(blockIdx.y / 3); rk = ind_kind(blockIdx.y % 3); }
    T* facb = fac + blockIdx.y * fac_stride;
    const PartLevel L = pl.lv[l];
    const int m = pl.m, P = L.len / m, tail = L.len - P * m, nblk = P + (tail > 0 ? 1 : 0);
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * kPfWarps + wi;
    if (c >= nblk) return;                                            // whole warps leave; only warp-level barriers below
    const int first = c * m, cnt = c < P ? m - 1 : tail, rows = c < P ? m : tail;
    const PartArrays<T> a(facb, L);
    T *lo = s_lo[wi], *mi = s_mi[wi], *uu = s_uu[wi], *wl = s_wl[wi], *rm = s_rm[wi];
    for (int t = lane; t < rows; t += 32) {
        T vl, vm, vu;
        part_row<T>(x, n, periodic, lk, rk, pl, facb, l, first + t, vl, vm, vu);
        lo[t] = vl; mi[t] = vm; uu[t] = vu;
        a.low[first + t] = vl; a.mid[first + t] = vm; a.up[first + t] = vu;
        if (l == 0) part_row_extras<T>(x, n, periodic, L.len, fac, facb, first + t);
        if (t == m - 1) { a.fr[first + t] = FacRow<T>{(T)0, (T)0, (T)0, (T)0}; a.p[first + t] = (T)0; a.q[first + t] = (T)0; }
    }
    __syncwarp();
    if (lane == 0) {
        T mp = mi[0];
        wl[0] = (T)0;
        for (int t = 1; t < cnt; ++t) {
            const T w = DIV(lo[t], mp);
            mi[t - 1] = mp;                                           // mi[] becomes the eliminated diagonal
            mp = FMA(-w, uu[t - 1], mi[t]);
            wl[t] = w;
        }
        mi[cnt - 1] = mp;
    }
    __syncwarp();
    const T one = (T)1;
    for (int t = lane; t < cnt; t += 32) {
        const T r = DIV(one, mi[t]);
        rm[t] = r;
        a.fr[first + t] = FacRow<T>{uu[t], mi[t], wl[t], r};
    }
    __syncwarp();
    if (lane == 0) {                                                  // p = A^-1 (low[first] e_first); lo[] becomes its forward sweep
        T f = lo[0];
        T* fw = s_lo[wi];
        fw[0] = f;
        for (int t = 1; t < cnt; ++t) { f = -MUL(wl[t], f); fw[t] = f; }
        T v = MUL(f, rm[cnt - 1]);
        a.p[first + cnt - 1] = v;
        for (int t = cnt - 2; t >= 0; --t) { v = MUL(FMA(-uu[t], v, fw[t]), rm[t]); a.p[first + t] = v; }
    } else if (lane == 1) {                                           // q = A^-1 (up[last] e_last)
        T v = MUL(uu[cnt - 1], rm[cnt - 1]);
        a.q[first + cnt - 1] = v;
        for (int t = cnt - 2; t >= 0; --t) { v = MUL(-MUL(uu[t], v), rm[t]); a.q[first + t] = v; }
    }
}

// The last level (at most kPartTopMax rows): rows formed by all threads, the elimination by one, the reciprocals by all.
template <class T>
__global__ void __launch_bounds__(kPartTopMax) part_top_factor_kernel(const T* __restrict__ x, int n, int periodic, int lk, int rk,
                                                                     const PartPlan pl, T* fac, size_t fac_stride) {
    __shared__ T sl[kPartTopMax], sm[kPartTopMax], su[kPartTopMax];
    if (gridDim.y > 1) { lk = ind_kind(blockIdx.y / 3); rk = ind_kind(blockIdx.y % 3); }
    T* facb = fac + blockIdx.y * fac_stride;
    const int l = pl.nsplit;
    const PartLevel L = pl.lv[l];
    const PartArrays<T> a(facb, L);
    const int j = threadIdx.x;
    if (j < L.len) {
        T lo, mi, uu;
        part_row<T>(x, n, periodic, lk, rk, pl, facb, l, j, lo, mi, uu);
        sl[j] = lo; sm[j] = mi; su[j] = uu;
        a.low[j] = lo; a.mid[j] = mi; a.up[j] = uu;
        if (l == 0) part_row_extras<T>(x, n, periodic, L.len, fac, facb, j);
    }
    __syncthreads();
    if (j == 0) {
        T mp = sm[0];
        sl[0] = (T)0;
        for (int i = 1; i < L.len; ++i) {
            const T w = DIV(sl[i], mp);
            sm[i - 1] = mp;
            mp = FMA(-w, su[i - 1], sm[i]);
            sl[i] = w;
        }
        sm[L.len - 1] = mp;
    }
    __syncthreads();
    if (j < L.len) a.fr[j] = FacRow<T>{su[j], sm[j], sl[j], DIV((T)1, sm[j])};
}

// ---- per column -------------------------------------------------------------------------------------------------
// Individual boundaries: a column's matrix is one of nine (ndi_spline.cuh)
__device__ __forceinline__ int part_group(const int32_t* __restrict__ lks, const int32_t* __restrict__ rks, long long col) {
    auto var = [](int k) { return k == SB_NAK ? 0 : ((k == SB_FIRST || k == SB_CLAMPED) ? 1 : 2); };
    return 3 * var(lks[col]) + var(rks[col]);
}

enum { PART_SRC_R = 0, PART_SRC_LOWER = 1 };
// right-hand side of row j of level l for one column: read from R (PART_SRC_R), or, on level l >= 1, formed from
// the block solutions g of level l - 1 around the separator (step 3)
template <class T, int SRC>
struct PartRhs {
    const T* R; long long w, col; int stride;       // of level l
    const T *plow, *pup; int plen, pstride, m;      // level l - 1
    __device__ __forceinline__ PartRhs(const PartPlan& pl, int l, T* facb, const T* R_, long long w_, long long col_)
        : R(R_), w(w_), col(col_), stride(pl.lv[l].stride), plow(nullptr), pup(nullptr), plen(0), pstride(0), m(pl.m) {
        if (SRC == PART_SRC_LOWER) {
            const PartArrays<T> a(facb, pl.lv[l - 1]);
            plow = a.low; pup = a.up; plen = pl.lv[l - 1].len; pstride = pl.lv[l - 1].stride;
        }
    }
    __device__ __forceinline__ long long at(int j) const { return ((long long)(j + 1) * stride - 1) * w + col; }
    __device__ __forceinline__ T operator()(int j) const {
        if (SRC != PART_SRC_LOWER) return R[at(j)];
        const int s = j * m + m - 1;
        const long long rs = at(j), d = (long long)pstride * w;
        T v = FMA(-__ldg(plow + s), R[rs - d], R[rs]);
        if (s + 1 < plen) v = FMA(-__ldg(pup + s), R[rs + d], v);
        return v;
    }
};

// One thread per (block, column): g = A_block^-1 rhs, in registers, written over the right-hand sides in R.
template <class T, int SRC, int MAXB>
__global__ void __launch_bounds__(128) part_local_kernel(const PartPlan pl, int l, T* fac, size_t fac_stride, T* __restrict__ R,
                                                         long long w, const int32_t* __restrict__ lks, const int32_t* __restrict__ rks) {
    const PartLevel L = pl.lv[l];
    const int m = pl.m, P = L.len / m, tail = L.len - P * m, nblk = P + (tail > 0 ? 1 : 0);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c64 = gid / w, col = gid - c64 * w;
    if (c64 >= nblk) return;
    const int c = (int)c64;
    T* facb = fac + (lks ? (size_t)part_group(lks, rks, col) * fac_stride : 0);
    const int first = c * m, cnt = c < P ? m - 1 : tail;
    const PartRhs<T, SRC> rhs(pl, l, facb, R, w, col);
    const FacRow<T>* fr = PartArrays<T>(facb, L).fr + first;
    T v[MAXB - 1];
#pragma unroll
    for (int t = 0; t < MAXB - 1; ++t) v[t] = t < cnt ? rhs(first + t) : (T)0;
    if (SRC == PART_SRC_LOWER && c < P) R[rhs.at(first + m - 1)] = rhs(first + m - 1);   // the separator's own right-hand side, for the level above
    // forward (:698) and backward (:704-720) recurrences of the block; the factors are warp-uniform loads out of L1
#pragma unroll
    for (int t = 1; t < MAXB - 1; ++t)
        if (t < cnt) v[t] = FMA(-ld_fac<T>(fr + t).wl, v[t - 1], v[t]);
    T nxt = (T)0;
#pragma unroll
    for (int t = MAXB - 2; t >= 0; --t) {
        if (t < cnt) {
            const FacRow<T> f = ld_fac<T>(fr + t);
            const T val = MUL(t == cnt - 1 ? v[t] : FMA(-f.up, nxt, v[t]), f.rmid);
            R[rhs.at(first + t)] = val;
            nxt = val;
        }
    }
}

// The last level: a block takes kTopCols columns (few -- kTopColsNarrow -- so that many SMs take part: the kernel is a chain of 2 len
// dependent steps per column whatever the block does); all its threads form the right-hand sides into shared memory,
// one lane per column runs the two recurrences there (factors staged in shared memory unless the columns have
// matrices of their own), all threads write k back.
// (With enough columns to fill the machine at 32 per block -- kTopColsWide -- the wider block stages the factors a
// quarter as often and needs a quarter of the blocks: 4096 x 16384 f32 build 0.544 -> 0.530 ms, 512 x 262144 1.126 ->
// 1.066 ms; profiles/r02/partition_build.md.)
constexpr int kTopColsNarrow = 8, kTopColsWide = 32, kTopThreads = 256;
template <class T, int kTopCols>
__global__ void __launch_bounds__(kTopThreads) part_top_kernel(const PartPlan pl, T* fac, size_t fac_stride, T* __restrict__ R, long long w,
                                                               const int32_t* __restrict__ lks, const int32_t* __restrict__ rks) {
    constexpr int kTopRowLanes = kTopThreads / kTopCols;
    __shared__ T tile[kPartTopMax][kTopCols + 1];
    __shared__ T s_wl[kPartTopMax], s_up[kPartTopMax], s_rm[kPartTopMax];
    const int l = pl.nsplit;
    const PartLevel L = pl.lv[l];
    const int cx = threadIdx.x % kTopCols, ry = threadIdx.x / kTopCols;
    const long long col = (long long)blockIdx.x * kTopCols + cx;
    const bool live = col < w;
    T* facb = fac + ((lks && live) ? (size_t)part_group(lks, rks, col) * fac_stride : 0);
    const FacRow<T>* fr = PartArrays<T>(facb, L).fr;
    if (!lks)
        for (int j = threadIdx.x; j < L.len; j += kTopThreads) { const FacRow<T> f = ld_fac<T>(fr + j); s_wl[j] = f.wl; s_up[j] = f.up; s_rm[j] = f.rmid; }
    if (live) {
        if (l == 0) { const PartRhs<T, PART_SRC_R> rhs(pl, l, facb, R, w, col); for (int j = ry; j < L.len; j += kTopRowLanes) tile[j][cx] = rhs(j); }
        else { const PartRhs<T, PART_SRC_LOWER> rhs(pl, l, facb, R, w, col); for (int j = ry; j < L.len; j += kTopRowLanes) tile[j][cx] = rhs(j); }
    }
    __syncthreads();
    if (ry == 0 && live) {
        const int len = L.len;
        if (!lks) {
            T prev = tile[0][cx];
#pragma unroll 4
            for (int i = 1; i < len; ++i) { prev = FMA(-s_wl[i], prev, tile[i][cx]); tile[i][cx] = prev; }
            T k = MUL(prev, s_rm[len - 1]);
            tile[len - 1][cx] = k;
#pragma unroll 4
            for (int i = len - 2; i >= 0; --i) { k = MUL(FMA(-s_up[i], k, tile[i][cx]), s_rm[i]); tile[i][cx] = k; }
        } else {
            T prev = tile[0][cx];
            for (int i = 1; i < len; ++i) { prev = FMA(-ld_fac<T>(fr + i).wl, prev, tile[i][cx]); tile[i][cx] = prev; }
            T k = MUL(prev, ld_fac<T>(fr + len - 1).rmid);
            tile[len - 1][cx] = k;
            for (int i = len - 2; i >= 0; --i) {
                const FacRow<T> f = ld_fac<T>(fr + i);
                k = MUL(FMA(-f.up, k, tile[i][cx]), f.rmid);
                tile[i][cx] = k;
            }
        }
    }
    __syncthreads();
    if (live) {
        const PartRhs<T, PART_SRC_R> at(pl, l, facb, R, w, col);
        for (int j = ry; j < L.len; j += kTopRowLanes) R[at.at(j)] = tile[j][cx];
    }
}

// k = g - p k_left - q k_right on the block rows of level l (separator rows hold their final k already)
template <class T>
__global__ void __launch_bounds__(256) part_corr_kernel(const PartPlan pl, int l, T* fac, size_t fac_stride, T* __restrict__ R, long long w,
                                                        const int32_t* __restrict__ lks, const int32_t* __restrict__ rks) {
    const PartLevel L = pl.lv[l];
    const int m = pl.m, P = L.len / m;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long j64 = gid / w, col = gid - j64 * w;
    if (j64 >= L.len) return;
    const int j = (int)j64, c = j / m;
    if (c < P && j - c * m == m - 1) return;
    T* facb = fac + (lks ? (size_t)part_group(lks, rks, col) * fac_stride : 0);
    const PartArrays<T> a(facb, L);
    const PartRhs<T, PART_SRC_R> at(pl, l, facb, R, w, col);
    const T kl = c > 0 ? R[at.at(c * m - 1)] : (T)0, kr = c < P ? R[at.at(c * m + m - 1)] : (T)0;
    const long long o = at.at(j);
    R[o] = FMA(-__ldg(a.q + j), kr, FMA(-__ldg(a.p + j), kl, R[o]));
}

// Level 0, non-periodic: the correction of step 4 and a[i] = k[i] dx - dy, b[i] = dy - k[i+1] dx (calc_coefficients
// :354-365) in one pass; k is formed in registers only.  A thread takes kRowGroup consecutive intervals of one column (the
// item layout of spline_ab_kernel) and the k of their kRowGroup + 1 rows; the separator rows it needs are read by every
// thread of the block's eight row groups, out of L2.
template <class T, int V>
__global__ void __launch_bounds__(256) part_ab_kernel(const PartPlan pl, T* fac, size_t fac_stride, const T* __restrict__ R,
                                                      const T* __restrict__ y, int n, long long w, T* __restrict__ a, T* __restrict__ b,
                                                      const int32_t* __restrict__ lks, const int32_t* __restrict__ rks) {
    const PartLevel L = pl.lv[0];
    const int m = pl.m, P = L.len / m;
    const T* dx = fac;
    const long long wv = w / V, ntasks = row_task_count(wv, n - 1, blockDim.x);
    for (long long task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const RowTask t = row_task(task, wv, n - 1);
        if (!t.live) continue;
        const long long col = t.col * V;
        const int row = t.row;
        T* facb = fac + (lks ? (size_t)part_group(lks, rks, col) * fac_stride : 0);     // V == 1 with Individual boundaries
        const PartArrays<T> pa(facb, L);
        T kv[kRowGroup + 1][V], yv[kRowGroup + 1][V];
#pragma unroll
        for (int j = 0; j <= kRowGroup; ++j) {
            const int i = min(row + j, n - 1);
            ld_vec<T, V>(R + (long long)i * w + col, kv[j]);
            ld_vec<T, V>(y + (long long)i * w + col, yv[j]);
        }
        // the rows of a group lie in at most two blocks: separators of the first row's block, and the one after them
        const int c0 = row / m;
        T s0[V], s1[V], s2[V];
#pragma unroll
        for (int c = 0; c < V; ++c) s0[c] = s1[c] = s2[c] = (T)0;
        if (c0 > 0) ld_vec<T, V>(R + (long long)(c0 * m - 1) * w + col, s0);
        if (c0 < P) ld_vec<T, V>(R + (long long)(c0 * m + m - 1) * w + col, s1);
        if (c0 + 1 < P) ld_vec<T, V>(R + (long long)(c0 * m + 2 * m - 1) * w + col, s2);
#pragma unroll
        for (int j = 0; j <= kRowGroup; ++j) {
            const int i = min(row + j, n - 1);
            const int cb = i / m;
            if (!(cb < P && i - cb * m == m - 1)) {
                const T pi = __ldg(pa.p + i), qi = __ldg(pa.q + i);
#pragma unroll
                for (int c = 0; c < V; ++c) {
                    const T kl = cb == c0 ? s0[c] : s1[c], kr = cb == c0 ? s1[c] : s2[c];
                    kv[j][c] = FMA(-qi, kr, FMA(-pi, kl, kv[j][c]));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kRowGroup; ++j) {
            const int i = row + j;
            if (i >= n - 1) break;
            const T d = __ldg(dx + i);
            T av[V], bv[V];
#pragma unroll
            for (int c = 0; c < V; ++c) {
                const T dy = SUB(yv[j + 1][c], yv[j][c]);
                av[c] = SUB(MUL(kv[j][c], d), dy);
                bv[c] = SUB(dy, MUL(kv[j + 1][c], d));
            }
            st_vec<T, V>(a + (long long)i * w + col, av);
            st_vec<T, V>(b + (long long)i * w + col, bv);
        }
    }
}

template <class T, int SRC>
static void part_launch_local(const PartPlan& pl, int l, T* fac, size_t fac_stride, T* R, long long w, const int32_t* lk,
                              const int32_t* rk, unsigned blocks, cudaStream_t st) {
    if (pl.m <= 32) part_local_kernel<T, SRC, 32><<<blocks, 128, 0, st>>>(pl, l, fac, fac_stride, R, w, lk, rk);
    else part_local_kernel<T, SRC, kPartBlockMax><<<blocks, 128, 0, st>>>(pl, l, fac, fac_stride, R, w, lk, rk);
    count_launch();
}

// the solve on R, whose level-0 rows hold the right-hand sides.
// final0: 0 leaves level 0 uncorrected (g in the block rows, k in the separator rows) for part_ab_kernel.
// ev0 / join: waited for before the first kernel that needs the factorisation of level 0 / of the levels above it.
template <class T>
static cudaError_t part_solve(const PartPlan& pl, T* fac, size_t fac_stride, T* R, long long w, const int32_t* lk, const int32_t* rk,
                              bool final0, cudaEvent_t ev0, cudaEvent_t join, cudaStream_t st) {
    cudaError_t e;
    if (ev0 && (e = cudaStreamWaitEvent(st, ev0, 0)) != cudaSuccess) return e;
    for (int l = 0; l < pl.nsplit; ++l) {
        const PartLevel& L = pl.lv[l];
        const long long nblk = L.len / pl.m + (L.len % pl.m ? 1 : 0);
        const long long blocks = (nblk * w + 127) / 128;
        if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
        if (l == 1 && join && (e = cudaStreamWaitEvent(st, join, 0)) != cudaSuccess) return e;
        if (l == 0) part_launch_local<T, PART_SRC_R>(pl, l, fac, fac_stride, R, w, lk, rk, (unsigned)blocks, st);
        else part_launch_local<T, PART_SRC_LOWER>(pl, l, fac, fac_stride, R, w, lk, rk, (unsigned)blocks, st);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (pl.nsplit <= 1 && join && (e = cudaStreamWaitEvent(st, join, 0)) != cudaSuccess) return e;
    if (w >= (long long)kTopColsWide * device_info().sm_count)
        part_top_kernel<T, kTopColsWide><<<(unsigned)((w + kTopColsWide - 1) / kTopColsWide), kTopThreads, 0, st>>>(pl, fac, fac_stride, R, w, lk, rk);
    else
        part_top_kernel<T, kTopColsNarrow><<<(unsigned)((w + kTopColsNarrow - 1) / kTopColsNarrow), kTopThreads, 0, st>>>(pl, fac, fac_stride, R, w, lk, rk);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    for (int l = pl.nsplit - 1; l >= (final0 ? 0 : 1); --l) {
        const long long blocks = ((long long)pl.lv[l].len * w + 255) / 256;
        if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
        part_corr_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(pl, l, fac, fac_stride, R, w, lk, rk);
        count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// second stream of the calling thread on the current device (matrix work of the upper levels beside the block solves)
struct PartSide { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, lvl0 = nullptr, join = nullptr; };
static PartSide& part_side() {
    static thread_local std::map<int, PartSide> per_dev;
    int dev = 0;
    cudaGetDevice(&dev);
    PartSide& sd = per_dev[dev];
    if (!sd.s) {
        if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sd.lvl0, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) != cudaSuccess) { sd.s = nullptr; cudaGetLastError(); }
    }
    return sd;
}

// n >= 4.  scratch: ngroups * partition_fac_elems(n, block) elements of factorisations, then R (n x w).
template <class T>
cudaError_t launch_partition_build(const T* x, int64_t n, const T* data, int64_t w, int bc_kind, int block, const int32_t* lk,
                                   const T* lv, const int32_t* rk, const T* rv, T* a, T* b, T* scratch, unsigned long long* err,
                                   cudaStream_t st) {
    const int periodic = bc_kind == BC_PERIODIC;
    const bool individual = bc_kind == BC_INDIVIDUAL;
    Side<T> l{SB_NAK, (T)0}, r{SB_NAK, (T)0};
    if (bc_kind == BC_NATURAL) l = r = Side<T>{SB_NATURAL, (T)0};
    if (bc_kind == BC_CLAMPED) l = r = Side<T>{SB_CLAMPED, (T)0};
    const Side<T> ls = specialize(l), rs = specialize(r);
    const int64_t len = periodic ? n - 2 : n;
    const PartPlan pl = part_plan(n, len, block);
    if (pl.lv[pl.nsplit].len > kPartTopMax) return cudaErrorInvalidConfiguration;
    const size_t fac_stride = partition_fac_elems(n, block);
    const int ngroups = individual ? 9 : 1;
    T* fac = scratch;
    T* R = scratch + (size_t)ngroups * fac_stride;
    const int32_t* ilk = individual ? lk : nullptr;
    cudaError_t e;
    auto factor = [&](int lvl, cudaStream_t s) -> cudaError_t {
        const int nblk = pl.lv[lvl].len / pl.m + (pl.lv[lvl].len % pl.m ? 1 : 0);
        part_factor_kernel<T><<<dim3((unsigned)((nblk + kPfWarps - 1) / kPfWarps), ngroups), 32 * kPfWarps, 0, s>>>(
            x, (int)n, periodic, ls.kind, rs.kind, pl, lvl, fac, fac_stride);
        count_launch();
        return cudaGetLastError();
    };
    auto top_factor = [&](cudaStream_t s) -> cudaError_t {
        part_top_factor_kernel<T><<<dim3(1, ngroups), kPartTopMax, 0, s>>>(x, (int)n, periodic, ls.kind, rs.kind, pl, fac, fac_stride);
        count_launch();
        return cudaGetLastError();
    };
    // All matrix work (it depends on x only) -- and, periodic, the shared second solution k2 (:535-550: one more column
    // through the same solve, in place at fac + 4n) -- runs on a second stream of this thread beside the right-hand sides.
    PartSide& side = part_side();
    const bool forked = side.s && cudaEventRecord(side.fork, st) == cudaSuccess && cudaStreamWaitEvent(side.s, side.fork, 0) == cudaSuccess;
    cudaStream_t ms = forked ? side.s : st;
    for (int lvl = 0; lvl < pl.nsplit; ++lvl) {
        if ((e = factor(lvl, ms)) != cudaSuccess) return e;
        if (lvl == 0 && forked && (e = cudaEventRecord(side.lvl0, ms)) != cudaSuccess) return e;
    }
    if ((e = top_factor(ms)) != cudaSuccess) return e;
    if (periodic && (e = part_solve<T>(pl, fac, fac_stride, fac + 4 * (size_t)n, 1, nullptr, nullptr, true, nullptr, nullptr, ms)) != cudaSuccess) return e;
    if (forked && (e = cudaEventRecord(side.join, ms)) != cudaSuccess) return e;
    if ((e = launch_spline_rhs<T>(x, (int)n, data, (long long)w, periodic, l, r, R, err, ilk, lv, rk, rv, st)) != cudaSuccess) return e;
    const bool fused_ab = !periodic && pl.nsplit > 0 && pl.m >= kRowGroup + 1;   // part_ab_kernel: a row group (kRowGroup + 1 rows from a multiple of kRowGroup) lies in at most two blocks
    if ((e = part_solve<T>(pl, fac, fac_stride, R, (long long)w, ilk, rk, !fused_ab, (forked && pl.nsplit > 0) ? side.lvl0 : nullptr,
                           forked ? side.join : nullptr, st)) != cudaSuccess) return e;
    if (!fused_ab) {
        if (periodic && (e = launch_spline_periodic_close<T>(x, (int)n, (long long)w, fac, R, st)) != cudaSuccess) return e;
        return launch_spline_ab<T>(x, (int)n, data, (long long)w, periodic, fac, R, a, b, nullptr, st);
    }
    constexpr int V = 16 / sizeof(T);
    if (!ilk && vec_ok<T>(w, data, R, a, b))
        part_ab_kernel<T, V><<<row_group_grid(w / V, n - 1), 256, 0, st>>>(pl, fac, fac_stride, R, data, (int)n, (long long)w, a, b, nullptr, nullptr);
    else
        part_ab_kernel<T, 1><<<row_group_grid(w, n - 1), 256, 0, st>>>(pl, fac, fac_stride, R, data, (int)n, (long long)w, a, b, ilk, rk);
    count_launch();
    return cudaGetLastError();
}

template cudaError_t launch_partition_build<float>(const float*, int64_t, const float*, int64_t, int, int, const int32_t*, const float*,
                                                   const int32_t*, const float*, float*, float*, float*, unsigned long long*, cudaStream_t);
template cudaError_t launch_partition_build<double>(const double*, int64_t, const double*, int64_t, int, int, const int32_t*, const double*,
                                                    const int32_t*, const double*, double*, double*, double*, unsigned long long*, cudaStream_t);

}  // namespace ndi
