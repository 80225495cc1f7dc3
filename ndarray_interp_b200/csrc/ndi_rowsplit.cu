// ndi_rowsplit.cu -- K6, second build mode: parallel cyclic reduction + Thomas ("row-split").
//
// The reference solves its tridiagonal system with one serial Thomas sweep per column
// (cubic_spline.rs:678-721 under solve_for_k, :409-674).  On few long columns (BASELINE C2: 4096 rows x
// 1024 columns) that is 2 x 4096 dependent steps per column and a 4096-step division chain for the matrix:
// latency, not bandwidth (profiles/r01: 0.35 ms + 0.49 ms for a build that moves 100 MB).  This mode
// shortens every chain by 2^L:
//
//   1. L steps of parallel cyclic reduction (PCR).  Step with stride s replaces row i by
//          row[i] + alpha[i] row[i-s] + gamma[i] row[i+s],   alpha = -(low[i] / mid[i-s]),  gamma = -(up[i] / mid[i+s])
//      which couples row i to rows i +- 2s only.  After L steps the system has fallen apart into S = 2^L
//      independent tridiagonal systems (rows j, j+S, j+2S, ... for j < S).  The matrix depends only on x, so
//      alpha / gamma and the reduced matrix are formed ONCE (rowsplit_level_kernel, one launch per step); per column only
//          rhs'[i] = (rhs[i] + alpha[i] rhs[i-s]) + gamma[i] rhs[i+s]
//      remains, an elementwise pass.  All L passes run on a tile held in shared memory
//      (rowsplit_reduce_kernel: rows of the tile plus a halo of 2^L - 1 rows on either side, right-hand sides
//      formed in the same kernel from y), so the matrix R is written once.
//   2. Thomas on the S interleaved systems: S short division chains for the matrix (one thread each,
//      rowsplit_chain_kernel) and S * w chains of n / S rows for the columns (spline_sweep_kernel of ndi_spline.cu
//      with nsys = S).
//   3. a, b from k: the same kernel as the reference-order build.
//
// The rounding differs from the reference's elimination order, so this is NOT bit-identical to the reference
// arithmetic; it is held bit for bit to the oracle's operation-by-operation specification of the same scheme
// (the CPU checker's rowsplit_thomas, see tests/test_rowsplit_gpu.py) and to north_star's 1e-12 (f64) / 1e-5 (f32) bars against the
// reference-order checker.  Every product and sum is rounded on its own (no FMA),
// the i-s term before the i+s term, exactly as the specification writes them.
#include <map>

#include "ndi_spline.cuh"

namespace ndi {

// fac layout per matrix (elements of T):
//   [0, 4n)             FacRow per row: up, eliminated mid, elimination weight, 1/mid   (as the reference-order build)
//   [4n, 5n)            k2, the shared second solution of the periodic system
//   [5n, 5n + 2Ln)      {alpha, gamma} per (level, row)
//   [.., .. + 8n)       two sets of (low, mid, up, rhs2) for the level-by-level reduction of the matrix
size_t rowsplit_fac_elems(int64_t n, int levels) { return ((size_t)(13 + 2 * levels) * (size_t)n + 3) & ~(size_t)3; }

int rowsplit_levels_for(int64_t rows, int requested, bool force) {
    int most = 0;
    while (most < kMaxRowsplitLevels && (rows >> (most + 1)) >= 2) ++most;      // every system keeps >= 2 rows
    if (requested > 0) return requested < most ? requested : most;
    if (most == 0 || (!force && rows < kRowsplitAutoRows)) return 0;             // short chains: keep the reference's order
    int lv = 1;
    while (lv < most && (rows >> lv) > 256) ++lv;
    return lv;
}

constexpr int kRsBlock = 256;
constexpr int kRsChain = 8;              // rows a chain thread loads ahead of its dependent steps

// The matrix side of the reduction runs level by level; a level is elementwise over the rows (row i reads rows
// i - s, i, i + s of the previous level), so each level is one launch over all rows -- with a single block looping
// over the levels a 65536-row system spent 0.4 ms here.  blockIdx.y selects the matrix (Individual: nine).
template <class T>
struct RsFac {
    T* fac; size_t fac_stride; int n, len, levels, periodic, lk, rk;
    __device__ __forceinline__ T* base() const { return fac + blockIdx.y * fac_stride; }
    __device__ __forceinline__ T* set(int which) const { return base() + (5 + 2 * (size_t)levels) * (size_t)n + (size_t)which * 4 * (size_t)n; }
};

// level 0: the matrix of solve_for_k (:440-451, boundary rows :599-669; periodic: the condensed system :512-518)
template <class T>
__global__ void __launch_bounds__(kRsBlock) rowsplit_matrix_kernel(const T* __restrict__ x, const RsFac<T> f) {
    const int i = blockIdx.x * kRsBlock + threadIdx.x;
    if (i >= f.len) return;
    const size_t N = (size_t)f.n;
    const int n = f.n;
    int lk = f.lk, rk = f.rk;
    if (gridDim.y > 1) { lk = ind_kind(blockIdx.y / 3); rk = ind_kind(blockIdx.y % 3); }
    T* cur = f.set(0);
    T u, m, l;
    if (f.periodic) matrix_row_periodic<T>(x, n, i, u, m, l); else matrix_row<T>(x, n, i, lk, rk, false, u, m, l);
    cur[i] = l; cur[N + i] = m; cur[2 * N + i] = u;
    if (f.periodic) {                                                 // rhs2 (:535-538)
        const T dx0 = SUB(x[1], x[0]), dx_3 = SUB(x[n - 3], x[n - 4]);
        cur[3 * N + i] = i == 0 ? -dx0 : (i == f.len - 1 ? -dx_3 : (T)0);
    }
}

// one reduction level with stride s = 2^lv: set (lv & 1) -> set ((lv + 1) & 1), and {alpha, gamma} of the level
template <class T>
__global__ void __launch_bounds__(kRsBlock) rowsplit_level_kernel(const RsFac<T> f, int lv) {
    const int i = blockIdx.x * kRsBlock + threadIdx.x;
    if (i >= f.len) return;
    const size_t N = (size_t)f.n;
    const int s = 1 << lv, len = f.len;
    const T* cur = f.set(lv & 1);
    T* nxt = f.set((lv + 1) & 1);
    const T *low = cur, *mid = cur + N, *up = cur + 2 * N, *r2 = cur + 3 * N;
    T* cf = f.base() + 5 * N + 2 * (size_t)lv * N;
    const bool hm = i - s >= 0, hp = i + s <= len - 1;
    const T alpha = hm ? -DIV(low[i], mid[i - s]) : (T)0;
    const T gamma = hp ? -DIV(up[i], mid[i + s]) : (T)0;
    cf[2 * (size_t)i] = alpha; cf[2 * (size_t)i + 1] = gamma;
    nxt[i] = hm ? MUL(alpha, low[i - s]) : (T)0;
    nxt[2 * N + i] = hp ? MUL(gamma, up[i + s]) : (T)0;
    T m = mid[i];
    if (hm) m = ADD(m, MUL(alpha, up[i - s]));
    if (hp) m = ADD(m, MUL(gamma, low[i + s]));
    nxt[N + i] = m;
    if (f.periodic) {
        T v = r2[i];
        if (hm) v = ADD(v, MUL(alpha, r2[i - s]));
        if (hp) v = ADD(v, MUL(gamma, r2[i + s]));
        nxt[3 * N + i] = v;
    }
}

// Thomas elimination (:690-692) of the S = 2^levels interleaved systems, one thread each; rows of a system are S
// apart.  Periodic: the shared second solution k2 as well (forward :698, back substitution :704-720).
template <class T>
__global__ void __launch_bounds__(64) rowsplit_chain_kernel(const RsFac<T> f) {
    const int S = 1 << f.levels, len = f.len, periodic = f.periodic;
    const int j = blockIdx.x * 64 + threadIdx.x;
    if (j >= S || j >= len) return;
    const size_t N = (size_t)f.n;
    FacRow<T>* rows = reinterpret_cast<FacRow<T>*>(f.base());
    T* k2 = f.base() + 4 * N;
    const T* cur = f.set(f.levels & 1);
    const T *low = cur, *mid = cur + N, *up = cur + 2 * N, *r2 = cur + 3 * N;
    const int m_rows = (len - j + S - 1) / S;
    T m_prev = mid[j], u_prev = up[j], r_prev = periodic ? r2[j] : (T)0;
    rows[j] = FacRow<T>{u_prev, m_prev, (T)0, Hoisted<T>::rcp(m_prev)};
    if (periodic) k2[j] = r_prev;
    for (int t0 = 1; t0 < m_rows; t0 += kRsChain) {
        T lo[kRsChain], mi[kRsChain], uu[kRsChain], rr[kRsChain];
#pragma unroll
        for (int q = 0; q < kRsChain; ++q) {                          // loads first: they do not depend on the chain
            const int i = j + min(t0 + q, m_rows - 1) * S;
            lo[q] = low[i]; mi[q] = mid[i]; uu[q] = up[i]; rr[q] = periodic ? r2[i] : (T)0;
        }
#pragma unroll
        for (int q = 0; q < kRsChain; ++q) {
            if (t0 + q < m_rows) {
                const int i = j + (t0 + q) * S;
                const T wgt = DIV(lo[q], m_prev);
                const T mm = SUB(mi[q], MUL(wgt, u_prev));
                rows[i] = FacRow<T>{uu[q], mm, wgt, Hoisted<T>::rcp(mm)};
                if (periodic) { r_prev = SUB(rr[q], MUL(wgt, r_prev)); k2[i] = r_prev; }
                m_prev = mm; u_prev = uu[q];
            }
        }
    }
    if (periodic) {
        T kr = DIV(r_prev, m_prev);
        k2[j + (m_rows - 1) * S] = kr;
        for (int t = m_rows - 2; t >= 0; --t) {
            const int i = j + t * S;
            kr = DIV(SUB(k2[i], MUL(rows[i].up, kr)), rows[i].mid);
            k2[i] = kr;
        }
    }
}

// Right-hand sides + all reduction levels for one tile of rows x 32 columns, in shared memory.
// Buffer row r holds global row i0 + r, i0 = tile start - H, H = 2^levels - 1.  A row's value after the last level
// depends on rows at most H away, so rows [H, H + rt) of the buffer come out exact and are the ones written;
// halo rows are computed as far as their inputs lie inside the buffer and never leave the block.
// Non-periodic: R[0 .. n-1] = reduced right-hand sides.  Periodic: R[0 .. n-3] = reduced right-hand sides of the
// condensed system, R[n-2] = the last condensed equation's right-hand side (:531-532), kept for k_m1.
constexpr int kRsLoadAhead = 8;          // rows of y a warp requests before it waits for any of them
template <class T>
__global__ void __launch_bounds__(512) rowsplit_reduce_kernel(const T* __restrict__ x, int n, const T* __restrict__ y, long long w,
                                                              int periodic, Side<T> left, Side<T> right, int levels, int rt,
                                                              int col_tiles, const T* __restrict__ fac, size_t fac_stride,
                                                              T* __restrict__ R, unsigned long long* err,
                                                              const int32_t* __restrict__ lks, const T* __restrict__ lvs,
                                                              const int32_t* __restrict__ rks, const T* __restrict__ rvs,
                                                              const int32_t* __restrict__ pos) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int H = (1 << levels) - 1;
    const int nb = rt + 2 * H;
    T* bufA = reinterpret_cast<T*>(rs_smem);
    T* bufB = bufA + (size_t)(nb + 2) * 32;
    const int slen = periodic ? n - 2 : n;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ty = blockIdx.x / col_tiles, tx = blockIdx.x - ty * col_tiles;
    const long long c = (long long)tx * 32 + lane;
    const bool colok = c < w;
    const int i0 = ty * rt - H;
    const T three = (T)3;
    Side<T> l = specialize(left), r = specialize(right);
    size_t fo = 0;
    long long cout = c;
    if (lks && colok) {                                               // Individual: this column's own boundary rows and matrix
        l = specialize(Side<T>{lks[c], lvs[c]}); r = specialize(Side<T>{rks[c], rvs[c]});
        fo = (size_t)(3 * ind_variant(l.kind) + ind_variant(r.kind)) * fac_stride;
        cout = pos[c];
    }
    const T* coef = fac + fo + 5 * (size_t)n;
    const T* ycol = y + (colok ? c : 0);
    auto Y = [&](int row) -> T { return __ldg(ycol + (long long)row * w); };
    // y rows [i0 - 1, i0 + nb + 1) -> bufB; kRsLoadAhead independent loads in flight per thread (a block that waits
    // for one row at a time spends its life in DRAM latency: 0.55 ms for the 65536 x 64 build, profiles/r02)
    for (int r0 = wi; r0 < nb + 2; r0 += nw * kRsLoadAhead) {
        T v[kRsLoadAhead];
#pragma unroll
        for (int k = 0; k < kRsLoadAhead; ++k) {
            const int rr = r0 + k * nw, gi = i0 - 1 + rr;
            v[k] = (rr < nb + 2 && gi >= 0 && gi < n && colok) ? Y(gi) : (T)0;
        }
#pragma unroll
        for (int k = 0; k < kRsLoadAhead; ++k) {
            const int rr = r0 + k * nw;
            if (rr < nb + 2) bufB[rr * 32 + lane] = v[k];
        }
    }
    __syncthreads();
    // right-hand sides (:456-471 interior, :599-669 boundary rows, :529-530 periodic first row) -> bufA
#pragma unroll 2
    for (int rr = wi; rr < nb; rr += nw) {
        const int i = i0 + rr;
        T v = (T)0;
        if (i >= 0 && i < slen && colok) {
            const bool interior = periodic ? i > 0 : (i > 0 && i < n - 1);
            if (interior) {
                const T xm = __ldg(x + i - 1), xi = __ldg(x + i), xp = __ldg(x + i + 1);
                v = rhs_interior<T>(bufB[rr * 32 + lane], bufB[(rr + 1) * 32 + lane], bufB[(rr + 2) * 32 + lane], SUB(xp, xi), SUB(xi, xm));
            } else if (periodic) {
                const T dx0 = SUB(x[1], x[0]), dx_1 = SUB(x[n - 1], x[n - 2]);
                const T y0 = Y(0), yN = Y(n - 1);
                if (y0 != yN) atomicMin(err, (unsigned long long)c);                  // :499-507
                const T slope0 = DIV(SUB(Y(1), y0), dx0);                             // :521
                const T slope_1 = DIV(SUB(yN, Y(n - 2)), dx_1);                       // :526
                v = MUL(ADD(MUL(slope_1, dx0), MUL(slope0, dx_1)), three);            // :529-530
            } else if (i == 0) {
                v = rhs_left<T>(x, l, Y(0), Y(1), Y(2));
            } else {
                v = rhs_right<T>(x, n, r, Y(n - 1), Y(n - 2), Y(n - 3));
            }
        }
        bufA[rr * 32 + lane] = v;
    }
    __syncthreads();
    T *cur = bufA, *nxt = bufB;
    for (int lv = 0, s = 1; lv < levels; ++lv, s <<= 1) {
        const T* cf = coef + 2 * (size_t)lv * (size_t)n;
#pragma unroll 4
        for (int rr = wi; rr < nb; rr += nw) {
            const int i = i0 + rr;
            T v = cur[rr * 32 + lane];
            if (i >= 0 && i < slen) {
                const T alpha = __ldg(cf + 2 * (size_t)i), gamma = __ldg(cf + 2 * (size_t)i + 1);
                if (i - s >= 0 && rr - s >= 0) v = ADD(v, MUL(alpha, cur[(rr - s) * 32 + lane]));
                if (i + s <= slen - 1 && rr + s < nb) v = ADD(v, MUL(gamma, cur[(rr + s) * 32 + lane]));
            }
            nxt[rr * 32 + lane] = v;
        }
        __syncthreads();
        T* t = cur; cur = nxt; nxt = t;
    }
    for (int rr = H + wi; rr < H + rt; rr += nw) {
        const int i = i0 + rr;
        if (i < slen && colok) R[(long long)i * w + cout] = cur[rr * 32 + lane];
    }
    if (periodic && wi == 0 && colok && n - 2 >= ty * rt && n - 2 < ty * rt + rt) {
        const T dx_1 = SUB(x[n - 1], x[n - 2]), dx_2 = SUB(x[n - 2], x[n - 3]);
        const T yn2 = Y(n - 2);
        const T slope_1 = DIV(SUB(Y(n - 1), yn2), dx_1), slope_2 = DIV(SUB(yn2, Y(n - 3)), dx_2);   // :526-527
        R[(long long)(n - 2) * w + cout] = MUL(ADD(MUL(slope_2, dx_1), MUL(slope_1, dx_2)), three);   // :531-532
    }
}

// second stream of the calling thread on the current device (chains beside the reduction)
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream& side_stream() {
    static thread_local std::map<int, SideStream> per_dev;
    int dev = 0;
    cudaGetDevice(&dev);
    SideStream& sd = per_dev[dev];
    if (!sd.s) {
        if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) != cudaSuccess) { sd.s = nullptr; cudaGetLastError(); }
    }
    return sd;
}

static int rowsplit_tile_rows(int levels, size_t elem) {
    const int H = (1 << levels) - 1;
    int rt = 4 << levels;
    if (rt < 64) rt = 64;
    while (rt > 8 && 2 * (size_t)(rt + 2 * H + 2) * 32 * elem > (size_t)200 * 1024) rt >>= 1;
    return rt;
}

template <class T>
cudaError_t launch_rowsplit_front(const T* x, int64_t n, const T* data, int64_t w, int bc_kind, int levels, const int32_t* lk,
                                  const T* lv, const int32_t* rk, const T* rv, const int32_t* pos, T* fac, size_t fac_stride,
                                  T* R, unsigned long long* err, cudaStream_t st) {
    const int periodic = bc_kind == BC_PERIODIC;
    const bool individual = bc_kind == BC_INDIVIDUAL;
    Side<T> l{SB_NAK, (T)0}, r{SB_NAK, (T)0};
    if (bc_kind == BC_NATURAL) l = r = Side<T>{SB_NATURAL, (T)0};
    if (bc_kind == BC_CLAMPED) l = r = Side<T>{SB_CLAMPED, (T)0};
    const Side<T> ls = specialize(l), rs = specialize(r);
    const int len = periodic ? (int)n - 2 : (int)n;
    const RsFac<T> f{fac, fac_stride, (int)n, len, levels, periodic, ls.kind, rs.kind};
    const dim3 rows_grid((unsigned)((len + kRsBlock - 1) / kRsBlock), individual ? 9 : 1);
    rowsplit_matrix_kernel<T><<<rows_grid, kRsBlock, 0, st>>>(x, f);
    count_launch();
    for (int lv = 0; lv < levels; ++lv) {
        rowsplit_level_kernel<T><<<rows_grid, kRsBlock, 0, st>>>(f, lv);
        count_launch();
    }
    // The division chains of the matrix (a few threads, pure latency) and the reduction of the right-hand sides (the
    // whole machine, bandwidth) need nothing from each other -- both follow the level kernels, the sweeps follow
    // both -- so the chains run beside the reduction on a second stream of this thread.
    SideStream& side = side_stream();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const bool forked = side.s && cudaEventRecord(side.fork, st) == cudaSuccess && cudaStreamWaitEvent(side.s, side.fork, 0) == cudaSuccess;
    rowsplit_chain_kernel<T><<<dim3((unsigned)(((1 << levels) + 63) / 64), individual ? 9 : 1), 64, 0, forked ? side.s : st>>>(f);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (forked && (e = cudaEventRecord(side.join, side.s)) != cudaSuccess) return e;
    const int rt = rowsplit_tile_rows(levels, sizeof(T));
    const int H = (1 << levels) - 1;
    const size_t smem = 2 * (size_t)(rt + 2 * H + 2) * 32 * sizeof(T);
    if (smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(rowsplit_reduce_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return e;
    const long long rows_total = periodic ? n - 1 : n;
    const long long col_tiles = (w + 31) / 32, row_tiles = (rows_total + rt - 1) / rt;
    if (col_tiles * row_tiles > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    // a tile that leaves room for one block per SM only gets sixteen warps to hide its latencies with
    const int threads = smem > 100 * 1024 ? 512 : 256;
    rowsplit_reduce_kernel<T><<<(unsigned)(col_tiles * row_tiles), threads, smem, st>>>(
        x, (int)n, data, (long long)w, periodic, l, r, levels, rt, (int)col_tiles, fac, fac_stride, R, err,
        individual ? lk : nullptr, lv, rk, rv, pos);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return forked ? cudaStreamWaitEvent(st, side.join, 0) : cudaSuccess;
}

template cudaError_t launch_rowsplit_front<float>(const float*, int64_t, const float*, int64_t, int, int, const int32_t*,
                                                  const float*, const int32_t*, const float*, const int32_t*, float*, size_t,
                                                  float*, unsigned long long*, cudaStream_t);
template cudaError_t launch_rowsplit_front<double>(const double*, int64_t, const double*, int64_t, int, int, const int32_t*,
                                                   const double*, const int32_t*, const double*, const int32_t*, double*, size_t,
                                                   double*, unsigned long long*, cudaStream_t);

}  // namespace ndi
