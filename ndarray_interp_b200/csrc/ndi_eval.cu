// ndi_eval.cu -- K2 lower_index, K3 interp1d_linear, K4 interp2d_bilinear, K5 interp1d_cubic,
// with K7 (first-error reduction) fused in, plus the query validation pre-pass.
//
// One launch replaces the reference's per-query loop (interp1d/mod.rs:300-343,
// interp2d/mod.rs:255-307).  All three evaluation kernels share one work decomposition:
//
//   * a TILE is 32 consecutive queries.  Each lane of a warp loads one query of the tile
//     (coalesced), does the bounds check and the lower-index search for it, and keeps the
//     per-query scalars (interval index, x - x1, x2 - x1, t ...) in registers.
//   * the per-query scalars are then handed round the warp with shuffles and the trailing axis
//     is processed by LPQ lanes per query with V-element (16-byte when possible) vector loads
//     and stores, so that a warp's stores cover whole contiguous output rows:
//       - thin rows (w <= 32*V):  LPQ = pow2ceil(w / V) lanes per query, 32/LPQ queries per
//         round, LPQ rounds per tile -- one thread per query when w == 1;
//       - wide rows (w > 32*V):   LPQ = 32, the row is cut into SLICES of 32*V columns and a
//         (tile, slice) pair is one warp task; consecutive warps take consecutive slices of the
//         same tile, so a block writes whole rows.  Within a task the table rows stay in
//         registers while consecutive queries fall into the same interval (sorted batches).
//   * a persistent grid (SM count x resident blocks) strides over the tasks.
//
// HBM traffic per query is the compulsory s*c (query) + s*w (output); table rows come out of
// L1/L2 unless the table exceeds L2.  Output uses streaming stores so it does not evict the table.
#include <initializer_list>

#include "ndi_device.cuh"
#include "ndi_internal.h"

namespace ndi {

constexpr int kBlock = 256;          // 8 warps
constexpr int kWarpsPerBlock = kBlock / 32;

template <class T>
struct Eval1 {                       // 1-D kernels
    const T* grid; int n; SearchCfg sc;
    const T* data; const T* a; const T* b; long long w;
    const T* q; long long nq; int mode;
    T* out; unsigned long long* err;
    long long ntasks; int nslices;
};
template <class T>
struct Eval2 {                       // bilinear
    const T* gx; int n; SearchCfg scx;
    const T* gy; int m; SearchCfg scy;
    const T* data; long long w;
    const T* qx; const T* qy; long long nq; int extrapolate;
    T* out; unsigned long long* err;
    long long ntasks; int nslices;
};

template <int LPQ> struct Rounds { static constexpr int QPR = 32 / LPQ; };

__device__ __forceinline__ bool shfl_b(bool v, int src) { return __shfl_sync(0xffffffffu, (int)v, src) != 0; }
template <class T> __device__ __forceinline__ T shfl_t(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ------------------------------------------------------------------------------------------------
// K3: Linear::interp_into x batch (linear.rs:73-98)
// ------------------------------------------------------------------------------------------------
template <class T, int V, int LPQ>
__global__ void __launch_bounds__(kBlock) interp1d_linear_kernel(const Eval1<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const T* g = p.grid;
    if (p.sc.smem) g = stage_grid<T>(reinterpret_cast<T*>(smem_raw), p.grid, p.n, &bar);
    const T g0 = g[0], gl = g[p.n - 1];
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    for (long long task = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < p.ntasks; task += nwarps) {
        const long long tile = (LPQ == 32) ? task / p.nslices : task;
        const int slice = (LPQ == 32) ? (int)(task - tile * p.nslices) : 0;
        const long long qbase = tile * 32;
        const long long qi = qbase + lane;
        const bool live = qi < p.nq;
        const T x = live ? ld_query(p.q + qi) : g0;
        // linear.rs:80-84: out of range (or NaN) without extrapolation is an error;
        // with extrapolation only NaN fails (vector_extensions.rs:83-84)
        const bool bad = live && (p.mode ? Ar<T>::is_nan(x) : !in_range(g0, gl, x));
        const int idx = lower_index<T>(g, p.n, x, p.sc.top_step, p.sc.guess != 0);   // linear.rs:87
        const T x1 = g[idx], x2 = g[idx + 1];                                          // linear.rs:90-91
        const T dx21 = Ar<T>::sub(x2, x1), dxq = Ar<T>::sub(x, x1);
        if (slice == 0) report_first_bad(p.err, bad, (unsigned long long)qi);
        const bool skip = bad || !live;

        if constexpr (LPQ < 32) {
            constexpr int QPR = 32 / LPQ;
            const int sub = lane % LPQ;
#pragma unroll
            for (int r = 0; r < LPQ; ++r) {
                const int src = r * QPR + lane / LPQ;
                const int is = __shfl_sync(0xffffffffu, idx, src);
                const T d21 = shfl_t(dx21, src), dq = shfl_t(dxq, src);
                const bool sk = shfl_b(skip, src);
                if (!sk) {
                    const T* row = p.data + (long long)is * p.w;
                    T* o = p.out + (qbase + src) * p.w;
                    for (long long col = (long long)sub * V; col < p.w; col += LPQ * V) {
                        const Vec<T, V> y1 = ld_table<T, V>(row + col);
                        const Vec<T, V> y2 = ld_table<T, V>(row + p.w + col);
                        Vec<T, V> res;
#pragma unroll
                        for (int e = 0; e < V; ++e) res.v[e] = calc_frac_pre<T>(y1.v[e], y2.v[e], d21, dq);  // linear.rs:94-96
                        st_stream<T, V>(o + col, res);
                    }
                }
            }
        } else {
            const long long col = ((long long)slice * 32 + lane) * V;
            const bool colok = col < p.w;
            int cur = -1;
            Vec<T, V> y1, y2;
            const int nlive = (int)min((long long)32, p.nq - qbase);
            for (int s = 0; s < nlive; ++s) {
                const int is = __shfl_sync(0xffffffffu, idx, s);
                const T d21 = shfl_t(dx21, s), dq = shfl_t(dxq, s);
                const bool sk = shfl_b(skip, s);
                if (sk || !colok) continue;
                if (is != cur) {
                    const T* row = p.data + (long long)is * p.w + col;
                    y1 = ld_table<T, V>(row);
                    y2 = ld_table<T, V>(row + p.w);
                    cur = is;
                }
                Vec<T, V> res;
#pragma unroll
                for (int e = 0; e < V; ++e) res.v[e] = calc_frac_pre<T>(y1.v[e], y2.v[e], d21, dq);
                st_stream<T, V>(p.out + (qbase + s) * p.w + col, res);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5: CubicSplineStrategy::interp_into x batch (cubic_spline.rs:791-830)
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T rem_euclid(T a, T b) {            // num_traits::Euclid for floats
    T r = fmod(a, b);
    return r < (T)0 ? Ar<T>::add(r, fabs(b)) : r;
}

template <class T>
__device__ __forceinline__ T cubic_point(T yl, T yr, T al, T bl, T t, T omt, T tt) {
    // (1 - t) * yl + t * yr + t * (1 - t) * (a * (1 - t) + b * t)        cubic_spline.rs:825-827
    const T lin = Ar<T>::add(Ar<T>::mul(omt, yl), Ar<T>::mul(t, yr));
    const T cur = Ar<T>::add(Ar<T>::mul(al, omt), Ar<T>::mul(bl, t));
    return Ar<T>::add(lin, Ar<T>::mul(tt, cur));
}

template <class T, int V, int LPQ>
__global__ void __launch_bounds__(kBlock) interp1d_cubic_kernel(const Eval1<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const T* g = p.grid;
    if (p.sc.smem) g = stage_grid<T>(reinterpret_cast<T*>(smem_raw), p.grid, p.n, &bar);
    const T g0 = g[0], gl = g[p.n - 1];
    const T one = (T)1;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    for (long long task = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < p.ntasks; task += nwarps) {
        const long long tile = (LPQ == 32) ? task / p.nslices : task;
        const int slice = (LPQ == 32) ? (int)(task - tile * p.nslices) : 0;
        const long long qbase = tile * 32;
        const long long qi = qbase + lane;
        const bool live = qi < p.nq;
        T x = live ? ld_query(p.q + qi) : g0;
        const bool inr = in_range(g0, gl, x);                                          // :797
        bool bad = false;
        if (p.mode == 0) bad = !inr;                                                   // :798-802
        else if (p.mode == 2 && !inr) x = Ar<T>::add(rem_euclid<T>(Ar<T>::sub(x, g0), Ar<T>::sub(gl, g0)), g0);  // :805-809
        if (p.mode != 0) bad = Ar<T>::is_nan(x);                                       // NaN reaches get_lower_index
        bad = bad && live;
        const int idx = lower_index<T>(g, p.n, x, p.sc.top_step, p.sc.guess != 0);   // :811
        const T xl = g[idx], xr = g[idx + 1];
        const T t = Ar<T>::div(Ar<T>::sub(x, xl), Ar<T>::sub(xr, xl));                 // :818
        const T omt = Ar<T>::sub(one, t);
        const T tt = Ar<T>::mul(t, omt);
        if (slice == 0) report_first_bad(p.err, bad, (unsigned long long)qi);
        const bool skip = bad || !live;

        if constexpr (LPQ < 32) {
            constexpr int QPR = 32 / LPQ;
            const int sub = lane % LPQ;
#pragma unroll
            for (int r = 0; r < LPQ; ++r) {
                const int src = r * QPR + lane / LPQ;
                const int is = __shfl_sync(0xffffffffu, idx, src);
                const T ts = shfl_t(t, src), os = shfl_t(omt, src), tts = shfl_t(tt, src);
                const bool sk = shfl_b(skip, src);
                if (!sk) {
                    const long long ro = (long long)is * p.w;
                    T* o = p.out + (qbase + src) * p.w;
                    for (long long col = (long long)sub * V; col < p.w; col += LPQ * V) {
                        const Vec<T, V> yl = ld_table<T, V>(p.data + ro + col);
                        const Vec<T, V> yr = ld_table<T, V>(p.data + ro + p.w + col);
                        const Vec<T, V> al = ld_table<T, V>(p.a + ro + col);
                        const Vec<T, V> bl = ld_table<T, V>(p.b + ro + col);
                        Vec<T, V> res;
#pragma unroll
                        for (int e = 0; e < V; ++e) res.v[e] = cubic_point<T>(yl.v[e], yr.v[e], al.v[e], bl.v[e], ts, os, tts);
                        st_stream<T, V>(o + col, res);
                    }
                }
            }
        } else {
            const long long col = ((long long)slice * 32 + lane) * V;
            const bool colok = col < p.w;
            int cur = -1;
            Vec<T, V> yl, yr, al, bl;
            const int nlive = (int)min((long long)32, p.nq - qbase);
            for (int s = 0; s < nlive; ++s) {
                const int is = __shfl_sync(0xffffffffu, idx, s);
                const T ts = shfl_t(t, s), os = shfl_t(omt, s), tts = shfl_t(tt, s);
                const bool sk = shfl_b(skip, s);
                if (sk || !colok) continue;
                if (is != cur) {
                    const long long ro = (long long)is * p.w + col;
                    yl = ld_table<T, V>(p.data + ro);
                    yr = ld_table<T, V>(p.data + ro + p.w);
                    al = ld_table<T, V>(p.a + ro);
                    bl = ld_table<T, V>(p.b + ro);
                    cur = is;
                }
                Vec<T, V> res;
#pragma unroll
                for (int e = 0; e < V; ++e) res.v[e] = cubic_point<T>(yl.v[e], yr.v[e], al.v[e], bl.v[e], ts, os, tts);
                st_stream<T, V>(p.out + (qbase + s) * p.w + col, res);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4: Bilinear::interp_into x batch (bilinear.rs:64-99)
// ------------------------------------------------------------------------------------------------
template <class T, int V, int LPQ>
__global__ void __launch_bounds__(kBlock) interp2d_bilinear_kernel(const Eval2<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar[2];
    const T* gx = p.gx;
    const T* gy = p.gy;
    if (p.scx.smem) gx = stage_grid<T>(reinterpret_cast<T*>(smem_raw), p.gx, p.n, &bar[0]);
    if (p.scy.smem) {
        size_t off = p.scx.smem ? (((size_t)p.n * sizeof(T) + 15) & ~(size_t)15) : 0;
        gy = stage_grid<T>(reinterpret_cast<T*>(smem_raw + off), p.gy, p.m, &bar[1]);
    }
    const T gx0 = gx[0], gxl = gx[p.n - 1], gy0 = gy[0], gyl = gy[p.m - 1];
    const int lane = threadIdx.x & 31;
    const long long rowx = (long long)p.m * p.w;      // elements between x-rows
    const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
    for (long long task = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); task < p.ntasks; task += nwarps) {
        const long long tile = (LPQ == 32) ? task / p.nslices : task;
        const int slice = (LPQ == 32) ? (int)(task - tile * p.nslices) : 0;
        const long long qbase = tile * 32;
        const long long qi = qbase + lane;
        const bool live = qi < p.nq;
        const T x = live ? ld_query(p.qx + qi) : gx0;
        const T y = live ? ld_query(p.qy + qi) : gy0;
        // bilinear.rs:71-80: x is checked before y
        bool badx, bady;
        if (p.extrapolate) { badx = Ar<T>::is_nan(x); bady = Ar<T>::is_nan(y); }
        else { badx = !in_range(gx0, gxl, x); bady = !in_range(gy0, gyl, y); }
        const bool bad = live && (badx || bady);
        const int ix = lower_index<T>(gx, p.n, x, p.scx.top_step, p.scx.guess != 0);  // :82
        const int iy = lower_index<T>(gy, p.m, y, p.scy.top_step, p.scy.guess != 0);
        const T x1 = gx[ix], x2 = gx[ix + 1], y1 = gy[iy], y2 = gy[iy + 1];
        const T dx21 = Ar<T>::sub(x2, x1), dxq = Ar<T>::sub(x, x1);
        const T dy21 = Ar<T>::sub(y2, y1), dyq = Ar<T>::sub(y, y1);
        const long long cell = ((long long)ix * p.m + iy) * p.w;                      // z11 (:83)
        if (slice == 0) report_first_bad(p.err, bad, 2ull * (unsigned long long)qi + (badx ? 0ull : 1ull));
        const bool skip = bad || !live;

        auto point = [&](T z11, T z12, T z21, T z22, T d21x, T dqx, T d21y, T dqy) -> T {
            const T z1 = calc_frac_pre<T>(z11, z21, d21x, dqx);                       // :94
            const T z2 = calc_frac_pre<T>(z12, z22, d21x, dqx);                       // :95
            return calc_frac_pre<T>(z1, z2, d21y, dqy);                               // :96
        };

        if constexpr (LPQ < 32) {
            constexpr int QPR = 32 / LPQ;
            const int sub = lane % LPQ;
#pragma unroll
            for (int r = 0; r < LPQ; ++r) {
                const int src = r * QPR + lane / LPQ;
                const long long cs = __shfl_sync(0xffffffffu, cell, src);
                const T ax = shfl_t(dx21, src), bx = shfl_t(dxq, src), ay = shfl_t(dy21, src), by = shfl_t(dyq, src);
                const bool sk = shfl_b(skip, src);
                if (!sk) {
                    T* o = p.out + (qbase + src) * p.w;
                    for (long long col = (long long)sub * V; col < p.w; col += LPQ * V) {
                        const T* c0 = p.data + cs + col;
                        const Vec<T, V> z11 = ld_table<T, V>(c0);
                        const Vec<T, V> z12 = ld_table<T, V>(c0 + p.w);
                        const Vec<T, V> z21 = ld_table<T, V>(c0 + rowx);
                        const Vec<T, V> z22 = ld_table<T, V>(c0 + rowx + p.w);
                        Vec<T, V> res;
#pragma unroll
                        for (int e = 0; e < V; ++e) res.v[e] = point(z11.v[e], z12.v[e], z21.v[e], z22.v[e], ax, bx, ay, by);
                        st_stream<T, V>(o + col, res);
                    }
                }
            }
        } else {
            const long long col = ((long long)slice * 32 + lane) * V;
            const bool colok = col < p.w;
            long long cur = -1;
            Vec<T, V> z11, z12, z21, z22;
            const int nlive = (int)min((long long)32, p.nq - qbase);
            for (int s = 0; s < nlive; ++s) {
                const long long cs = __shfl_sync(0xffffffffu, cell, s);
                const T ax = shfl_t(dx21, s), bx = shfl_t(dxq, s), ay = shfl_t(dy21, s), by = shfl_t(dyq, s);
                const bool sk = shfl_b(skip, s);
                if (sk || !colok) continue;
                if (cs != cur) {
                    const T* c0 = p.data + cs + col;
                    z11 = ld_table<T, V>(c0);
                    z12 = ld_table<T, V>(c0 + p.w);
                    z21 = ld_table<T, V>(c0 + rowx);
                    z22 = ld_table<T, V>(c0 + rowx + p.w);
                    cur = cs;
                }
                Vec<T, V> res;
#pragma unroll
                for (int e = 0; e < V; ++e) res.v[e] = point(z11.v[e], z12.v[e], z21.v[e], z22.v[e], ax, bx, ay, by);
                st_stream<T, V>(p.out + (qbase + s) * p.w + col, res);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 standalone: get_lower_index for a batch (vector_extensions.rs:55-111)
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kBlock) lower_index_kernel(const T* __restrict__ grid, int n, SearchCfg sc,
                                                            const T* __restrict__ q, long long nq,
                                                            long long* __restrict__ idx, unsigned long long* err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const T* g = grid;
    if (sc.smem) g = stage_grid<T>(reinterpret_cast<T*>(smem_raw), grid, n, &bar);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nq_pad = (nq + 31) & ~31ll;       // whole warps, so the ballot in report_first_bad is full
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq_pad; i += stride) {
        const bool live = i < nq;
        const T x = live ? ld_query(q + i) : g[0];
        const bool bad = live && Ar<T>::is_nan(x);
        const int r = lower_index<T>(g, n, x, sc.top_step, sc.guess != 0);
        report_first_bad(err, bad, (unsigned long long)i);
        if (live && !bad) idx[i] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// K7 pre-pass for the host entry points: first failing query, reading only the queries.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kBlock) validate_queries_kernel(const T* __restrict__ gx, int n, const T* __restrict__ gy,
                                                                 int m, const T* __restrict__ qx,
                                                                 const T* __restrict__ qy, long long nq, int check,
                                                                 unsigned long long* err) {
    const T gx0 = gx[0], gxl = gx[n - 1];
    const T gy0 = gy ? gy[0] : (T)0, gyl = gy ? gy[m - 1] : (T)0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nq_pad = (nq + 31) & ~31ll;
    unsigned long long first = ~0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq_pad; i += stride) {
        if (i >= nq) continue;
        const T x = ld_query(qx + i);
        bool badx, bady = false;
        if (check == CHECK_IN_RANGE) badx = !in_range(gx0, gxl, x);
        else if (check == CHECK_NOT_NAN) badx = Ar<T>::is_nan(x);
        else badx = !in_range(gx0, gxl, x) && !Ar<T>::is_finite(x);   // periodic wrap of +-inf / NaN gives NaN
        if (qy) {
            const T y = ld_query(qy + i);
            bady = (check == CHECK_IN_RANGE) ? !in_range(gy0, gyl, y) : Ar<T>::is_nan(y);
        }
        if (badx || bady) {
            unsigned long long word = qy ? 2ull * (unsigned long long)i + (badx ? 0ull : 1ull) : (unsigned long long)i;
            first = min(first, word);
        }
    }
    // warp min, then one atomic per warp that saw a failure
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if ((threadIdx.x & 31) == 0 && first != ~0ull) atomicMin(err, first);
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
static int pow2ceil(long long v) { int p = 1; while (p < v && p < 32) p <<= 1; return p; }

static bool aligned(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }

template <class T>
static int pick_vec(long long w, std::initializer_list<const void*> ptrs) {
    constexpr int vmax = 16 / (int)sizeof(T);
    for (int v = vmax; v > 1; v >>= 1) {
        bool ok = (w % v) == 0;
        for (const void* p : ptrs) ok = ok && aligned(p, sizeof(T) * v);
        if (ok) return v;
    }
    return 1;
}

struct Shape { int v; int lpq; int nslices; long long ntasks; };

template <class T>
static Shape pick_shape(long long w, long long nq, int v) {
    Shape s;
    s.v = v;
    long long groups = (w + v - 1) / v;
    s.lpq = pow2ceil(groups);
    s.nslices = s.lpq == 32 ? (int)((groups + 31) / 32) : 1;
    long long tiles = (nq + 31) / 32;
    s.ntasks = tiles * s.nslices;
    return s;
}

template <class K>
static cudaError_t prep_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

template <class K>
static int persistent_grid(K kernel, size_t smem, long long ntasks) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long want = (ntasks + kWarpsPerBlock - 1) / kWarpsPerBlock;
    long long cap = (long long)device_info().sm_count * per_sm;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

template <class P, class K>
static cudaError_t launch_eval(K kernel, const P& p, size_t smem, cudaStream_t st) {
    cudaError_t e = prep_smem(kernel, smem);
    if (e != cudaSuccess) return e;
    int grid = persistent_grid(kernel, smem, p.ntasks);
    kernel<<<grid, kBlock, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

#define NDI_LPQ_SWITCH(KERNEL, T, V, LPQ, P, SMEM, ST)                                   \
    switch (LPQ) {                                                                        \
    case 1:  return launch_eval(KERNEL<T, V, 1>, P, SMEM, ST);                            \
    case 2:  return launch_eval(KERNEL<T, V, 2>, P, SMEM, ST);                            \
    case 4:  return launch_eval(KERNEL<T, V, 4>, P, SMEM, ST);                            \
    case 8:  return launch_eval(KERNEL<T, V, 8>, P, SMEM, ST);                            \
    case 16: return launch_eval(KERNEL<T, V, 16>, P, SMEM, ST);                           \
    default: return launch_eval(KERNEL<T, V, 32>, P, SMEM, ST);                           \
    }

#define NDI_VEC_SWITCH(KERNEL, T, SH, P, SMEM, ST)                                        \
    if constexpr (sizeof(T) == 4) {                                                       \
        if (SH.v == 4) { NDI_LPQ_SWITCH(KERNEL, T, 4, SH.lpq, P, SMEM, ST) }              \
        if (SH.v == 2) { NDI_LPQ_SWITCH(KERNEL, T, 2, SH.lpq, P, SMEM, ST) }              \
        NDI_LPQ_SWITCH(KERNEL, T, 1, SH.lpq, P, SMEM, ST)                                 \
    } else {                                                                              \
        if (SH.v == 2) { NDI_LPQ_SWITCH(KERNEL, T, 2, SH.lpq, P, SMEM, ST) }              \
        NDI_LPQ_SWITCH(KERNEL, T, 1, SH.lpq, P, SMEM, ST)                                 \
    }

static size_t grid_smem_bytes(int smem, long long n, size_t elem) { return smem ? (((size_t)n * elem + 15) & ~(size_t)15) : 0; }

template <class T>
cudaError_t launch_interp1d_linear(const T* grid, int64_t n, SearchCfg sc, const T* data, int64_t w, const T* q,
                                   int64_t nq, int extrapolate, T* out, unsigned long long* err, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    Shape sh = pick_shape<T>(w, nq, pick_vec<T>(w, {data, out}));
    Eval1<T> p{grid, (int)n, sc, data, nullptr, nullptr, (long long)w, q, (long long)nq, extrapolate, out, err, sh.ntasks, sh.nslices};
    size_t smem = grid_smem_bytes(sc.smem, n, sizeof(T));
    NDI_VEC_SWITCH(interp1d_linear_kernel, T, sh, p, smem, st)
}

template <class T>
cudaError_t launch_interp1d_cubic(const T* grid, int64_t n, SearchCfg sc, const T* data, const T* a, const T* b,
                                  int64_t w, const T* q, int64_t nq, int extrap_mode, T* out,
                                  unsigned long long* err, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    Shape sh = pick_shape<T>(w, nq, pick_vec<T>(w, {data, a, b, out}));
    Eval1<T> p{grid, (int)n, sc, data, a, b, (long long)w, q, (long long)nq, extrap_mode, out, err, sh.ntasks, sh.nslices};
    size_t smem = grid_smem_bytes(sc.smem, n, sizeof(T));
    NDI_VEC_SWITCH(interp1d_cubic_kernel, T, sh, p, smem, st)
}

template <class T>
cudaError_t launch_interp2d_bilinear(const T* gx, int64_t n, SearchCfg scx, const T* gy, int64_t m, SearchCfg scy,
                                     const T* data, int64_t w, const T* qx, const T* qy, int64_t nq, int extrapolate,
                                     T* out, unsigned long long* err, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    Shape sh = pick_shape<T>(w, nq, pick_vec<T>(w, {data, out}));
    Eval2<T> p{gx, (int)n, scx, gy, (int)m, scy, data, (long long)w, qx, qy, (long long)nq, extrapolate, out, err, sh.ntasks, sh.nslices};
    size_t smem = grid_smem_bytes(scx.smem, n, sizeof(T)) + grid_smem_bytes(scy.smem, m, sizeof(T));
    NDI_VEC_SWITCH(interp2d_bilinear_kernel, T, sh, p, smem, st)
}

template <class T>
cudaError_t launch_lower_index(const T* grid, int64_t n, SearchCfg sc, const T* q, int64_t nq, int64_t* idx,
                               unsigned long long* err, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    size_t smem = grid_smem_bytes(sc.smem, n, sizeof(T));
    cudaError_t e = prep_smem(lower_index_kernel<T>, smem);
    if (e != cudaSuccess) return e;
    long long blocks = (nq + kBlock - 1) / kBlock;
    long long cap = (long long)device_info().sm_count * 8;
    lower_index_kernel<T><<<(int)(blocks < cap ? blocks : cap), kBlock, smem, st>>>(grid, (int)n, sc, q, (long long)nq,
                                                                                   (long long*)idx, err);
    count_launch();
    return cudaGetLastError();
}

template <class T>
cudaError_t launch_validate_queries(const T* gx, int64_t n, const T* gy, int64_t m, const T* qx, const T* qy,
                                    int64_t nq, int check, unsigned long long* err, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    long long blocks = (nq + kBlock - 1) / kBlock;
    long long cap = (long long)device_info().sm_count * 8;
    validate_queries_kernel<T><<<(int)(blocks < cap ? blocks : cap), kBlock, 0, st>>>(gx, (int)n, gy, (int)m, qx, qy,
                                                                                     (long long)nq, check, err);
    count_launch();
    return cudaGetLastError();
}

#define NDI_INST_COMMON(T)                                                                                             \
    template cudaError_t launch_interp1d_linear<T>(const T*, int64_t, SearchCfg, const T*, int64_t, const T*, int64_t, \
                                                   int, T*, unsigned long long*, cudaStream_t);                        \
    template cudaError_t launch_interp2d_bilinear<T>(const T*, int64_t, SearchCfg, const T*, int64_t, SearchCfg,       \
                                                     const T*, int64_t, const T*, const T*, int64_t, int, T*,          \
                                                     unsigned long long*, cudaStream_t);                               \
    template cudaError_t launch_lower_index<T>(const T*, int64_t, SearchCfg, const T*, int64_t, int64_t*,              \
                                               unsigned long long*, cudaStream_t);                                     \
    template cudaError_t launch_validate_queries<T>(const T*, int64_t, const T*, int64_t, const T*, const T*, int64_t, \
                                                    int, unsigned long long*, cudaStream_t);
NDI_INST_COMMON(float)
NDI_INST_COMMON(double)
NDI_INST_COMMON(int32_t)
template cudaError_t launch_interp1d_cubic<float>(const float*, int64_t, SearchCfg, const float*, const float*,
                                                  const float*, int64_t, const float*, int64_t, int, float*,
                                                  unsigned long long*, cudaStream_t);
template cudaError_t launch_interp1d_cubic<double>(const double*, int64_t, SearchCfg, const double*, const double*,
                                                   const double*, int64_t, const double*, int64_t, int, double*,
                                                   unsigned long long*, cudaStream_t);

}  // namespace ndi
