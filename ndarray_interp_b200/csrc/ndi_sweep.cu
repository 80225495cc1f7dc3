// ndi_sweep.cu -- K9: bilinear evaluation in BAND SWEEPS, for thin rows on a table a few times the size of L2.
//
// Why.  Bilinear::interp_into (bilinear.rs:83-97) gathers four cells per query.  With random queries on a table
// that just misses L2 (C4: 134 MB of 32-byte cells against a 126 MB L2 that random gathers from both dies use
// about half of) nearly every gather is a DRAM access: ncu shows 2.6 GB read for a 0.13 GB table and the direct
// kernel sits at the DRAM roofline of that pattern (profiles/r01/gather_ceiling.md).  Grouping the queries by table
// band first (K8, ndi_bin.cu) fixes the traffic but costs two passes over the batch plus scattered output rows,
// which on 32-byte rows is what the gathers cost in the first place (profiles/r01/binning.md).
//
// What.  The batch loop of the reference (interp2d/mod.rs:255-307) is order-independent: every query writes its own
// output row.  So the launch walks the batch P times; sweep b evaluates exactly the queries whose x lies in band b
// of the table (bound[b-1] <= x < bound[b], the bounds being grid values, so that the interval index of a member
// lies in the band) and leaves the others alone.  Nothing is reordered in memory and nothing extra is written:
//   * a warp takes a stretch of 32*S consecutive queries, reads their x (coalesced), and COMPACTS the members of the
//     current band into a per-warp ring in shared memory (ballot + popc); whenever 32 members have collected it
//     evaluates them as one full tile -- so the evaluation runs at full lane efficiency although only 1/P of a
//     stretch belongs to the sweep (a pass that merely predicates the direct kernel pays the whole latency chain
//     per stretch: 0.54 / 0.78 / 1.08 / 1.39 ms for P = 1..4, profiles/r01/binning.md).  Members left over travel
//     on into the next stretch -- and across a band boundary, a tile may mix bands; only locality depends on it;
//   * stretches are handed out IN ORDER through an atomic counter, sweep after sweep, so all SMs work in the same
//     band (two at a boundary) and the band's rows (about 32 MB) are read from DRAM once and gathered from L2;
//   * extra HBM traffic: (P - 1) * s bytes per query (x is read P times; y once per member), against up to
//     4 * 64 bytes of gather traffic per query saved.
// Rows of exactly 32 bytes (C4) take the PAIR form of the gathers: the two lanes of a query load the two
// y-neighbours z11 | z12 (one contiguous 64-byte segment: one L1TEX wavefront instead of two) with one 256-bit load
// each, then z21 | z22; each lane runs the x-interpolation (bilinear.rs:94-95) on the whole row it holds, the lanes
// swap halves, and each finishes the y-interpolation (:96) for its 16 bytes of the output row.  Same operations on
// the same operands as bilerp_vec, hence the same bits.
#include <stdlib.h>

#include "ndi_device.cuh"
#include "ndi_internal.h"

namespace ndi {

constexpr int kSwBlock = 256;
constexpr int kSwWarps = kSwBlock / 32;

template <class T>
struct Sweep2 {
    const T* gx; int n; SearchCfg scx;
    const T* gy; int m; SearchCfg scy;
    const T* data; long long w;
    const T* qx; const T* qy; long long nq; int extrapolate;
    T* out; unsigned long long* err;
    int fast_tables;
    unsigned long long* next_task;      // zeroed before the launch
    int nsweeps, band_rows;             // sweep b: interval indices [b * band_rows, (b + 1) * band_rows)
    long long nstretch;                 // stretches of 32 * S queries per sweep
    int stage_x;                        // 1: the x-grid is copied to shared memory (n elements after the staged search tables)
};

template <class T> struct SweepShape {
    static constexpr int S = sizeof(T) == 4 ? 8 : 4;        // 32-query loads per stretch
    static constexpr int QCAP = 64 * S;                     // ring of collected members (power of two): a stretch on top of at most 63 waiting
};

template <class T> struct alignas(16) SwRec { long long cs; T bx, by; Slope<T> sx, sy; unsigned orow; };

template <class T> struct alignas(32) Row32 { T v[32 / sizeof(T)]; };
template <class T>
__device__ __forceinline__ Row32<T> ld_row32(const T* p) {            // one 256-bit load (LDG.E.256), read-only path
    unsigned long long a, b, c, d;
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    unsigned long long t[4] = {a, b, c, d};
    return *reinterpret_cast<const Row32<T>*>(t);
}

// Linear::calc_frac (linear.rs:29-36) on N elements that share one divisor, the per-query differences already formed:
// out[e] = (b[e] - a[e]) / s.d * dq + a[e].  `fast`: both divisors of the query have a hoisted reciprocal -- the
// condition under which bilerp_vec divides that way; its second stage (checked) also needs numerators that are 0 or
// at least 2^-80, else the whole vector is redone with IEEE divisions.  Same operations as bilerp_vec, same bits.
template <class T, int N>
__device__ __forceinline__ void sweep_stage(const T (&a)[N], const T (&b)[N], const Slope<T>& s, T dq, bool fast, bool checked, T (&out)[N]) {
    if constexpr (std::is_same<T, float>::value) {
        if (fast) {
            bool all_ok = true;
            const F2 r2 = F2::both(s.r), nb2 = F2::both(-s.d), dq2 = F2::both(dq);
#pragma unroll
            for (int e = 0; e < N; e += 2) {
                const F2 lo{a[e], a[e + 1]}, hi{b[e], b[e + 1]};
                const F2 num = sub2(hi, lo);
                if (checked) all_ok = all_ok && numer_ok(num.lo) && numer_ok(num.hi);
                const F2 o = add_halves(mul2(div_by2(num, nb2, r2), dq2), lo);
                out[e] = o.lo; out[e + 1] = o.hi;
            }
            if (all_ok) return;
        }
    } else if constexpr (std::is_same<T, double>::value) {
        if (fast) {
#pragma unroll
            for (int e = 0; e < N; ++e) out[e] = __dadd_rn(__dmul_rn(s.div(__dsub_rn(b[e], a[e])), dq), a[e]);
            return;
        }
    }
#pragma unroll
    for (int e = 0; e < N; ++e) out[e] = calc_frac_pre<T>(a[e], b[e], s.d, dq);
}
template <class T> __device__ __forceinline__ bool sweep_fast(const Slope<T>&, const Slope<T>&) { return false; }
template <> __device__ __forceinline__ bool sweep_fast<float>(const Slope<float>& a, const Slope<float>& b) { return a.r != 0.0f && b.r != 0.0f; }
template <> __device__ __forceinline__ bool sweep_fast<double>(const Slope<double>& a, const Slope<double>& b) { return a.r != 0.0 && b.r != 0.0; }

#ifndef NDI_SWEEP_MINBLOCKS
#define NDI_SWEEP_MINBLOCKS 4
#endif

template <class T, int LPQ>
__global__ void __launch_bounds__(kSwBlock, sizeof(T) == 4 ? NDI_SWEEP_MINBLOCKS : 3) interp2d_bilinear_sweep_kernel(const Sweep2<T> p) {
    constexpr int V = 16 / (int)sizeof(T), QPR = 32 / LPQ;
    constexpr int S = SweepShape<T>::S, QCAP = SweepShape<T>::QCAP;
    constexpr bool kBcast = LPQ >= 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar[2];
    __shared__ unsigned q_idx[kSwWarps][QCAP];
    __shared__ T q_x[kSwWarps][QCAP];
    __shared__ T s_bound[kMaxSweeps + 1];
    __shared__ SwRec<T> recs[kBcast ? kSwWarps : 1][kBcast ? 32 : 1];
    GridView<T> gx = make_grid_view<T>(p.gx, p.n, p.scx, smem_raw, &bar[0]);
    const GridView<T> gy = make_grid_view<T>(p.gy, p.m, p.scy, smem_raw + stage_bytes(p.scx, sizeof(T)), &bar[1]);
    if (p.stage_x) {
        // within a sweep the x-indices of a tile are spread over the whole band, so the two verification reads of the
        // even-spacing guess touch up to 32 different lines each; out of shared memory they are a few wavefronts
        T* sx = reinterpret_cast<T*>(smem_raw + stage_bytes(p.scx, sizeof(T)) + stage_bytes(p.scy, sizeof(T)));
        for (int i = threadIdx.x; i < p.n; i += kSwBlock) sx[i] = p.gx[i];
        gx.fine = sx;
        if (gx.shift == 0 && !p.scx.smem) gx.top = sx;
    }
    // bound[b] = first grid value of band b + 1: sweep b takes bound[b-1] <= x < bound[b]; the first sweep also
    // takes everything below (and NaN), the last everything above
    for (int i = threadIdx.x; i < p.nsweeps - 1; i += kSwBlock) s_bound[i] = p.gx[(long long)(i + 1) * p.band_rows];
    __syncthreads();
    const T gx0 = gx.g0, gxl = gx.gl, gy0 = gy.g0, gyl = gy.gl;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long rowx = (long long)p.m * p.w;
    const long long total = p.nstretch * p.nsweeps;
    const int sub = lane % LPQ, qsel = lane / LPQ;
    const long long col = (long long)sub * V;

    auto fetch = [&]() -> long long {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(p.next_task, 1ull);
        return (long long)__shfl_sync(0xffffffffu, t, 0);
    };
    auto load_stretch = [&](long long ticket, T (&xs)[S]) {
        const long long base = (ticket % p.nstretch) * (32 * S);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const long long qi = base + s * 32 + lane;
            xs[s] = (ticket < total && qi < p.nq) ? ld_query(p.qx + qi) : gx0;
        }
    };

    // one tile of (up to) 32 collected queries: lane holds query `qi` with coordinates `x`, `y`
    auto evaluate = [&](bool live, unsigned qi, T x, T y) {
        T xq[1] = {live ? x : gx0}, yq[1] = {live ? y : gy0};
        int ix[1], iy[1]; T ax[1], bx[1], ay[1], by[1];
        search_multi<T, 1>(gx, xq, ix, ax, bx);                                           // bilinear.rs:82 (+ x1, x2)
        search_multi<T, 1>(gy, yq, iy, ay, by);
        bool badx, bady;                                                                  // :71-80: x is checked before y
        if (p.extrapolate) { badx = Ar<T>::is_nan(xq[0]); bady = Ar<T>::is_nan(yq[0]); }
        else { badx = !in_range(gx0, gxl, xq[0]); bady = !in_range(gy0, gyl, yq[0]); }
        const bool bad = live && (badx || bady);
        const T dqx = Ar<T>::sub(xq[0], ax[0]), dqy = Ar<T>::sub(yq[0], ay[0]);
        const Slope<T> slx = Slope<T>::make(Ar<T>::sub(bx[0], ax[0]), dqx, p.fast_tables != 0);
        const Slope<T> sly = Slope<T>::make(Ar<T>::sub(by[0], ay[0]), dqy, p.fast_tables != 0);
        if (bad && p.err != nullptr) atomicMin(p.err, 2ull * qi + (badx ? 0ull : 1ull));   // lanes are not in query order
        const bool skip = bad || !live;
        const int cellno = ix[0] * p.m + iy[0];                                           // z11 (:83); n * m < 2^31 (sweep_shape_ok)
        if constexpr (LPQ == 2) {
            constexpr int E = 32 / (int)sizeof(T);                                        // elements of a 32-byte row; V = E / 2
#pragma unroll 1
            for (int r = 0; r < 2; ++r) {
                const int src = r * 16 + qsel;
                const int cs = __shfl_sync(0xffffffffu, skip ? -1 : cellno, src);
                const T sbx = __shfl_sync(0xffffffffu, dqx, src), sby = __shfl_sync(0xffffffffu, dqy, src);
                const Slope<T> ssx = slx.from_lane(src), ssy = sly.from_lane(src);
                const unsigned srow = __shfl_sync(0xffffffffu, qi, src);
                // this lane's y-neighbour of the cell: z11 and z21 (sub 0) or z12 and z22 (sub 1)
                Row32<T> lo, hi;
                if (cs >= 0) {
                    const T* c0 = p.data + (long long)(cs + sub) * E;
                    lo = ld_row32<T>(c0); hi = ld_row32<T>(c0 + rowx);
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) { lo.v[e] = (T)0; hi.v[e] = (T)0; }
                }
                const bool fast = sweep_fast<T>(ssx, ssy);
                T zz[E];                                                                  // :94 (sub 0) / :95 (sub 1), all columns
                sweep_stage<T, E>(lo.v, hi.v, ssx, sbx, fast, false, zz);
                // sub 0 produces columns 0 .. V-1 and needs z2 of those; sub 1 produces V .. 2V-1 and needs z1
                T z1[V], z2[V];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const T mine = sub ? zz[V + e] : zz[e], send = sub ? zz[e] : zz[V + e];
                    const T got = __shfl_xor_sync(0xffffffffu, send, 1);
                    z1[e] = sub ? got : mine; z2[e] = sub ? mine : got;
                }
                Vec<T, V> res;
                sweep_stage<T, V>(z1, z2, ssy, sby, fast, true, res.v);                   // :96
                if (cs >= 0) st_stream<T, V>(p.out + (long long)srow * p.w + col, res);
            }
        } else {
            if constexpr (kBcast) {
                recs[wid][lane] = SwRec<T>{skip ? -1ll : (long long)cellno * p.w, dqx, dqy, slx, sly, qi};
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < LPQ; ++r) {
                const int src = r * QPR + qsel;
                long long cs; T sbx, sby; Slope<T> ssx, ssy; unsigned srow;
                if constexpr (kBcast) {
                    const SwRec<T> rc = recs[wid][src];
                    cs = rc.cs; sbx = rc.bx; sby = rc.by; ssx = rc.sx; ssy = rc.sy; srow = rc.orow;
                } else {                                                                 // LPQ == 1: nothing to hand round
                    cs = skip ? -1ll : (long long)cellno * p.w; sbx = dqx; sby = dqy; ssx = slx; ssy = sly; srow = qi;
                }
                if (cs >= 0) {
                    const T* c0 = p.data + cs + col;
                    const Vec<T, V> z11 = ld_table<T, V>(c0), z12 = ld_table<T, V>(c0 + p.w);
                    const Vec<T, V> z21 = ld_table<T, V>(c0 + rowx), z22 = ld_table<T, V>(c0 + rowx + p.w);
                    st_stream<T, V>(p.out + (long long)srow * p.w + col, bilerp_vec<T, V, (LPQ > 2)>(z11, z12, z21, z22, ssx, sbx, ssy, sby));
                }
            }
            if constexpr (kBcast) __syncwarp();                        // the slab is rewritten by the next tile
        }
    };

    // The collected members wait in a ring (FIFO), so the tile after the current one is known as soon as it is
    // complete: its y-coordinates (a gather from the batch, usually a DRAM access) are loaded while the current tile
    // is evaluated, and stretches are filtered ahead of need so that there always is a next tile.
    int head = 0, tail = 0;                                            // warp-uniform
    long long ticket = fetch();
    T xs[S];
    load_stretch(ticket, xs);
    T y_next = gy0; bool have_next = false;
    for (;;) {
        while (tail - head < 64 && ticket < total) {                   // filter the next stretch
            const long long next = fetch();
            const int sweep = (int)(ticket / p.nstretch);
            const long long base = (ticket - (long long)sweep * p.nstretch) * (32 * S);
            const bool first = sweep == 0, last = sweep == p.nsweeps - 1;
            const T lo = first ? gx0 : s_bound[sweep - 1], hi = last ? gxl : s_bound[sweep];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const long long qi = base + s * 32 + lane;
                const bool member = qi < p.nq && (first || xs[s] >= lo) && (last || !(xs[s] >= hi));
                const unsigned mask = __ballot_sync(0xffffffffu, member);
                if (member) {
                    const int at = (tail + __popc(mask & ((1u << lane) - 1u))) & (QCAP - 1);
                    q_idx[wid][at] = (unsigned)qi; q_x[wid][at] = xs[s];
                }
                tail += __popc(mask);
            }
            __syncwarp();
            load_stretch(next, xs);                                    // in flight while tiles are evaluated
            ticket = next;
        }
        const int avail = tail - head;
        if (avail == 0) break;
        const bool live = lane < avail;                                // fewer than 32 only at the very end
        const int at = (head + lane) & (QCAP - 1);
        const unsigned qi = live ? q_idx[wid][at] : 0u; const T x = live ? q_x[wid][at] : gx0;
        const T y = have_next ? y_next : (live ? ld_query(p.qy + qi) : gy0);
        head += min(avail, 32);
        have_next = tail - head > 0;                                   // 32 or more unless the batch is exhausted
        if (have_next) {
            const bool live2 = lane < tail - head;
            y_next = live2 ? ld_query(p.qy + q_idx[wid][(head + lane) & (QCAP - 1)]) : gy0;
        }
        __syncwarp();                                                  // the slots may be pushed to again
        evaluate(live, qi, x, y);
    }
}

bool sweep_shape_ok(int64_t n, int64_t m, int64_t w, size_t elem, const void* data, const void* out, int64_t nq) {
    const int64_t row = w * (int64_t)elem;
    if (!(row == 16 || row == 32 || row == 64 || row == 128)) return false;
    if (n < 3 || n * m >= (1ll << 31) || nq >= (1ll << 32) || nq < 1) return false;
    return ((uintptr_t)data % 32) == 0 && ((uintptr_t)out % 16) == 0;
}

SweepPlan plan_sweeps(int64_t n, int64_t m, int64_t w, size_t elem, size_t band_bytes, int band_rows) {
    const size_t row = (size_t)m * (size_t)w * elem;                   // bytes of one x-row of the table
    int64_t rows = band_rows > 0 ? band_rows : (int64_t)(band_bytes / (row ? row : 1));
    if (rows < 1) rows = 1;
    const int64_t intervals = n - 1;
    int64_t sweeps = (intervals + rows - 1) / rows;
    if (sweeps > kMaxSweeps) { sweeps = kMaxSweeps; rows = (intervals + sweeps - 1) / sweeps; sweeps = (intervals + rows - 1) / rows; }
    else if (band_rows <= 0) { rows = (intervals + sweeps - 1) / sweeps; sweeps = (intervals + rows - 1) / rows; }   // even bands
    return SweepPlan{(int)rows, (int)sweeps};
}

template <class T>
cudaError_t launch_interp2d_bilinear_sweep(const T* gx, int64_t n, SearchCfg scx, const T* gy, int64_t m, SearchCfg scy,
                                           const T* data, int64_t w, const T* qx, const T* qy, int64_t nq, int extrapolate,
                                           T* out, unsigned long long* err, int fast_tables, SweepPlan sp,
                                           unsigned long long* next_task, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    constexpr int S = SweepShape<T>::S;
    Sweep2<T> p{gx, (int)n, scx, gy, (int)m, scy, data, (long long)w, qx, qy, (long long)nq, extrapolate, out, err,
                fast_tables, next_task, sp.nsweeps, sp.band_rows, ((long long)nq + 32 * S - 1) / (32 * S), 0};
    size_t smem = stage_bytes(scx, sizeof(T)) + stage_bytes(scy, sizeof(T));
    static const bool env_stage = !(getenv("NDI_SWEEP_STAGE_X") && atoi(getenv("NDI_SWEEP_STAGE_X")) == 0);   // measurement switch
    if (env_stage && scx.guess && !scx.smem && (size_t)n * sizeof(T) <= 32 * 1024) {
        p.stage_x = 1;
        smem += ((size_t)n * sizeof(T) + 15) & ~(size_t)15;
    }
    auto go = [&](auto kernel) -> cudaError_t {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cudaGetLastError(); fa.sharedSizeBytes = 40 * 1024; }
        if (fa.sharedSizeBytes + smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSwBlock, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        const long long want = (p.nstretch + kSwWarps - 1) / kSwWarps, cap = (long long)device_info().sm_count * per_sm;
        kernel<<<(int)(want < cap ? (want < 1 ? 1 : want) : cap), kSwBlock, smem, st>>>(p);
        count_launch();
        return cudaGetLastError();
    };
    switch ((int)(w * (int64_t)sizeof(T) / 16)) {
    case 1: return go(interp2d_bilinear_sweep_kernel<T, 1>);
    case 2: return go(interp2d_bilinear_sweep_kernel<T, 2>);
    case 4: return go(interp2d_bilinear_sweep_kernel<T, 4>);
    default: return go(interp2d_bilinear_sweep_kernel<T, 8>);
    }
}

#define NDI_INST_SWEEP(T)                                                                                                   \
    template cudaError_t launch_interp2d_bilinear_sweep<T>(const T*, int64_t, SearchCfg, const T*, int64_t, SearchCfg,      \
                                                           const T*, int64_t, const T*, const T*, int64_t, int, T*,         \
                                                           unsigned long long*, int, SweepPlan, unsigned long long*,        \
                                                           cudaStream_t);
NDI_INST_SWEEP(float)
NDI_INST_SWEEP(double)
NDI_INST_SWEEP(int32_t)
NDI_INST_SWEEP(int64_t)
NDI_INST_SWEEP(uint32_t)
NDI_INST_SWEEP(uint64_t)

}  // namespace ndi
