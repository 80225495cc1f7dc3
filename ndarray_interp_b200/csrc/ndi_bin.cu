// ndi_bin.cu -- K8: locality binning of a 2-D query batch by table band.
//
// Why.  Bilinear::interp_into (bilinear.rs:83-97) gathers four cells per query.  With random
// queries on a table that does not fit in L2 (C4: 134 MB, C5a: 2.1 GB) nearly every gather is a
// DRAM access of its own: ncu shows 2.5 GB (C4) / 17 GB (C5a) of table reads for tables of
// 0.13 / 2.1 GB (profiles/r01/ncu_c4_*.txt, ncu_c5a_*.txt).  The batch loop of the reference
// (interp2d/mod.rs:255-307) is order-independent -- every query writes its own output row -- so the
// launch may evaluate the queries in any order as long as row q of the output belongs to query q.
//
// What.  The x-axis is cut into BANDS of 2^band_shift grid intervals, sized so that the table rows
// of one band fit comfortably in L2.  A counting sort groups the queries by band:
//   pass 1  bin_totals_kernel   reads qx, counts queries per band (shared-memory histogram per block,
//                               one global atomic per band per block)
//   pass 2  bin_scatter_kernel  reads qx, qy; every block sorts its chunk of 2048 queries by band in
//                               shared memory, reserves a run in each band with one atomic per band,
//                               and writes (original index, qx, qy) with coalesced stores
//   pass 3  the bilinear kernel (ndi_eval.cu) walks the binned arrays in order; at any time the
//           persistent grid works inside one band (two at a boundary), so the table is read from
//           DRAM once and gathered from L2, and row perm[i] of the output receives the result.
// Extra HBM traffic: s*Q (pass 1) + 2s*Q + (4+2s)*Q (pass 2) + (4+2s)*Q - 2s*Q (pass 3 reads the
// binned copy instead of the original), i.e. 28 B per f32 query, against up to 4 * max(32, s*w) B of
// gather traffic saved.  The order of queries inside a band depends on block scheduling; the
// results do not (each output row depends on its own query only).
#include "ndi_device.cuh"
#include "ndi_internal.h"

namespace ndi {

constexpr int kBinBlock = 256;
constexpr int kBinPerThread = 8;
constexpr int kBinChunk = kBinBlock * kBinPerThread;

template <class T>
struct BinArgs {
    const T* gx; int n; SearchCfg scx;
    const T* qx; const T* qy; long long nq;
    int band_shift, nbands;
    long long nchunks;
    unsigned* totals;     // [nbands] queries per band (filled by pass 1)
    unsigned* cursor;     // [nbands] slots of each band already handed out (pass 2)
    unsigned* perm; T* bqx; T* bqy;
};

// Band of each of K queries.  Binning only has to keep neighbours together, so the band may come
// from an APPROXIMATE interval index as long as both passes use the same function (they do): on an
// evenly spaced grid the index is arithmetic, with a bucket table it is the table entry without the
// finishing bisection; otherwise the exact search runs.  NaN and out-of-range queries land in some
// valid band and are dealt with by the evaluation kernel.
template <class T, int K>
__device__ __forceinline__ void bands_of(const GridView<T>& g, const T (&x)[K], int shift, int (&band)[K]) {
    typedef typename std::conditional<std::is_same<T, float>::value, float, double>::type F;
    if (g.mode == SEARCH_GUESS) {
        const F g0 = (F)g.g0, inv = (F)(g.n - 1) / ((F)g.gl - (F)g.g0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const F f = ((F)x[k] - g0) * inv;
            band[k] = f > (F)0 ? (f < (F)(g.n - 2) ? (int)f : g.n - 2) : 0;      // NaN -> 0
        }
    } else if (g.mode == SEARCH_LUT) {
#pragma unroll
        for (int k = 0; k < K; ++k)
            band[k] = lut_tag_index(LutEntry<T>::load_tag(g.lut, bucket_of<T>(x[k], g.g0d, g.scale, g.nb)));
    } else {
        T vlo[K], vhi[K];
        search_multi<T, K>(g, x, band, vlo, vhi);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) band[k] = min(max(band[k], 0), g.n - 2) >> shift;
}

// Counting is one shared-memory atomic per query.  (Measured: aggregating equal bands of a warp
// with __match_any_sync first is slower -- MATCH.ANY serialises over the distinct values, 229 vs
// 95 us for the 2^25 queries / 64 bands of C5a.)  Dead lanes use the dummy counter kMaxBands.

template <class T>
__global__ void __launch_bounds__(kBinBlock) bin_totals_kernel(const BinArgs<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ unsigned hist[kMaxBands + 1];
    const GridView<T> g = make_grid_view<T>(p.gx, p.n, p.scx, smem_raw, &bar);
    for (int i = threadIdx.x; i <= kMaxBands; i += kBinBlock) hist[i] = 0;
    __syncthreads();
    for (long long chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
        const long long base = chunk * kBinChunk;
        T x[kBinPerThread]; int band[kBinPerThread];
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k) {
            const long long i = base + k * kBinBlock + threadIdx.x;
            x[k] = i < p.nq ? __ldg(p.qx + i) : g.g0;
        }
        bands_of<T, kBinPerThread>(g, x, p.band_shift, band);
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k)
            atomicAdd(&hist[base + k * kBinBlock + threadIdx.x < p.nq ? band[k] : kMaxBands], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.nbands; i += kBinBlock)
        if (hist[i]) atomicAdd(p.totals + i, hist[i]);
}

// exclusive scan of v over the 256 threads of the block (scratch: 8 words of shared memory)
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* warp_sums) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    unsigned prefix = 0;
#pragma unroll
    for (int i = 0; i < kBinBlock / 32; ++i) prefix += i < wid ? warp_sums[i] : 0u;
    __syncthreads();
    return prefix + inc - v;
}

template <class T>
__global__ void __launch_bounds__(kBinBlock) bin_scatter_kernel(const BinArgs<T> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ unsigned hist[kMaxBands + 1], lbase[kMaxBands], gbase[kMaxBands], band_base[kMaxBands], warp_sums[kBinBlock / 32];
    __shared__ unsigned sidx[kBinChunk];
    __shared__ T sx[kBinChunk], sy[kBinChunk];
    __shared__ unsigned char sband[kBinChunk];
    static_assert(kMaxBands <= kBinBlock, "one thread per band in the scans");
    const GridView<T> g = make_grid_view<T>(p.gx, p.n, p.scx, smem_raw, &bar);
    {   // first slot of every band = exclusive scan of the totals of pass 1
        const unsigned v = (int)threadIdx.x < p.nbands ? p.totals[threadIdx.x] : 0u;
        const unsigned e = block_exclusive_scan(v, warp_sums);
        if (threadIdx.x < kMaxBands) { band_base[threadIdx.x] = e; hist[threadIdx.x] = 0; }
    }
    __syncthreads();
    for (long long chunk = blockIdx.x; chunk < p.nchunks; chunk += gridDim.x) {
        const long long base = chunk * kBinChunk;
        T x[kBinPerThread], y[kBinPerThread]; int band[kBinPerThread]; unsigned rank[kBinPerThread];
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k) {
            const long long i = base + k * kBinBlock + threadIdx.x;
            const bool live = i < p.nq;
            x[k] = live ? ld_query(p.qx + i) : g.g0;
            y[k] = live ? ld_query(p.qy + i) : g.g0;
        }
        bands_of<T, kBinPerThread>(g, x, p.band_shift, band);
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k)
            rank[k] = atomicAdd(&hist[base + k * kBinBlock + threadIdx.x < p.nq ? band[k] : kMaxBands], 1u);
        __syncthreads();
        {   // where each band starts inside this chunk, and the run this chunk gets in each band
            const unsigned v = (int)threadIdx.x < p.nbands ? hist[threadIdx.x] : 0u;
            const unsigned e = block_exclusive_scan(v, warp_sums);
            if (threadIdx.x < kMaxBands) {
                lbase[threadIdx.x] = e;
                if (v) gbase[threadIdx.x] = band_base[threadIdx.x] + atomicAdd(p.cursor + threadIdx.x, v);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kBinPerThread; ++k) {
            const long long i = base + k * kBinBlock + threadIdx.x;
            if (i < p.nq) {
                const unsigned pos = lbase[band[k]] + rank[k];
                sidx[pos] = (unsigned)i; sx[pos] = x[k]; sy[pos] = y[k]; sband[pos] = (unsigned char)band[k];
            }
        }
        __syncthreads();
        const int cnt = (int)min((long long)kBinChunk, p.nq - base);
        for (int s = threadIdx.x; s < cnt; s += kBinBlock) {
            const int b = sband[s];
            const unsigned dst = gbase[b] + ((unsigned)s - lbase[b]);
            p.perm[dst] = sidx[s]; p.bqx[dst] = sx[s]; p.bqy[dst] = sy[s];
        }
        __syncthreads();
        if (threadIdx.x < kMaxBands) hist[threadIdx.x] = 0;
        __syncthreads();
    }
}

BandPlan plan_bands(int64_t n, int64_t m, int64_t w, size_t elem, size_t band_bytes, int band_rows) {
    BandPlan bp{0, 1};
    const size_t row = (size_t)m * (size_t)w * elem;              // bytes of one x-row of the table
    int shift = 0;
    if (band_rows > 0) { while ((1ll << (shift + 1)) <= band_rows) ++shift; }
    else { while (((size_t)2 << shift) * row <= band_bytes) ++shift; }
    auto bands = [&](int sh) { return (int)(((n - 1) + (1ll << sh) - 1) >> sh); };
    while (bands(shift) > kMaxBands) ++shift;
    bp.band_shift = shift; bp.nbands = bands(shift);
    return bp;
}

size_t bin_scratch_bytes(int64_t nq, size_t elem) {
    const size_t q = ((size_t)nq + 63) & ~(size_t)63;
    return (2 * kMaxBands + 4) * sizeof(unsigned) + q * (sizeof(unsigned) + 2 * elem);     // counters, task counter, perm, qx, qy
}

template <class T>
cudaError_t launch_bin_queries(const T* gx, int64_t n, SearchCfg scx, const T* qx, const T* qy, int64_t nq,
                               BandPlan bp, void* scratch, const unsigned** perm, const T** bqx, const T** bqy,
                               unsigned long long** next_task, cudaStream_t st) {
    const size_t q = ((size_t)nq + 63) & ~(size_t)63;
    unsigned* counters = static_cast<unsigned*>(scratch);
    unsigned* perm_w = counters + 2 * kMaxBands + 4;             // 16 bytes in between: the evaluation's tile counter
    T* bx = reinterpret_cast<T*>(perm_w + q);
    T* by = bx + q;
    *next_task = reinterpret_cast<unsigned long long*>(counters + 2 * kMaxBands);
    cudaError_t e = cudaMemsetAsync(counters, 0, (2 * kMaxBands + 4) * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    BinArgs<T> p{gx, (int)n, scx, qx, qy, (long long)nq, bp.band_shift, bp.nbands,
                 ((long long)nq + kBinChunk - 1) / kBinChunk, counters, counters + kMaxBands, perm_w, bx, by};
    const size_t smem = stage_bytes(scx, sizeof(T));
    auto blocks = [&](auto kernel) {
        // the kernels also hold their chunk in STATIC shared memory (up to 47 KB): static + dynamic beyond the
        // 48 KB default needs the opt-in, not only a large dynamic part
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cudaGetLastError(); fa.sharedSizeBytes = 48 * 1024; }
        if (fa.sharedSizeBytes + smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBinBlock, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        const long long cap = (long long)device_info().sm_count * per_sm;
        return (int)(p.nchunks < cap ? p.nchunks : cap);
    };
    bin_totals_kernel<T><<<blocks(bin_totals_kernel<T>), kBinBlock, smem, st>>>(p);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    bin_scatter_kernel<T><<<blocks(bin_scatter_kernel<T>), kBinBlock, smem, st>>>(p);
    count_launch();
    *perm = perm_w; *bqx = bx; *bqy = by;
    return cudaGetLastError();
}

#define NDI_INST_BIN(T)                                                                                              \
    template cudaError_t launch_bin_queries<T>(const T*, int64_t, SearchCfg, const T*, const T*, int64_t, BandPlan, \
                                               void*, const unsigned**, const T**, const T**, unsigned long long**, \
                                               cudaStream_t);
NDI_INST_BIN(float)
NDI_INST_BIN(double)
NDI_INST_BIN(int32_t)
NDI_INST_BIN(int64_t)
NDI_INST_BIN(uint32_t)
NDI_INST_BIN(uint64_t)

}  // namespace ndi
