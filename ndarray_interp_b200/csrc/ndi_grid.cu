// ndi_grid.cu -- K1: VectorExtensions::monotonic_prop on the device
// (src/vector_extensions.rs:40-53 and the MonotonicState machine at :115-198).
//
// The reference folds a state machine over windows(2) with an early exit.  The machine has
// seven states and the input alphabet is the ordering of one adjacent pair {<, ==, >, unordered},
// so the fold is a composition of maps {0..6} -> {0..6}.  Composition is associative, which makes
// the classification an ordered parallel reduction that is EXACT for every input, including the
// state-dependent NaN behaviour (an unordered pair means "Falling" from Init/NotStrict but
// "NotMonotonic" from a Likely state, :136-170) that a flag-OR reduction would get wrong.
// A map is packed 3 bits per state into one 32-bit word.
//
// The same launch also answers whether the even-spacing guess of get_lower_index
// (:68-90) hits on every cell, which the evaluation kernels use to pick the O(1) search path.
#include <type_traits>

#include "ndi_device.cuh"
#include "ndi_internal.h"

namespace ndi {

enum : uint32_t { S_INIT = 0, S_NOT_STRICT = 1, S_R_STRICT = 2, S_R = 3, S_F_STRICT = 4, S_F = 5, S_NM = 6 };

__host__ __device__ constexpr uint32_t pack7(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g) {
    return a | (b << 3) | (c << 6) | (d << 9) | (e << 12) | (f << 15) | (g << 18);
}
//                                        Init        NotStrict  R_strict   R        F_strict   F     NM
constexpr uint32_t MAP_ID = pack7(S_INIT,     S_NOT_STRICT, S_R_STRICT, S_R,  S_F_STRICT, S_F,  S_NM);
constexpr uint32_t MAP_LT = pack7(S_R_STRICT, S_R,          S_R_STRICT, S_R,  S_NM,       S_NM, S_NM);   // a <  b
constexpr uint32_t MAP_EQ = pack7(S_NOT_STRICT, S_NOT_STRICT, S_R,      S_R,  S_F,        S_F,  S_NM);   // a == b
constexpr uint32_t MAP_GT = pack7(S_F_STRICT, S_F,          S_NM,       S_NM, S_F_STRICT, S_F,  S_NM);   // a >  b
constexpr uint32_t MAP_UN = pack7(S_F_STRICT, S_F,          S_NM,       S_NM, S_NM,       S_NM, S_NM);   // unordered

__device__ __forceinline__ uint32_t map_apply(uint32_t f, uint32_t s) { return (f >> (3 * s)) & 7u; }
// (f then g)
__device__ __forceinline__ uint32_t map_compose(uint32_t f, uint32_t g) {
    uint32_t r = 0;
#pragma unroll
    for (uint32_t s = 0; s < 7; ++s) r |= map_apply(g, map_apply(f, s)) << (3 * s);
    return r;
}
template <class T>
__device__ __forceinline__ uint32_t pair_map(T a, T b) {
    if (a < b) return MAP_LT;
    if (a == b) return MAP_EQ;
    if (a > b) return MAP_GT;
    return MAP_UN;
}

constexpr int kGridBlock = 256;
constexpr int kGridMaxBlocks = 256;     // partial maps + 2 control words fit the scratch

size_t grid_classify_scratch_words() { return kGridMaxBlocks + 4; }

// ordered block reduction of maps and AND-reduction of the guess flag
__device__ __forceinline__ void block_reduce(uint32_t& f, uint32_t& ok, uint32_t* sh_f, uint32_t* sh_ok) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // ordered warp reduction: after step o, lane l holds the composition of lanes [l, l+2o)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t g = __shfl_down_sync(0xffffffffu, f, o);
        if (lane + o < 32) f = map_compose(f, g);
        ok &= __shfl_down_sync(0xffffffffu, ok, o) | (lane + o < 32 ? 0u : 1u);
    }
    if (lane == 0) { sh_f[warp] = f; sh_ok[warp] = ok; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = MAP_ID, a_ok = 1;
        for (int w = 0; w < kGridBlock / 32; ++w) { acc = map_compose(acc, sh_f[w]); a_ok &= sh_ok[w]; }
        f = acc; ok = a_ok;
    }
}

// scratch: [0..kGridMaxBlocks) partial maps, [kGridMaxBlocks] ticket, [kGridMaxBlocks+1] guess AND
template <class T>
__global__ void __launch_bounds__(kGridBlock) grid_classify_kernel(const T* __restrict__ x, long long n,
                                                                  int32_t* __restrict__ result, uint32_t* scratch) {
    __shared__ uint32_t sh_f[kGridBlock / 32], sh_ok[kGridBlock / 32];
    __shared__ bool last;
    const long long npairs = n - 1;
    // contiguous chunk of pairs per block, contiguous sub-chunk per thread (order matters)
    const long long per_block = (npairs + gridDim.x - 1) / gridDim.x;
    const long long b_lo = per_block * blockIdx.x, b_hi = min(npairs, b_lo + per_block);
    const long long per_thread = (per_block + kGridBlock - 1) / kGridBlock;
    const long long t_lo = min(b_hi, b_lo + per_thread * threadIdx.x), t_hi = min(b_hi, t_lo + per_thread);
    uint32_t f = MAP_ID, ok = 1;
    if (t_lo < t_hi) {
        T a = x[t_lo];
        for (long long i = t_lo; i < t_hi; ++i) {
            const T b = x[i + 1];
            f = map_compose(f, pair_map<T>(a, b));
            // does the O(1) guess of get_lower_index land in cell i for a point inside it?
            if (n <= 0x7fffffff) {
                T mid;
                if constexpr (std::is_same_v<T, double>) mid = 0.5 * (a + b);
                else if constexpr (std::is_same_v<T, float>) mid = 0.5f * (a + b);
                else if constexpr (sizeof(T) == 8) mid = a < b ? (T)(a + (T)(((unsigned long long)b - (unsigned long long)a) / 2)) : a;   // a + b may overflow
                else mid = (T)(((long long)a + (long long)b) / 2);
                if (a <= mid && mid < b) {
                    T g0 = x[0], gl = x[n - 1];
                    T est = calc_frac<T>(g0, (T)0, gl, (T)(n - 1), mid);
                    if (!(est >= (T)0 && est < (T)(n - 1) && (long long)est == i)) ok = 0;
                } else if (!(a < b)) ok = 0;
            } else ok = 0;
            a = b;
        }
    }
    block_reduce(f, ok, sh_f, sh_ok);
    if (threadIdx.x == 0) {
        scratch[blockIdx.x] = f;
        if (!ok) atomicAnd(&scratch[kGridMaxBlocks + 1], 0u);
        __threadfence();
        uint32_t ticket = atomicAdd(&scratch[kGridMaxBlocks], 1u);
        last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        uint32_t acc = MAP_ID;
        for (unsigned b = 0; b < gridDim.x; ++b) acc = map_compose(acc, ((volatile uint32_t*)scratch)[b]);
        const uint32_t s = map_apply(acc, S_INIT);
        int32_t mono;                                   // MonotonicState::finish (:191-197)
        switch (s) {
        case S_R_STRICT: mono = 1; break;
        case S_R:        mono = 2; break;
        case S_F_STRICT: mono = 3; break;
        case S_F:        mono = 4; break;
        default:         mono = 0; break;               // NotStrict (all equal) and NM -> NotMonotonic
        }
        result[0] = mono;
        result[1] = (mono == 1 && ((volatile uint32_t*)scratch)[kGridMaxBlocks + 1] != 0) ? 1 : 0;
    }
}

template <class T>
cudaError_t launch_grid_classify(const T* x, int64_t n, int32_t* result_dev, uint32_t* scratch_dev, cudaStream_t st) {
    // ticket = 0, guess flag = all ones; partial maps are fully overwritten
    cudaError_t e = cudaMemsetAsync(scratch_dev + kGridMaxBlocks, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(scratch_dev + kGridMaxBlocks + 1, 0xff, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    if (n <= 1) return cudaMemsetAsync(result_dev, 0, 2 * sizeof(int32_t), st);   // :41-43
    long long npairs = n - 1;
    long long want = (npairs + (long long)kGridBlock * 8 - 1) / ((long long)kGridBlock * 8);   // >= 8 pairs per thread
    int blocks = (int)(want < 1 ? 1 : (want > kGridMaxBlocks ? kGridMaxBlocks : want));
    grid_classify_kernel<T><<<blocks, kGridBlock, 0, st>>>(x, (long long)n, result_dev, scratch_dev);
    count_launch();
    return cudaGetLastError();
}

// ---- bucket table for the O(1) search (see ndi_device.cuh, SEARCH_LUT) ----------------------------
// count(b) = number of grid points whose bucket is < b; bucket_of(g[i]) is non-decreasing in i, so
// count(b) is a lower bound found by bisection.  One thread per bucket.
template <class T>
__global__ void __launch_bounds__(256) build_lut_kernel(const T* __restrict__ x, int n, double g0d, double scale, int nb,
                                                        typename LutEntry<T>::type* __restrict__ lut) {
    typedef LutEntry<T> L;
    typedef typename L::type Entry;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    auto count_below = [&](int bb) {
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (bucket_of<T>(x[mid], g0d, scale, nb) < bb) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int c0 = count_below(b), c1 = count_below(b + 1);
    Entry e;
    // the index every query of this bucket gets when the bucket holds no grid point, or only g[0] / g[n-1]
    // (get_lower_index clamps to 0 and n-2 on both sides of those two, vector_extensions.rs:61-66)
    int flat = -1;
    if (c0 == c1) flat = min(max(c0 - 1, 0), n - 2);
    else if (c1 == c0 + 1 && c0 == 0) flat = 0;
    else if (c1 == c0 + 1 && c0 == n - 1) flat = n - 2;
    if (flat >= 0 && flat < kLutOnePoint) {                                   // kind A
        e.tag = flat; e.v0 = x[flat]; e.v1 = x[flat + 1]; e.v2 = x[flat + 1];
    } else if (c1 == c0 + 1 && c0 < kLutOnePoint) {                           // kind B: 1 <= c0 <= n-2 here
        e.tag = c0 | kLutOnePoint; e.v0 = x[c0 - 1]; e.v1 = x[c0]; e.v2 = x[c0 + 1];
    } else {                                                                  // kind C
        e.tag = -(c0 + 1); e.v0 = L::as_count(c1); e.v1 = e.v0; e.v2 = e.v0;
    }
    lut[b] = e;
}

size_t lut_entry_bytes(size_t elem) { return elem == 4 ? 16 : 32; }

template <class T>
cudaError_t launch_build_lut(const T* x, int64_t n, double g0d, double scale, int nb, void* lut_dev, cudaStream_t st) {
    build_lut_kernel<T><<<(nb + 255) / 256, 256, 0, st>>>(x, (int)n, g0d, scale, nb,
                                                        static_cast<typename LutEntry<T>::type*>(lut_dev));
    count_launch();
    return cudaGetLastError();
}
template cudaError_t launch_build_lut<float>(const float*, int64_t, double, double, int, void*, cudaStream_t);
template cudaError_t launch_build_lut<double>(const double*, int64_t, double, double, int, void*, cudaStream_t);
template cudaError_t launch_build_lut<int32_t>(const int32_t*, int64_t, double, double, int, void*, cudaStream_t);
template cudaError_t launch_build_lut<int64_t>(const int64_t*, int64_t, double, double, int, void*, cudaStream_t);
template cudaError_t launch_build_lut<uint32_t>(const uint32_t*, int64_t, double, double, int, void*, cudaStream_t);
template cudaError_t launch_build_lut<uint64_t>(const uint64_t*, int64_t, double, double, int, void*, cudaStream_t);

template cudaError_t launch_grid_classify<float>(const float*, int64_t, int32_t*, uint32_t*, cudaStream_t);
template cudaError_t launch_grid_classify<double>(const double*, int64_t, int32_t*, uint32_t*, cudaStream_t);
template cudaError_t launch_grid_classify<int32_t>(const int32_t*, int64_t, int32_t*, uint32_t*, cudaStream_t);
template cudaError_t launch_grid_classify<int64_t>(const int64_t*, int64_t, int32_t*, uint32_t*, cudaStream_t);
template cudaError_t launch_grid_classify<uint32_t>(const uint32_t*, int64_t, int32_t*, uint32_t*, cudaStream_t);
template cudaError_t launch_grid_classify<uint64_t>(const uint64_t*, int64_t, int32_t*, uint32_t*, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// Dense copy of a strided view (ndi_interp*_create_strided): one thread per destination element,
// destination writes coalesced, source reads as coalesced as the view's innermost stride allows.
// ------------------------------------------------------------------------------------------------
template <class U>
__global__ void __launch_bounds__(256) pack_strided_kernel(const U* __restrict__ src, long long origin, StridedDesc d,
                                                           long long count, U* __restrict__ dst) {
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += step) {
        long long rest = i, off = origin;
#pragma unroll 1
        for (int k = d.ndim - 1; k >= 0; --k) {
            const long long q = rest / d.shape[k];
            off += (rest - q * d.shape[k]) * d.stride[k];
            rest = q;
        }
        dst[i] = src[off];
    }
}

cudaError_t launch_pack_strided(const void* src_dev, long long origin, const StridedDesc& d, size_t elem,
                                long long count, void* dst_dev, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const long long want = (count + 255) / 256, cap = (long long)device_info().sm_count * 16;
    const int blocks = (int)(want < cap ? want : cap);
    if (elem == 8) pack_strided_kernel<unsigned long long><<<blocks, 256, 0, st>>>((const unsigned long long*)src_dev, origin, d, count, (unsigned long long*)dst_dev);
    else pack_strided_kernel<unsigned><<<blocks, 256, 0, st>>>((const unsigned*)src_dev, origin, d, count, (unsigned*)dst_dev);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// May the evaluation kernels divide with a per-query reciprocal (ndi_device.cuh, div_by)?  Yes when
// every table value is finite and either 0 or of magnitude in [2^-56, 2^30].  *flag must be 1 on entry.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_fast_div_kernel(const float* __restrict__ data, size_t count, int32_t* flag) {
    bool ok = true;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint32_t u = __float_as_uint(__ldg(data + i)) & 0x7fffffffu;
        // biased exponents: 2^-56 -> 71, 2^30 -> 157
        ok = ok && (u == 0u || (u >= (71u << 23) && u <= (157u << 23)));
    }
    if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}

cudaError_t launch_table_fast_div(const float* data, size_t count, int32_t* flag_dev, cudaStream_t st) {
    const size_t want = (count + 255) / 256;
    const size_t cap = (size_t)device_info().sm_count * 16;
    table_fast_div_kernel<<<(unsigned)(want < cap ? (want ? want : 1) : cap), 256, 0, st>>>(data, count, flag_dev);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// ndi_selftest_fdiv: div_by(a, b, rcp_refined(b)) against __fdiv_rn(a, b) for
//   a = (1.m_a) * 2^a_exp with m_a in [a_mant_begin, a_mant_begin + a_mant_count), and
//   b = (1.m_b) * 2^b_exp for ALL 2^23 mantissas m_b.
// One block per 2^12 values of m_b; the reciprocal is formed once per b as in the kernels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) selftest_fdiv_kernel(uint32_t a_mant_begin, uint32_t a_mant_count, int a_exp, int b_exp,
                                                           unsigned long long* mismatches) {
    unsigned long long bad = 0;
    const uint32_t a_hi = (uint32_t)(a_exp + 127) << 23, b_hi = (uint32_t)(b_exp + 127) << 23;
    for (uint32_t k = 0; k < 16; ++k) {
        const uint32_t mb = blockIdx.x * 4096u + k * 256u + threadIdx.x;
        const float b = __uint_as_float(b_hi | mb);
        const float r = rcp_refined(b);
        // zero numerators and both signs through the spline sweeps' entry point (sign of zero included)
        for (uint32_t v = 0; v < 4; ++v) {
            const float z = __uint_as_float((v & 1u) << 31), bb = (v & 2u) ? -b : b;
            bad += __float_as_uint(Hoisted<float>::div(z, bb, Hoisted<float>::rcp(bb))) != __float_as_uint(__fdiv_rn(z, bb));
        }
        for (uint32_t ma = a_mant_begin; ma < a_mant_begin + a_mant_count; ++ma) {
            const float a = __uint_as_float(a_hi | ma);
            bad += __float_as_uint(div_by(a, b, r)) != __float_as_uint(__fdiv_rn(a, b));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

cudaError_t launch_selftest_fdiv(uint32_t a_mant_begin, uint32_t a_mant_count, int a_exp, int b_exp,
                                 unsigned long long* mismatches_dev, cudaStream_t st) {
    selftest_fdiv_kernel<<<(1u << 23) / 4096u, 256, 0, st>>>(a_mant_begin, a_mant_count, a_exp, b_exp, mismatches_dev);
    count_launch();
    return cudaGetLastError();
}

// ndi_selftest_ddiv: Hoisted<double>::div against __ddiv_rn on pseudo-random operand pairs (splitmix64
// bit patterns: every exponent, sign and mantissa; plus pairs with equal or nearly equal exponents).
__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) selftest_ddiv_kernel(uint64_t seed, int per_thread, unsigned long long* mismatches) {
    uint64_t st = seed + 0x632BE59BD9B4E019ull * ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long bad = 0;
    for (int it = 0; it < per_thread; ++it) {
        uint64_t ub = splitmix64(st);
        if (it & 1) ub = (ub & 0x800FFFFFFFFFFFFFull) | ((uint64_t)(1023 - 400 + (int)(splitmix64(st) % 800)) << 52);   // in Hoisted's range
        const double b = __longlong_as_double((long long)ub);
        const double r = Hoisted<double>::rcp(b);
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            uint64_t ua = splitmix64(st);
            if (j & 1) ua = (ua & 0x800FFFFFFFFFFFFFull) | (ub & 0x7FF0000000000000ull);     // same exponent as b
            const double a = __longlong_as_double((long long)ua);
            const double q = Hoisted<double>::div(a, b, r), ref = __ddiv_rn(a, b);
            bad += __double_as_longlong(q) != __double_as_longlong(ref) && !(q != q && ref != ref);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}
cudaError_t launch_selftest_ddiv(uint64_t seed, int blocks, int per_thread, unsigned long long* mismatches_dev, cudaStream_t st) {
    selftest_ddiv_kernel<<<blocks, 256, 0, st>>>(seed, per_thread, mismatches_dev);
    count_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Tail of the scalar / tiny-batch latency path: hands the error word to the host through mapped pinned
// memory, re-arms it, and raises the completion flag the host is spinning on (no driver synchronisation).
// ------------------------------------------------------------------------------------------------
__global__ void publish_kernel(unsigned long long* d_err, volatile unsigned long long* h_err,
                               volatile unsigned long long* h_flag, unsigned long long seq) {
    *h_err = *d_err;
    *d_err = ~0ull;
    __threadfence_system();
    *h_flag = seq;
}
cudaError_t launch_publish(unsigned long long* d_err, void* h_err, void* h_flag, unsigned long long seq, cudaStream_t st) {
    publish_kernel<<<1, 1, 0, st>>>(d_err, static_cast<volatile unsigned long long*>(h_err),
                                    static_cast<volatile unsigned long long*>(h_flag), seq);
    count_launch();
    return cudaGetLastError();
}

}  // namespace ndi
