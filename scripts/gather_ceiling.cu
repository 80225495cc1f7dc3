// gather_ceiling.cu -- what the B200 memory system gives the ACCESS PATTERN of the thin-row kernels, with the
// search and the arithmetic taken out: per query one 4-byte index read (coalesced), `segs` contiguous segments
// of `seg_bytes` gathered from a table at a random row, `out_bytes` written with a streaming store.  Same
// decomposition as csrc/ndi_eval.cu (32 queries per warp iteration, LPQ lanes per query, 16-byte vectors,
// persistent grid), so the number is the ceiling of that design, not of a different kernel.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/gather_ceiling.bin scripts/gather_ceiling.cu
//   scripts/gather_ceiling.bin            (prints one JSON line per shape)
//
// Shapes: C3 (4 MB table, two 64-byte rows = one 128-byte segment, 64-byte result), C4 (134 MB table, two
// 64-byte segments one x-row apart, 32-byte result), C5a (2.1 GB table, two 256-byte segments, 128-byte result;
// random and band-sorted indices).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// LPQ lanes per query; each lane moves 16 bytes per segment part.  seg_bytes = parts * LPQ * 16 where `parts` is
// how many 16-byte loads a lane issues per segment (1 or 2: two rows of a linear query are one segment of 2 parts).
template <int LPQ, int PARTS, int SEGS>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ idx, const int4* __restrict__ table,
                                                      long long row_vecs, long long seg2_vecs, int4* __restrict__ out,
                                                      long long nq) {
    constexpr int QPR = 32 / LPQ;
    const int lane = threadIdx.x & 31, sub = lane % LPQ, qsel = lane / LPQ;
    const long long nwarps = (long long)gridDim.x * 8, ntiles = (nq + 31) / 32;
    long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    uint32_t ahead = tile * 32 + lane < nq ? __ldcs(idx + tile * 32 + lane) : 0u;
    for (; tile < ntiles; tile += nwarps) {
        const uint32_t mine = ahead;
        const long long qn = (tile + nwarps) * 32 + lane;
        ahead = qn < nq ? __ldcs(idx + qn) : 0u;
#pragma unroll
        for (int r = 0; r < LPQ; ++r) {
            const int src = r * QPR + qsel;
            const uint32_t row = __shfl_sync(0xffffffffu, mine, src);
            const long long q = tile * 32 + src;
            if (q >= nq) continue;
            int4 acc = make_int4(0, 0, 0, 0);
#pragma unroll
            for (int s = 0; s < SEGS; ++s) {
#pragma unroll
                for (int p = 0; p < PARTS; ++p) {
                    const int4 v = __ldg(table + (long long)row * row_vecs + s * seg2_vecs + p * LPQ + sub);
                    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                }
            }
            __stcs(out + q * LPQ + sub, acc);
        }
    }
}

// Candidate table layouts for thin bilinear rows (C4: 8 f32 columns), one lane per output column:
//   QUAD  per cell (i, j) and column c the four corners {z11, z12, z21, z22} in 16 bytes: one aligned 128-byte
//         segment per query, 4 x the table;
//   PAIR  per cell and column {z[i][j][c], z[i+1][j][c]} in 8 bytes: cells (i, j) and (i, j+1) are one contiguous
//         128-byte segment at 64-byte alignment, 2 x the table.
// Each lane XORs what it loaded into 4 bytes and stores them (32 bytes per query, like C4's result row).
template <bool QUAD>
__global__ void __launch_bounds__(256) gather_layout_kernel(const uint32_t* __restrict__ idx, const void* __restrict__ table,
                                                            uint32_t* __restrict__ out, long long nq) {
    const int lane = threadIdx.x & 31, sub = lane & 7, qsel = lane >> 3;
    const long long nwarps = (long long)gridDim.x * 8, ntiles = (nq + 31) / 32;
    long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    uint32_t ahead = tile * 32 + lane < nq ? __ldcs(idx + tile * 32 + lane) : 0u;
    for (; tile < ntiles; tile += nwarps) {
        const uint32_t mine = ahead;
        const long long qn = (tile + nwarps) * 32 + lane;
        ahead = qn < nq ? __ldcs(idx + qn) : 0u;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int src = r * 4 + qsel;
            const uint32_t cell = __shfl_sync(0xffffffffu, mine, src);
            const long long q = tile * 32 + src;
            if (q >= nq) continue;
            uint32_t acc;
            if (QUAD) {
                const int4 v = __ldg(static_cast<const int4*>(table) + (long long)cell * 8 + sub);
                acc = v.x ^ v.y ^ v.z ^ v.w;
            } else {
                const int2 a = __ldg(static_cast<const int2*>(table) + (long long)cell * 8 + sub);
                const int2 b = __ldg(static_cast<const int2*>(table) + (long long)(cell + 1) * 8 + sub);
                acc = a.x ^ a.y ^ b.x ^ b.y;
            }
            __stcs(out + q * 8 + sub, acc);
        }
    }
}

template <bool QUAD>
static void run_layout(const char* name, size_t table_bytes, size_t cells, long long nq, int sms) {
    std::vector<uint32_t> h((size_t)nq);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s % (cells - 2)); }
    uint32_t *idx, *out; void* table;
    CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMalloc(&table, table_bytes + 4096)); CK(cudaMalloc(&out, (size_t)nq * 32));
    CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(table, 1, table_bytes + 4096));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_layout_kernel<QUAD>, 256, 0));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f, sum = 0; const int reps = 12;
    for (int i = 0; i < 3 + reps; ++i) {
        CK(cudaEventRecord(a));
        gather_layout_kernel<QUAD><<<sms * per_sm, 256>>>(idx, table, out, nq);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (i >= 3) { best = std::min(best, ms); sum += ms; }
    }
    CK(cudaGetLastError());
    printf("{\"shape\": \"%s\", \"ms_mean\": %.4f, \"ms_best\": %.4f, \"algorithmic_GBps\": %.1f, \"blocks_per_sm\": %d}\n",
           name, sum / reps, best, 0.805e9 / (sum / reps * 1e-3) / 1e9, per_sm);
    fflush(stdout);
    CK(cudaFree(idx)); CK(cudaFree(table)); CK(cudaFree(out));
}

struct Shape { const char* name; size_t table_bytes; int row_bytes; int lpq, parts, segs; long long seg2_rows; long long nq; int sorted_bands; double algo_bytes; };

template <int LPQ, int PARTS, int SEGS>
static float run(const Shape& sh, const uint32_t* idx, const int4* table, int4* out, int sms) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_kernel<LPQ, PARTS, SEGS>, 256, 0));
    const int grid = sms * per_sm;
    const long long row_vecs = sh.row_bytes / 16, seg2 = sh.seg2_rows * row_vecs;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f, sum = 0;
    const int reps = 12;
    for (int i = 0; i < 3 + reps; ++i) {
        CK(cudaEventRecord(a));
        gather_kernel<LPQ, PARTS, SEGS><<<grid, 256>>>(idx, table, row_vecs, seg2, out, sh.nq);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (i >= 3) { best = std::min(best, ms); sum += ms; }
    }
    CK(cudaGetLastError());
    printf("{\"shape\": \"%s\", \"ms_mean\": %.4f, \"ms_best\": %.4f, \"algorithmic_GBps\": %.1f, \"blocks_per_sm\": %d}\n",
           sh.name, sum / reps, best, sh.algo_bytes / (sum / reps * 1e-3) / 1e9, per_sm);
    fflush(stdout);
    return sum / reps;
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    // algorithmic bytes as in SURVEY.md 8(d): queries + output + table once
    const Shape shapes[] = {
        {"c3: 4 MB table, 128 B segment (two 64 B rows), 64 B out, 2^24 random", 65536ull * 64, 64, 4, 2, 1, 0, 1ll << 24, 0, 1.145e9},
        {"c4: 134 MB table, two 64 B segments 64 KB apart, 32 B out, 2^24 random", 2048ull * 2048 * 32, 32, 2, 2, 2, 2048, 1ll << 24, 0, 0.805e9},
        {"c5a direct: 2.1 GB table, two 256 B segments, 128 B out, 2^25 random", 4096ull * 4096 * 128, 128, 8, 2, 2, 4096, 1ll << 25, 0, 6.71e9},
        {"c5a binned: same, indices sorted into 128 bands of 16 MB", 4096ull * 4096 * 128, 128, 8, 2, 2, 4096, 1ll << 25, 128, 6.71e9},
    };
    for (const Shape& sh : shapes) {
        const size_t rows = sh.table_bytes / sh.row_bytes;
        // rows a query may start at: keep every segment inside the table
        const size_t usable = rows - (size_t)sh.seg2_rows * (sh.segs - 1) - (size_t)sh.parts;
        std::vector<uint32_t> h((size_t)sh.nq);
        uint64_t s = 0x9E3779B97F4A7C15ull;
        for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s % usable); }
        if (sh.sorted_bands) {
            const size_t band = (rows + sh.sorted_bands - 1) / sh.sorted_bands;
            std::stable_sort(h.begin(), h.end(), [&](uint32_t x, uint32_t y) { return x / band < y / band; });
        }
        uint32_t* idx; int4 *table, *out;
        CK(cudaMalloc(&idx, h.size() * 4));
        CK(cudaMalloc(&table, sh.table_bytes + 4096));
        CK(cudaMalloc(&out, (size_t)sh.nq * sh.lpq * 16));
        CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(table, 1, sh.table_bytes + 4096));
        if (sh.lpq == 4) run<4, 2, 1>(sh, idx, table, out, sms);
        else if (sh.lpq == 2) run<2, 2, 2>(sh, idx, table, out, sms);
        else run<8, 2, 2>(sh, idx, table, out, sms);
        CK(cudaFree(idx)); CK(cudaFree(table)); CK(cudaFree(out));
    }
    // C4 with other table layouts (algorithmic bytes as for C4: the layout is the library's business)
    run_layout<false>("c4 PAIR layout: 268 MB table, one 128 B segment at 64 B alignment, 32 B out, 2^24 random", 2048ull * 2048 * 64, 2048ull * 2048, 1ll << 24, sms);
    run_layout<true>("c4 QUAD layout: 537 MB table, one aligned 128 B segment, 32 B out, 2^24 random", 2048ull * 2048 * 128, 2048ull * 2048, 1ll << 24, sms);
    return 0;
}
