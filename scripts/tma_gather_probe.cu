// tma_gather_probe.cu -- can the bulk asynchronous copy engine (cp.async.bulk, SASS UBLKCP) do the table gathers of the
// binned bilinear kernel?  C5a gathers, per query, two contiguous 256-byte segments (z11 | z12 and z21 | z22) out of an
// L2-resident band; the shipped kernel does that with four 16-byte loads per lane and is bound by the bytes it can
// keep in flight in registers (32 warps x 2 KB per SM).  Here every lane issues its query's two 256-byte bulk copies
// into a per-warp slot of shared memory (no registers held), the warp waits on the slot's mbarrier, optionally reads
// the slot back (LDS.128, as the evaluation would) and goes on; two slots per warp, one tile ahead.
// Reported: rows (queries) per second and the L2 -> SM gigabytes per second, against the 1.23 ms / 2^25 queries of
// the register-gather pattern (profiles/r01/gather_ceiling.md: 27 G queries/s).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/tma_gather_probe.bin scripts/tma_gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr size_t kBand = 16u << 20;          // one band of the table: L2-resident
constexpr int kSeg = 256;                    // bytes per bulk copy

__device__ __forceinline__ uint32_t mix(uint32_t a) {
    a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16;
    return a;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// QPT: queries per slot (32: every lane issues its own two copies; 16: half tiles, lanes 0..15 issue)
template <int QPT, bool READBACK>
__global__ void __launch_bounds__(1024) gather_probe(const unsigned char* table, int tiles_per_warp, unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int kSlot = QPT * 2 * kSeg;
    unsigned char* my = smem + (size_t)wid * 2 * kSlot;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * 2 * kSlot) + wid * 2;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars), slot0 = (uint32_t)__cvta_generic_to_shared(my);
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const uint32_t warp = blockIdx.x * nw + wid;
    auto issue = [&](int t) {
        const uint32_t s = t & 1, bar = bar0 + 8 * s, dst = slot0 + s * kSlot;
        if (lane == 0) mbar_expect(bar, kSlot);
        __syncwarp();
        if (lane < QPT) {
            const uint32_t h = mix(warp * 0x9e3779b9u + t * 32 + lane);
            const size_t cell = (size_t)(h % (kBand / 2 / kSeg - 1)) * kSeg;        // z11 | z12 ...
            bulk_g2s(dst + lane * 2 * kSeg, table + cell, kSeg, bar);
            bulk_g2s(dst + lane * 2 * kSeg + kSeg, table + kBand / 2 + cell, kSeg, bar);   // ... and z21 | z22, one x-row further
        }
    };
    unsigned long long acc = 0;
    issue(0);
    for (int t = 0; t < tiles_per_warp; ++t) {
        if (t + 1 < tiles_per_warp) issue(t + 1);
        mbar_wait(bar0 + 8 * (t & 1), (t >> 1) & 1);
        if (READBACK) {
            // as the evaluation would: 8 lanes x 16 bytes per 128-byte row, 4 rows per query, 4 queries per round
            const unsigned char* slot = my + (t & 1) * kSlot;
#pragma unroll 4
            for (int r = 0; r < QPT / 4; ++r) {
                const unsigned char* q = slot + (r * 4 + (lane >> 3)) * 2 * kSeg + (lane & 7) * 16;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 v = *reinterpret_cast<const uint4*>(q + k * 128);
                    acc += v.x ^ v.y ^ v.z ^ v.w;
                }
            }
        }
        __syncwarp();
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
}

template <int QPT, bool RB>
static void run(const unsigned char* table, unsigned long long* sink, int warps_per_block, int blocks_per_sm, int tiles) {
    const size_t smem = (size_t)warps_per_block * 2 * QPT * 2 * kSeg + warps_per_block * 16;
    cudaFuncSetAttribute(gather_probe<QPT, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * blocks_per_sm;
    gather_probe<QPT, RB><<<grid, warps_per_block * 32, smem>>>(table, tiles, sink);
    cudaEventRecord(a);
    gather_probe<QPT, RB><<<grid, warps_per_block * 32, smem>>>(table, tiles, sink);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double q = (double)grid * warps_per_block * tiles * QPT;
    printf("{\"qpt\": %d, \"readback\": %d, \"warps_per_sm\": %d, \"smem_per_sm_kb\": %.0f, \"ms\": %.4f, \"queries_per_s\": %.4g, \"bulk_ops_per_s_per_sm\": %.4g, \"l2_to_sm_GBps\": %.1f, \"status\": \"%s\"}\n",
           QPT, (int)RB, warps_per_block * blocks_per_sm, smem * blocks_per_sm / 1024.0, ms, q / (ms * 1e-3), 2 * q / (ms * 1e-3) / 148,
           q * 2 * kSeg / (ms * 1e-3) / 1e9, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
}

int main() {
    unsigned char* table; unsigned long long* sink;
    cudaMalloc(&table, kBand + 4096); cudaMalloc(&sink, 8);
    cudaMemset(table, 1, kBand + 4096);
    const int tiles = 512;
    for (int w : {2, 4, 6}) { run<32, false>(table, sink, w, 1, tiles); run<32, true>(table, sink, w, 1, tiles); }
    for (int w : {4, 8, 12, 13}) { run<16, false>(table, sink, w, 1, tiles); run<16, true>(table, sink, w, 1, tiles); }
    for (int w : {4, 6}) { run<16, true>(table, sink, w, 2, tiles); }
    return 0;
}
