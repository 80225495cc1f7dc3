// wavefront_probe.cu -- what ONE warp-wide memory instruction costs on the L1TEX data pipe (B200), per access pattern.
//
// The thin-row kernels are bound by l1tex__data_pipe_lsu_wavefronts (profiles/r01/l1_wavefronts.md); their budgets were
// so far estimated as "one wavefront per distinct 128-byte line".  The counters of the product kernels disagree for
// stores (C5a: 2.5 wavefronts per scattered 128-byte row).  This program isolates the patterns the kernels use: every
// kernel below issues exactly ONE kind of instruction per loop iteration; run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts*.sum,smsp__inst_executed_op_*.sum
// and divide (scripts/gpu_r2_call9.sh prints wavefronts per instruction).  Tables are 32 MB (L2-resident).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/wavefront_probe.bin scripts/wavefront_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kIters = 256;
constexpr size_t kBytes = 32u << 20;

__device__ __forceinline__ uint32_t mix(uint32_t a) {
    a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16;
    return a;
}

enum Pattern {
    LDG128_ROWS4,       // 8 lanes x 16 B per aligned 128-byte row, 4 random rows              (C5a gather, one cell)
    LDG128_CONTIG,      // 32 lanes x 16 B = 512 contiguous aligned bytes
    LDG128_DISTINCT,    // every lane its own random 16-byte entry                           (bucket-table probe)
    LDG32_DISTINCT8K,   // every lane a random 4-byte word of an 8 KB array                    (guess verification, direct)
    LDG32_BAND,         // every lane a random word of 132 consecutive bytes                   (guess verification, binned)
    LDG128_PAIR32,      // 2 lanes x 16 B per random 32-byte cell, 16 cells                    (C4 gather as shipped)
    LDG256_PAIR64,      // 2 lanes x 32 B per random 64-byte segment at 32-byte alignment      (C4 gather, pair form)
    LDG256_ROWS8,       // 4 lanes x 32 B per aligned 128-byte row, 8 random rows
    STG128_ROWS4_CS, STG128_ROWS4, STG256_ROWS8_CS,                                         // scattered 128-byte output rows
    STG128_CONTIG_CS, STG128_CONTIG,                                                        // 512 contiguous bytes
    STG128_PAIR32_CS,   // 2 lanes x 16 B per scattered 32-byte row, 16 rows                   (C4 binned / swept output)
    STG128_ROWS8_64_CS, // 4 lanes x 16 B per scattered 64-byte row, 8 rows
    LDS128_BCAST4, LDS128_BCAST16, LDS128_DISTINCT, STS128_DISTINCT, LDS64_BCAST16, LDS32_RANDOM, SHFL
};

template <int P>
__global__ void __launch_bounds__(256) probe(unsigned char* buf, unsigned long long* sink) {
    __shared__ __align__(16) unsigned char sm[16384];
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    __syncthreads();
    unsigned long long acc = 0;
    const uint32_t sm_a = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        const uint32_t h = mix(warp * 0x9e3779b9u + it);
        unsigned long long a = 0, b = 0, c = 0, d = 0;
        uint32_t x = 0, y = 0, z = 0, w = 0;
        if constexpr (P == LDG128_ROWS4) {
            const unsigned char* p = buf + (size_t)(mix(h + (lane >> 3)) % (kBytes / 128)) * 128 + (lane & 7) * 16;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(p));
        } else if constexpr (P == LDG128_CONTIG) {
            const unsigned char* p = buf + (size_t)(h % (kBytes / 512)) * 512 + lane * 16;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(p));
        } else if constexpr (P == LDG128_DISTINCT) {
            const unsigned char* p = buf + (size_t)(mix(h + lane) % (kBytes / 16)) * 16;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(p));
        } else if constexpr (P == LDG32_DISTINCT8K) {
            const unsigned char* p = buf + (size_t)(mix(h + lane) % 2048) * 4;
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(p));
        } else if constexpr (P == LDG32_BAND) {
            const unsigned char* p = buf + (size_t)(h % 1024) * 128 + (mix(h + lane) % 33) * 4;
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(x) : "l"(p));
        } else if constexpr (P == LDG128_PAIR32) {
            const unsigned char* p = buf + (size_t)(mix(h + (lane >> 1)) % (kBytes / 32)) * 32 + (lane & 1) * 16;
            asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "l"(p));
        } else if constexpr (P == LDG256_PAIR64) {
            const unsigned char* p = buf + (size_t)(mix(h + (lane >> 1)) % (kBytes / 32 - 1)) * 32 + (lane & 1) * 32;
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        } else if constexpr (P == LDG256_ROWS8) {
            const unsigned char* p = buf + (size_t)(mix(h + (lane >> 2)) % (kBytes / 128)) * 128 + (lane & 3) * 32;
            asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
        } else if constexpr (P == STG128_ROWS4_CS) {
            unsigned char* p = buf + (size_t)(mix(h + (lane >> 3)) % (kBytes / 128)) * 128 + (lane & 7) * 16;
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(h) : "memory");
        } else if constexpr (P == STG128_ROWS4) {
            unsigned char* p = buf + (size_t)(mix(h + (lane >> 3)) % (kBytes / 128)) * 128 + (lane & 7) * 16;
            asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(h) : "memory");
        } else if constexpr (P == STG256_ROWS8_CS) {
            unsigned char* p = buf + (size_t)(mix(h + (lane >> 2)) % (kBytes / 128)) * 128 + (lane & 3) * 32;
            asm volatile("st.global.cs.v4.u64 [%0], {%1,%1,%1,%1};" :: "l"(p), "l"((unsigned long long)h) : "memory");
        } else if constexpr (P == STG128_CONTIG_CS) {
            unsigned char* p = buf + (size_t)(h % (kBytes / 512)) * 512 + lane * 16;
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(h) : "memory");
        } else if constexpr (P == STG128_CONTIG) {
            unsigned char* p = buf + (size_t)(h % (kBytes / 512)) * 512 + lane * 16;
            asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(h) : "memory");
        } else if constexpr (P == STG128_PAIR32_CS) {
            unsigned char* p = buf + (size_t)(mix(h + (lane >> 1)) % (kBytes / 32)) * 32 + (lane & 1) * 16;
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(h) : "memory");
        } else if constexpr (P == STG128_ROWS8_64_CS) {
            unsigned char* p = buf + (size_t)(mix(h + (lane >> 2)) % (kBytes / 64)) * 64 + (lane & 3) * 16;
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(h) : "memory");
        } else if constexpr (P == LDS128_BCAST4) {
            const uint32_t p = sm_a + ((h & 63) * 64 + (lane >> 3) * 16);
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(p));
        } else if constexpr (P == LDS128_BCAST16) {
            const uint32_t p = sm_a + ((h & 31) * 256 + (lane >> 1) * 16);
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(p));
        } else if constexpr (P == LDS128_DISTINCT) {
            const uint32_t p = sm_a + ((h & 15) * 512 + lane * 16);
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(p));
        } else if constexpr (P == STS128_DISTINCT) {
            const uint32_t p = sm_a + ((h & 15) * 512 + lane * 16);
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" :: "r"(p), "r"(h) : "memory");
        } else if constexpr (P == LDS64_BCAST16) {
            const uint32_t p = sm_a + ((h & 63) * 128 + (lane >> 1) * 8);
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(p));
        } else if constexpr (P == LDS32_RANDOM) {
            const uint32_t p = sm_a + (mix(h + lane) % 2048) * 4;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(p));
        } else {
            x = __shfl_sync(0xffffffffu, h + lane, (h >> 3) & 31);
        }
        acc += a ^ b ^ c ^ d ^ x ^ y ^ z ^ w;
    }
    if (acc == 0x1234567887654321ull) *sink = acc;
}

template <int P>
static void run(const char* name, unsigned char* buf, unsigned long long* sink) {
    probe<P><<<148 * 4, 256>>>(buf, sink);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-22s %s\n", name, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
}

int main() {
    unsigned char* buf; unsigned long long* sink;
    cudaMalloc(&buf, kBytes + 4096); cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, kBytes + 4096);
#define RUN(P) run<P>(#P, buf, sink)
    RUN(LDG128_ROWS4); RUN(LDG128_CONTIG); RUN(LDG128_DISTINCT); RUN(LDG32_DISTINCT8K); RUN(LDG32_BAND);
    RUN(LDG128_PAIR32); RUN(LDG256_PAIR64); RUN(LDG256_ROWS8);
    RUN(STG128_ROWS4_CS); RUN(STG128_ROWS4); RUN(STG256_ROWS8_CS); RUN(STG128_CONTIG_CS); RUN(STG128_CONTIG);
    RUN(STG128_PAIR32_CS); RUN(STG128_ROWS8_64_CS);
    RUN(LDS128_BCAST4); RUN(LDS128_BCAST16); RUN(LDS128_DISTINCT); RUN(STS128_DISTINCT); RUN(LDS64_BCAST16);
    RUN(LDS32_RANDOM); RUN(SHFL);
    return 0;
}
