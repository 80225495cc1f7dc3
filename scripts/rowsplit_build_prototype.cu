// rowsplit_build_prototype.cu -- PROTOTYPE, NOT PART OF THE LIBRARY, NOT YET RUN ON A GPU.
//
// Written at the end of round 1, after the round's GPU minutes were spent: it compiles here (nvcc, sm_100a) but no
// number in this repository comes from it.  Purpose: before the row-split spline build of DESIGN.md section 7 item 3
// is integrated into csrc/ndi_spline.cu, measure with the simplest possible kernels what splitting the rows buys:
// `levels` steps of parallel cyclic reduction turn the reference's one tridiagonal system into 2^levels interleaved
// systems (rows j, j+S, j+2S, ...) whose Thomas chains are 2^levels times shorter and 2^levels times as many.
//
// The arithmetic follows oracle/ndi_oracle.cpp rowsplit_thomas operation by operation (no FMA, IEEE division), so
// the device result is compared BIT FOR BIT with a host restatement in this file, and its distance from the
// reference's sequential order (levels = 0 runs exactly that order through the same kernels) is reported against
// north_star's bars (1e-12 relative for f64, 1e-5 for f32).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -ffp-contract=off \
//        -o scripts/rowsplit_build_prototype.bin scripts/rowsplit_build_prototype.cu
//   scripts/rowsplit_build_prototype.bin          (one JSON line per shape and level count)
//
// Only the solve for k is prototyped (Natural boundary): forming the right-hand sides and a / b from k is the
// same elementwise work as in the shipped spline_rhs_kernel / spline_ab_kernel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <class T> struct Ar;
template <> struct Ar<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Ar<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// ---- one reduction level on the matrix (depends on x only: n elements) -------------------------------------------------
template <class T>
__global__ void pcr_matrix_level(const T* __restrict__ low, const T* __restrict__ mid, const T* __restrict__ up, T* __restrict__ nlow,
                                 T* __restrict__ nmid, T* __restrict__ nup, T* __restrict__ alpha, T* __restrict__ gamma, int n, int s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool hm = i - s >= 0, hp = i + s <= n - 1;
    const T al = hm ? -Ar<T>::div(low[i], mid[i - s]) : (T)0;
    const T ga = hp ? -Ar<T>::div(up[i], mid[i + s]) : (T)0;
    nlow[i] = hm ? Ar<T>::mul(al, low[i - s]) : (T)0;
    nup[i] = hp ? Ar<T>::mul(ga, up[i + s]) : (T)0;
    T m = mid[i];
    if (hm) m = Ar<T>::add(m, Ar<T>::mul(al, up[i - s]));
    if (hp) m = Ar<T>::add(m, Ar<T>::mul(ga, low[i + s]));
    nmid[i] = m;
    alpha[i] = al; gamma[i] = ga;
}

// ---- the same level on the right-hand sides: one elementwise pass over (n, w) ---------------------------------------
template <class T>
__global__ void pcr_rhs_level(const T* __restrict__ r, T* __restrict__ nr, const T* __restrict__ alpha, const T* __restrict__ gamma,
                              int n, long long w, int s) {
    const long long total = (long long)n * w, step = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        const int i = (int)(e / w);
        T v = r[e];
        if (i - s >= 0) v = Ar<T>::add(v, Ar<T>::mul(alpha[i], r[e - (long long)s * w]));
        if (i + s <= n - 1) v = Ar<T>::add(v, Ar<T>::mul(gamma[i], r[e + (long long)s * w]));
        nr[e] = v;
    }
}

// ---- forward factorisation of the S interleaved systems: thread j owns rows j, j+S, ... (thomas(), :690-702) ----------
template <class T>
__global__ void factor_systems(const T* __restrict__ low, T* __restrict__ mid, const T* __restrict__ up, T* __restrict__ ww, int n, int S) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S || j >= n) return;
    T prev_mid = mid[j];
    for (int i = j + S; i < n; i += S) {
        const T w_ = Ar<T>::div(low[i], prev_mid);
        ww[i] = w_;
        prev_mid = Ar<T>::sub(mid[i], Ar<T>::mul(w_, up[i - S]));
        mid[i] = prev_mid;
    }
}

// ---- both sweeps, in place on the right-hand sides: thread (j, c) owns rows j, j+S, ... of column c --------------------
// Consecutive threads take consecutive columns, so a warp's loads and stores of one row are contiguous.
template <class T>
__global__ void sweep_systems(T* __restrict__ r, const T* __restrict__ ww, const T* __restrict__ mid, const T* __restrict__ up, int n,
                              long long w, int S) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long c = tid % w;
    const long long j = tid / w;
    if (j >= S || j >= n) return;
    T prev = r[j * w + c];
    int last = (int)j;
    for (int i = (int)j + S; i < n; i += S) {                                  // :690-702
        prev = Ar<T>::sub(r[(long long)i * w + c], Ar<T>::mul(ww[i], prev));
        r[(long long)i * w + c] = prev;
        last = i;
    }
    T k = Ar<T>::div(prev, mid[last]);                                          // :704-708
    r[(long long)last * w + c] = k;
    for (int i = last - S; i >= (int)j; i -= S) {                               // :711-720
        k = Ar<T>::div(Ar<T>::sub(r[(long long)i * w + c], Ar<T>::mul(up[i], k)), mid[i]);
        r[(long long)i * w + c] = k;
    }
}

// ---- host restatement (oracle/ndi_oracle.cpp: thomas, rowsplit_thomas) ---------------------------------------------------
template <class T>
static void host_thomas(T* k, const T* up, T* mid, const T* low, T* rhs, long long len, long long w) {
    for (long long i = 1; i < len; ++i) {
        const T ww = low[i] / mid[i - 1];
        mid[i] = mid[i] - ww * up[i - 1];
        for (long long c = 0; c < w; ++c) rhs[i * w + c] = rhs[i * w + c] - ww * rhs[(i - 1) * w + c];
    }
    for (long long c = 0; c < w; ++c) k[(len - 1) * w + c] = rhs[(len - 1) * w + c] / mid[len - 1];
    for (long long i = len - 2; i >= 0; --i)
        for (long long c = 0; c < w; ++c) k[i * w + c] = (rhs[i * w + c] - up[i] * k[(i + 1) * w + c]) / mid[i];
}

template <class T>
static std::vector<T> host_rowsplit(std::vector<T> low, std::vector<T> mid, std::vector<T> up, std::vector<T> r, long long n, long long w,
                                    int levels) {
    std::vector<T> nlow(n), nmid(n), nup(n), nr(r.size()), k(r.size());
    long long s = 1;
    for (int lv = 0; lv < levels; ++lv, s *= 2) {
        for (long long i = 0; i < n; ++i) {
            const bool hm = i - s >= 0, hp = i + s <= n - 1;
            const T al = hm ? -(low[i] / mid[i - s]) : (T)0, ga = hp ? -(up[i] / mid[i + s]) : (T)0;
            nlow[i] = hm ? al * low[i - s] : (T)0;
            nup[i] = hp ? ga * up[i + s] : (T)0;
            T m = mid[i];
            if (hm) m = m + al * up[i - s];
            if (hp) m = m + ga * low[i + s];
            nmid[i] = m;
            for (long long c = 0; c < w; ++c) {
                T v = r[i * w + c];
                if (hm) v = v + al * r[(i - s) * w + c];
                if (hp) v = v + ga * r[(i + s) * w + c];
                nr[i * w + c] = v;
            }
        }
        low.swap(nlow); mid.swap(nmid); up.swap(nup); r.swap(nr);
    }
    for (long long j = 0; j < s && j < n; ++j) {
        const long long m = (n - j + s - 1) / s;
        std::vector<T> sl(m), sm(m), su(m), sr(m * w), sk(m * w);
        for (long long t = 0; t < m; ++t) {
            sl[t] = low[j + t * s]; sm[t] = mid[j + t * s]; su[t] = up[j + t * s];
            for (long long c = 0; c < w; ++c) sr[t * w + c] = r[(j + t * s) * w + c];
        }
        host_thomas(sk.data(), su.data(), sm.data(), sl.data(), sr.data(), m, w);
        for (long long t = 0; t < m; ++t)
            for (long long c = 0; c < w; ++c) k[(j + t * s) * w + c] = sk[t * w + c];
    }
    return k;
}

// ---- one shape ---------------------------------------------------------------------------------------------------------------
template <class T>
static void run_shape(const char* name, int n, long long w, std::initializer_list<int> level_list, long long host_cols) {
    // Natural system as the reference forms it (cubic_spline.rs:440-471, :619-631, :656-668)
    std::vector<T> x(n), low(n, 0), mid(n, 0), up(n, 0), y((size_t)n * w), rhs((size_t)n * w);
    uint64_t s = 0x243F6A8885A308D3ull;
    auto uni = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); };
    double acc = 0;
    for (int i = 0; i < n; ++i) { acc += 0.5 + uni(); x[i] = (T)acc; }
    for (auto& v : y) v = (T)(2.0 * uni() - 1.0);
    const T two = 2, three = 3;
    for (int i = 1; i + 1 < n; ++i) {
        const T dxm = x[i] - x[i - 1], dxp = x[i + 1] - x[i];
        up[i] = dxm; mid[i] = two * (dxp + dxm); low[i] = dxp;
        for (long long c = 0; c < w; ++c)
            rhs[i * w + c] = three * (dxp * (y[i * w + c] - y[(i - 1) * w + c]) / dxm + dxm * (y[(i + 1) * w + c] - y[i * w + c]) / dxp);
    }
    const T dx0 = x[1] - x[0], dx1 = x[n - 1] - x[n - 2];
    up[0] = dx0; mid[0] = two * dx0; low[n - 1] = dx1; mid[n - 1] = two * dx1;
    for (long long c = 0; c < w; ++c) {
        rhs[c] = three * (y[w + c] - y[c]);
        rhs[(long long)(n - 1) * w + c] = three * (y[(long long)(n - 1) * w + c] - y[(long long)(n - 2) * w + c]);
    }
    // host reference on the first host_cols columns only (the host solve is slow; columns are independent)
    const long long hc = std::min(w, host_cols);
    std::vector<T> rhs_h((size_t)n * hc), ymax(hc, 0);
    for (int i = 0; i < n; ++i)
        for (long long c = 0; c < hc; ++c) { rhs_h[i * hc + c] = rhs[i * w + c]; ymax[c] = std::max(ymax[c], (T)std::fabs(y[i * w + c])); }
    const std::vector<T> k_seq = host_rowsplit<T>(low, mid, up, rhs_h, n, hc, 0);

    T *d_low[2], *d_mid[2], *d_up[2], *d_r[2], *d_alpha, *d_gamma, *d_ww;
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&d_low[b], n * sizeof(T))); CK(cudaMalloc(&d_mid[b], n * sizeof(T))); CK(cudaMalloc(&d_up[b], n * sizeof(T)));
        CK(cudaMalloc(&d_r[b], (size_t)n * w * sizeof(T)));
    }
    CK(cudaMalloc(&d_alpha, n * sizeof(T))); CK(cudaMalloc(&d_gamma, n * sizeof(T))); CK(cudaMalloc(&d_ww, n * sizeof(T)));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t ev[4]; for (auto& e : ev) CK(cudaEventCreate(&e));

    for (int levels : level_list) {
        const int S = 1 << levels;
        float t_reduce = 0, t_factor = 0, t_sweep = 0;
        const int reps = 5;
        int cur = 0;
        for (int rep = 0; rep < reps + 1; ++rep) {                              // rep 0 is the warm-up
            CK(cudaMemcpy(d_low[0], low.data(), n * sizeof(T), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_mid[0], mid.data(), n * sizeof(T), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_up[0], up.data(), n * sizeof(T), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_r[0], rhs.data(), (size_t)n * w * sizeof(T), cudaMemcpyHostToDevice));
            cur = 0;
            CK(cudaEventRecord(ev[0]));
            for (int lv = 0, st = 1; lv < levels; ++lv, st *= 2) {
                pcr_matrix_level<T><<<(n + 255) / 256, 256>>>(d_low[cur], d_mid[cur], d_up[cur], d_low[cur ^ 1], d_mid[cur ^ 1], d_up[cur ^ 1],
                                                            d_alpha, d_gamma, n, st);
                pcr_rhs_level<T><<<sms * 8, 256>>>(d_r[cur], d_r[cur ^ 1], d_alpha, d_gamma, n, w, st);
                cur ^= 1;
            }
            CK(cudaEventRecord(ev[1]));
            factor_systems<T><<<(S + 127) / 128, 128>>>(d_low[cur], d_mid[cur], d_up[cur], d_ww, n, S);
            CK(cudaEventRecord(ev[2]));
            const long long threads = (long long)S * w;
            sweep_systems<T><<<(unsigned)((threads + 127) / 128), 128>>>(d_r[cur], d_ww, d_mid[cur], d_up[cur], n, w, S);
            CK(cudaEventRecord(ev[3]));
            CK(cudaEventSynchronize(ev[3]));
            CK(cudaGetLastError());
            if (rep > 0) {
                float a, b, c;
                CK(cudaEventElapsedTime(&a, ev[0], ev[1])); CK(cudaEventElapsedTime(&b, ev[1], ev[2])); CK(cudaEventElapsedTime(&c, ev[2], ev[3]));
                t_reduce += a; t_factor += b; t_sweep += c;
            }
        }
        std::vector<T> k_dev((size_t)n * w);
        CK(cudaMemcpy(k_dev.data(), d_r[cur], (size_t)n * w * sizeof(T), cudaMemcpyDeviceToHost));
        const std::vector<T> k_ref = host_rowsplit<T>(low, mid, up, rhs_h, n, hc, levels);
        long long mismatches = 0;
        double dev_seq = 0;
        for (int i = 0; i < n; ++i)
            for (long long c = 0; c < hc; ++c) {
                const T g = k_dev[i * w + c], r = k_ref[i * hc + c], q = k_seq[i * hc + c];
                if (std::memcmp(&g, &r, sizeof(T)) != 0) ++mismatches;
                // k is a slope: scale by max|y| / typical spacing (spacing is O(1) here)
                dev_seq = std::max(dev_seq, (double)std::fabs(g - q) / std::max((double)std::fabs(q), (double)ymax[c]));
            }
        printf("{\"shape\": \"%s\", \"rows\": %d, \"columns\": %lld, \"dtype\": \"%s\", \"levels\": %d, \"chains\": %lld, \"chain_length\": %d, "
               "\"ms_reduce\": %.4f, \"ms_factor\": %.4f, \"ms_sweep\": %.4f, \"ms_total\": %.4f, "
               "\"bit_mismatches_vs_host_restatement\": %lld, \"checked_columns\": %lld, \"max_dev_vs_sequential\": %.3e}\n",
               name, n, w, sizeof(T) == 8 ? "f64" : "f32", levels, (long long)S * w, (n + S - 1) / S, t_reduce / reps, t_factor / reps,
               t_sweep / reps, (t_reduce + t_factor + t_sweep) / reps, mismatches, hc, dev_seq);
        fflush(stdout);
    }
    for (int b = 0; b < 2; ++b) { cudaFree(d_low[b]); cudaFree(d_mid[b]); cudaFree(d_up[b]); cudaFree(d_r[b]); }
    cudaFree(d_alpha); cudaFree(d_gamma); cudaFree(d_ww);
}

int main() {
    run_shape<double>("c2", 4096, 1024, {0, 1, 2, 3, 4, 5}, 16);
    run_shape<float>("c5b-shard", 4096, 16384, {0, 2, 4}, 16);
    run_shape<double>("long", 65536, 64, {0, 3, 6, 8}, 8);
    return 0;
}
