// ndi_spline.cuh -- device-side pieces shared by the two spline builds: the reference-order build
// (ndi_spline.cu) and the row-split PCR + Thomas build (ndi_rowsplit.cu).  Matrix rows, right-hand
// sides and boundary rows follow CubicSpline::solve_for_k (cubic_spline.rs:409-674) operation by
// operation, one rounding each.
#pragma once

#include "ndi_device.cuh"
#include "ndi_internal.h"

namespace ndi {

enum { SB_NAK = 0, SB_NATURAL = 1, SB_CLAMPED = 2, SB_FIRST = 3, SB_SECOND = 4 };
enum { BC_NAK = 0, BC_NATURAL = 1, BC_CLAMPED = 2, BC_PERIODIC = 3, BC_INDIVIDUAL = 4 };

template <class T>
struct Side { int kind; T val; };

// SingleBoundary::specialize (:287-296)
template <class T>
__host__ __device__ inline Side<T> specialize(Side<T> s) {
    if (s.kind == SB_NATURAL) return {SB_SECOND, (T)0};
    if (s.kind == SB_CLAMPED) return {SB_FIRST, (T)0};
    return s;
}

template <class T> struct A : Ar<T> {};
#define ADD A<T>::add
#define SUB A<T>::sub
#define MUL A<T>::mul
#define DIV A<T>::div

// matrix row i of the full (non-periodic) system: (:440-451) interior, (:584-590) the 3-point
// NotAKnot parabola system, (:599-669) boundary rows.
template <class T>
__device__ __forceinline__ void matrix_row(const T* __restrict__ x, int n, int i, int lk, int rk, bool nak3, T& up,
                                           T& mid, T& low) {
    const T two = (T)2, one = (T)1, zero = (T)0;
    if (i > 0 && i < n - 1) {
        const T dxn = SUB(x[i + 1], x[i]), dxn_1 = SUB(x[i], x[i - 1]);
        up = dxn_1; mid = MUL(two, ADD(dxn, dxn_1)); low = dxn;
    } else if (i == 0) {
        low = zero;
        const T dx0 = SUB(x[1], x[0]);
        if (nak3) { mid = one; up = one; }
        else if (lk == SB_NAK) { mid = SUB(x[2], x[1]); up = SUB(x[2], x[0]); }
        else if (lk == SB_FIRST) { mid = one; up = zero; }
        else { up = dx0; mid = MUL(two, dx0); }
    } else {
        up = zero;
        const T dx_1 = SUB(x[n - 1], x[n - 2]);
        if (nak3) { low = one; mid = one; }
        else if (rk == SB_NAK) { mid = dx_1; low = SUB(x[n - 1], x[n - 3]); }
        else if (rk == SB_FIRST) { mid = one; low = zero; }
        else { mid = MUL(two, dx_1); low = dx_1; }
    }
}
// row i of the condensed periodic system (:512-518), i in [0, n-3]
template <class T>
__device__ __forceinline__ void matrix_row_periodic(const T* __restrict__ x, int n, int i, T& up, T& mid, T& low) {
    const T two = (T)2;
    if (i == 0) {
        const T dx0 = SUB(x[1], x[0]), dx_1 = SUB(x[n - 1], x[n - 2]);
        mid = MUL(two, ADD(dx_1, dx0)); up = dx_1; low = (T)0;
    } else {
        const T dxn = SUB(x[i + 1], x[i]), dxn_1 = SUB(x[i], x[i - 1]);
        up = dxn_1; mid = MUL(two, ADD(dxn, dxn_1)); low = dxn;
    }
}

// right-hand sides ---------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T rhs_interior(T yl, T ym, T yr, T dxn, T dxn_1) {       // :468
    const T three = (T)3;
    return MUL(three, ADD(DIV(MUL(dxn, SUB(ym, yl)), dxn_1), DIV(MUL(dxn_1, SUB(yr, ym)), dxn)));
}
template <class T>
__device__ __forceinline__ T rhs_left(const T* __restrict__ x, Side<T> l, T y0, T y1, T y2) {
    const T two = (T)2, three = (T)3;
    const T dx0 = SUB(x[1], x[0]), dx1 = SUB(x[2], x[1]);
    if (l.kind == SB_NAK) {                                                           // :600-610
        const T d = SUB(x[2], x[0]);
        const T tmp1 = MUL(ADD(dx0, MUL(two, d)), dx1);
        return DIV(ADD(DIV(MUL(tmp1, SUB(y1, y0)), dx0), DIV(MUL(MUL(dx0, dx0), SUB(y2, y1)), dx1)), d);
    }
    if (l.kind == SB_FIRST) return l.val;                                             // :614-618
    return SUB(MUL(three, SUB(y1, y0)), DIV(MUL(l.val, MUL(dx0, dx0)), two));         // :629
}
template <class T>
__device__ __forceinline__ T rhs_right(const T* __restrict__ x, int n, Side<T> r, T y_1, T y_2, T y_3) {
    const T two = (T)2, three = (T)3;
    const T dx_1 = SUB(x[n - 1], x[n - 2]), dx_2 = SUB(x[n - 2], x[n - 3]);
    if (r.kind == SB_NAK) {                                                           // :635-647
        const T d = SUB(x[n - 1], x[n - 3]);
        const T tmp1 = MUL(ADD(MUL(two, d), dx_1), dx_2);
        return DIV(ADD(DIV(MUL(MUL(dx_1, dx_1), SUB(y_2, y_3)), dx_2), DIV(MUL(tmp1, SUB(y_1, y_2)), dx_1)), d);
    }
    if (r.kind == SB_FIRST) return r.val;                                             // :651-655
    return ADD(MUL(three, SUB(y_1, y_2)), DIV(MUL(r.val, MUL(dx_1, dx_1)), two));     // :666
}

// ---- shared-matrix factorisation -------------------------------------------------------------------
// The elimination w[i] = low[i] / mid'[i-1]; mid'[i] = mid[i] - w[i] * up[i-1] (thomas :690-692) is a
// serial chain by definition (division -> multiply -> subtract, in the reference's order).  One
// block: all threads form the matrix rows of a tile in shared memory, thread 0 runs the chain over
// the tile out of shared memory (so its only latency is the arithmetic), all threads write the
// tile back.  fac layout: up[n] | mid[n] (eliminated) | wl[n] | k2[n]
constexpr int kFacTile = 1024, kFacBlock = 256;
template <class T> struct alignas(4 * sizeof(T)) FacRow { T up, mid, wl, rmid; };
template <class T>
__device__ __forceinline__ FacRow<T> ld_fac(const FacRow<T>* p) {
    FacRow<T> f;
    if constexpr (sizeof(T) == 4) { const int4 v = __ldg(reinterpret_cast<const int4*>(p)); f = *reinterpret_cast<const FacRow<T>*>(&v); }
    else {
        const int4 v0 = __ldg(reinterpret_cast<const int4*>(p)), v1 = __ldg(reinterpret_cast<const int4*>(p) + 1);
        int4 t[2] = {v0, v1};
        f = *reinterpret_cast<const FacRow<T>*>(t);
    }
    return f;
}
// Individual boundaries: the matrix differs between columns only through the KIND of the two boundary
// rows (three kinds each after specialize()), so there are at most nine matrices.  Block g = 3*l + r
// factorises the one with left kind ind_kind(l) and right kind ind_kind(r) into fac + g * fac_stride.
__host__ __device__ inline int ind_kind(int v) { return v == 0 ? SB_NAK : (v == 1 ? SB_FIRST : SB_SECOND); }
__host__ __device__ inline int ind_variant(int specialized_kind) { return specialized_kind == SB_NAK ? 0 : (specialized_kind == SB_FIRST ? 1 : 2); }

// ---- the two serial recurrences of the solve (ndi_spline.cu), over nsys interleaved systems ----------
// Individual boundaries sweep all (up to nine) groups of columns in ONE launch: the groups are
// independent and each is bound by its chain latency, not by throughput.
struct SweepGroups {
    int ngroups;                 // 0: one group = all columns, factorisation at fac
    int first_block[10];         // blocks [first_block[g], first_block[g+1]) sweep group g
    long long col_off[9], count[9];
    unsigned long long fac_stride;
};
// Rows i = j, j + nsys, j + 2 nsys, ... of R form system j (nsys == 1: the whole matrix is one system,
// the reference's order); every (system, column) pair is one chain.  fac: FacRow per row i (per group).
// counts: nine group sizes (Individual) or nullptr.
template <class T>
cudaError_t launch_spline_sweep(int len, int nsys, long long w, const T* fac, size_t fac_stride, T* R,
                                const int64_t* counts, cudaStream_t st);

#ifndef NDI_ROW_GROUP
#define NDI_ROW_GROUP 4
#endif
constexpr int kRowGroup = NDI_ROW_GROUP;  // rhs / ab kernels: rows per thread, sharing their loads
int row_group_grid(long long wv, long long nrows);      // grid for items of kRowGroup rows x one vector of columns
// V consecutive columns as one 16-byte access (V == 1: a plain element)
template <class T, int V>
__device__ __forceinline__ void ld_vec(const T* __restrict__ p, T (&out)[V]) {
    if constexpr (V == 1) out[0] = __ldg(p);
    else {
        static_assert(sizeof(T) * V == 16, "one 16-byte vector");
        const int4 q = __ldg(reinterpret_cast<const int4*>(p));
        memcpy(out, &q, 16);
    }
}
template <class T, int V>
__device__ __forceinline__ void st_vec(T* __restrict__ p, const T (&in)[V]) {
    if constexpr (V == 1) *p = in[0];
    else { int4 q; memcpy(&q, in, 16); *reinterpret_cast<int4*>(p) = q; }
}
template <class T>
inline bool vec_ok(long long w, const void* p0, const void* p1, const void* p2, const void* p3) {
    return w % (long long)(16 / sizeof(T)) == 0 &&
           (((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)p3) & 15) == 0;
}
// The elementwise kernels (right-hand sides, a / b) work on items of kRowGroup rows x V columns, V columns being one
// 16-byte vector (V = 1 when the row length or a pointer does not allow it): a thread that requests 4-byte words
// keeps too few bytes in flight to fill the memory system (4096 x 16384 f32: 363 us for 537 MB, profiles/r02).  A
// task is blockDim.x items: a chunk of columns of one row group, or -- tables narrower than a block -- consecutive
// items wrapping into the next row groups, so that every thread stays busy; grid-stride over the tasks.
struct RowTask { int row; long long col; bool live; };      // col: in vectors
__host__ __device__ inline long long row_task_count(long long wv, long long nrows, int block) {
    const long long groups = (nrows + kRowGroup - 1) / kRowGroup;
    return wv >= block ? groups * ((wv + block - 1) / block) : (groups * wv + block - 1) / block;
}
__device__ __forceinline__ RowTask row_task(long long task, long long wv, long long nrows) {
    if (wv >= blockDim.x) {
        const long long chunks = (wv + blockDim.x - 1) / blockDim.x, g = task / chunks;
        const long long col = (task - g * chunks) * blockDim.x + threadIdx.x;
        return RowTask{(int)g * kRowGroup, col, col < wv};
    }
    const long long item = task * blockDim.x + threadIdx.x, g = item / wv;
    return RowTask{(int)g * kRowGroup, item - g * wv, g * kRowGroup < nrows};
}

template <class T>
cudaError_t launch_spline_ab(const T* x, int n, const T* y, long long w, int periodic, const T* fac, const T* R, T* a, T* b,
                             const int32_t* pos, cudaStream_t st);
template <class T>
cudaError_t launch_spline_periodic_close(const T* x, int n, long long w, const T* fac, T* R, cudaStream_t st);

template <class T>
cudaError_t launch_spline_rhs(const T* x, int n, const T* y, long long w, int periodic, Side<T> left, Side<T> right, T* R,
                              unsigned long long* err, const int32_t* lks, const T* lvs, const int32_t* rks, const T* rvs,
                              cudaStream_t st);

// partition build (ndi_partition.cu): blocks of `block` rows (separator included, 3 .. kPartBlockMax) solved in registers,
// the separators' system recursively; the periodic close (and, for the short systems solved directly, right-hand sides and
// a / b) are the kernels above
constexpr int kPartBlockMax = 64, kPartBlockDefault = 32;
int partition_block_for(int requested);
size_t partition_fac_elems(int64_t n, int block);
template <class T>
cudaError_t launch_partition_build(const T* x, int64_t n, const T* data, int64_t w, int bc_kind, int block, const int32_t* lk,
                                   const T* lv, const int32_t* rk, const T* rv, T* a, T* b, T* scratch, unsigned long long* err,
                                   cudaStream_t st);

// row-split build (ndi_rowsplit.cu): factorisation with `levels` steps of cyclic reduction, right-hand sides
// reduced in shared memory; the sweeps, the periodic close and a / b are the kernels above
size_t rowsplit_fac_elems(int64_t n, int levels);
template <class T>
cudaError_t launch_rowsplit_front(const T* x, int64_t n, const T* data, int64_t w, int bc_kind, int levels, const int32_t* lk,
                                  const T* lv, const int32_t* rk, const T* rv, const int32_t* pos,
                                  T* fac, size_t fac_stride, T* R, unsigned long long* err, cudaStream_t st);

}  // namespace ndi
