// ndi_oracle.cpp -- CPU restatement of ndarray-interp's batched interpolation hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (ndarray_interp_b200/) never links, imports or calls anything in this directory and
// has no CPU fallback.
//
// Parity status: PINNED.  The reference is Rust and there is no Rust toolchain in this
// image, so the reference itself cannot be executed here.  Every function below restates
// the reference algorithm operation-by-operation (same operand order, true division, no
// FMA contraction: build with -ffp-contract=off and without -ffast-math) and is pinned
// against every golden vector the reference's own tests hold for this path
// (tests/test_oracle_golden.py; vectors transcribed to tests/golden/*.json):
//   * src/vector_extensions.rs:221-302,318-402 (index + monotonic unit tests)
//   * tests/interp1d.rs:21-205, tests/interp2d.rs:27-265 (exact linear / bilinear values)
//   * src/interp1d/strategies/cubic_spline.rs:62-82 (doctest, f64::EPSILON absolute)
//   * tests/cubic_spline_strat.rs (scipy-derived vectors, max_relative 1e-3)
// plus an independent cross-check against scipy.interpolate.CubicSpline.
//
// Reference citations are relative to /root/reference/.
//
// Two functions here are NOT restatements of the reference: rowsplit_thomas (ora_spline_build_rowsplit_*) and
// partition_thomas (ora_spline_build_partition_*), the operation-by-operation specifications of the two other solves the
// product offers for the spline build (csrc/ndi_rowsplit.cu: cyclic reduction + Thomas; csrc/ndi_partition.cu: block
// partition).  Each is checked against the sequential solve above at north_star's tolerances
// (tests/test_oracle_golden.py) and is what those kernels are compared with bit for bit; the reference-order functions
// remain the statement of the reference's own arithmetic, and parity with the reference is claimed through them only.
//
// Element types: f32, f64 (all entry points) and i32 / i64 / u32 / u64 (everything except splines, which
// the reference restricts to float types through SplineNum, cubic_spline.rs:34-49).
// Integer arithmetic wraps on overflow like a Rust release build (pinned against big-integer arithmetic modulo 2^bits,
// tests/test_oracle_golden.py::test_integer_arithmetic_*).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <type_traits>
#include <vector>

namespace {

// ---- status codes (shared with include/ndi_b200.h) -------------------------------------
constexpr int32_t ST_OK = 0;
constexpr int32_t ST_OUT_OF_BOUNDS = 1;     // InterpolateError::OutOfBounds (lib.rs:142-146)
constexpr int32_t ST_NAN_QUERY = 2;         // panic "failed to convert NaN to usize" (vector_extensions.rs:83-84)
constexpr int32_t ST_PERIODIC_MISMATCH = 3; // BuilderError::ValueError (cubic_spline.rs:483-507)
constexpr int32_t ST_INVALID_ARGUMENT = 4;

// Monotonic enum encoding (vector_extensions.rs:24-29)
constexpr int32_t MONO_NOT = 0, MONO_RISING_STRICT = 1, MONO_RISING = 2, MONO_FALLING_STRICT = 3,
                  MONO_FALLING = 4;

// boundary kinds
constexpr int32_t BC_NOT_A_KNOT = 0, BC_NATURAL = 1, BC_CLAMPED = 2, BC_PERIODIC = 3,
                  BC_INDIVIDUAL = 4;
constexpr int32_t SB_NOT_A_KNOT = 0, SB_NATURAL = 1, SB_CLAMPED = 2, SB_FIRST_DERIV = 3,
                  SB_SECOND_DERIV = 4;

template <class T>
inline bool is_nan(T v) {
    if constexpr (std::is_floating_point_v<T>) return v != v;
    return false;
}

// wrapping integer arithmetic / plain float arithmetic --------------------------------------
// (integers: what a release build of the reference computes; a debug build panics on overflow -- for the unsigned
// types that includes every negative difference y2 - y1 or x - x1)
template <class T> inline T t_add(T a, T b) {
    if constexpr (std::is_integral_v<T>) { using U = std::make_unsigned_t<T>; return (T)((U)a + (U)b); } else return a + b;
}
template <class T> inline T t_sub(T a, T b) {
    if constexpr (std::is_integral_v<T>) { using U = std::make_unsigned_t<T>; return (T)((U)a - (U)b); } else return a - b;
}
template <class T> inline T t_mul(T a, T b) {
    if constexpr (std::is_integral_v<T>) { using U = std::make_unsigned_t<T>; return (T)((U)a * (U)b); } else return a * b;
}
template <class T> inline T t_div(T a, T b) {
    if constexpr (std::is_integral_v<T>) {
        if (b == 0) return 0;              // Rust panics; unreachable on a strictly rising grid
        if constexpr (std::is_signed_v<T>) { if (a == std::numeric_limits<T>::min() && b == (T)-1) return a; }
        return a / b;                      // truncating, like Rust (unsigned types: unsigned division)
    } else return a / b;
}

// Linear::calc_frac -- src/interp1d/strategies/linear.rs:29-36
//   b = y1; m = (y2 - y1) / (x2 - x1); m * (x - x1) + b
template <class T>
inline T calc_frac(T x1, T y1, T x2, T y2, T x) {
    T b = y1;
    T m = t_div(t_sub(y2, y1), t_sub(x2, x1));
    return t_add(t_mul(m, t_sub(x, x1)), b);
}

// MonotonicState + monotonic_prop -- src/vector_extensions.rs:40-53, :115-198
template <class T>
int32_t monotonic_prop(const T* x, int64_t n, int64_t stride) {
    if (n <= 1) return MONO_NOT;                                  // :41-43
    enum { INIT, NOT_STRICT, LIKELY } st = INIT;
    int32_t mon = MONO_NOT;
    for (int64_t i = 0; i + 1 < n; ++i) {
        T a = x[i * stride], b = x[(i + 1) * stride];
        switch (st) {
        case INIT:                                                // :135-143
            if (a < b) { st = LIKELY; mon = MONO_RISING_STRICT; }
            else if (a == b) { st = NOT_STRICT; }
            else { st = LIKELY; mon = MONO_FALLING_STRICT; }
            break;
        case NOT_STRICT:                                          // :144-152
            if (a < b) { st = LIKELY; mon = MONO_RISING; }
            else if (a == b) { st = NOT_STRICT; }
            else { st = LIKELY; mon = MONO_FALLING; }
            break;
        case LIKELY:
            if (mon == MONO_RISING_STRICT || mon == MONO_RISING) { // :153-161
                if (a == b) mon = MONO_RISING;
                else if (a < b) { /* unchanged */ }
                else mon = MONO_NOT;
            } else if (mon == MONO_FALLING_STRICT || mon == MONO_FALLING) { // :162-170
                if (a == b) mon = MONO_FALLING;
                else if (a > b) { /* unchanged */ }
                else mon = MONO_NOT;
            }
            break;
        }
        if (st == LIKELY && mon == MONO_NOT) return MONO_NOT;     // short_circuit :180-185
    }
    if (st == NOT_STRICT) return MONO_NOT;                        // finish :191-197
    return mon;
}

// get_lower_index -- src/vector_extensions.rs:55-111 (literal, including the even-spacing
// guess).  *status = ST_NAN_QUERY where the reference panics on the NaN -> usize cast.
template <class T>
int64_t get_lower_index(const T* g, int64_t n, T x, int32_t* status) {
    if (x <= g[0]) return 0;                                      // :61-63
    if (x >= g[n - 1]) return n - 2;                              // :64-66
    int64_t lo = 0, hi = n - 1;                                   // :70
    T mid = calc_frac<T>(g[lo], (T)lo, g[hi], (T)hi, x);          // :71-82
    if (is_nan(mid)) { *status = ST_NAN_QUERY; return 0; }        // :83-84
    int64_t mid_idx;                                              // truncation, NumCast
    if (!(mid < (T)(n - 1))) mid_idx = n - 1;                     // (Rust would panic past n-1;
    else if (std::is_signed_v<T> && mid < (T)0) mid_idx = 0;      //  unreachable for g[0] < x < g[n-1])
    else mid_idx = (int64_t)mid;
    T mid_x = g[mid_idx];                                         // :86
    if (mid_x <= x && x < g[mid_idx + 1]) return mid_idx;         // :88-90 (&& short-circuits)
    if (mid_x <= x) lo = mid_idx; else hi = mid_idx;              // :91-96
    while (lo + 1 < hi) {                                         // :100-109
        int64_t m = (hi - lo) / 2 + lo;
        if (g[m] <= x) lo = m; else hi = m;
    }
    return lo;                                                    // :110
}

template <class T>
inline bool in_range(const T* g, int64_t n, T x) {               // interp1d/mod.rs:384-386
    return g[0] <= x && x <= g[n - 1];
}

// Linear::interp_into x batch loop -- linear.rs:73-98 x interp1d/mod.rs:326-343
// Stops at the first error, later rows untouched (like the reference).
template <class T>
int32_t interp1d_linear(const T* g, int64_t n, const T* data, int64_t w, const T* q, int64_t nq,
                        int32_t extrapolate, T* out, int64_t* first_bad) {
    for (int64_t i = 0; i < nq; ++i) {
        T x = q[i];
        if (!extrapolate && !in_range(g, n, x)) { *first_bad = i; return ST_OUT_OF_BOUNDS; } // :80-84
        int32_t st = ST_OK;
        int64_t idx = get_lower_index(g, n, x, &st);             // :87
        if (st != ST_OK) { *first_bad = i; return st; }
        T x1 = g[idx], x2 = g[idx + 1];                           // :90-91
        const T* y1 = data + idx * w;
        const T* y2 = data + (idx + 1) * w;
        T* t = out + i * w;
        for (int64_t c = 0; c < w; ++c) t[c] = calc_frac<T>(x1, y1[c], x2, y2[c], x); // :94-96
    }
    return ST_OK;
}

// Bilinear::interp_into x batch loop -- bilinear.rs:64-99 x interp2d/mod.rs:287-307
template <class T>
int32_t interp2d_bilinear(const T* gx, int64_t n, const T* gy, int64_t m, const T* data, int64_t w,
                          const T* qx, const T* qy, int64_t nq, int32_t extrapolate, T* out,
                          int64_t* first_bad, int32_t* bad_axis) {
    for (int64_t i = 0; i < nq; ++i) {
        T x = qx[i], y = qy[i];
        if (!extrapolate && !in_range(gx, n, x)) { *first_bad = i; *bad_axis = 0; return ST_OUT_OF_BOUNDS; } // :71-75
        if (!extrapolate && !in_range(gy, m, y)) { *first_bad = i; *bad_axis = 1; return ST_OUT_OF_BOUNDS; } // :76-80
        int32_t st = ST_OK;
        int64_t xi = get_lower_index(gx, n, x, &st);              // :82 -> interp2d/mod.rs:370-372
        if (st != ST_OK) { *first_bad = i; *bad_axis = 0; return st; }
        int64_t yi = get_lower_index(gy, m, y, &st);
        if (st != ST_OK) { *first_bad = i; *bad_axis = 1; return st; }
        T x1 = gx[xi], x2 = gx[xi + 1], y1 = gy[yi], y2 = gy[yi + 1];
        const T* z11 = data + (xi * m + yi) * w;                  // :83-86
        const T* z12 = data + (xi * m + yi + 1) * w;
        const T* z21 = data + ((xi + 1) * m + yi) * w;
        const T* z22 = data + ((xi + 1) * m + yi + 1) * w;
        T* z = out + i * w;
        for (int64_t c = 0; c < w; ++c) {                         // :88-97
            T z1 = calc_frac<T>(x1, z11[c], x2, z21[c], x);
            T z2 = calc_frac<T>(x1, z12[c], x2, z22[c], x);
            z[c] = calc_frac<T>(y1, z1, y2, z2, y);
        }
    }
    return ST_OK;
}

// ---- cubic spline ---------------------------------------------------------------------

// thomas -- cubic_spline.rs:678-721.  Arrays are taken by value in the reference, so the
// caller's a_mid / rhs are scratch here as well.  rhs, k: (len, w) row-major.
template <class T>
void thomas(T* k, const T* a_up, T* a_mid, const T* a_low, T* rhs, int64_t len, int64_t w) {
    for (int64_t i = 1; i < len; ++i) {                           // :690-702
        T ww = a_low[i] / a_mid[i - 1];
        a_mid[i] -= ww * a_up[i - 1];
        T* r = rhs + i * w;
        const T* rl = rhs + (i - 1) * w;
        for (int64_t c = 0; c < w; ++c) r[c] = r[c] - ww * rl[c];
    }
    for (int64_t c = 0; c < w; ++c) k[(len - 1) * w + c] = rhs[(len - 1) * w + c] / a_mid[len - 1]; // :704-708
    for (int64_t i = len - 2; i >= 0; --i) {                      // :711-720
        for (int64_t c = 0; c < w; ++c)
            k[i * w + c] = (rhs[i * w + c] - a_up[i] * k[(i + 1) * w + c]) / a_mid[i];
    }
}

// Row-split variant of the solve (NOT the reference's order): `levels` steps of
// parallel cyclic reduction, then thomas() on each of the 2^levels interleaved systems (rows j, j+S, j+2S, ...).
// It is the operation-by-operation specification the row-split build kernels (csrc/ndi_rowsplit.cu,
// NDI_BUILD_ROWSPLIT) are compared with bit for bit, the way every other kernel is compared with the functions
// above; the functions above remain the statement of the reference's own arithmetic.  One reduction step with stride s:
//     alpha = -(low[i] / mid[i-s])   (0 when i-s < 0)        gamma = -(up[i] / mid[i+s])   (0 when i+s > len-1)
//     low'[i] = alpha * low[i-s]     up'[i] = gamma * up[i+s]
//     mid'[i] = (mid[i] + alpha * up[i-s]) + gamma * low[i+s]
//     rhs'[i] = (rhs[i] + alpha * rhs[i-s]) + gamma * rhs[i+s]          (per column)
// every product and sum rounded on its own (no FMA), the i-s term before the i+s term.
template <class T>
void rowsplit_thomas(T* k, const T* a_up, const T* a_mid, const T* a_low, const T* rhs, int64_t len, int64_t w,
                     int32_t levels) {
    std::vector<T> low(a_low, a_low + len), mid(a_mid, a_mid + len), up(a_up, a_up + len), r(rhs, rhs + len * w);
    std::vector<T> nlow((size_t)len), nmid((size_t)len), nup((size_t)len), nr((size_t)len * w);
    int64_t s = 1;
    for (int32_t lv = 0; lv < levels; ++lv, s *= 2) {
        for (int64_t i = 0; i < len; ++i) {
            const bool hm = i - s >= 0, hp = i + s <= len - 1;
            const T alpha = hm ? -(low[i] / mid[i - s]) : (T)0, gamma = hp ? -(up[i] / mid[i + s]) : (T)0;
            nlow[i] = hm ? alpha * low[i - s] : (T)0;
            nup[i] = hp ? gamma * up[i + s] : (T)0;
            T m = mid[i];
            if (hm) m = m + alpha * up[i - s];
            if (hp) m = m + gamma * low[i + s];
            nmid[i] = m;
            for (int64_t c = 0; c < w; ++c) {
                T v = r[i * w + c];
                if (hm) v = v + alpha * r[(i - s) * w + c];
                if (hp) v = v + gamma * r[(i + s) * w + c];
                nr[i * w + c] = v;
            }
        }
        low.swap(nlow); mid.swap(nmid); up.swap(nup); r.swap(nr);
    }
    for (int64_t j = 0; j < s && j < len; ++j) {                  // the interleaved systems
        const int64_t m = (len - j + s - 1) / s;
        std::vector<T> sl((size_t)m), sm((size_t)m), su((size_t)m), sr((size_t)m * w), sk((size_t)m * w);
        for (int64_t t = 0; t < m; ++t) {
            sl[t] = low[j + t * s]; sm[t] = mid[j + t * s]; su[t] = up[j + t * s];
            for (int64_t c = 0; c < w; ++c) sr[t * w + c] = r[(j + t * s) * w + c];
        }
        thomas(sk.data(), su.data(), sm.data(), sl.data(), sr.data(), m, w);
        for (int64_t t = 0; t < m; ++t)
            for (int64_t c = 0; c < w; ++c) k[(j + t * s) * w + c] = sk[t * w + c];
    }
}
// Partition (substructuring) variant of the solve -- like rowsplit_thomas NOT the reference's order, but the
// operation-by-operation specification of the third build mode (csrc/ndi_partition.cu, NDI_BUILD_PARTITION), compared
// with the kernels bit for bit.  The rows are cut into blocks of m-1 rows separated by single rows (rows m-1, 2m-1, ...):
// with the separators' unknowns known the blocks are independent, so every block is solved on its own -- for the
// right-hand side (g) and for its two couplings (the "spikes" p, q: the block's response to its left and right
// separator) -- the separators' equations become a tridiagonal system of len / m rows, solved the same way
// (recursively; directly once it has at most `top` rows), and k = g - p k_left - q k_right.  The matrix work (block
// factorisations, spikes, reduced matrices) depends on x only and is shared by all columns.  Every multiply-add below is
// ONE fused operation (std::fma), divisions by an eliminated diagonal are multiplications by its IEEE reciprocal.
template <class T>
void partition_thomas(T* k, const T* up, const T* mid, const T* low, const T* rhs, int64_t len, int64_t w, int32_t m,
                      int32_t top) {
    const T one = (T)1, zero = (T)0;
    auto solve_block = [&](int64_t first, int64_t cnt, std::vector<T>& wl, std::vector<T>& rm) {   // Thomas factors of rows [first, first + cnt)
        T mp = mid[first];
        wl[first] = zero; rm[first] = one / mp;
        for (int64_t i = first + 1; i < first + cnt; ++i) {
            wl[i] = low[i] / mp;
            mp = std::fma(-wl[i], up[i - 1], mid[i]);
            rm[i] = one / mp;
        }
    };
    std::vector<T> wl((size_t)len, zero), rm((size_t)len, zero);
    if (len <= top) {                                                                          // direct
        solve_block(0, len, wl, rm);
        for (int64_t c = 0; c < w; ++c) {
            std::vector<T> f((size_t)len);
            f[0] = rhs[c];
            for (int64_t i = 1; i < len; ++i) f[i] = std::fma(-wl[i], f[i - 1], rhs[i * w + c]);
            T kr = f[len - 1] * rm[len - 1];
            k[(len - 1) * w + c] = kr;
            for (int64_t i = len - 2; i >= 0; --i) { kr = std::fma(-up[i], kr, f[i]) * rm[i]; k[i * w + c] = kr; }
        }
        return;
    }
    const int64_t P = len / m;                                   // separators s_c = c m + m - 1, c < P
    const int64_t tail = len - P * m;                            // rows after the last separator
    const int64_t nblk = P + (tail > 0 ? 1 : 0);
    std::vector<T> p((size_t)len, zero), q((size_t)len, zero), g((size_t)len * w, zero);
    for (int64_t c = 0; c < nblk; ++c) {
        const int64_t first = c * m, cnt = c < P ? m - 1 : tail, last = first + cnt - 1;
        solve_block(first, cnt, wl, rm);
        {   // p = A^-1 (low[first] e_first)
            std::vector<T> f((size_t)cnt);
            f[0] = low[first];
            for (int64_t t = 1; t < cnt; ++t) f[t] = -(wl[first + t] * f[t - 1]);
            T v = f[cnt - 1] * rm[last];
            p[last] = v;
            for (int64_t t = cnt - 2; t >= 0; --t) { v = std::fma(-up[first + t], v, f[t]) * rm[first + t]; p[first + t] = v; }
        }
        {   // q = A^-1 (up[last] e_last)
            T v = up[last] * rm[last];
            q[last] = v;
            for (int64_t t = cnt - 2; t >= 0; --t) { v = -(up[first + t] * v) * rm[first + t]; q[first + t] = v; }
        }
        for (int64_t col = 0; col < w; ++col) {                  // g = A^-1 rhs
            std::vector<T> f((size_t)cnt);
            f[0] = rhs[first * w + col];
            for (int64_t t = 1; t < cnt; ++t) f[t] = std::fma(-wl[first + t], f[t - 1], rhs[(first + t) * w + col]);
            T v = f[cnt - 1] * rm[last];
            g[last * w + col] = v;
            for (int64_t t = cnt - 2; t >= 0; --t) { v = std::fma(-up[first + t], v, f[t]) * rm[first + t]; g[(first + t) * w + col] = v; }
        }
    }
    // the separators' equations
    std::vector<T> rlow((size_t)P), rmid((size_t)P), rup((size_t)P), rrhs((size_t)P * w), ks((size_t)P * w);
    for (int64_t c = 0; c < P; ++c) {
        const int64_t s = c * m + m - 1, lastL = s - 1, firstR = s + 1;
        const bool right = firstR < len;
        rlow[c] = -(low[s] * p[lastL]);
        T md = std::fma(-low[s], q[lastL], mid[s]);
        if (right) md = std::fma(-up[s], p[firstR], md);
        rmid[c] = md;
        rup[c] = right ? -(up[s] * q[firstR]) : zero;
        for (int64_t col = 0; col < w; ++col) {
            T v = std::fma(-low[s], g[lastL * w + col], rhs[s * w + col]);
            if (right) v = std::fma(-up[s], g[firstR * w + col], v);
            rrhs[c * w + col] = v;
        }
    }
    partition_thomas(ks.data(), rup.data(), rmid.data(), rlow.data(), rrhs.data(), P, w, m, top);
    for (int64_t c = 0; c < nblk; ++c) {
        const int64_t first = c * m, cnt = c < P ? m - 1 : tail;
        for (int64_t col = 0; col < w; ++col) {
            const T kl = c > 0 ? ks[(c - 1) * w + col] : zero, kr = c < P ? ks[c * w + col] : zero;
            for (int64_t i = first; i < first + cnt; ++i)
                k[i * w + col] = std::fma(-q[i], kr, std::fma(-p[i], kl, g[i * w + col]));
            if (c < P) k[(first + m - 1) * w + col] = kr;
        }
    }
}
// > 0: solve_for_k's solves use partition_thomas with blocks of that many rows (separator included), direct solve from
// min(128, 4 x that many) rows down (set by ora_spline_build_partition_* for the duration of one call)
thread_local int32_t g_partition_block = 0;
inline int32_t partition_top(int32_t block) { return 4 * block < 128 ? 4 * block : 128; }

// > 0: solve_for_k's full-system solve (:672) uses rowsplit_thomas with that many levels (set by
// ora_spline_build_rowsplit_* for the duration of one call; periodic: both solves of the condensed system)
thread_local int32_t g_rowsplit_levels = 0;

template <class T>
struct SingleBc { int32_t kind; T val; };

template <class T>
SingleBc<T> specialize(SingleBc<T> b) {                          // cubic_spline.rs:287-296
    if (b.kind == SB_NATURAL) return {SB_SECOND_DERIV, (T)0.0};
    if (b.kind == SB_CLAMPED) return {SB_FIRST_DERIV, (T)0.0};
    return b;
}

// solve_for_k -- cubic_spline.rs:409-674.  data, k: (len, w) row-major with row stride ld
// (ld >= w lets solve_for_k_individual address one column of a wider array).
// periodic != 0 selects InternalBoundary::Periodic; otherwise Mixed{left,right}.
template <class T>
int32_t solve_for_k(T* k, const T* x, const T* data, int64_t len, int64_t w, int64_t ld,
                    bool periodic, SingleBc<T> left, SingleBc<T> right) {
    const T zero = (T)0.0, one = (T)1.0, two = (T)2.0, three = (T)3.0;   // :435-438
    auto D = [&](int64_t r, int64_t c) -> T { return data[r * ld + c]; };
    auto K = [&](int64_t r, int64_t c) -> T& { return k[r * ld + c]; };

    std::vector<T> a_up(len, zero), a_mid(len, zero), a_low(len, zero);  // :431-433
    for (int64_t n = 1; n + 1 < len; ++n) {                       // :440-451 (x.windows(3))
        T dxn = x[n + 1] - x[n];
        T dxn_1 = x[n] - x[n - 1];
        a_up[n] = dxn_1;
        a_mid[n] = two * (dxn + dxn_1);
        a_low[n] = dxn;
    }
    std::vector<T> rhs((size_t)len * w, zero);                    // :454
    for (int64_t n = 1; n + 1 < len; ++n) {                       // :456-471
        T dxn = x[n + 1] - x[n];
        T dxn_1 = x[n] - x[n - 1];
        for (int64_t c = 0; c < w; ++c) {
            T yl = D(n - 1, c), ym = D(n, c), yr = D(n + 1, c);
            rhs[n * w + c] = three * (dxn * (ym - yl) / dxn_1 + dxn_1 * (yr - ym) / dxn);
        }
    }
    T dx0 = x[1] - x[0];                                          // :473-476
    T dx1 = x[2] - x[1];
    T dx_1 = x[len - 1] - x[len - 2];
    T dx_2 = x[len - 2] - x[len - 3];

    if (periodic && len == 3) {                                   // :480-496
        for (int64_t c = 0; c < w; ++c)
            if (D(0, c) != D(2, c)) return ST_PERIODIC_MISMATCH;
        for (int64_t c = 0; c < w; ++c) {
            T slope0 = (D(1, c) - D(0, c)) / dx0;
            T slope1 = (D(2, c) - D(1, c)) / dx1;
            T v = (slope0 / dx0 + slope1 / dx1) / (one / dx0 + one / dx1);
            K(0, c) = v; K(1, c) = v; K(2, c) = v;
        }
        return ST_OK;
    }
    if (periodic) {                                               // :498-565
        for (int64_t c = 0; c < w; ++c)
            if (D(0, c) != D(len - 1, c)) return ST_PERIODIC_MISMATCH;
        // condensed system: matrix rows 0..len-3, rhs rows 0..len-2  (:512-515)
        const int64_t m = len - 2;
        a_mid[0] = two * (dx_1 + dx0);                            // :517
        a_up[0] = dx_1;                                           // :518
        std::vector<T> slope_1(w), slope_2(w);
        for (int64_t c = 0; c < w; ++c) {
            T slope0 = (D(1, c) - D(0, c)) / dx0;                 // :521
            slope_1[c] = (D(len - 1, c) - D(len - 2, c)) / dx_1;  // :526
            slope_2[c] = (D(len - 2, c) - D(len - 3, c)) / dx_2;  // :527
            rhs[0 * w + c] = (slope_1[c] * dx0 + slope0 * dx_1) * three;            // :529-530
            rhs[(len - 2) * w + c] = (slope_2[c] * dx_1 + slope_1[c] * dx_2) * three; // :531-532
        }
        std::vector<T> rhs1(rhs.begin(), rhs.begin() + (size_t)m * w);   // :534
        std::vector<T> rhs2((size_t)m * w, zero);                 // :535
        T dx_3 = x[len - 3] - x[len - 4];                         // :537
        for (int64_t c = 0; c < w; ++c) rhs2[0 * w + c] = -dx0;   // :536
        for (int64_t c = 0; c < w; ++c) rhs2[(len - 3) * w + c] = -dx_3;  // :538
        std::vector<T> k1((size_t)m * w, zero), k2((size_t)m * w, zero);
        {
            std::vector<T> mid1(a_mid.begin(), a_mid.begin() + m), mid2(mid1);
            if (g_partition_block > 0) {                          // partition specification: both solves of the condensed system
                partition_thomas(k1.data(), a_up.data(), mid1.data(), a_low.data(), rhs1.data(), m, w, g_partition_block, partition_top(g_partition_block));
                partition_thomas(k2.data(), a_up.data(), mid2.data(), a_low.data(), rhs2.data(), m, w, g_partition_block, partition_top(g_partition_block));
            } else if (g_rowsplit_levels > 0) {                   // row-split specification: both solves of the condensed system
                rowsplit_thomas(k1.data(), a_up.data(), mid1.data(), a_low.data(), rhs1.data(), m, w, g_rowsplit_levels);
                rowsplit_thomas(k2.data(), a_up.data(), mid2.data(), a_low.data(), rhs2.data(), m, w, g_rowsplit_levels);
            } else {
                thomas(k1.data(), a_up.data(), mid1.data(), a_low.data(), rhs1.data(), m, w); // :543-549
                thomas(k2.data(), a_up.data(), mid2.data(), a_low.data(), rhs2.data(), m, w); // :550
            }
        }
        for (int64_t c = 0; c < w; ++c) {
            T k_m1 = (rhs[(len - 2) * w + c] - k1[0 * w + c] * dx_2 - k1[(len - 3) * w + c] * dx_1)
                     / (k2[0 * w + c] * dx_2 + k2[(len - 3) * w + c] * dx_1 + two * (dx_1 + dx_2)); // :552-557
            for (int64_t i = 0; i < m; ++i) K(i, c) = k1[i * w + c] + k_m1 * k2[i * w + c]; // :559-560
            K(len - 2, c) = k_m1;                                 // :561
            K(len - 1, c) = K(0, c);                              // :562-563
        }
        return ST_OK;
    }
    if (left.kind == SB_NOT_A_KNOT && right.kind == SB_NOT_A_KNOT && len == 3) { // :569-596
        a_mid[0] = one; a_up[0] = one;
        a_low[1] = dx1; a_mid[1] = two * (dx0 + dx1); a_up[1] = dx0;
        a_low[2] = one; a_mid[2] = one;
        for (int64_t c = 0; c < w; ++c) {
            T slope0 = (D(1, c) - D(0, c)) / dx0;
            T slope1 = (D(2, c) - D(1, c)) / dx1;
            rhs[0 * w + c] = slope0 * two;
            rhs[1 * w + c] = (slope1 * dx0 + slope0 * dx1) * three;
            rhs[2 * w + c] = slope1 * two;
        }
    } else {                                                      // :597-670
        SingleBc<T> l = specialize(left), r = specialize(right);
        if (l.kind == SB_NOT_A_KNOT) {                            // :599-611
            a_mid[0] = dx1;
            T d = x[2] - x[0];
            a_up[0] = d;
            T tmp1 = (dx0 + two * d) * dx1;
            for (int64_t c = 0; c < w; ++c) {
                T y0 = D(0, c), y1 = D(1, c), y2 = D(2, c);
                rhs[0 * w + c] = (tmp1 * (y1 - y0) / dx0 + (dx0 * dx0) * (y2 - y1) / dx1) / d;
            }
        } else if (l.kind == SB_FIRST_DERIV) {                    // :614-618
            a_mid[0] = one; a_up[0] = zero;
            for (int64_t c = 0; c < w; ++c) rhs[0 * w + c] = l.val;
        } else {                                                  // SecondDeriv :619-631
            a_up[0] = dx0; a_mid[0] = two * dx0;
            for (int64_t c = 0; c < w; ++c)
                rhs[0 * w + c] = three * (D(1, c) - D(0, c)) - l.val * (dx0 * dx0) / two;
        }
        if (r.kind == SB_NOT_A_KNOT) {                            // :634-648
            a_mid[len - 1] = dx_1;
            T d = x[len - 1] - x[len - 3];
            a_low[len - 1] = d;
            T tmp1 = (two * d + dx_1) * dx_2;
            for (int64_t c = 0; c < w; ++c) {
                T y_1 = D(len - 1, c), y_2 = D(len - 2, c), y_3 = D(len - 3, c);
                rhs[(len - 1) * w + c] = ((dx_1 * dx_1) * (y_2 - y_3) / dx_2 + tmp1 * (y_1 - y_2) / dx_1) / d;
            }
        } else if (r.kind == SB_FIRST_DERIV) {                    // :651-655
            a_mid[len - 1] = one; a_low[len - 1] = zero;
            for (int64_t c = 0; c < w; ++c) rhs[(len - 1) * w + c] = r.val;
        } else {                                                  // SecondDeriv :656-668
            a_mid[len - 1] = two * dx_1; a_low[len - 1] = dx_1;
            for (int64_t c = 0; c < w; ++c)
                rhs[(len - 1) * w + c] = three * (D(len - 1, c) - D(len - 2, c)) + r.val * (dx_1 * dx_1) / two;
        }
    }
    // thomas on the full system (:672); k has row stride ld, so solve into a dense temp
    std::vector<T> kk((size_t)len * w);
    if (g_partition_block > 0) partition_thomas(kk.data(), a_up.data(), a_mid.data(), a_low.data(), rhs.data(), len, w, g_partition_block, partition_top(g_partition_block));
    else if (g_rowsplit_levels > 0) rowsplit_thomas(kk.data(), a_up.data(), a_mid.data(), a_low.data(), rhs.data(), len, w, g_rowsplit_levels);
    else thomas(kk.data(), a_up.data(), a_mid.data(), a_low.data(), rhs.data(), len, w);
    for (int64_t i = 0; i < len; ++i)
        for (int64_t c = 0; c < w; ++c) K(i, c) = kk[i * w + c];
    return ST_OK;
}

// calc_coefficients -- cubic_spline.rs:310-368 (+ solve_for_k_individual :370-403, which
// recurses down to single columns; equivalent to a loop over the flattened columns).
template <class T>
int32_t spline_build(const T* x, int64_t len, const T* data, int64_t w, int32_t bc_kind,
                     const int32_t* left_kind, const T* left_val, const int32_t* right_kind,
                     const T* right_val, T* a, T* b) {
    if (len < 3 || w < 0) return ST_INVALID_ARGUMENT;             // MINIMUM_DATA_LENGHT :751
    std::vector<T> k((size_t)len * w, (T)0.0);                    // :321
    int32_t st = ST_OK;
    switch (bc_kind) {
    case BC_PERIODIC:
        st = solve_for_k<T>(k.data(), x, data, len, w, w, true, {0, (T)0}, {0, (T)0}); break;
    case BC_NATURAL:                                              // InternalBoundary::specialize :255-274
        st = solve_for_k<T>(k.data(), x, data, len, w, w, false, {SB_NATURAL, (T)0}, {SB_NATURAL, (T)0}); break;
    case BC_CLAMPED:
        st = solve_for_k<T>(k.data(), x, data, len, w, w, false, {SB_CLAMPED, (T)0}, {SB_CLAMPED, (T)0}); break;
    case BC_NOT_A_KNOT:
        st = solve_for_k<T>(k.data(), x, data, len, w, w, false, {SB_NOT_A_KNOT, (T)0}, {SB_NOT_A_KNOT, (T)0}); break;
    case BC_INDIVIDUAL:
        if (!left_kind || !right_kind || !left_val || !right_val) return ST_INVALID_ARGUMENT;
        for (int64_t c = 0; c < w && st == ST_OK; ++c)            // :379-402
            st = solve_for_k<T>(k.data() + c, x, data + c, len, 1, w, false,
                                {left_kind[c], left_val[c]}, {right_kind[c], right_val[c]});
        break;
    default: return ST_INVALID_ARGUMENT;
    }
    if (st != ST_OK) return st;
    for (int64_t i = 0; i + 1 < len; ++i) {                       // :354-365
        T dx = x[i + 1] - x[i];
        for (int64_t c = 0; c < w; ++c) {
            T kl = k[i * w + c], kr = k[(i + 1) * w + c];
            T y = data[i * w + c], yr = data[(i + 1) * w + c];
            a[i * w + c] = kl * dx - (yr - y);
            b[i * w + c] = (yr - y) - kr * dx;
        }
    }
    return ST_OK;
}

template <class T>
inline T rem_euclid(T a, T b) {                                  // f64::rem_euclid via num_traits::Euclid
    T r = std::fmod(a, b);
    return r < (T)0.0 ? r + std::fabs(b) : r;
}

// CubicSplineStrategy::interp_into x batch loop -- cubic_spline.rs:791-830
// extrap_mode: 0 = No, 1 = Yes, 2 = Periodic (:219-224)
template <class T>
int32_t interp1d_cubic(const T* g, int64_t n, const T* data, const T* a, const T* b, int64_t w,
                       const T* q, int64_t nq, int32_t extrap_mode, T* out, int64_t* first_bad) {
    const T one = (T)1.0;
    for (int64_t i = 0; i < nq; ++i) {
        T x = q[i];
        bool inr = in_range(g, n, x);                             // :797
        if (extrap_mode == 0 && !inr) { *first_bad = i; return ST_OUT_OF_BOUNDS; } // :798-802
        if (extrap_mode == 2 && !inr) {                           // :805-809
            T x0 = g[0], xn = g[n - 1];
            x = rem_euclid<T>(x - x0, xn - x0) + x0;
        }
        int32_t st = ST_OK;
        int64_t idx = get_lower_index(g, n, x, &st);             // :811
        if (st != ST_OK) { *first_bad = i; return st; }
        T xl = g[idx], xr = g[idx + 1];
        const T* yl = data + idx * w;
        const T* yr = data + (idx + 1) * w;
        const T* al = a + idx * w;
        const T* bl = b + idx * w;
        T t = (x - xl) / (xr - xl);                               // :818
        T* y = out + i * w;
        for (int64_t c = 0; c < w; ++c)                           // :825-827
            y[c] = (one - t) * yl[c] + t * yr[c] + t * (one - t) * (al[c] * (one - t) + bl[c] * t);
    }
    return ST_OK;
}

// ---- query-sharded multi-thread drivers (CPU baseline only) -----------------------------
// The reference has no internal threading (README.md:17-18); its own MT benches shard the
// queries from outside with rayon (benches/bench_interp1d.rs:49-79).  Same thing here.
template <class F>
int32_t shard_queries(int64_t nq, int32_t nthreads, int64_t* first_bad, F&& fn) {
    if (nthreads < 1) nthreads = 1;
    std::vector<int32_t> st(nthreads, ST_OK);
    std::vector<int64_t> bad(nthreads, -1);
    std::vector<std::thread> th;
    for (int32_t t = 0; t < nthreads; ++t) {
        int64_t lo = nq * t / nthreads, hi = nq * (t + 1) / nthreads;
        th.emplace_back([&, t, lo, hi] {
            int64_t fb = -1;
            st[t] = fn(lo, hi, &fb);
            bad[t] = fb < 0 ? -1 : lo + fb;
        });
    }
    for (auto& t : th) t.join();
    for (int32_t t = 0; t < nthreads; ++t)
        if (st[t] != ST_OK) { *first_bad = bad[t]; return st[t]; }
    return ST_OK;
}

}  // namespace

// ---- C entry points ---------------------------------------------------------------------
#define ORA_COMMON(SFX, T)                                                                         \
    extern "C" int32_t ora_monotonic_prop_##SFX(const T* x, int64_t n, int64_t stride) {          \
        return monotonic_prop<T>(x, n, stride);                                                    \
    }                                                                                              \
    extern "C" int32_t ora_lower_index_##SFX(const T* g, int64_t n, const T* q, int64_t nq,       \
                                             int64_t* out, int64_t* first_bad) {                   \
        for (int64_t i = 0; i < nq; ++i) {                                                         \
            int32_t st = ST_OK;                                                                    \
            out[i] = get_lower_index<T>(g, n, q[i], &st);                                          \
            if (st != ST_OK) { *first_bad = i; return st; }                                        \
        }                                                                                          \
        return ST_OK;                                                                              \
    }                                                                                              \
    extern "C" T ora_calc_frac_##SFX(T x1, T y1, T x2, T y2, T x) {                               \
        return calc_frac<T>(x1, y1, x2, y2, x);                                                    \
    }                                                                                              \
    extern "C" int32_t ora_interp1d_linear_##SFX(const T* g, int64_t n, const T* data, int64_t w, \
                                                 const T* q, int64_t nq, int32_t extrapolate,      \
                                                 T* out, int64_t* first_bad) {                     \
        return interp1d_linear<T>(g, n, data, w, q, nq, extrapolate, out, first_bad);              \
    }                                                                                              \
    extern "C" int32_t ora_interp1d_linear_mt_##SFX(const T* g, int64_t n, const T* data,         \
                                                    int64_t w, const T* q, int64_t nq,             \
                                                    int32_t extrapolate, T* out,                   \
                                                    int64_t* first_bad, int32_t nthreads) {        \
        return shard_queries(nq, nthreads, first_bad, [&](int64_t lo, int64_t hi, int64_t* fb) {   \
            return interp1d_linear<T>(g, n, data, w, q + lo, hi - lo, extrapolate, out + lo * w, fb); \
        });                                                                                        \
    }                                                                                              \
    extern "C" int32_t ora_interp2d_bilinear_##SFX(const T* gx, int64_t n, const T* gy, int64_t m, \
                                                   const T* data, int64_t w, const T* qx,          \
                                                   const T* qy, int64_t nq, int32_t extrapolate,   \
                                                   T* out, int64_t* first_bad, int32_t* bad_axis) {\
        return interp2d_bilinear<T>(gx, n, gy, m, data, w, qx, qy, nq, extrapolate, out,           \
                                    first_bad, bad_axis);                                          \
    }                                                                                              \
    extern "C" int32_t ora_interp2d_bilinear_mt_##SFX(const T* gx, int64_t n, const T* gy,        \
                                                      int64_t m, const T* data, int64_t w,         \
                                                      const T* qx, const T* qy, int64_t nq,        \
                                                      int32_t extrapolate, T* out,                 \
                                                      int64_t* first_bad, int32_t* bad_axis,       \
                                                      int32_t nthreads) {                          \
        int32_t st = shard_queries(nq, nthreads, first_bad, [&](int64_t lo, int64_t hi, int64_t* fb) { \
            int32_t ax = 0;                                                                        \
            int32_t s = interp2d_bilinear<T>(gx, n, gy, m, data, w, qx + lo, qy + lo, hi - lo,     \
                                             extrapolate, out + lo * w, fb, &ax);                  \
            if (s != ST_OK) *bad_axis = ax;                                                        \
            return s;                                                                              \
        });                                                                                        \
        return st;                                                                                 \
    }

#define ORA_SPLINE(SFX, T)                                                                         \
    extern "C" int32_t ora_spline_build_##SFX(const T* x, int64_t n, const T* data, int64_t w,    \
                                              int32_t bc_kind, const int32_t* left_kind,           \
                                              const T* left_val, const int32_t* right_kind,        \
                                              const T* right_val, T* a, T* b) {                    \
        return spline_build<T>(x, n, data, w, bc_kind, left_kind, left_val, right_kind, right_val, \
                               a, b);                                                              \
    }                                                                                              \
    extern "C" int32_t ora_spline_build_rowsplit_##SFX(const T* x, int64_t n, const T* data,      \
                                                       int64_t w, int32_t bc_kind,                 \
                                                       const int32_t* left_kind,                   \
                                                       const T* left_val,                          \
                                                       const int32_t* right_kind,                  \
                                                       const T* right_val, int32_t levels, T* a,   \
                                                       T* b) {                                     \
        g_rowsplit_levels = levels;                                                                \
        int32_t st = spline_build<T>(x, n, data, w, bc_kind, left_kind, left_val, right_kind,      \
                                     right_val, a, b);                                             \
        g_rowsplit_levels = 0;                                                                     \
        return st;                                                                                 \
    }                                                                                              \
    extern "C" int32_t ora_spline_build_partition_##SFX(const T* x, int64_t n, const T* data,     \
                                                        int64_t w, int32_t bc_kind,                \
                                                        const int32_t* left_kind,                  \
                                                        const T* left_val,                         \
                                                        const int32_t* right_kind,                 \
                                                        const T* right_val, int32_t block, T* a,   \
                                                        T* b) {                                    \
        g_partition_block = block;                                                                 \
        int32_t st = spline_build<T>(x, n, data, w, bc_kind, left_kind, left_val, right_kind,      \
                                     right_val, a, b);                                             \
        g_partition_block = 0;                                                                     \
        return st;                                                                                 \
    }                                                                                              \
    extern "C" int32_t ora_interp1d_cubic_##SFX(const T* g, int64_t n, const T* data, const T* a, \
                                                const T* b, int64_t w, const T* q, int64_t nq,     \
                                                int32_t extrap_mode, T* out, int64_t* first_bad) { \
        return interp1d_cubic<T>(g, n, data, a, b, w, q, nq, extrap_mode, out, first_bad);         \
    }                                                                                              \
    extern "C" int32_t ora_interp1d_cubic_mt_##SFX(const T* g, int64_t n, const T* data,          \
                                                   const T* a, const T* b, int64_t w, const T* q,  \
                                                   int64_t nq, int32_t extrap_mode, T* out,        \
                                                   int64_t* first_bad, int32_t nthreads) {         \
        return shard_queries(nq, nthreads, first_bad, [&](int64_t lo, int64_t hi, int64_t* fb) {   \
            return interp1d_cubic<T>(g, n, data, a, b, w, q + lo, hi - lo, extrap_mode,            \
                                     out + lo * w, fb);                                            \
        });                                                                                        \
    }

ORA_COMMON(f32, float)
ORA_COMMON(f64, double)
ORA_COMMON(i32, int32_t)
ORA_COMMON(i64, int64_t)
ORA_COMMON(u32, uint32_t)
ORA_COMMON(u64, uint64_t)
ORA_SPLINE(f32, float)
ORA_SPLINE(f64, double)

extern "C" int32_t ora_hardware_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int32_t)n : 1;
}
