// device_math.cuh -- FP64 building blocks shared by every kernel: Boys function (reference-faithful
// and accurate), the Hermite E rows and the Coulomb R table of the McMurchie-Davidson scheme,
// written iteratively but with the reference's recursion relations and branch order.
//
// Reference: boys/src/lib.rs:13-21,58-81 (Boys), basis/src/utils.rs:33-50 (hermite_expansion),
// basis/src/utils.rs:73-106 (coulomb_auxiliary, incl. the x-branch quirk F2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FOCK_PI 3.14159265358979323846264338327950288
#define FOCK_TAU 6.28318530717958647692528676655900577

#define FOCK_MODE_BOYS_ACCURATE_BIT 1
#define FOCK_MODE_FIX_R_BIT 2

// Accurate-mode Boys table: F_n(k*DT), n = 0..BOYS_TAB_ORDERS-1, k = 0..BOYS_TAB_POINTS-1
#define BOYS_TAB_DT 0.0625
#define BOYS_TAB_INV_DT 16.0
#define BOYS_TAB_TMAX 40.0
#define BOYS_TAB_POINTS 641
#define BOYS_TAB_ORDERS 18  // max order 10 + 7 Taylor terms

namespace fock {

struct BoysTables {
    const double *lut;  // [200] reference F_10 table (boys/src/lut.rs)
    const double *tab;  // [BOYS_TAB_POINTS][BOYS_TAB_ORDERS] accurate-mode grid
};

// f64::powi (compiler-rt __powidf2)
__device__ __forceinline__ double powi_rt(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

// Reference Boys function for all orders 0..L (L <= 10) in one downward pass.  The reference
// restarts the recursion from the F_10 table for every order it needs; the intermediate values
// are the same numbers, so one pass reproduces every call.  boys/src/lib.rs:58-81.
__device__ __forceinline__ void boys_faithful(double t, int L, double *F, const double *lut) {
    const double at = fabs(t);
    if (at < 1e-16) {
        for (int n = 0; n <= L; ++n) F[n] = 1.0 / (double)(2 * n + 1);
    } else if (at >= 30.0) {
        double df = 1.0;  // (2n-1)!!
        for (int n = 0; n <= L; ++n) {
            if (n > 0) df *= (double)(2 * n - 1);
            F[n] = df * (1.0 / powi_rt(2.0, n + 1)) * sqrt(FOCK_PI * powi_rt(t, -2 * n - 1));
        }
    } else {
        const double index_f = 200.0 * t / 30.0;
        const double lower = floor(index_f);
        const double w = index_f - lower;
        const int lo = (int)lower;
        int up = (int)ceil(index_f);
        if (up > 199) up = 199;
        double c = w * __ldg(lut + up) + (1.0 - w) * __ldg(lut + lo);
        const double e = exp(-t);
        if (L >= 10) F[10] = c;
        for (int m = 10; m >= 1; --m) {
            c = (2.0 * t * c + e) / (double)(2 * m - 1);
            if (m - 1 <= L) F[m - 1] = c;
        }
    }
}

// Accurate Boys function, orders 0..L: 7-term Taylor expansion of F_L about the nearest grid
// point + stable downward recursion for T < 40; asymptotic F_0 + upward recursion beyond
// (exp(-40) < 5e-18).  Relative error ~1e-14.
__device__ __forceinline__ void boys_accurate(double t, int L, double *F, const double *tab) {
    t = fabs(t);
    if (t < BOYS_TAB_TMAX) {
        const int k = (int)rint(t * BOYS_TAB_INV_DT);
        const double d = (double)k * BOYS_TAB_DT - t;  // = -(t - t_k)
        const double *row = tab + (size_t)k * BOYS_TAB_ORDERS + L;
        double term = 1.0, s = __ldg(row);
#pragma unroll
        for (int j = 1; j <= 6; ++j) {
            term *= d / (double)j;
            s += __ldg(row + j) * term;
        }
        F[L] = s;
        const double e = exp(-t);
        for (int n = L; n >= 1; --n) F[n - 1] = (2.0 * t * F[n] + e) / (double)(2 * n - 1);
    } else {
        F[0] = 0.5 * sqrt(FOCK_PI / t);
        const double h = 0.5 / t;
        for (int n = 0; n < L; ++n) F[n + 1] = (double)(2 * n + 1) * F[n] * h;
    }
}

// One row E^{la,lb}_t, t = 0..la+lb, of the Hermite expansion coefficients in one dimension,
// following basis/src/utils.rs:33-50: raise i first with j = 0, then raise j at i = la
// (the recursion the reference unwinds in the opposite direction).  `dist` = B - A component.
// Negative la or lb (kinetic-energy terms) give an all-zero row, like the reference's
// `t > i + j` guard does for every case the reference actually multiplies by a nonzero factor.
#define FOCK_EROW 12
__device__ __forceinline__ void hermite_row(int la, int lb, double dist, double a, double b, double *row) {
#pragma unroll
    for (int t = 0; t < FOCK_EROW; ++t) row[t] = 0.0;
    if (la < 0 || lb < 0) return;
    const double p = a + b;
    const double q = a * b / p;
    const double half_ip = 1.0 / (2.0 * p);
    row[0] = exp(-q * (dist * dist));
    const double fa = q * dist / a, fb = q * dist / b;
    double tmp[FOCK_EROW];
    for (int i = 1; i <= la; ++i) {
        for (int t = 0; t <= i; ++t)
            tmp[t] = half_ip * (t > 0 ? row[t - 1] : 0.0) - fa * row[t] + (double)(t + 1) * row[t + 1];
        for (int t = 0; t <= i; ++t) row[t] = tmp[t];
    }
    for (int j = 1; j <= lb; ++j) {
        for (int t = 0; t <= la + j; ++t)
            tmp[t] = half_ip * (t > 0 ? row[t - 1] : 0.0) + fb * row[t] + (double)(t + 1) * row[t + 1];
        for (int t = 0; t <= la + j; ++t) row[t] = tmp[t];
    }
}

// R^0_{tuv} for t<=T, u<=U, v<=V, built level by level in the auxiliary index n (two ping-pong
// buffers of (T+1)(U+1)(V+1) doubles) with the reference's reduction order: v while t=u=0, then u
// while t=0, then t -- and the t step WITHOUT the (t-1) factor unless `fix_r`
// (basis/src/utils.rs:73-106).  F holds Boys values F[0..L], L = T+U+V.  Returns the buffer
// holding level 0; entry (t,u,v) is at (t*(U+1)+u)*(V+1)+v.
#define FOCK_RBUF 80  // max (T+1)(U+1)(V+1) for T+U+V <= 10
__device__ __forceinline__ double *coulomb_table(int T, int U, int V, double alpha, double X, double Y, double Z,
                                                 const double *F, bool fix_r, double *bufA, double *bufB) {
    const int L = T + U + V;
    const int su = V + 1, st = (U + 1) * (V + 1);
    double *cur = bufA, *prev = bufB;
    for (int n = L; n >= 0; --n) {
        double *sw = cur;
        cur = prev;
        prev = sw;  // prev = level n+1, cur = level n
        const int lim = L - n;
        cur[0] = powi_rt(-2.0 * alpha, n) * F[n];
        for (int v = 1; v <= V && v <= lim; ++v)
            cur[v] = Z * prev[v - 1] + (v > 1 ? (double)(v - 1) * prev[v - 2] : 0.0);
        for (int u = 1; u <= U && u <= lim; ++u)
            for (int v = 0; v <= V && u + v <= lim; ++v)
                cur[u * su + v] = Y * prev[(u - 1) * su + v] + (u > 1 ? (double)(u - 1) * prev[(u - 2) * su + v] : 0.0);
        for (int t = 1; t <= T && t <= lim; ++t)
            for (int u = 0; u <= U && t + u <= lim; ++u)
                for (int v = 0; v <= V && t + u + v <= lim; ++v) {
                    double second = 0.0;
                    if (t > 1) {
                        second = prev[(t - 2) * st + u * su + v];
                        if (fix_r) second *= (double)(t - 1);
                    }
                    cur[t * st + u * su + v] = X * prev[(t - 1) * st + u * su + v] + second;
                }
    }
    return cur;
}

__host__ __device__ __forceinline__ size_t tri(size_t x, size_t y) { return x * (x + 1) / 2 + y; }  // x >= y

// Canonical 8-fold slot of (xy|zw): basis/src/electron_tensor.rs:37-51.
__host__ __device__ __forceinline__ size_t packed_index(size_t x, size_t y, size_t z, size_t w) {
    if (x < y) { size_t s = x; x = y; y = s; }
    if (z < w) { size_t s = z; z = w; w = s; }
    const size_t xy = tri(x, y), zw = tri(z, w);
    return xy >= zw ? tri(xy, zw) : tri(zw, xy);
}

// Supermatrix storage: packed lower triangle whose ROWS are padded to a multiple of 4 doubles, so that
// every row starts on a 32-byte sector boundary and can be streamed with aligned 128-bit loads.
// Row r = 4a + b starts at 4(a+1)(2a+b); element (r, c), c <= r, sits at prow(r) + c.
__host__ __device__ __forceinline__ size_t prow(size_t r) {
    const size_t a = r >> 2, b = r & 3;
    return 4 * (a + 1) * (2 * a + b);
}
__host__ __device__ __forceinline__ size_t ptri(size_t r, size_t c) { return prow(r) + c; }

// Inverse of tri(): row r with tri(r,0) <= k < tri(r+1,0), and the column.
__host__ __device__ __forceinline__ void untri(size_t k, size_t &r, size_t &c) {
    size_t g = (size_t)((sqrt(8.0 * (double)k + 1.0) - 1.0) * 0.5);
    while (g * (g + 1) / 2 > k) --g;
    while ((g + 1) * (g + 2) / 2 <= k) ++g;
    r = g;
    c = k - g * (g + 1) / 2;
}

}  // namespace fock
