// boys_tables.cpp -- see boys_tables.h.
#include "boys_tables.h"

#include <cmath>

namespace fock {

namespace {

// The reference's boys(T, n), |T| < 30, in long double (boys/src/lib.rs:70-80, 13-21).
long double boys_reference_ld(const double *lut, long double t, int n) {
    const long double index_f = 200.0L * t / 30.0L;
    const long double lower = floorl(index_f);
    const long double w = index_f - lower;
    const int lo = (int)lower;
    int up = (int)ceill(index_f);
    if (up > 199) up = 199;
    long double c = w * (long double)lut[up] + (1.0L - w) * (long double)lut[lo];
    const long double e = expl(-t);
    for (int m = 10; m > n; --m) c = (2.0L * t * c + e) / (long double)(2 * m - 1);
    return c;
}

// Exact F_n(T): Kummer series (all terms positive).
long double boys_exact_ld(long double t, int n) {
    long double term = 1.0L / (2 * n + 1), sum = term;
    for (int j = 1; j < 800; ++j) {
        term *= 2.0L * t / (2 * n + 2 * j + 1);
        sum += term;
        if (term < 1e-23L * sum) break;
    }
    return expl(-t) * sum;
}

// Degree-(N-1) interpolant through N Chebyshev nodes of [-h, h], returned as monomial coefficients in x.
template <class F>
void fit_interval(F &&f, long double mid, long double h, double *coef_out, int stride) {
    constexpr int N = kBoysPolyCoefs;
    long double A[N][N + 1];
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < N; ++i) {
        const long double u = cosl(pi * (2 * i + 1) / (2.0L * N));  // node in [-1, 1]
        long double p = 1.0L;
        for (int j = 0; j < N; ++j) {
            A[i][j] = p;
            p *= u;
        }
        A[i][N] = f(mid + u * h);
    }
    for (int c = 0; c < N; ++c) {  // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int r = c + 1; r < N; ++r)
            if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
        for (int j = 0; j <= N; ++j) {
            const long double t = A[c][j];
            A[c][j] = A[piv][j];
            A[piv][j] = t;
        }
        for (int r = 0; r < N; ++r) {
            if (r == c) continue;
            const long double m = A[r][c] / A[c][c];
            for (int j = c; j <= N; ++j) A[r][j] -= m * A[c][j];
        }
    }
    long double hp = 1.0L;
    for (int j = 0; j < N; ++j) {
        coef_out[(size_t)j * stride] = (double)(A[j][N] / A[j][j] / hp);  // coefficient of x^j, x = T - mid
        hp *= h;
    }
}

}  // namespace

void make_boys_poly_faithful(const double *lut, std::vector<double> &out) {
    const int NI = kBoysPolyIntervals;
    out.assign(boys_poly_faithful_len(), 0.0);
    for (int k = 0; k < NI; ++k) {
        const double mid = ((double)k + 0.5) * 0.15;  // the device computes the midpoint with this same expression
        for (int n = 0; n < kBoysPolyOrders; ++n)
            fit_interval([&](long double t) { return boys_reference_ld(lut, t, n); }, (long double)mid, 0.0749999L,
                         out.data() + ((size_t)n * kBoysPolyCoefs) * NI + k, NI);
        out[(size_t)kBoysPolyOrders * kBoysPolyCoefs * NI + k] = (double)expl(-(long double)mid);
    }
}

void make_boys_poly_accurate(std::vector<double> &out) {
    const int NI = kBoysAccIntervals;
    out.assign(boys_poly_accurate_len(), 0.0);
    for (int k = 0; k < NI; ++k) {
        const double mid = ((double)k + 0.5) * 0.15;
        for (int n = 0; n < kBoysPolyOrders; ++n)
            fit_interval([&](long double t) { return boys_exact_ld(t, n); }, (long double)mid, 0.075L,
                         out.data() + ((size_t)n * kBoysPolyCoefs) * NI + k, NI);
        out[(size_t)kBoysPolyOrders * kBoysPolyCoefs * NI + k] = (double)expl(-(long double)mid);
    }
}

}  // namespace fock
