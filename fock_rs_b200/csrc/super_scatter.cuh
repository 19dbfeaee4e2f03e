// super_scatter.cuh -- where one unique integral lands in the J - 1/2 K supermatrix
//   P[ab,cd] = (ab|cd) - 1/4 (ac|bd) - 1/4 (ad|bc)        (canonical a>=b, c>=d, ab>=cd)
// which has the integrals' full 8-fold symmetry and satisfies, for a symmetric density,
//   G_ab = sum_{c>=d} P[ab,cd] * D_cd * (2 - delta_cd)  ==  sum_xy D[x,y] ((ab|xy) - 1/2 (ax|by))
// (the reference's electron_terms / G loops, scf/src/lib.rs:167-185).
//
// A unique integral v = (ij|kl) appears in at most three supermatrix elements: its own slot
// (coefficient +1) and the slots of (ik|jl) and (il|jk) (coefficient -1/4 per matching exchange
// term of that slot).  The coefficients are obtained by evaluating the definition of the target
// element, so every index coincidence (i==j, ij==kl, ...) is handled by construction.
#pragma once
#include "device_math.cuh"

namespace fock {

struct CanonQuartet {
    int a, b, c, d;  // a>=b, c>=d, tri(a,b) >= tri(c,d)
    size_t ab, cd;
    __device__ __forceinline__ size_t slot() const { return tri(ab, cd); }
    __device__ __forceinline__ bool same(const CanonQuartet &o) const { return ab == o.ab && cd == o.cd; }
};

__device__ __forceinline__ CanonQuartet canon_quartet(int i, int j, int k, int l) {
    CanonQuartet q;
    q.a = i >= j ? i : j;
    q.b = i >= j ? j : i;
    q.c = k >= l ? k : l;
    q.d = k >= l ? l : k;
    // 32-bit triangular indices (N <= 65535 functions), widened afterwards
    q.ab = (unsigned)q.a * ((unsigned)q.a + 1u) / 2u + (unsigned)q.b;
    q.cd = (unsigned)q.c * ((unsigned)q.c + 1u) / 2u + (unsigned)q.d;
    if (q.ab < q.cd) {
        int t = q.a; q.a = q.c; q.c = t;
        t = q.b; q.b = q.d; q.d = t;
        size_t s = q.ab; q.ab = q.cd; q.cd = s;
    }
    return q;
}

// -1/4 * (number of exchange terms of supermatrix element `t` that are the integral `key`)
__device__ __forceinline__ double exchange_coefficient(const CanonQuartet &t, const CanonQuartet &key) {
    const CanonQuartet e1 = canon_quartet(t.a, t.c, t.b, t.d);
    const CanonQuartet e2 = canon_quartet(t.a, t.d, t.b, t.c);
    return -0.25 * (double)((e1.same(key) ? 1 : 0) + (e2.same(key) ? 1 : 0));
}

// Calls f(ab, cd, value) for every supermatrix element the unique integral v = (ij|kl) feeds.
template <class F>
__device__ __forceinline__ void for_each_super_target(int i, int j, int k, int l, double v, F &&f) {
    const CanonQuartet key = canon_quartet(i, j, k, l);
    f(key.ab, key.cd, v);
    const CanonQuartet t1 = canon_quartet(key.a, key.c, key.b, key.d);
    const CanonQuartet t2 = canon_quartet(key.a, key.d, key.b, key.c);
    // Fast path (the overwhelmingly common case in anything but tiny molecules): four distinct
    // functions.  Then (ac|bd) and (ad|bc) are two different slots, each containing the integral in
    // exactly one of its two exchange terms.
    if (key.a != key.b && key.c != key.d && key.a != key.c && key.a != key.d && key.b != key.c && key.b != key.d) {
        const double x = -0.25 * v;
        f(t1.ab, t1.cd, x);
        f(t2.ab, t2.cd, x);
        return;
    }
    const double c1 = exchange_coefficient(t1, key);
    if (c1 != 0.0) f(t1.ab, t1.cd, c1 * v);
    if (!t2.same(t1)) {
        const double c2 = exchange_coefficient(t2, key);
        if (c2 != 0.0) f(t2.ab, t2.cd, c2 * v);
    }
}

// P[slot] += contribution (in-core supermatrix formation fused into the integral kernels); P uses the
// row-padded layout ptri() of device_math.cuh
__device__ __forceinline__ void scatter_super(double *P, int i, int j, int k, int l, double v, size_t mask = ~(size_t)0) {
    for_each_super_target(i, j, k, l, v, [&](size_t ab, size_t cd, double c) { atomicAdd(P + (ptri(ab, cd) & mask), c); });
}

// y += P_contribution * d (integral-direct digestion): the same elements applied on the fly.
__device__ __forceinline__ void digest_super(double *y, const double *d, int i, int j, int k, int l, double v) {
    for_each_super_target(i, j, k, l, v, [&](size_t ab, size_t cd, double c) {
        atomicAdd(y + ab, c * __ldg(d + cd));
        if (ab != cd) atomicAdd(y + cd, c * __ldg(d + ab));
    });
}

}  // namespace fock
