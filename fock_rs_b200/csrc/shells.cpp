// shells.cpp -- see shells.h.
#include "shells.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <sstream>

namespace fock {

// one-dimensional Hermite table of a general record: { E^{ij}_t : i <= la, j <= lb, t <= i+j }, i-major
static int e1d_count(int la, int lb) {
    int n = 0;
    for (int i = 0; i <= la; ++i)
        for (int j = 0; j <= lb; ++j) n += i + j + 1;
    return n;
}
static int e1d_off(int lb, int i, int j) {
    int n = 0;
    for (int ii = 0; ii < i; ++ii)
        for (int jj = 0; jj <= lb; ++jj) n += ii + jj + 1;
    for (int jj = 0; jj < j; ++jj) n += i + jj + 1;
    return n;
}

int pair_fields(int cls) {
    switch (cls) {
        case 0: return 6;   // p Px Py Pz K (pad)
        case 1: return 8;   // p Px Py Pz KWx KWy KWz Kh
        case 2: return 20;  // p Px Py Pz KAB[3][3] KhA[3] KhB[3] Khh
        default:            // pairs with a d shell: p Px Py Pz Ex[] Ey[] Ez[] (general record), padded to even
            return (4 + 3 * e1d_count(kPairClassLA[cls], kPairClassLB[cls]) + 1) / 2 * 2;
    }
}

static bool same_bits(double a, double b) { return std::memcmp(&a, &b, sizeof(double)) == 0; }

bool build_shells(int n, const double *center, const int32_t *angular, const int32_t *prim_offset,
                  const double *exponent, const double *coefficient, ShellData &out, std::string &err) {
    out = ShellData();
    out.bf_shell.assign(n, -1);
    out.bf_comp.assign(n, -1);
    int cur = -1;  // open shell
    for (int f = 0; f < n; ++f) {
        const int lx = angular[3 * f], ly = angular[3 * f + 1], lz = angular[3 * f + 2];
        if (lx < 0 || ly < 0 || lz < 0) {
            err = "negative angular momentum component";
            return false;
        }
        const int L = lx + ly + lz;
        const int np = prim_offset[f + 1] - prim_offset[f];
        if (np <= 0) {
            err = "basis function without primitives";
            return false;
        }
        const double *fe = exponent + prim_offset[f], *fc = coefficient + prim_offset[f];
        const int comp = L <= 2 ? cart_index(lx, ly, lz) : 0;
        bool joined = false;
        if (cur >= 0 && L <= 2) {
            Shell &s = out.shells[cur];
            bool ok = s.L == L && s.nprim == np && s.bf[comp] < 0;
            for (int k = 0; ok && k < 3; ++k) ok = same_bits(s.c[k], center[3 * f + k]);
            for (int i = 0; ok && i < np; ++i) ok = same_bits(out.exps[s.prim_begin + i], fe[i]);
            // one radial coefficient list per shell: a component may differ from it by a constant factor
            // (Cartesian d components differ by the sqrt(3) norm ratio, basis/src/primitives.rs:19-27);
            // the ratio must be the same for every primitive to within the rounding of the two norm chains
            // (sqrt((8a)^2/12) against sqrt((8a)^2/4), the c*norm products and the division: ~16 ulp worst case;
            // the bound is 32 ulp so that a legitimate Cartesian d shell is never split into partial shells)
            double ratio = 1.0;
            if (ok) {
                const double r0 = out.coefs[s.prim_begin];
                ok = r0 != 0.0 && std::isfinite(fc[0] / r0);
                if (ok) ratio = fc[0] / r0;
                for (int i = 0; ok && i < np; ++i) {
                    const double want = ratio * out.coefs[s.prim_begin + i];
                    ok = std::fabs(fc[i] - want) <= 4e-15 * std::fabs(want);
                }
                bool identical = true;
                for (int i = 0; identical && i < np; ++i) identical = same_bits(out.coefs[s.prim_begin + i], fc[i]);
                if (identical) ratio = 1.0;
            }
            if (ok) {
                s.bf[comp] = f;
                s.scale[comp] = ratio;
                if (ratio != 1.0 && s.L <= 1) s.class_ok = false;
                out.bf_shell[f] = cur;
                out.bf_comp[f] = comp;
                joined = true;
            }
        }
        if (!joined) {
            Shell s;
            s.L = L;
            for (int k = 0; k < 3; ++k) s.c[k] = center[3 * f + k];
            s.nprim = np;
            s.prim_begin = (int)out.exps.size();
            for (int i = 0; i < np; ++i) {
                out.exps.push_back(fe[i]);
                out.coefs.push_back(fc[i]);
            }
            if (L <= 2) s.bf[comp] = f; else s.bf[0] = f;
            s.class_ok = (L <= kMaxClassL);
            out.shells.push_back(s);
            cur = (int)out.shells.size() - 1;
            out.bf_shell[f] = cur;
            out.bf_comp[f] = L <= 2 ? comp : 0;
        }
    }
    out.all_class_ok = true;
    out.n_class_shells = 0;
    out.n_partial_shells = 0;
    for (const Shell &s : out.shells) {
        if (s.class_ok) ++out.n_class_shells; else out.all_class_ok = false;
        if (s.L <= 2) {  // an s/p/d shell that did not collect all its Cartesian components (reported in fock_stats[12])
            int have = 0;
            for (int c = 0; c < 6; ++c) have += s.bf[c] >= 0;
            if (have != (s.L + 1) * (s.L + 2) / 2) ++out.n_partial_shells;
        }
    }
    return true;
}

void build_pairs(ShellData &sd) {
    sd.pairs.clear();
    sd.sublists.clear();
    const int ns = (int)sd.shells.size();
    std::map<std::pair<int, int>, int> key_to_sub;  // (cls, K) -> sublist
    for (int a = 0; a < ns; ++a) {
        if (!sd.shells[a].class_ok) continue;
        for (int b = 0; b <= a; ++b) {
            if (!sd.shells[b].class_ok) continue;
            int A = a, B = b;  // shell index order a >= b; put the higher L first
            if (sd.shells[B].L > sd.shells[A].L) std::swap(A, B);
            ShellPair p;
            p.A = A;
            p.B = B;
            p.cls = pair_class(sd.shells[A].L, sd.shells[B].L);
            p.K = sd.shells[A].nprim * sd.shells[B].nprim;
            sd.pairs.push_back(p);
        }
    }
    for (int i = 0; i < (int)sd.pairs.size(); ++i) {
        const ShellPair &p = sd.pairs[i];
        auto key = std::make_pair(p.cls, p.K);
        auto it = key_to_sub.find(key);
        if (it == key_to_sub.end()) {
            Sublist s;
            s.cls = p.cls;
            s.K = p.K;
            s.nf = pair_fields(p.cls);
            sd.sublists.push_back(s);
            it = key_to_sub.emplace(key, (int)sd.sublists.size() - 1).first;
        }
        sd.sublists[it->second].pair_ids.push_back(i);
    }
    // deterministic order: by class, then by K descending
    std::sort(sd.sublists.begin(), sd.sublists.end(), [](const Sublist &x, const Sublist &y) {
        if (x.cls != y.cls) return x.cls < y.cls;
        return x.K > y.K;
    });
}

void sort_sublists(ShellData &sd) {
    for (Sublist &s : sd.sublists) {
        std::stable_sort(s.pair_ids.begin(), s.pair_ids.end(),
                         [&](int x, int y) { return sd.pairs[x].maxdiag > sd.pairs[y].maxdiag; });
    }
}

// One primitive pair record (see pair_fields()).  E coefficients follow the reference's
// hermite_expansion recursion (basis/src/utils.rs:33-50) written out for l <= 1:
//   E^{00}_0 = exp(-q X^2);  E^{10}_0 = -(q X / a) E^{00}_0;  E^{10}_1 = E^{00}_0 / (2p)
//   E^{01}_0 = +(q X / b) E^{00}_0;  E^{01}_1 = E^{00}_0 / (2p)
//   E^{11}_0 = (qX/b) E^{10}_0 + E^{10}_1;  E^{11}_1 = E^{10}_0/(2p) + (qX/b) E^{10}_1;  E^{11}_2 = E^{10}_1/(2p)
// with X = B - A per dimension.  K carries cA cB, the three E^{00}_0 and sqrt(2 pi^{5/2}) / p, so the
// product of a bra and a ket record times 1/sqrt(p+q) is the reference's prefactor
// 2 pi^{5/2} / (p q sqrt(p+q)) (basis/src/contracted_gaussian.rs:169-171).
static void fill_record(const ShellData &sd, const ShellPair &pr, int i, int j, double *rec) {
    const Shell &A = sd.shells[pr.A], &B = sd.shells[pr.B];
    const double a = sd.exps[A.prim_begin + i], b = sd.exps[B.prim_begin + j];
    const double ca = sd.coefs[A.prim_begin + i], cb = sd.coefs[B.prim_begin + j];
    const double p = a + b, q = a * b / p;
    const double kSqrtPref = std::sqrt(2.0 * std::sqrt(std::pow(3.14159265358979323846, 5)));
    double e00 = 1.0, X[3];
    for (int k = 0; k < 3; ++k) {
        X[k] = B.c[k] - A.c[k];
        e00 *= std::exp(-q * (X[k] * X[k]));
    }
    const double K = ca * cb * e00 * kSqrtPref / p;
    rec[0] = p;
    for (int k = 0; k < 3; ++k) rec[1 + k] = (a * A.c[k] + b * B.c[k]) / p;
    if (pr.cls >= 3) {
        // general record: the one-dimensional tables E^{ij}_t of basis/src/utils.rs:33-50 (i lowered while
        // j = 0, then j), the x table carrying c_a c_b sqrt(2 pi^{5/2}) / p
        const int la = A.L, lb = B.L, cnt = e1d_count(la, lb);
        const double half_ip = 1.0 / (2.0 * p);
        for (int dim = 0; dim < 3; ++dim) {
            double E[3][3][5] = {};
            const double fa = q * X[dim] / a, fb = q * X[dim] / b;
            E[0][0][0] = std::exp(-q * (X[dim] * X[dim])) * (dim == 0 ? ca * cb * kSqrtPref / p : 1.0);
            for (int i = 1; i <= la; ++i)
                for (int t = 0; t <= i; ++t)
                    E[i][0][t] = half_ip * (t > 0 ? E[i - 1][0][t - 1] : 0.0) - fa * E[i - 1][0][t] +
                                 (t + 1 <= i - 1 ? (double)(t + 1) * E[i - 1][0][t + 1] : 0.0);
            for (int i = 0; i <= la; ++i)
                for (int j = 1; j <= lb; ++j)
                    for (int t = 0; t <= i + j; ++t)
                        E[i][j][t] = half_ip * (t > 0 ? E[i][j - 1][t - 1] : 0.0) + fb * E[i][j - 1][t] +
                                     (t + 1 <= i + j - 1 ? (double)(t + 1) * E[i][j - 1][t + 1] : 0.0);
            for (int i = 0; i <= la; ++i)
                for (int j = 0; j <= lb; ++j)
                    for (int t = 0; t <= i + j; ++t) rec[4 + dim * cnt + e1d_off(lb, i, j) + t] = E[i][j][t];
        }
        if ((4 + 3 * cnt) & 1) rec[4 + 3 * cnt] = 0.0;
        return;
    }
    if (pr.cls == 0) {
        rec[4] = K;
        rec[5] = 0.0;
        return;
    }
    const double h = 1.0 / (2.0 * p);
    double WA[3], WB[3];
    for (int k = 0; k < 3; ++k) {
        WA[k] = -(q * X[k] / a);
        WB[k] = q * X[k] / b;
    }
    if (pr.cls == 1) {
        for (int k = 0; k < 3; ++k) rec[4 + k] = K * WA[k];
        rec[7] = K * h;
        return;
    }
    for (int da = 0; da < 3; ++da)
        for (int db = 0; db < 3; ++db) rec[4 + da * 3 + db] = K * (WA[da] * WB[db] + (da == db ? h : 0.0));
    for (int k = 0; k < 3; ++k) {
        rec[13 + k] = K * h * WA[k];
        rec[16 + k] = K * h * WB[k];
    }
    rec[19] = K * h * h;
}

void build_pair_tables(ShellData &sd, std::vector<double> &bra, std::vector<double> &ket, std::vector<int> &meta) {
    size_t bra_n = 0, ket_n = 0, meta_n = 0;
    for (Sublist &s : sd.sublists) {
        const size_t np = s.pair_ids.size();
        s.npad = (int)((np + 31) / 32 * 32);
        s.bra_off = bra_n;
        s.ket_off = ket_n;
        s.meta_off = meta_n;
        bra_n += np * s.K * s.nf;
        bra_n = (bra_n + 1) / 2 * 2;
        ket_n += (size_t)s.K * s.nf * s.npad;
        meta_n += np * 2;
    }
    bra.assign(bra_n, 0.0);
    ket.assign(ket_n, 0.0);
    meta.assign(meta_n, 0);
    std::vector<double> rec(96);
    for (const Sublist &s : sd.sublists) {
        for (size_t pi = 0; pi < s.pair_ids.size(); ++pi) {
            const ShellPair &pr = sd.pairs[s.pair_ids[pi]];
            const int KB = sd.shells[pr.B].nprim;
            meta[s.meta_off + 2 * pi] = pr.A;
            meta[s.meta_off + 2 * pi + 1] = pr.B;
            for (int ij = 0; ij < s.K; ++ij) {
                fill_record(sd, pr, ij / KB, ij % KB, rec.data());
                for (int f = 0; f < s.nf; ++f) {
                    bra[s.bra_off + (pi * s.K + ij) * s.nf + f] = rec[f];
                    ket[s.ket_off + ((size_t)ij * s.nf + f) * s.npad + pi] = rec[f];
                }
            }
        }
        // pad ket columns with a harmless record (p=1, K=0) so out-of-range lanes stay finite
        for (size_t pi = s.pair_ids.size(); pi < (size_t)s.npad; ++pi)
            for (int ij = 0; ij < s.K; ++ij) ket[s.ket_off + ((size_t)ij * s.nf + 0) * s.npad + pi] = 1.0;
    }
}

static double binom(int n, int k) {
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

double model_flops_per_prim_quartet(int la, int lb, int lc, int ld) {
    // convention of SURVEY.md section 8d: within a pair l_a >= l_b, bra = lexicographically larger pair
    if (la < lb) std::swap(la, lb);
    if (lc < ld) std::swap(lc, ld);
    if (std::make_pair(la, lb) < std::make_pair(lc, ld)) {
        std::swap(la, lc);
        std::swap(lb, ld);
    }
    const int L = la + lb + lc + ld;
    const double nab = ncart(la) * ncart(lb), ncd = ncart(lc) * ncart(ld);
    const double hab = binom(la + lb + 3, 3), hcd = binom(lc + ld + 3, 3);
    return 84.0 + 3.0 * binom(L + 4, 4) + 2.0 * ncd * hab * hcd + 2.0 * nab * ncd * hab;
}

std::string ShellData::describe() const {
    std::ostringstream os;
    os << shells.size() << " shells, " << pairs.size() << " class pairs, " << sublists.size() << " sublists:";
    for (const Sublist &s : sublists) os << " (cls" << s.cls << ",K" << s.K << ",n" << s.pair_ids.size() << ")";
    return os.str();
}

}  // namespace fock
