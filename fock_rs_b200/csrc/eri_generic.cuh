// eri_generic.cuh -- the generic per-basis-function evaluators (any angular momentum with
// per-function l <= 4 and total L <= 10, the reference's Boys-order limit).  One thread
// evaluates one contracted integral the way the reference does (contracted_gaussian.rs:264-318
// over :140-205), but with the E rows and the R table built once per primitive pair / quartet
// instead of re-expanded per term.  These kernels are the correctness anchor on the device,
// serve fock_eri_batch / fock_one_electron, and cover whatever the class kernels do not.
#pragma once
#include "device_math.cuh"

namespace fock {

struct BasisDev {
    int n_basis;
    const double *center;    // [N][3]
    const int *angular;      // [N][3]
    const int *prim_offset;  // [N+1]
    const double *exponent;
    const double *coefficient;
};

// Contracted (ab|cd) for basis functions a,b,c,d.  basis/src/contracted_gaussian.rs:264-318.
__device__ inline double eri_function_quartet(const BasisDev &bs, int ia, int ib, int ic, int id, int mode,
                                              const BoysTables &bt) {
    const double *A = bs.center + 3 * ia, *B = bs.center + 3 * ib, *C = bs.center + 3 * ic, *D = bs.center + 3 * id;
    int la[3], lb[3], lc[3], ld[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        la[k] = bs.angular[3 * ia + k];
        lb[k] = bs.angular[3 * ib + k];
        lc[k] = bs.angular[3 * ic + k];
        ld[k] = bs.angular[3 * id + k];
    }
    const int T = la[0] + lb[0] + lc[0] + ld[0], U = la[1] + lb[1] + lc[1] + ld[1], V = la[2] + lb[2] + lc[2] + ld[2];
    const int L = T + U + V;
    const bool fix_r = (mode & FOCK_MODE_FIX_R_BIT) != 0;
    const bool accurate = (mode & FOCK_MODE_BOYS_ACCURATE_BIT) != 0;
    const double dab[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    const double dcd[3] = {D[0] - C[0], D[1] - C[1], D[2] - C[2]};
    const int nbx = la[0] + lb[0], nby = la[1] + lb[1], nbz = la[2] + lb[2];
    const int nkx = lc[0] + ld[0], nky = lc[1] + ld[1], nkz = lc[2] + ld[2];
    const int su = V + 1, st = (U + 1) * (V + 1);
    const double pref = 2.0 * sqrt(powi_rt(FOCK_PI, 5));

    double Ebx[FOCK_EROW], Eby[FOCK_EROW], Ebz[FOCK_EROW];
    double Ekx[FOCK_EROW], Eky[FOCK_EROW], Ekz[FOCK_EROW];
    double bufA[FOCK_RBUF], bufB[FOCK_RBUF];
    double F[11];
    double sum = 0.0;
    for (int i = bs.prim_offset[ia]; i < bs.prim_offset[ia + 1]; ++i) {
        for (int j = bs.prim_offset[ib]; j < bs.prim_offset[ib + 1]; ++j) {
            const double a = bs.exponent[i], b = bs.exponent[j];
            const double p = a + b;
            const double Px = (a * A[0] + b * B[0]) / p, Py = (a * A[1] + b * B[1]) / p, Pz = (a * A[2] + b * B[2]) / p;
            hermite_row(la[0], lb[0], dab[0], a, b, Ebx);
            hermite_row(la[1], lb[1], dab[1], a, b, Eby);
            hermite_row(la[2], lb[2], dab[2], a, b, Ebz);
            const double cab = bs.coefficient[i] * bs.coefficient[j];
            for (int k = bs.prim_offset[ic]; k < bs.prim_offset[ic + 1]; ++k) {
                for (int l = bs.prim_offset[id]; l < bs.prim_offset[id + 1]; ++l) {
                    const double c = bs.exponent[k], d = bs.exponent[l];
                    const double q = c + d;
                    const double X = (c * C[0] + d * D[0]) / q - Px;
                    const double Y = (c * C[1] + d * D[1]) / q - Py;
                    const double Z = (c * C[2] + d * D[2]) / q - Pz;
                    const double dist_sq = X * X + Y * Y + Z * Z;
                    const double alpha = p * q / (p + q);
                    hermite_row(lc[0], ld[0], dcd[0], c, d, Ekx);
                    hermite_row(lc[1], ld[1], dcd[1], c, d, Eky);
                    hermite_row(lc[2], ld[2], dcd[2], c, d, Ekz);
                    if (accurate)
                        boys_accurate(alpha * dist_sq, L, F, bt.tab);
                    else
                        boys_faithful(alpha * dist_sq, L, F, bt.lut);
                    const double *R = coulomb_table(T, U, V, alpha, X, Y, Z, F, fix_r, bufA, bufB);
                    double s = 0.0;
                    for (int t1 = 0; t1 <= nbx; ++t1)
                        for (int u1 = 0; u1 <= nby; ++u1)
                            for (int v1 = 0; v1 <= nbz; ++v1) {
                                const double eb = Ebx[t1] * Eby[u1] * Ebz[v1];
                                for (int t2 = 0; t2 <= nkx; ++t2)
                                    for (int u2 = 0; u2 <= nky; ++u2)
                                        for (int v2 = 0; v2 <= nkz; ++v2) {
                                            const double r = R[(t1 + t2) * st + (u1 + u2) * su + (v1 + v2)];
                                            const double term = eb * Ekx[t2] * Eky[u2] * Ekz[v2] * r;
                                            s += ((t2 + u2 + v2) & 1) ? -term : term;
                                        }
                            }
                    sum += cab * bs.coefficient[k] * bs.coefficient[l] * (pref * (1.0 / (p * q * sqrt(p + q))) * s);
                }
            }
        }
    }
    return sum;
}

}  // namespace fock
