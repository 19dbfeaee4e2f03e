// tensor_kernels.cuh -- kernels around the packed tensor: generic per-function ERI passes, the
// J - 1/2 K supermatrix formation, and the bandwidth-bound digestion (packed symmetric
// matrix-vector product) that replaces the reference's per-iteration G loop.
#pragma once
#include "device_math.cuh"
#include "eri_class.cuh"
#include "eri_generic.cuh"
#include "super_scatter.cuh"

namespace fock {

// ---- generic diagonal pass: (xy|xy) for basis-function pairs not covered by the class kernels
// (basis/src/electron_tensor.rs:90-101).  `skip_class` != 0 skips pairs whose two functions both
// sit in class-eligible shells.
__global__ void generic_diag_kernel(BasisDev bs, BoysTables bt, int mode, const unsigned char *bf_class_ok,
                                    int skip_class, size_t npair, double *diag) {
    const size_t xy = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (xy >= npair) return;
    size_t x, y;
    untri(xy, x, y);
    if (skip_class && bf_class_ok[x] && bf_class_ok[y]) return;
    diag[xy] = eri_function_quartet(bs, (int)x, (int)y, (int)x, (int)y, mode, bt);
}

// ---- generic tensor pass: every canonical slot e = tri(xy, zw) this rank owns and the class
// kernels do not cover, with the reference's screening rule (electron_tensor.rs:129-158).
__global__ void generic_tensor_kernel(BasisDev bs, BoysTables bt, int mode, const unsigned char *bf_class_ok,
                                      int skip_class, int screen, double thr, const double *diag, size_t nelem,
                                      int rank, int world, double *V, double *P, unsigned long long *n_eval) {
    const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * (size_t)world + (size_t)rank;
    if (e >= nelem) return;
    size_t xy, zw, x, y, z, w;
    untri(e, xy, zw);
    untri(xy, x, y);
    untri(zw, z, w);
    if (xy == zw) return;  // diagonal slots: diag_to_tensor_kernel
    if (skip_class && bf_class_ok[x] && bf_class_ok[y] && bf_class_ok[z] && bf_class_ok[w]) return;
    if (screen != 1 && !(sqrt(diag[xy] * diag[zw]) > thr)) return;  // slot stays 0.0
    const double v = eri_function_quartet(bs, (int)x, (int)y, (int)z, (int)w, mode, bt);
    atomicAdd(n_eval, 1ULL);
    if (V) V[e] = v;
    if (P) scatter_super(P, (int)x, (int)y, (int)z, (int)w, v);
}

// ---- integral-direct remainder: the same loop digesting into y instead of storing
__global__ void generic_direct_kernel(BasisDev bs, BoysTables bt, int mode, const unsigned char *bf_class_ok,
                                      int skip_class, int screen, double thr, const double *diag, size_t nelem,
                                      int rank, int world, const double *d, double *y) {
    const size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * (size_t)world + (size_t)rank;
    if (e >= nelem) return;
    size_t xy, zw, x, y_, z, w;
    untri(e, xy, zw);
    untri(xy, x, y_);
    untri(zw, z, w);
    if (xy == zw) return;
    if (skip_class && bf_class_ok[x] && bf_class_ok[y_] && bf_class_ok[z] && bf_class_ok[w]) return;
    if (screen != 1 && !(sqrt(diag[xy] * diag[zw]) > thr)) return;
    const double v = eri_function_quartet(bs, (int)x, (int)y_, (int)z, (int)w, mode, bt);
    digest_super(y, d, (int)x, (int)y_, (int)z, (int)w, v);
}

// ---- the diagonal pass's results go into the tensor unconditionally (the reference computes
// (ij|ij) before and outside the screened loop, electron_tensor.rs:90-101).  With several ranks each
// diagonal slot is written by exactly one of them so that partial tensors sum correctly.
__global__ void diag_to_tensor_kernel(const double *__restrict__ diag, size_t npair, int rank, int world,
                                      double *__restrict__ V, double *__restrict__ P) {
    const size_t xy = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (xy >= npair || (int)(xy % (size_t)world) != rank) return;
    const double v = diag[xy];
    if (V) V[tri(xy, xy)] = v;
    if (P) {
        size_t x, y;
        untri(xy, x, y);
        scatter_super(P, (int)x, (int)y, (int)x, (int)y, v);
    }
}

// integral-direct: the diagonal integrals digested from the stored diagonal
__global__ void diag_digest_kernel(const double *__restrict__ diag, size_t npair, int rank, int world,
                                   const double *__restrict__ d, double *__restrict__ y) {
    const size_t xy = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (xy >= npair || (int)(xy % (size_t)world) != rank) return;
    size_t x, yy;
    untri(xy, x, yy);
    digest_super(y, d, (int)x, (int)yy, (int)x, (int)yy, diag[xy]);
}

// ---- shell-block direct digestion (BlockDirectSink, eri_class.cuh): its O(N^2) companions.
// Ds = (D + D^T)/2 (D column-major N x N), A = 0.
__global__ void symm_density_kernel(const double *__restrict__ D, int n, double *__restrict__ Ds, double *__restrict__ A) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const size_t i = idx % n, j = idx / n;
    Ds[idx] = 0.5 * (D[i + j * (size_t)n] + D[j + i * (size_t)n]);
    A[idx] = 0.0;
}

// the diagonal integrals (xy|xy) in the same weighted form: w = v (x==y ? 1/4 : 1) / 2
// (fx_scale != 0: A holds 64-bit fixed-point integers, deterministic mode -- see BlockDirectSink::add)
__device__ __forceinline__ void acc_add(double *p, double v, double fx_scale) {
    if (fx_scale != 0.0)
        atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)__double2ll_rn(v * fx_scale));
    else
        atomicAdd(p, v);
}

__global__ void diag_digest_block_kernel(const double *__restrict__ diag, size_t npair, int rank, int world,
                                         const double *__restrict__ Ds, int n, double *__restrict__ A, double fx_scale) {
    const size_t xy = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (xy >= npair || (int)(xy % (size_t)world) != rank) return;
    size_t x, y;
    untri(xy, x, y);
    const double w = diag[xy] * (x == y ? 0.125 : 0.5);
    if (w == 0.0) return;
    const size_t N = (size_t)n;
    const double Dxy = Ds[x * N + y];
    acc_add(A + x * N + y, 4.0 * w * Dxy - 0.5 * w * Dxy, fx_scale);  // both Coulomb terms and the (xy|yx) exchange term
    acc_add(A + y * N + x, -0.5 * w * Dxy, fx_scale);
    acc_add(A + x * N + x, -0.5 * w * Ds[y * N + y], fx_scale);
    acc_add(A + y * N + y, -0.5 * w * Ds[x * N + x], fx_scale);
}

// max |v| over a device array as the bit pattern of a non-negative double (they order like integers); NaN -> +inf
__global__ void max_abs_kernel(const double *__restrict__ v, size_t n, unsigned long long *__restrict__ out) {
    double m = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double a = fabs(v[i]);
        if (!(a <= m)) m = (a == a) ? a : __longlong_as_double(0x7ff0000000000000LL);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// G (column-major N x N) = packed y (general-path contributions) + A + A^T
// (inv_scale != 0: A holds fixed-point integers; the two triangles are added as integers, then scaled -- exact)
__global__ void combine_g_kernel(const double *__restrict__ y, const double *__restrict__ A, int n, double *__restrict__ G, double inv_scale) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const size_t i = idx % n, j = idx / n;
    double a;
    if (inv_scale != 0.0) {
        const long long *Ai = reinterpret_cast<const long long *>(A);
        a = (double)(Ai[i * (size_t)n + j] + Ai[j * (size_t)n + i]) * inv_scale;
    } else {
        a = A[i * (size_t)n + j] + A[j * (size_t)n + i];
    }
    G[idx] = y[i >= j ? tri(i, j) : tri(j, i)] + a;  // exactly symmetric
}

// ---- batch of arbitrary function quartets (fock_eri_batch)
__global__ void generic_batch_kernel(BasisDev bs, BoysTables bt, int mode, const int *quartets, size_t n, double *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = eri_function_quartet(bs, quartets[4 * i], quartets[4 * i + 1], quartets[4 * i + 2], quartets[4 * i + 3], mode, bt);
}

// ---- supermatrix: P[ij,kl] = (ij|kl) - 1/4 (ik|jl) - 1/4 (il|jk).  It has the full 8-fold symmetry
// of the integrals, and for a symmetric density  G_ij = sum_{k>=l} P[ij,kl] * D_kl * (2 - delta_kl)
// equals the reference's sum_xy D[x,y] * ((ij|xy) - 1/2 (ix|jy)) (scf/src/lib.rs:167-185): this
// one-off pass plays the role of the reference's `electron_terms` precompute.
__global__ void supermatrix_kernel(const double *__restrict__ V, double *__restrict__ P, size_t nelem) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nelem) return;
    size_t ij, kl, i, j, k, l;
    untri(e, ij, kl);
    untri(ij, i, j);
    untri(kl, k, l);
    P[ptri(ij, kl)] = V[e] - 0.25 * V[packed_index(i, k, j, l)] - 0.25 * V[packed_index(i, l, j, k)];
}

// ---- density packing: d[tri(k,l)] = D_kl + D_lk (k != l), D_kk (k == l); D column-major N x N.
__global__ void pack_density_kernel(const double *__restrict__ D, int n, double *__restrict__ d, double *__restrict__ y,
                                    size_t npair) {
    const size_t kl = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kl >= npair) return;
    size_t k, l;
    untri(kl, k, l);
    d[kl] = (k == l) ? D[k + k * (size_t)n] : D[k + l * (size_t)n] + D[l + k * (size_t)n];
    y[kl] = 0.0;
}

__global__ void unpack_result_kernel(const double *__restrict__ y, int n, double *__restrict__ G) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const size_t i = idx % n, j = idx / n;
    G[idx] = y[i >= j ? tri(i, j) : tri(j, i)];
}

// ---- digestion: y += P_sym * d over the row-padded packed lower triangle (ptri layout), padded
// elements [e_begin, e_end).  A block takes a strip of RB rows and a chunk of THREADS*2*M columns;
// consecutive threads read consecutive 16-byte column pairs of one row (aligned, coalesced 128-bit
// streaming loads), keep RB row partials and 2*M column partials in registers, and flush with one
// atomic per column per strip and one per row per block.
// Algorithmic traffic: 8 bytes per stored element, read exactly once.
constexpr int kSymvRB = 16;
constexpr int kSymvM = 2;  // 128-bit loads per thread per row
constexpr int kSymvThreads = 256;
constexpr int kSymvCB = kSymvThreads * kSymvM * 2;

__global__ void __launch_bounds__(kSymvThreads) symv_packed_kernel(const double *__restrict__ P, const double *__restrict__ d,
                                                                  double *__restrict__ y, size_t np, size_t e_begin,
                                                                  size_t e_end, size_t row_first) {
    const size_t r0 = row_first + (size_t)blockIdx.y * kSymvRB;
    const size_t c0 = (size_t)blockIdx.x * kSymvCB;
    if (r0 >= np) return;
    const size_t rlast = (r0 + kSymvRB <= np ? r0 + kSymvRB : np) - 1;
    if (c0 > rlast) return;
    // skip blocks entirely outside [e_begin, e_end)
    if (ptri(rlast, rlast) < e_begin || prow(r0) >= e_end) return;
    const int tid = threadIdx.x;
    __shared__ double s_drow[kSymvRB];
    __shared__ double s_red[kSymvThreads / 32][kSymvRB];
    if (tid < kSymvRB) s_drow[tid] = (r0 + tid < np) ? d[r0 + tid] : 0.0;
    double dcol[kSymvM][2], colsum[kSymvM][2], rowsum[kSymvRB];
    size_t col[kSymvM];
#pragma unroll
    for (int m = 0; m < kSymvM; ++m) {
        col[m] = c0 + ((size_t)m * kSymvThreads + tid) * 2;
        dcol[m][0] = col[m] < np ? d[col[m]] : 0.0;
        dcol[m][1] = col[m] + 1 < np ? d[col[m] + 1] : 0.0;
        colsum[m][0] = colsum[m][1] = 0.0;
    }
#pragma unroll
    for (int r = 0; r < kSymvRB; ++r) rowsum[r] = 0.0;
    __syncthreads();
    const bool interior = (c0 + kSymvCB <= r0) && prow(r0) >= e_begin && ptri(rlast, rlast) < e_end && (r0 + kSymvRB <= np);
    if (interior) {
#pragma unroll
        for (int r = 0; r < kSymvRB; ++r) {
            const double *rowp = P + prow(r0 + r);
            const double dr = s_drow[r];
#pragma unroll
            for (int m = 0; m < kSymvM; ++m) {
                const double2 v = __ldcs(reinterpret_cast<const double2 *>(rowp + col[m]));
                rowsum[r] = fma(v.x, dcol[m][0], rowsum[r]);
                rowsum[r] = fma(v.y, dcol[m][1], rowsum[r]);
                colsum[m][0] = fma(v.x, dr, colsum[m][0]);
                colsum[m][1] = fma(v.y, dr, colsum[m][1]);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < kSymvRB; ++r) {
            const size_t row = r0 + r;
            if (row < np) {
                const size_t off = prow(row);
                const double dr = s_drow[r];
#pragma unroll
                for (int m = 0; m < kSymvM; ++m) {
                    const size_t c = col[m];
                    // rows are padded to 4 doubles, so the 16-byte pair starting at an even c <= row lies inside it
                    const size_t e0 = off + c;
                    if (c <= row && e0 + 1 >= e_begin && e0 < e_end) {
                        const double2 v = __ldcs(reinterpret_cast<const double2 *>(P + e0));
                        const bool in0 = e0 >= e_begin;
                        const bool in1 = c + 1 <= row && e0 + 1 < e_end;
                        if (in0) {
                            rowsum[r] = fma(v.x, dcol[m][0], rowsum[r]);
                            if (c != row) colsum[m][0] = fma(v.x, dr, colsum[m][0]);
                        }
                        if (in1) {
                            rowsum[r] = fma(v.y, dcol[m][1], rowsum[r]);
                            if (c + 1 != row) colsum[m][1] = fma(v.y, dr, colsum[m][1]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < kSymvM; ++m) {
        if (col[m] < np && colsum[m][0] != 0.0) atomicAdd(y + col[m], colsum[m][0]);
        if (col[m] + 1 < np && colsum[m][1] != 0.0) atomicAdd(y + col[m] + 1, colsum[m][1]);
    }
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int r = 0; r < kSymvRB; ++r) {
        double v = rowsum[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][r] = v;
    }
    __syncthreads();
    if (tid < kSymvRB && r0 + tid < np) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kSymvThreads / 32; ++w) v += s_red[w][tid];
        if (v != 0.0) atomicAdd(y + r0 + tid, v);
    }
}

// Boys function for a batch of arguments through the same tables the class kernels use (tests).
__global__ void boys_poly_batch_kernel(const double *__restrict__ poly, int ni, const double *__restrict__ T, size_t n,
                                       int order, double *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double F0[1], F1[2];
    // global-memory table with the shared-memory layout of the order requested
    switch (order) {
        case 0: boys_poly_eval<0>(T[i], F0, poly, ni); out[i] = F0[0]; break;
        case 1: boys_poly_eval<1>(T[i], F1, poly, ni); out[i] = F1[1]; break;
        default: {
            // L >= 2 reads one table + the exp(-mid) vector laid out behind it: assemble on the fly
            const int L = order;
            const double tmax = (double)ni * 0.15;
            const double t = T[i];
            if (t < tmax) {
                const int k = (int)(t * (200.0 / 30.0));
                const double x = t - ((double)k + 0.5) * 0.15;
                const double *c = poly + (size_t)L * 8 * ni + k;
                double r = c[7 * ni];
                for (int j = 6; j >= 0; --j) r = fma(r, x, c[j * ni]);
                out[i] = r;
            } else {  // asymptotic form, boys/src/lib.rs:65-68
                const double r = rsqrt(t), h = 0.5 * r * r;
                double f = 0.88622692545275801364908374167057 * r;
                for (int m = 1; m <= L; ++m) f *= (double)(2 * m - 1) * h;
                out[i] = f;
            }
        }
    }
}

// ---- Boys function through the generic evaluators (orders the class tables do not carry)
__global__ void boys_generic_batch_kernel(BoysTables bt, int mode, const double *__restrict__ T, size_t n, int order,
                                          double *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double F[11];
    if (mode & FOCK_MODE_BOYS_ACCURATE_BIT) boys_accurate(T[i], order, F, bt.tab); else boys_faithful(T[i], order, F, bt.lut);
    out[i] = F[order];
}

// ---- FP64 FMA peak probe: 8 independent register chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

// ---- one-electron matrices (off the hot path).  One thread per i<=j.
// basis/src/contracted_gaussian.rs:51-137 (primitive S, T, V) and :209-262 (contraction).
__global__ void one_electron_kernel(BasisDev bs, BoysTables bt, int mode, int n_nuc, const double *nuc_pos,
                                    const double *nuc_charge, double *S, double *T, double *V) {
    const int n = bs.n_basis;
    const size_t ij = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ij >= (size_t)n * (n + 1) / 2) return;
    size_t jj, ii;
    untri(ij, jj, ii);  // ii <= jj
    const int i = (int)ii, j = (int)jj;
    const double *A = bs.center + 3 * i, *B = bs.center + 3 * j;
    const double diff[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    int la[3], lb[3];
    for (int k = 0; k < 3; ++k) {
        la[k] = bs.angular[3 * i + k];
        lb[k] = bs.angular[3 * j + k];
    }
    const bool fix_r = (mode & FOCK_MODE_FIX_R_BIT) != 0, accurate = (mode & FOCK_MODE_BOYS_ACCURATE_BIT) != 0;
    const int Tn = la[0] + lb[0], Un = la[1] + lb[1], Vn = la[2] + lb[2], L = Tn + Un + Vn;
    const int su = Vn + 1, st = (Un + 1) * (Vn + 1);
    double s_acc = 0.0, t_acc = 0.0;
    double row[FOCK_EROW];
    double e0[3], ep[3], em[3];  // E(l1,l2,0), E(l1,l2+2,0), E(l1,l2-2,0) per dimension
    for (int pa = bs.prim_offset[i]; pa < bs.prim_offset[i + 1]; ++pa)
        for (int pb = bs.prim_offset[j]; pb < bs.prim_offset[j + 1]; ++pb) {
            const double a = bs.exponent[pa], b = bs.exponent[pb];
            const double cc = bs.coefficient[pa] * bs.coefficient[pb];
            for (int k = 0; k < 3; ++k) {
                hermite_row(la[k], lb[k], diff[k], a, b, row);
                e0[k] = row[0];
                hermite_row(la[k], lb[k] + 2, diff[k], a, b, row);
                ep[k] = row[0];
                hermite_row(la[k], lb[k] - 2, diff[k], a, b, row);
                em[k] = row[0];
            }
            const double c3 = sqrt(powi_rt(FOCK_PI / (a + b), 3));
            const double ov = e0[0] * e0[1] * e0[2] * c3;
            const double term0 = b * (double)(2 * (lb[0] + lb[1] + lb[2]) + 3) * ov;
            const double term1 = -2.0 * (b * b) * (ep[0] * e0[1] * e0[2] * c3 + e0[0] * ep[1] * e0[2] * c3 + e0[0] * e0[1] * ep[2] * c3);
            const double term2 = -0.5 * ((double)(lb[0] * (lb[0] - 1)) * (em[0] * e0[1] * e0[2] * c3) +
                                         (double)(lb[1] * (lb[1] - 1)) * (e0[0] * em[1] * e0[2] * c3) +
                                         (double)(lb[2] * (lb[2] - 1)) * (e0[0] * e0[1] * em[2] * c3));
            s_acc += cc * ov;
            t_acc += cc * (term0 + term1 + term2);
        }
    double v_acc = 0.0;
    double Ex[FOCK_EROW], Ey[FOCK_EROW], Ez[FOCK_EROW], bufA[FOCK_RBUF], bufB[FOCK_RBUF], F[11];
    for (int c = 0; c < n_nuc; ++c)
        for (int pa = bs.prim_offset[i]; pa < bs.prim_offset[i + 1]; ++pa)
            for (int pb = bs.prim_offset[j]; pb < bs.prim_offset[j + 1]; ++pb) {
                const double a = bs.exponent[pa], b = bs.exponent[pb], p = a + b;
                const double X = nuc_pos[3 * c] - (a * A[0] + b * B[0]) / p;
                const double Y = nuc_pos[3 * c + 1] - (a * A[1] + b * B[1]) / p;
                const double Z = nuc_pos[3 * c + 2] - (a * A[2] + b * B[2]) / p;
                const double dsq = X * X + Y * Y + Z * Z;
                hermite_row(la[0], lb[0], diff[0], a, b, Ex);
                hermite_row(la[1], lb[1], diff[1], a, b, Ey);
                hermite_row(la[2], lb[2], diff[2], a, b, Ez);
                if (accurate)
                    boys_accurate(p * dsq, L, F, bt.tab);
                else
                    boys_faithful(p * dsq, L, F, bt.lut);
                const double *R = coulomb_table(Tn, Un, Vn, p, X, Y, Z, F, fix_r, bufA, bufB);
                double s = 0.0;
                for (int t = 0; t <= Tn; ++t)
                    for (int u = 0; u <= Un; ++u)
                        for (int v = 0; v <= Vn; ++v) s += Ex[t] * Ey[u] * Ez[v] * R[t * st + u * su + v];
                v_acc += bs.coefficient[pa] * bs.coefficient[pb] * (-nuc_charge[c] * FOCK_TAU / p * s);
            }
    S[i + (size_t)j * n] = S[j + (size_t)i * n] = s_acc;
    T[i + (size_t)j * n] = T[j + (size_t)i * n] = t_acc;
    V[i + (size_t)j * n] = V[j + (size_t)i * n] = v_acc;
}

}  // namespace fock
