// orbital.cu -- molecular-orbital evaluation on point batches and on the renderer's voxel grid
// (SURVEY.md section 8 row f4).
//
// Replaces, behind fock_orbital_eval / fock_orbital_voxelize (include/fock_b200.h):
//   ContractedGaussian::evaluate          basis/src/contracted_gaussian.rs:320-331
//   MolecularOrbital::evaluate            scf/src/molecular_orbitals.rs:37-43
//   MolecularOrbitals::new (the |c| > 0.05 filter)   scf/src/molecular_orbitals.rs:70-79
//   State::create_wave_texture_data       volumetric/src/main.rs:100-125
//
// One thread = one point.  The reference recomputes exp(-alpha r^2) for every basis FUNCTION; the
// functions of one shell (and the s and p parts of an SP shell) share centre and exponents, so the
// host groups consecutive selected functions with bit-identical centre and exponent list and the
// kernel evaluates each exponential once per group -- the value is the same double, so the result
// does not change.  All other arithmetic keeps the reference's order of operations with explicit
// round-to-nearest intrinsics (no FMA contraction: rustc does not fuse), so the only difference
// from a Rust run is the last-ulp behaviour of exp().
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "context.h"

namespace fock {

constexpr int kOrbCacheK = 8;  // exponentials kept in registers per group

struct OrbGroup {
    double cx, cy, cz;
    int fn_begin, fn_count;  // into OrbFn
    int exp_off, nprim;      // into the basis exponent array
};
struct OrbFn {
    double cmo;  // molecular-orbital coefficient of this function
    int lx, ly, lz;
    int coef_off;  // into the basis coefficient array (nprim entries)
};

// f64::powi -> compiler-rt __powidf2: square-and-multiply from the low bit.
__device__ __forceinline__ double powi_rn(double a, int b) {
    double r = 1.0;
    while (true) {
        if (b & 1) r = __dmul_rn(r, a);
        b >>= 1;
        if (b == 0) break;
        a = __dmul_rn(a, a);
    }
    return r;
}

__device__ __forceinline__ double orbital_at(const OrbGroup *__restrict__ grp, int ngrp, const OrbFn *__restrict__ fns,
                                             const double *__restrict__ expo, const double *__restrict__ coef, double px,
                                             double py, double pz) {
    double total = 0.0;
    for (int g = 0; g < ngrp; ++g) {
        const OrbGroup G = grp[g];  // uniform across the warp: broadcast loads
        // let diff = self.position - at;  (contracted_gaussian.rs:321)
        const double dx = __dsub_rn(G.cx, px), dy = __dsub_rn(G.cy, py), dz = __dsub_rn(G.cz, pz);
        const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (G.nprim <= kOrbCacheK) {
            double e[kOrbCacheK];
#pragma unroll
            for (int k = 0; k < kOrbCacheK; ++k) e[k] = (k < G.nprim) ? exp(__dmul_rn(-expo[G.exp_off + k], r2)) : 0.0;
            for (int f = 0; f < G.fn_count; ++f) {
                const OrbFn F = fns[G.fn_begin + f];
                const double ax = powi_rn(dx, F.lx), ay = powi_rn(dy, F.ly), az = powi_rn(dz, F.lz);
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < kOrbCacheK; ++k)
                    if (k < G.nprim) {
                        // acc + coefficient * dx^i * dy^j * dz^k * exp(-alpha r^2)   (:325-329)
                        const double t = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(coef[F.coef_off + k], ax), ay), az), e[k]);
                        acc = __dadd_rn(acc, t);
                    }
                total = __dadd_rn(total, __dmul_rn(F.cmo, acc));  // molecular_orbitals.rs:41
            }
        } else {  // deep contraction: exponentials recomputed per function, as the reference does
            for (int f = 0; f < G.fn_count; ++f) {
                const OrbFn F = fns[G.fn_begin + f];
                const double ax = powi_rn(dx, F.lx), ay = powi_rn(dy, F.ly), az = powi_rn(dz, F.lz);
                double acc = 0.0;
                for (int k = 0; k < G.nprim; ++k) {
                    const double ek = exp(__dmul_rn(-expo[G.exp_off + k], r2));
                    const double t = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(coef[F.coef_off + k], ax), ay), az), ek);
                    acc = __dadd_rn(acc, t);
                }
                total = __dadd_rn(total, __dmul_rn(F.cmo, acc));
            }
        }
    }
    return total;
}

__global__ void __launch_bounds__(256) orbital_points_kernel(const OrbGroup *__restrict__ grp, int ngrp, const OrbFn *__restrict__ fns,
                                                             const double *__restrict__ expo, const double *__restrict__ coef,
                                                             const double *__restrict__ pts, size_t n, double *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = orbital_at(grp, ngrp, fns, expo, coef, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}

// volumetric/src/main.rs:108-118: voxel i -> x = i % n, y = (i / n) % n, z = i / n^2,
// at = (x as f64 / n as f64 - 0.5) * BOX_SIZE, value stored `as f32`.  x is fastest, so a warp writes
// 128 contiguous bytes.
__global__ void __launch_bounds__(256) orbital_voxel_kernel(const OrbGroup *__restrict__ grp, int ngrp, const OrbFn *__restrict__ fns,
                                                            const double *__restrict__ expo, const double *__restrict__ coef,
                                                            int nv, double box, size_t n, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t nv2 = (size_t)nv * nv;
    const int x = (int)(i % nv), y = (int)((i / nv) % nv), z = (int)(i / nv2);
    const double dn = (double)nv;
    const double px = __dmul_rn(__dsub_rn(__ddiv_rn((double)x, dn), 0.5), box);
    const double py = __dmul_rn(__dsub_rn(__ddiv_rn((double)y, dn), 0.5), box);
    const double pz = __dmul_rn(__dsub_rn(__ddiv_rn((double)z, dn), 0.5), box);
    out[i] = __double2float_rn(orbital_at(grp, ngrp, fns, expo, coef, px, py, pz));
}

namespace {

// MolecularOrbitals::new keeps element.abs() > MIN_COEFFICIENT_MAGNITUDE (strict); NaN fails the test.
// Consecutive kept functions with bit-identical centre and exponent list form one group.
void plan_orbital(const fock_ctx *ctx, const double *coeff, double min_coeff, std::vector<OrbGroup> &grp, std::vector<OrbFn> &fns) {
    const int n = ctx->N;
    for (int i = 0; i < n; ++i) {
        if (!(std::fabs(coeff[i]) > min_coeff)) continue;
        const int off = ctx->h_primoff[i], np = ctx->h_primoff[i + 1] - off;
        OrbFn f{coeff[i], ctx->h_angular[3 * i], ctx->h_angular[3 * i + 1], ctx->h_angular[3 * i + 2], off};
        bool join = false;
        if (!grp.empty()) {
            const OrbGroup &g = grp.back();
            join = g.nprim == np && np <= kOrbCacheK && std::memcmp(&g.cx, &ctx->h_center[3 * i], 3 * sizeof(double)) == 0 &&
                   std::memcmp(&ctx->h_exp[g.exp_off], &ctx->h_exp[off], np * sizeof(double)) == 0;
        }
        if (join) {
            grp.back().fn_count++;
        } else {
            grp.push_back(OrbGroup{ctx->h_center[3 * i], ctx->h_center[3 * i + 1], ctx->h_center[3 * i + 2], (int)fns.size(), 1, off, np});
        }
        fns.push_back(f);
    }
}

struct OrbDev {
    OrbGroup *grp = nullptr;
    OrbFn *fns = nullptr;
    void *out = nullptr;
    double *pts = nullptr;
    ~OrbDev() {
        cudaFree(grp);
        cudaFree(fns);
        cudaFree(out);
        cudaFree(pts);
    }
};

int upload_plan(fock_ctx *ctx, const std::vector<OrbGroup> &grp, const std::vector<OrbFn> &fns, OrbDev &d) {
    FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&d.grp, std::max<size_t>(grp.size(), 1) * sizeof(OrbGroup)));
    FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&d.fns, std::max<size_t>(fns.size(), 1) * sizeof(OrbFn)));
    if (!grp.empty()) {
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(d.grp, grp.data(), grp.size() * sizeof(OrbGroup), cudaMemcpyHostToDevice, ctx->stream));
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(d.fns, fns.data(), fns.size() * sizeof(OrbFn), cudaMemcpyHostToDevice, ctx->stream));
    }
    return FOCK_OK;
}

int negative_angular(const fock_ctx *ctx) {
    for (int v : ctx->h_angular)
        if (v < 0) return 1;
    return 0;
}

}  // namespace
}  // namespace fock

using namespace fock;

extern "C" int fock_orbital_eval(fock_ctx *ctx, const double *coeff, double min_coeff, const double *points, size_t n, double *values) {
    if (!ctx || !coeff || (n && (!points || !values))) return FOCK_ERR_BAD_ARG;
    if (negative_angular(ctx)) return FOCK_ERR_BAD_ARG;
    if (n == 0) return FOCK_OK;
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<OrbGroup> grp;
    std::vector<OrbFn> fns;
    plan_orbital(ctx, coeff, min_coeff, grp, fns);
    OrbDev d;
    FOCK_TRY(upload_plan(ctx, grp, fns, d));
    FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&d.pts, 3 * n * sizeof(double)));
    FOCK_CUDA_TRY(ctx, cudaMalloc(&d.out, n * sizeof(double)));
    cudaStream_t st = ctx->stream;
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(d.pts, points, 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
    orbital_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.grp, (int)grp.size(), d.fns, ctx->d_exp, ctx->d_coef, d.pts, n,
                                                                      (double *)d.out);
    FOCK_CUDA_TRY(ctx, cudaGetLastError());
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(values, d.out, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return FOCK_OK;
}

extern "C" int fock_orbital_voxelize(fock_ctx *ctx, const double *coeff, double min_coeff, int32_t n_voxels, double box_size, float *grid) {
    if (!ctx || !coeff || n_voxels < 0 || n_voxels > 2048 || (n_voxels && !grid)) return FOCK_ERR_BAD_ARG;
    if (negative_angular(ctx)) return FOCK_ERR_BAD_ARG;
    if (n_voxels == 0) return FOCK_OK;
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)n_voxels * n_voxels * n_voxels;
    std::vector<OrbGroup> grp;
    std::vector<OrbFn> fns;
    plan_orbital(ctx, coeff, min_coeff, grp, fns);
    OrbDev d;
    FOCK_TRY(upload_plan(ctx, grp, fns, d));
    FOCK_CUDA_TRY(ctx, cudaMalloc(&d.out, n * sizeof(float)));
    cudaStream_t st = ctx->stream;
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_ORB0], st));
    orbital_voxel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.grp, (int)grp.size(), d.fns, ctx->d_exp, ctx->d_coef, n_voxels, box_size,
                                                                     n, (float *)d.out);
    FOCK_CUDA_TRY(ctx, cudaGetLastError());
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_ORB1], st));
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(grid, d.out, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[EV_ORB0], ctx->ev[EV_ORB1]);
    ctx->orbital_ms = ms;
    ctx->orbital_groups = (int)grp.size();
    ctx->orbital_functions = (int)fns.size();
    return FOCK_OK;
}

extern "C" int fock_orbital_stats(fock_ctx *ctx, double *kernel_ms, int32_t *functions, int32_t *groups) {
    if (!ctx) return FOCK_ERR_BAD_ARG;
    if (kernel_ms) *kernel_ms = ctx->orbital_ms;
    if (functions) *functions = ctx->orbital_functions;
    if (groups) *groups = ctx->orbital_groups;
    return FOCK_OK;
}
