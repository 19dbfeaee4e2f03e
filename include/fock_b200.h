/*
 * fock_b200.h -- C ABI of the B200-native ERI + Fock-build path for fock-rs.
 *
 * The reference (iggedi-ig-ig/fock-rs) is pure Rust with no FFI/plugin seam of its own
 * (SURVEY.md section 8b).  This header is the seam a maintainer binds from Rust (see
 * INTEGRATION.md for the `extern "C"` block and the three call sites in scf/src/lib.rs).
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions: plain C, opaque handle, int status (0 = OK), caller-owned HOST buffers unless
 * the name says `_device`, no callbacks, nothing thrown across the boundary.  A context is
 * used from one thread at a time (the reference calls this path from one thread too).
 * All matrices are column-major N x N doubles exactly as nalgebra's DMatrix<f64> stores
 * them (scf/src/lib.rs uses DMatrix throughout).
 *
 * There is no CPU fallback: every compute entry point returns FOCK_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef FOCK_B200_H
#define FOCK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOCK_B200_ABI_VERSION 5

typedef struct fock_ctx fock_ctx;

/* Status codes.  The reference panics / returns None where these are nonzero:
 *   FOCK_ERR_BOYS_ORDER  <->  assert!(n <= MAX_TABULATED_N) in boys/src/lib.rs:59-62 */
enum {
    FOCK_OK = 0,
    FOCK_ERR_BAD_ARG = 1,
    FOCK_ERR_BOYS_ORDER = 2,
    FOCK_ERR_CUDA = 3,
    FOCK_ERR_OOM = 4,
    FOCK_ERR_STATE = 5,
    FOCK_ERR_UNSUPPORTED = 6
};

/* Structure-of-arrays view of `&[ContractedGaussian]` (basis/src/contracted_gaussian.rs:9-15,
 * basis/src/primitives.rs:7-11), one entry per basis FUNCTION in the reference's order
 * (scf/src/lib.rs:79-83).  Built by the Rust shim through the public accessors position(),
 * primitives(), angular(), exponent(), coefficient().  `coefficient` already carries the
 * primitive norm (GaussianPrimitive::new, primitives.rs:41-48).  Shell structure is
 * reconstructed inside fock_create (the reference discards it, basis_set/src/lib.rs:84). */
typedef struct {
    int32_t n_basis;            /* N */
    const double *center;       /* [N][3] bohr */
    const int32_t *angular;     /* [N][3] (l,m,n) as GaussianPrimitive.angular */
    const int32_t *prim_offset; /* [N+1] */
    const double *exponent;     /* [prim_offset[N]] */
    const double *coefficient;  /* [prim_offset[N]] */
} fock_basis;

/* mode bits (0 = reproduce the reference's arithmetic) */
#define FOCK_MODE_FAITHFUL 0
#define FOCK_MODE_BOYS_ACCURATE 1 /* accurate Boys function instead of boys/src/lib.rs:58-81 (F1) */
#define FOCK_MODE_FIX_R 2         /* restore the (t-1) factor of basis/src/utils.rs:98-105 (F2)  */
#define FOCK_MODE_ACCURATE 3

/* screen */
#define FOCK_SCREEN_REFERENCE 0 /* basis/src/electron_tensor.rs:129-155: zero unless sqrt((ij|ij)(kl|kl)) > 1e-6 */
#define FOCK_SCREEN_OFF 1       /* every unique integral evaluated */
#define FOCK_SCREEN_SCHWARZ 2   /* same estimate, user threshold `tau` */

/* storage */
#define FOCK_STORAGE_INCORE 0 /* symmetry-packed tensor resident in HBM, streamed every fock_build_g */
#define FOCK_STORAGE_DIRECT 1 /* integrals regenerated and digested inside every fock_build_g */

typedef struct {
    int32_t mode;
    int32_t screen;
    double tau;       /* FOCK_SCREEN_SCHWARZ threshold */
    int32_t storage;
    int32_t device;   /* CUDA ordinal; -1 = current device */
    int32_t rank;     /* this process's shard of the shell-quartet tiles, 0 <= rank < world */
    int32_t world;    /* number of shards (1 = whole job) */
    int32_t keep_eri; /* in-core: also keep the raw (ij|kl) tensor for fock_eri_get/export (default 1) */
    int32_t kernel;   /* 0 = class kernels (default), 1 = force the generic per-function kernel,
                       * 2 = class kernels for s/p shells only, d shells through the general shell-quartet kernel */
    int32_t incore_shard; /* world > 1, in-core: 0 = replicate the build, shard the per-iteration streaming
                             (best when a build is followed by more than ~T_build/T_G iterations);
                             1 = shard the build, every rank streams its own partial supermatrix */
    int32_t deterministic; /* integral-direct only: 1 = the Fock accumulation runs in 64-bit fixed point (integer
                              atomics are associative), so G is bit-identical from run to run, for any launch
                              geometry and any number of ranks; the reference's serial loop (scf/src/lib.rs:177-185)
                              is deterministic, FP64 atomics are not.  Default 0. */
} fock_opts;

void fock_default_opts(fock_opts *opts);

/* Upload the basis once, regroup it into shells / angular-momentum classes, build primitive-pair
 * data.  Replaces the `&[ContractedGaussian]` argument handling of
 * ElectronTensor::from_basis (basis/src/electron_tensor.rs:81-87). */
int fock_create(const fock_basis *basis, const fock_opts *opts, fock_ctx **out);

/* ElectronTensor::from_basis (basis/src/electron_tensor.rs:81-176): diagonal pass, then all
 * unique integrals with the screening rule of `opts.screen`; in-core mode also forms the
 * J-1/2K supermatrix that replaces the `electron_terms` precompute (scf/src/lib.rs:167-173).
 * In direct mode only the diagonal (screening estimates) is computed. */
int fock_eri_build(fock_ctx *ctx);

/* `tensor[(i,j,k,l)]` -- Index<(usize,usize,usize,usize)> for ElectronTensor
 * (basis/src/electron_tensor.rs:179-187).  Needs in-core storage with keep_eri. */
int fock_eri_get(fock_ctx *ctx, int32_t i, int32_t j, int32_t k, int32_t l, double *value);

/* All unique integrals in canonical order: slot tri(tri(x,y),tri(z,w)), x>=y, z>=w,
 * tri(x,y)>=tri(z,w), tri(a,b)=a(a+1)/2+b -- the canonical representative of
 * IntegralIndex::correct_order (basis/src/electron_tensor.rs:37-51).  `n` must equal
 * fock_packed_len(N). */
int fock_eri_export_packed(fock_ctx *ctx, double *host, size_t n);
size_t fock_packed_len(int32_t n_basis);

/* BasisFunction::electron_repulsion_int(a,b,c,d) (basis/src/lib.rs:41,
 * basis/src/contracted_gaussian.rs:264-318) for a batch of basis-function quartets
 * `quartets[n][4]`; no screening.  Evaluated on the GPU by the generic per-function kernel. */
int fock_eri_batch(fock_ctx *ctx, const int32_t *quartets, size_t n, double *values);

/* boys::boys(t, n) (boys/src/lib.rs:58-81) for a batch of arguments, on the device, in the context's
 * Boys mode.  Orders 0..4 go through the per-interval tables the class kernels use, higher orders
 * through the generic evaluator.  order > 10 -> FOCK_ERR_BOYS_ORDER (the reference asserts). */
int fock_boys_batch(fock_ctx *ctx, const double *t, size_t n, int32_t order, double *values);

/* The `diagonal` vector of ElectronTensor::from_basis (basis/src/electron_tensor.rs:90-101): (xy|xy) for every
 * x >= y at slot tri(x,y), the screening estimates of the cut.  `n` must equal N(N+1)/2; after fock_eri_build. */
int fock_diagonal_export(fock_ctx *ctx, double *host, size_t n);

/* The per-iteration G build, scf/src/lib.rs:177-185 (inside utils::hermitian, scf/src/utils.rs:25-28):
 *   G_ij = sum_xy D[(x,y)] * ((ij|xy) - 1/2 (ix|jy)),   G symmetric.
 * D must be symmetric (the reference's density is, by construction).  With world > 1 and a communicator
 * attached (fock_comm_init) the result is the COMPLETE G on every rank -- the partial matrices are summed by
 * one NCCL all-reduce of N*N doubles on the context's stream, inside this call; without a communicator it is
 * this rank's PARTIAL G and the caller sums the partials. */
int fock_build_g(fock_ctx *ctx, const double *D, double *G);

/* Page-locked host matrices for fock_build_g (ABI v5).  When D and/or G lie in page-locked memory (these, or any buffer the host
 * registered with the CUDA runtime), fock_build_g copies straight between them and the device; pageable buffers go through the
 * context's own staging area, one extra host copy each way (about 1 ms per 5 MB matrix).  The reference's DMatrix owns its
 * storage (nalgebra, scf/src/lib.rs:95-96), so a binding keeps one such pair per SCF run and maps/copies the matrix views onto
 * it.  Free with fock_pinned_free; the buffers are not tied to a context. */
int fock_pinned_alloc(size_t bytes, void **out);
int fock_pinned_free(void *p);

/* Same with device-resident D and G (N*N doubles each) on the context's stream. */
int fock_build_g_device(fock_ctx *ctx, const double *D_dev, double *G_dev);

/* One-electron matrices S, T, V (BasisFunction::overlap_int / kinetic_int /
 * nuclear_attraction_int, basis/src/contracted_gaussian.rs:209-262; scf/src/lib.rs:118-128),
 * `nuclei_pos[n][3]`, `nuclei_charge[n]` = PointCharge (basis/src/lib.rs:8-12). */
int fock_one_electron(fock_ctx *ctx, int32_t n_nuclei, const double *nuclei_pos, const double *nuclei_charge,
                      double *S, double *T, double *V);

/* Molecular-orbital evaluation (the renderers' hot loop).
 *   psi(at) = sum_f coeff[f] * chi_f(at)         MolecularOrbital::evaluate, scf/src/molecular_orbitals.rs:37-43
 *   chi_f(at) = sum_k c_k dx^l dy^m dz^n exp(-alpha_k |diff|^2), diff = position - at
 *                                                 ContractedGaussian::evaluate, basis/src/contracted_gaussian.rs:320-331
 * `coeff[N]` is one column of the coefficient matrix; functions with |coeff| <= min_coeff are dropped
 * exactly as MolecularOrbitals::new does (molecular_orbitals.rs:70-79, MIN_COEFFICIENT_MAGNITUDE = 0.05,
 * strict '>'); pass a negative min_coeff to keep every function.
 * fock_orbital_eval: a batch of points `points[n][3]` (bohr) -> `values[n]` in FP64 -- the call
 * pointcloud/src/main.rs:108 makes per Monte-Carlo sample.
 * fock_orbital_voxelize: State::create_wave_texture_data (volumetric/src/main.rs:100-125): voxel
 * i = x + n*y + n*n*z sits at ((x,y,z)/n - 0.5) * box_size; `grid[n^3]` receives psi `as f32`
 * (the R32F texture image; the reference uses n = 200, box_size = 20.0). */
int fock_orbital_eval(fock_ctx *ctx, const double *coeff, double min_coeff, const double *points, size_t n, double *values);
int fock_orbital_voxelize(fock_ctx *ctx, const double *coeff, double min_coeff, int32_t n_voxels, double box_size, float *grid);
/* Kernel time (ms, CUDA events) and the number of functions kept / exponent groups of the last fock_orbital_voxelize. */
int fock_orbital_stats(fock_ctx *ctx, double *kernel_ms, int32_t *functions, int32_t *groups);

/* Use a caller-provided cudaStream_t for all launches (NULL = the context's own stream). */
int fock_set_stream(fock_ctx *ctx, void *cuda_stream);

/* Multi-process plumbing (one process per GPU, `rank`/`world` in fock_opts).  Integral-direct
 * mode shards the shell-quartet tiles over ranks.  In-core mode by default replicates the (one-off)
 * build and shards the per-iteration streaming: after fock_eri_build rank r streams the r-th of `world`
 * equal contiguous element ranges of the packed supermatrix; fock_set_digest_range overrides that
 * range, fock_supermatrix_device exposes the buffer and its length (rows padded to 4 doubles).  With
 * opts.incore_shard = 1 the build is sharded instead (tiles dealt to ranks) and each rank streams its
 * own partial supermatrix in full.  Either way
 * fock_build_g returns a PARTIAL G that the caller sums with one all-reduce of N*N doubles. */
int fock_supermatrix_device(fock_ctx *ctx, void **dev_ptr, size_t *n_elements);
int fock_set_digest_range(fock_ctx *ctx, size_t begin, size_t end);

/* The exchange step behind the ABI (nothing in the reference -- its try_scf is one process).
 * One process (or thread) per GPU, each with its own context created with the same basis/options and its
 * rank/world.  Rank 0 obtains an id, the host ships those FOCK_COMM_ID_BYTES bytes to the other ranks by any means it
 * has, and EVERY rank calls fock_comm_init (collective: it returns when all ranks have joined).  From then on
 * fock_build_g / fock_build_g_device return the complete G on every rank.  NCCL is loaded at run time
 * (libnccl.so.2; FOCK_NCCL_LIB overrides the name); FOCK_ERR_UNSUPPORTED if it cannot be found. */
#define FOCK_COMM_ID_BYTES 128
int fock_comm_unique_id(void *id_out);
int fock_comm_init(fock_ctx *ctx, const void *id);
int fock_comm_destroy(fock_ctx *ctx);   /* also done by fock_destroy */
int fock_comm_attached(fock_ctx *ctx);  /* 1 if fock_build_g returns the complete G */
const char *fock_comm_last_error(void); /* for fock_comm_unique_id, which has no context */

/* Timings of the last fock_eri_build / fock_build_g in milliseconds (CUDA events):
 *   [0] pair/diagonal pass  [1] tensor build after the diagonal launch (memsets, host sort/plan, class kernels)  [2] supermatrix gather (kernel=1 only)
 *   [3] last G build (device part)  [4] last G build incl. host<->device copies
 *   [5] host: sublist sorting  [6] host: job/tile plan  [7] class ERI kernels only (events around
 *   the class launches of the last build or direct G build) */
int fock_timings(fock_ctx *ctx, double *ms, int32_t n);

/* Counters: [0] N  [1] shells  [2] shell pairs  [3] unique shell quartets (unscreened)
 *   [4] shell quartets evaluated by the last build / direct G  [5] primitive quartets evaluated
 *   [6] kernel launches of the last fock_eri_build  [7] kernel launches of the last fock_build_g
 *   [8] model FP64 flops of the class quartets evaluated (SURVEY.md section 8d formula)
 *   [9] integrals evaluated by the per-function generic kernel  [10] warp tiles of the class launches
 *   [11] shell quartets evaluated by the general shell-quartet kernel (irregular shells; d shells with kernel = 2)
 *   [12] s/p/d shells that came out of the regrouping with Cartesian components missing (0 for every shipped basis:
 *        results are right either way, but fragmented d shells run far below the class kernels' speed) */
int fock_stats(fock_ctx *ctx, int64_t *out, int32_t n);

/* Measurement aid (no reference counterpart): sustained FP64 FMA rate of this GPU from a
 * register-resident DFMA-chain kernel, in TFLOP/s (2 flops per FMA).  It is the denominator the
 * ERI kernels' "fraction of FP64 peak" is quoted against (MEASURED_PEAKS.json has no FP64 figure). */
int fock_measure_fp64_peak(fock_ctx *ctx, double *tflops);

const char *fock_last_error(fock_ctx *ctx); /* ctx may be NULL: last fock_create error */
void fock_destroy(fock_ctx *ctx);
int fock_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FOCK_B200_H */
