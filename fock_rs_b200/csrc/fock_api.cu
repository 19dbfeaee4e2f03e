// fock_api.cu -- implementation of the C ABI in include/fock_b200.h.
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>

#include "boys_tables.h"
#include "class_launch.h"
#include "context.h"
#include "eri_class.cuh"
#include "sq_launch.cuh"
#include "tensor_kernels.cuh"

using namespace fock;

namespace {

thread_local std::string g_create_error;

const double kBoysLut[200] = {
#include "boys_lut.inc"
};

// Accurate-mode grid F_n(k*DT): series  F_n(T) = exp(-T) sum_k (2T)^k / ((2n+1)(2n+3)...(2n+2k+1)).
void make_boys_table(std::vector<double> &tab) {
    tab.resize((size_t)BOYS_TAB_POINTS * BOYS_TAB_ORDERS);
    for (int k = 0; k < BOYS_TAB_POINTS; ++k) {
        const long double t = (long double)k * (long double)BOYS_TAB_DT;
        for (int n = 0; n < BOYS_TAB_ORDERS; ++n) {
            long double term = 1.0L / (2 * n + 1), sum = term;
            for (int j = 1; j < 600; ++j) {
                term *= 2.0L * t / (2 * n + 2 * j + 1);
                sum += term;
                if (term < 1e-22L * sum) break;
            }
            tab[(size_t)k * BOYS_TAB_ORDERS + n] = (double)(expl(-t) * sum);
        }
    }
}

struct ClassState {
    ClassEngine engine;
    ClassStats stats;
    unsigned char *d_bf_class_ok = nullptr;  // function sits in an s/p class shell (its diagonal comes from eri_diag_kernel)
    bool all_sp_class = false;               // ... true for every function
    unsigned char *d_bf_covered = nullptr;   // function is handled by a shell-level path (class or shell-quartet kernel)
    bool all_covered = false;
    SqState sq;
    unsigned long long *d_counter = nullptr;
    double *d_boys_poly = nullptr;
    double *h_diag = nullptr;  // pinned image of d_diag
    cudaEvent_t ev_diag = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // bracket the class-kernel launches
};

ClassState *cls(fock_ctx *ctx);

}  // namespace

// the class engine lives beside the context (kept out of context.h to keep that header light)
struct fock_ctx_ext : fock_ctx {
    ClassState cs;
};
namespace {
ClassState *cls(fock_ctx *ctx) { return &static_cast<fock_ctx_ext *>(ctx)->cs; }

double screen_threshold(const fock_opts &o) { return o.screen == FOCK_SCREEN_SCHWARZ ? o.tau : 1e-6; }

int ensure_g_buffers(fock_ctx *ctx) {
    const size_t nn = (size_t)ctx->N * ctx->N;
    if (!ctx->d_dvec) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_dvec, ctx->npair * sizeof(double)));
    if (!ctx->d_yvec) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_yvec, ctx->npair * sizeof(double)));
    if (!ctx->d_Ds) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_Ds, nn * sizeof(double)));
    if (!ctx->d_A) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_A, nn * sizeof(double)));
    if (!ctx->d_D) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_D, nn * sizeof(double)));
    if (!ctx->d_G) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_G, nn * sizeof(double)));
    if (!ctx->h_pinned) FOCK_CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->h_pinned, 2 * nn * sizeof(double)));
    return FOCK_OK;
}

}  // namespace

extern "C" {

int fock_abi_version(void) { return FOCK_B200_ABI_VERSION; }

void fock_default_opts(fock_opts *o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->mode = FOCK_MODE_FAITHFUL;
    o->screen = FOCK_SCREEN_REFERENCE;
    o->tau = 1e-6;
    o->storage = FOCK_STORAGE_INCORE;
    o->device = -1;
    o->rank = 0;
    o->world = 1;
    o->keep_eri = 1;
    o->kernel = 0;
    o->incore_shard = 0;
    o->deterministic = 0;
}

size_t fock_packed_len(int32_t n) {
    const size_t p = (size_t)n * ((size_t)n + 1) / 2;
    return p * (p + 1) / 2;
}

const char *fock_last_error(fock_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fock_create(const fock_basis *b, const fock_opts *opts_in, fock_ctx **out) {
    g_create_error.clear();
    if (!b || !out) {
        g_create_error = "null argument";
        return FOCK_ERR_BAD_ARG;
    }
    *out = nullptr;
    fock_opts o;
    if (opts_in) o = *opts_in; else fock_default_opts(&o);
    if (b->n_basis <= 0 || !b->center || !b->angular || !b->prim_offset || !b->exponent || !b->coefficient) {
        g_create_error = "empty or incomplete basis";
        return FOCK_ERR_BAD_ARG;
    }
    if (o.world < 1 || o.rank < 0 || o.rank >= o.world || (o.mode & ~3) || o.screen < 0 || o.screen > 2 ||
        o.storage < 0 || o.storage > 1) {
        g_create_error = "bad options";
        return FOCK_ERR_BAD_ARG;
    }
    if (o.deterministic && o.storage != FOCK_STORAGE_DIRECT) {
        g_create_error = "the deterministic (fixed-point) accumulation exists for integral-direct storage only";
        return FOCK_ERR_UNSUPPORTED;
    }
    const int n = b->n_basis;
    if (n > 65535) {
        g_create_error = "more than 65535 basis functions are not supported (32-bit pair indices)";
        return FOCK_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < n; ++i) {
        const int l = b->angular[3 * i] + b->angular[3 * i + 1] + b->angular[3 * i + 2];
        if (b->angular[3 * i] < 0 || b->angular[3 * i + 1] < 0 || b->angular[3 * i + 2] < 0 ||
            b->prim_offset[i + 1] <= b->prim_offset[i]) {
            g_create_error = "malformed basis function";
            return FOCK_ERR_BAD_ARG;
        }
        if (l > 4) {
            g_create_error = "angular momentum above g is not supported";
            return FOCK_ERR_UNSUPPORTED;
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device (this library has no CPU fallback)";
        return FOCK_ERR_CUDA;
    }
    fock_ctx_ext *ctx = new (std::nothrow) fock_ctx_ext();
    if (!ctx) return FOCK_ERR_OOM;
    std::unique_ptr<fock_ctx_ext, void (*)(fock_ctx_ext *)> guard(ctx, [](fock_ctx_ext *c) { fock_destroy(c); });
    auto fail = [&](int code) {
        g_create_error = ctx->err;
        return code;
    };
    ctx->opts = o;
    if (o.device >= 0) {
        if (cudaSetDevice(o.device) != cudaSuccess) {
            ctx->err = "cudaSetDevice failed";
            return fail(FOCK_ERR_CUDA);
        }
    }
    cudaGetDevice(&ctx->device);
    ctx->N = n;
    ctx->npair = (size_t)n * (n + 1) / 2;
    ctx->nelem = ctx->npair * (ctx->npair + 1) / 2;
    ctx->nprim = b->prim_offset[n];
    ctx->h_center.assign(b->center, b->center + 3 * n);
    ctx->h_angular.assign(b->angular, b->angular + 3 * n);
    ctx->h_primoff.assign(b->prim_offset, b->prim_offset + n + 1);
    ctx->h_exp.assign(b->exponent, b->exponent + ctx->nprim);
    ctx->h_coef.assign(b->coefficient, b->coefficient + ctx->nprim);

    auto body = [&]() -> int {
        FOCK_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        for (int i = 0; i < EV_COUNT; ++i) FOCK_CUDA_TRY(ctx, cudaEventCreate(&ctx->ev[i]));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_center, 3 * n * sizeof(double)));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_angular, 3 * n * sizeof(int)));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_primoff, (n + 1) * sizeof(int)));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_exp, ctx->nprim * sizeof(double)));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_coef, ctx->nprim * sizeof(double)));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_boys_lut, 200 * sizeof(double)));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_boys_tab, (size_t)BOYS_TAB_POINTS * BOYS_TAB_ORDERS * sizeof(double)));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_diag, ctx->npair * sizeof(double)));
        std::vector<double> tab;
        make_boys_table(tab);
        cudaStream_t st = ctx->stream;
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_center, ctx->h_center.data(), 3 * n * sizeof(double), cudaMemcpyHostToDevice, st));
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_angular, ctx->h_angular.data(), 3 * n * sizeof(int), cudaMemcpyHostToDevice, st));
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_primoff, ctx->h_primoff.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_exp, ctx->h_exp.data(), ctx->nprim * sizeof(double), cudaMemcpyHostToDevice, st));
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_coef, ctx->h_coef.data(), ctx->nprim * sizeof(double), cudaMemcpyHostToDevice, st));
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_boys_lut, kBoysLut, sizeof(kBoysLut), cudaMemcpyHostToDevice, st));
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_boys_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));

        // shells, pairs, class tables
        std::string err;
        if (!build_shells(n, b->center, b->angular, b->prim_offset, b->exponent, b->coefficient, ctx->shells, err)) {
            ctx->err = err;
            return FOCK_ERR_BAD_ARG;
        }
        if (o.kernel == 1) {
            for (Shell &s : ctx->shells.shells) s.class_ok = false;
            ctx->shells.all_class_ok = false;
            ctx->shells.n_class_shells = 0;
        } else if (o.kernel == 2) {  // d shells through the general warp-per-quartet list instead of the class engine
            for (Shell &s : ctx->shells.shells)
                if (s.L == 2) {
                    s.class_ok = false;
                    ctx->shells.all_class_ok = false;
                    --ctx->shells.n_class_shells;
                }
        }
        build_pairs(ctx->shells);
        ClassState *cs = &ctx->cs;
        cs->engine.sd = &ctx->shells;
        cs->engine.dev = &ctx->pairs;
        {
            std::vector<double> poly;
            const bool accurate = (o.mode & FOCK_MODE_BOYS_ACCURATE) != 0;
            if (accurate) make_boys_poly_accurate(poly); else make_boys_poly_faithful(kBoysLut, poly);
            FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&cs->d_boys_poly, poly.size() * sizeof(double)));
            FOCK_CUDA_TRY(ctx, cudaMemcpy(cs->d_boys_poly, poly.data(), poly.size() * sizeof(double), cudaMemcpyHostToDevice));
            cs->engine.boys_poly = cs->d_boys_poly;
            cs->engine.boys_ni = accurate ? kBoysAccIntervals : kBoysPolyIntervals;
            cudaDeviceProp prop;
            FOCK_CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
            cs->engine.num_sms = prop.multiProcessorCount;
        }
        cs->engine.stream = ctx->stream;
        cs->engine.mode = o.mode;
        if (!ctx->shells.pairs.empty() && cs->engine.upload_master(err)) {
            ctx->err = err;
            return FOCK_ERR_CUDA;
        }
        std::vector<unsigned char> ok(n), covered(n);
        if (o.kernel != 1 && cs->sq.init(ctx->shells, err)) {
            ctx->err = err;
            return FOCK_ERR_CUDA;
        }
        cs->engine.sq_shells = cs->sq.d_shells;
        cs->engine.sq_exps = cs->sq.d_exps;
        cs->engine.sq_coefs = cs->sq.d_coefs;
        cs->engine.sq_boys = ctx->boys();
        cs->engine.sq_tab = cs->sq.d_tab;
        cs->all_covered = true;
        cs->all_sp_class = true;
        for (int i = 0; i < n; ++i) {
            const Shell &sh = ctx->shells.shells[ctx->shells.bf_shell[i]];
            ok[i] = (sh.class_ok && sh.L <= 1) ? 1 : 0;
            if (!ok[i]) cs->all_sp_class = false;
            covered[i] = (o.kernel != 1 && sh.L <= 2) ? 1 : 0;
            if (!covered[i]) cs->all_covered = false;
        }
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&cs->d_bf_class_ok, n));
        FOCK_CUDA_TRY(ctx, cudaMemcpy(cs->d_bf_class_ok, ok.data(), n, cudaMemcpyHostToDevice));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&cs->d_bf_covered, n));
        FOCK_CUDA_TRY(ctx, cudaMemcpy(cs->d_bf_covered, covered.data(), n, cudaMemcpyHostToDevice));
        FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&cs->d_counter, sizeof(unsigned long long)));
        ctx->digest_begin = 0;
        ctx->digest_end = prow(ctx->npair);
        ctx->stats[0] = n;
        ctx->stats[1] = (int64_t)ctx->shells.shells.size();
        ctx->stats[2] = (int64_t)(ctx->shells.shells.size() * (ctx->shells.shells.size() + 1) / 2);
        {
            const int64_t sp = ctx->stats[2];
            ctx->stats[3] = sp * (sp + 1) / 2;
        }
        ctx->stats[12] = ctx->shells.n_partial_shells;
        return FOCK_OK;
    };
    const int status = body();
    if (status != FOCK_OK) return fail(status);
    *out = guard.release();
    return FOCK_OK;
}

void fock_destroy(fock_ctx *c) {
    if (!c) return;
    fock_ctx_ext *ctx = static_cast<fock_ctx_ext *>(c);
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    fock_comm_destroy(ctx);
    if (ctx->d_det) cudaFree(ctx->d_det);
    void *ptrs[] = {ctx->d_center, ctx->d_angular, ctx->d_primoff, ctx->d_exp,  ctx->d_coef, ctx->d_boys_lut,
                    ctx->d_boys_tab, ctx->d_diag,   ctx->d_V,       ctx->d_P,    ctx->d_dvec, ctx->d_yvec,
                    ctx->d_D,      ctx->d_G,       ctx->d_Ds,      ctx->d_A,    ctx->pairs.d_bra, ctx->pairs.d_ket, ctx->pairs.d_meta,
                    ctx->pairs.d_shell_bf, ctx->pairs.d_maxdiag, ctx->cs.d_bf_class_ok, ctx->cs.d_bf_covered, ctx->cs.d_counter, ctx->cs.d_boys_poly};
    ctx->cs.sq.destroy();
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->cs.h_diag) cudaFreeHost(ctx->cs.h_diag);
    if (ctx->cs.ev_diag) cudaEventDestroy(ctx->cs.ev_diag);
    if (ctx->cs.ev_k0) cudaEventDestroy(ctx->cs.ev_k0);
    if (ctx->cs.ev_k1) cudaEventDestroy(ctx->cs.ev_k1);
    for (int i = 0; i < EV_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int fock_set_stream(fock_ctx *ctx, void *s) {
    if (!ctx) return FOCK_ERR_BAD_ARG;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    cls(ctx)->engine.stream = ctx->stream;
    return FOCK_OK;
}

// Diagonal pass shared by both storage modes, in two halves so that independent device work (the
// tensor memsets) can be queued between them and overlap the host-side sorting:
//   diagonal_launch : kernels filling d_diag + asynchronous copy to the pinned host image
//   diagonal_finish : wait for that copy, order the class sublists by their largest diagonal
static int diagonal_launch(fock_ctx *ctx, int *launches) {
    ClassState *cs = cls(ctx);
    cudaStream_t st = ctx->stream;
    const int T = 128;
    std::string err;
    if (!cs->ev_diag) FOCK_CUDA_TRY(ctx, cudaEventCreateWithFlags(&cs->ev_diag, cudaEventDisableTiming));
    if (!cs->ev_k0) FOCK_CUDA_TRY(ctx, cudaEventCreate(&cs->ev_k0));
    if (!cs->ev_k1) FOCK_CUDA_TRY(ctx, cudaEventCreate(&cs->ev_k1));
    FOCK_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_diag, 0, ctx->npair * sizeof(double), st));
    if (!ctx->shells.pairs.empty()) {
        const int l = cs->engine.run_diagonal(ctx->d_diag, err);
        if (l < 0) {
            ctx->err = err;
            return FOCK_ERR_CUDA;
        }
        *launches += l;
    }
    if (!cs->all_sp_class) {  // function pairs outside the s/p class blocks (d shells, irregular shells, f/g functions)
        generic_diag_kernel<<<(unsigned)((ctx->npair + T - 1) / T), T, 0, st>>>(ctx->basis_dev(), ctx->boys(), ctx->opts.mode,
                                                                              cs->d_bf_class_ok, 1, ctx->npair, ctx->d_diag);
        *launches += 1;
    }
    if (!ctx->shells.pairs.empty()) {
        const int l = cs->engine.finish_diagonal(ctx->d_diag, err);
        if (l < 0) {
            ctx->err = err;
            return FOCK_ERR_CUDA;
        }
        *launches += l;
    }
    cs->sq.launch_pairmax(ctx->d_diag, st);
    FOCK_CUDA_TRY(ctx, cudaGetLastError());
    FOCK_CUDA_TRY(ctx, cudaEventRecord(cs->ev_diag, st));  // after the copies of the per-pair maxima
    return FOCK_OK;
}

static int diagonal_finish(fock_ctx *ctx, int *launches) {
    ClassState *cs = cls(ctx);
    FOCK_CUDA_TRY(ctx, cudaEventSynchronize(cs->ev_diag));
    if (!ctx->shells.pairs.empty()) {
        const auto t0 = std::chrono::steady_clock::now();
        std::string err;
        const int l = cs->engine.sort_by_diagonal(err);
        if (l < 0) {
            ctx->err = err;
            return FOCK_ERR_CUDA;
        }
        *launches += l;
        ctx->timings[5] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return FOCK_OK;
}

static int make_plan_timed(fock_ctx *ctx, double thr, int rank, int world) {
    ClassState *cs = cls(ctx);
    const auto t0 = std::chrono::steady_clock::now();
    std::string err;
    if (cs->engine.make_plan(ctx->opts.screen, thr, rank, world, err)) {
        ctx->err = err;
        return FOCK_ERR_CUDA;
    }
    ctx->timings[6] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return FOCK_OK;
}

int fock_eri_build(fock_ctx *ctx) {
    if (!ctx) return FOCK_ERR_BAD_ARG;
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // reference: boys() asserts n <= 10 (boys/src/lib.rs:59-62); total L of a quartet is the max order
    {
        int lmax = 0;
        for (int i = 0; i < ctx->N; ++i)
            lmax = std::max(lmax, ctx->h_angular[3 * i] + ctx->h_angular[3 * i + 1] + ctx->h_angular[3 * i + 2]);
        if (4 * lmax > 10) {
            ctx->err = "Boys-Function order is higher than max tabulated order 10";
            return FOCK_ERR_BOYS_ORDER;
        }
    }
    ClassState *cs = cls(ctx);
    cudaStream_t st = ctx->stream;
    int launches = 0;
    const int T = 128;
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_A], st));
    FOCK_TRY(diagonal_launch(ctx, &launches));
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_B], st));
    ctx->stats[4] = ctx->stats[5] = 0;
    ctx->stats[8] = 0;
    if (ctx->opts.storage == FOCK_STORAGE_INCORE) {
        // kernel == 1 (forced generic) forms the supermatrix by the gather definition from V; the
        // default path scatters every integral into P as it is produced (V only if keep_eri).
        const bool gather = ctx->opts.kernel == 1;
        const bool want_v = ctx->opts.keep_eri || gather;
        const size_t padded = prow(ctx->npair) + 8;  // row-padded supermatrix (ptri layout) + slack
        if (want_v && !ctx->d_V) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_V, ctx->nelem * sizeof(double)));
        if (!ctx->d_P) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_P, padded * sizeof(double)));
        if (want_v) FOCK_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_V, 0, ctx->nelem * sizeof(double), st));
        FOCK_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_P, 0, padded * sizeof(double), st));
        FOCK_CUDA_TRY(ctx, cudaMemsetAsync(cs->d_counter, 0, sizeof(unsigned long long), st));
        FOCK_TRY(diagonal_finish(ctx, &launches));  // host sorting overlaps the memsets above
        const double thr = screen_threshold(ctx->opts);
        double *Vp = want_v ? ctx->d_V : nullptr, *Pp = gather ? nullptr : ctx->d_P;
        if (getenv("FOCK_NULL_SINK")) Vp = Pp = nullptr;  // measurement aid: integrals are computed and dropped
        // In-core builds are REPLICATED across ranks (every rank forms the whole supermatrix, ~ms);
        // what is sharded is the per-iteration streaming of it (fock_set_digest_range).  Sharding the
        // build instead would need a reduce-scatter of the full packed tensor (1.9 GB at N=208, more
        // than the build itself costs).  The integral-direct path shards the quartet tiles.
        const bool shard_build = ctx->opts.incore_shard == 1 && ctx->opts.world > 1;
        const int b_rank = shard_build ? ctx->opts.rank : 0, b_world = shard_build ? ctx->opts.world : 1;
        diag_to_tensor_kernel<<<(unsigned)((ctx->npair + 255) / 256), 256, 0, st>>>(ctx->d_diag, ctx->npair, b_rank, b_world, Vp, Pp);
        launches += 1;
        if (!ctx->shells.pairs.empty()) {
            std::string err;
            FOCK_TRY(make_plan_timed(ctx, thr, b_rank, b_world));
            TensorSink sink{Vp, Pp, ctx->d_diag, thr, ctx->opts.screen};
            if (getenv("FOCK_SINK_MASK")) sink.mask = (size_t)strtoull(getenv("FOCK_SINK_MASK"), nullptr, 0);  // wrong results, timing only
            FOCK_CUDA_TRY(ctx, cudaEventRecord(cs->ev_k0, st));
            const int l = cs->engine.run_plan(sink, err);
            FOCK_CUDA_TRY(ctx, cudaEventRecord(cs->ev_k1, st));
            if (l < 0) {
                ctx->err = err;
                return FOCK_ERR_CUDA;
            }
            launches += l;
            cs->stats = cs->engine.plan_stats;
            ctx->stats[4] = cs->stats.shell_quartets;
            ctx->stats[5] = cs->stats.prim_quartets;
            ctx->stats[8] = (int64_t)cs->stats.model_flops;
            ctx->stats[10] = cs->stats.warp_tiles;
        }
        if (cs->sq.enabled) {  // shell quartets with d shells / irregular shells
            cs->sq.prepare(ctx->shells, ctx->opts.screen, thr, ctx->boys(), ctx->opts.mode, b_rank, b_world, st);
            TensorSink sink{Vp, Pp, ctx->d_diag, thr, ctx->opts.screen};
            launches += cs->sq.run(sink, cs->engine.num_sms, st);
            ctx->stats[4] += cs->sq.quartets;
            ctx->stats[11] = cs->sq.quartets;
        }
        if (!cs->all_covered) {  // functions above d: per-function generic kernel
            generic_tensor_kernel<<<(unsigned)((ctx->nelem + T - 1) / T), T, 0, st>>>(
                ctx->basis_dev(), ctx->boys(), ctx->opts.mode, cs->d_bf_covered, 1, ctx->opts.screen, thr, ctx->d_diag,
                ctx->nelem, b_rank, b_world, Vp, Pp, cs->d_counter);
            launches += 1;
        }
        FOCK_CUDA_TRY(ctx, cudaGetLastError());
        FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_C], st));
        if (gather) {
            supermatrix_kernel<<<(unsigned)((ctx->nelem + 255) / 256), 256, 0, st>>>(ctx->d_V, ctx->d_P, ctx->nelem);
            launches += 1;
        }
        FOCK_CUDA_TRY(ctx, cudaGetLastError());
        FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_D], st));
        unsigned long long generic_eval = 0;
        FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(&generic_eval, cs->d_counter, sizeof(generic_eval), cudaMemcpyDeviceToHost, st));
        FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        ctx->stats[9] = (int64_t)generic_eval;
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[EV_A], ctx->ev[EV_B]);
        ctx->timings[0] = ms;
        cudaEventElapsedTime(&ms, ctx->ev[EV_B], ctx->ev[EV_C]);
        ctx->timings[1] = ms;
        cudaEventElapsedTime(&ms, ctx->ev[EV_C], ctx->ev[EV_D]);
        ctx->timings[2] = ms;
    } else {
        FOCK_TRY(diagonal_finish(ctx, &launches));
        FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        if (ctx->opts.deterministic) {
            // bound of every |A| element per unit of max|D| from Schwarz, |(xy|zw)| <= sqrt(d_xy d_zw):
            //   Coulomb  <= 2 max sqrt(d) * sum_pairs sqrt(d),  exchange (4 forms) <= 2 (max_x sum_y sqrt(d_xy))^2
            std::vector<double> hd(ctx->npair);
            FOCK_CUDA_TRY(ctx, cudaMemcpy(hd.data(), ctx->d_diag, ctx->npair * sizeof(double), cudaMemcpyDeviceToHost));
            std::vector<double> rowsum(ctx->N, 0.0);
            double t1 = 0.0, dm = 0.0;
            size_t xy = 0;
            for (int x = 0; x < ctx->N; ++x)
                for (int y = 0; y <= x; ++y, ++xy) {
                    const double r = std::sqrt(std::fabs(hd[xy]));
                    t1 += r;
                    dm = std::max(dm, r);
                    rowsum[x] += r;
                    if (y != x) rowsum[y] += r;
                }
            double rmax = 0.0;
            for (double r : rowsum) rmax = std::max(rmax, r);
            ctx->det_bound_jk = 1.5 * (2.0 * dm * t1 + 2.0 * rmax * rmax);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[EV_A], ctx->ev[EV_B]);
        ctx->timings[0] = ms;
        ctx->timings[1] = ctx->timings[2] = 0;
    }
    if (ctx->opts.storage == FOCK_STORAGE_INCORE && ctx->opts.world > 1 && ctx->opts.incore_shard != 1) {
        // default sharding of the digestion: equal contiguous element ranges of the packed supermatrix
        const size_t total = prow(ctx->npair);
        const size_t chunk = ((total + ctx->opts.world - 1) / ctx->opts.world + 3) / 4 * 4;  // whole 32-byte sectors
        ctx->digest_begin = std::min(total, chunk * (size_t)ctx->opts.rank);
        ctx->digest_end = std::min(total, chunk * ((size_t)ctx->opts.rank + 1));
    }
    ctx->stats[6] = launches;
    ctx->built = true;
    return FOCK_OK;
}

int fock_eri_get(fock_ctx *ctx, int32_t i, int32_t j, int32_t k, int32_t l, double *value) {
    if (!ctx || !value) return FOCK_ERR_BAD_ARG;
    const int n = ctx->N;
    if (i < 0 || j < 0 || k < 0 || l < 0 || i >= n || j >= n || k >= n || l >= n) {
        ctx->err = "index out of range";
        return FOCK_ERR_BAD_ARG;
    }
    if (!ctx->built || !ctx->d_V) {
        ctx->err = "fock_eri_get needs fock_eri_build with in-core storage and keep_eri";
        return FOCK_ERR_STATE;
    }
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(value, ctx->d_V + packed_index(i, j, k, l), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FOCK_OK;
}

int fock_eri_export_packed(fock_ctx *ctx, double *host, size_t n) {
    if (!ctx || !host) return FOCK_ERR_BAD_ARG;
    if (n != ctx->nelem) {
        ctx->err = "export size must be fock_packed_len(N)";
        return FOCK_ERR_BAD_ARG;
    }
    if (!ctx->built || !ctx->d_V) {
        ctx->err = "fock_eri_export_packed needs fock_eri_build with in-core storage and keep_eri";
        return FOCK_ERR_STATE;
    }
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(host, ctx->d_V, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FOCK_OK;
}

int fock_diagonal_export(fock_ctx *ctx, double *host, size_t n) {
    if (!ctx || !host) return FOCK_ERR_BAD_ARG;
    if (n != ctx->npair) {
        ctx->err = "export size must be N(N+1)/2";
        return FOCK_ERR_BAD_ARG;
    }
    if (!ctx->built) {
        ctx->err = "fock_diagonal_export before fock_eri_build";
        return FOCK_ERR_STATE;
    }
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(host, ctx->d_diag, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return FOCK_OK;
}

int fock_eri_batch(fock_ctx *ctx, const int32_t *quartets, size_t n, double *values) {
    if (!ctx || (!quartets && n) || (!values && n)) return FOCK_ERR_BAD_ARG;
    if (n == 0) return FOCK_OK;
    for (size_t i = 0; i < 4 * n; ++i)
        if (quartets[i] < 0 || quartets[i] >= ctx->N) {
            ctx->err = "quartet index out of range";
            return FOCK_ERR_BAD_ARG;
        }
    for (size_t i = 0; i < n; ++i) {
        int L = 0;
        for (int k = 0; k < 4; ++k) {
            const int f = quartets[4 * i + k];
            L += ctx->h_angular[3 * f] + ctx->h_angular[3 * f + 1] + ctx->h_angular[3 * f + 2];
        }
        if (L > 10) {
            ctx->err = "Boys-Function order is higher than max tabulated order 10";
            return FOCK_ERR_BOYS_ORDER;
        }
    }
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int *d_q = nullptr;
    double *d_o = nullptr;
    FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&d_q, 4 * n * sizeof(int)));
    cudaError_t e = cudaMalloc((void **)&d_o, n * sizeof(double));
    if (e != cudaSuccess) {
        cudaFree(d_q);
        ctx->err = cudaGetErrorString(e);
        return FOCK_ERR_OOM;
    }
    cudaStream_t st = ctx->stream;
    cudaMemcpyAsync(d_q, quartets, 4 * n * sizeof(int), cudaMemcpyHostToDevice, st);
    generic_batch_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(ctx->basis_dev(), ctx->boys(), ctx->opts.mode, d_q, n, d_o);
    cudaMemcpyAsync(values, d_o, n * sizeof(double), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    cudaFree(d_q);
    cudaFree(d_o);
    if (e != cudaSuccess) {
        ctx->err = std::string("fock_eri_batch: ") + cudaGetErrorString(e);
        return FOCK_ERR_CUDA;
    }
    return FOCK_OK;
}

int fock_build_g_device(fock_ctx *ctx, const double *D_dev, double *G_dev) {
    if (!ctx || !D_dev || !G_dev) return FOCK_ERR_BAD_ARG;
    if (!ctx->built) {
        ctx->err = "fock_build_g before fock_eri_build";
        return FOCK_ERR_STATE;
    }
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FOCK_TRY(ensure_g_buffers(ctx));
    cudaStream_t st = ctx->stream;
    int launches = 0;
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_E], st));
    if (ctx->opts.storage == FOCK_STORAGE_INCORE) {
        if (!ctx->d_P) {
            ctx->err = "supermatrix missing";
            return FOCK_ERR_STATE;
        }
        pack_density_kernel<<<(unsigned)((ctx->npair + 255) / 256), 256, 0, st>>>(D_dev, ctx->N, ctx->d_dvec, ctx->d_yvec, ctx->npair);
        // rows that intersect [digest_begin, digest_end)
        size_t r_first = 0, r_last = ctx->npair - 1;
        if (ctx->digest_end > ctx->digest_begin) {
            // rows that intersect the padded element range: largest r with prow(r) <= e
            auto row_of = [&](size_t e) {
                size_t lo = 0, hi = ctx->npair - 1;
                while (lo < hi) {
                    const size_t mid = (lo + hi + 1) / 2;
                    if (prow(mid) <= e) lo = mid; else hi = mid - 1;
                }
                return lo;
            };
            r_first = row_of(ctx->digest_begin);
            r_last = row_of(ctx->digest_end - 1);
            r_first = r_first / kSymvRB * kSymvRB;
            const size_t nstrips = (r_last - r_first) / kSymvRB + 1;
            const size_t nchunks = r_last / kSymvCB + 1;
            dim3 grid((unsigned)nchunks, (unsigned)nstrips);
            if (nstrips > 65535) {
                // split the strip range over several launches (grid.y limit)
                for (size_t s0 = 0; s0 < nstrips; s0 += 65535) {
                    const size_t ns = std::min<size_t>(65535, nstrips - s0);
                    dim3 g((unsigned)nchunks, (unsigned)ns);
                    symv_packed_kernel<<<g, kSymvThreads, 0, st>>>(ctx->d_P, ctx->d_dvec, ctx->d_yvec, ctx->npair, ctx->digest_begin,
                                                                  ctx->digest_end, r_first + s0 * kSymvRB);
                    ++launches;
                }
            } else {
                symv_packed_kernel<<<grid, kSymvThreads, 0, st>>>(ctx->d_P, ctx->d_dvec, ctx->d_yvec, ctx->npair, ctx->digest_begin,
                                                                 ctx->digest_end, r_first);
                ++launches;
            }
        }
        unpack_result_kernel<<<(unsigned)(((size_t)ctx->N * ctx->N + 255) / 256), 256, 0, st>>>(ctx->d_yvec, ctx->N, G_dev);
        launches += 2;
    } else {
        // integral-direct: regenerate every surviving integral and digest it on the fly
        ClassState *cs = cls(ctx);
        const double thr = screen_threshold(ctx->opts);
        pack_density_kernel<<<(unsigned)((ctx->npair + 255) / 256), 256, 0, st>>>(D_dev, ctx->N, ctx->d_dvec, ctx->d_yvec, ctx->npair);
        // the class kernels digest shell blocks into the full matrix A (BlockDirectSink); the general paths
        // below still digest integral by integral into the packed y; combine_g_kernel adds the two
        const unsigned nn_blocks = (unsigned)(((size_t)ctx->N * ctx->N + 255) / 256);
        double fx_scale = 0.0;
        if (ctx->opts.deterministic) {
            // fixed-point scale: the largest power of two that keeps every partial sum of A inside 62 bits
            if (cs->sq.enabled || !cs->all_covered) {
                ctx->err = "deterministic mode needs a basis the class kernels cover entirely (s, p, d shells)";
                return FOCK_ERR_UNSUPPORTED;
            }
            if (!ctx->d_det) FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_det, 2 * sizeof(double)));
            FOCK_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_det, 0, 2 * sizeof(double), st));
            max_abs_kernel<<<64, 256, 0, st>>>(D_dev, (size_t)ctx->N * ctx->N, reinterpret_cast<unsigned long long *>(ctx->d_det));
            double dmax = 0.0;
            FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(&dmax, ctx->d_det, sizeof(double), cudaMemcpyDeviceToHost, st));
            FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            double bound = dmax * ctx->det_bound_jk;
            if (!(bound > 0.0) || !std::isfinite(bound)) bound = 1.0;
            int e = 0;
            std::frexp(bound, &e);  // bound < 2^e
            fx_scale = std::ldexp(1.0, 62 - e);
            ctx->det_scale = fx_scale;
            launches += 1;
        }
        symm_density_kernel<<<nn_blocks, 256, 0, st>>>(D_dev, ctx->N, ctx->d_Ds, ctx->d_A);
        diag_digest_block_kernel<<<(unsigned)((ctx->npair + 255) / 256), 256, 0, st>>>(ctx->d_diag, ctx->npair, ctx->opts.rank,
                                                                                     ctx->opts.world, ctx->d_Ds, ctx->N, ctx->d_A, fx_scale);
        launches += 3;
        if (!ctx->shells.pairs.empty()) {
            std::string err;
            FOCK_TRY(make_plan_timed(ctx, thr, ctx->opts.rank, ctx->opts.world));
            FOCK_CUDA_TRY(ctx, cudaEventRecord(cs->ev_k0, st));
            BlockDirectSink sink{ctx->d_A, ctx->d_Ds, ctx->d_diag, thr, ctx->opts.screen, ctx->N, fx_scale};
            const int l = cs->engine.run_plan(sink, err);
            FOCK_CUDA_TRY(ctx, cudaEventRecord(cs->ev_k1, st));
            if (l < 0) {
                ctx->err = err;
                return FOCK_ERR_CUDA;
            }
            launches += l;
            cs->stats = cs->engine.plan_stats;
            ctx->stats[4] = cs->stats.shell_quartets;
            ctx->stats[5] = cs->stats.prim_quartets;
            ctx->stats[8] = (int64_t)cs->stats.model_flops;
            ctx->stats[10] = cs->stats.warp_tiles;
        }
        if (cs->sq.enabled) {
            cs->sq.prepare(ctx->shells, ctx->opts.screen, thr, ctx->boys(), ctx->opts.mode, ctx->opts.rank, ctx->opts.world, st);
            DirectSink sink{ctx->d_yvec, ctx->d_dvec, ctx->d_diag, thr, ctx->opts.screen};
            launches += cs->sq.run(sink, cs->engine.num_sms, st);
            ctx->stats[4] += cs->sq.quartets;
            ctx->stats[11] = cs->sq.quartets;
        }
        if (!cs->all_covered) {
            const size_t mine = (ctx->nelem + ctx->opts.world - 1) / ctx->opts.world;
            generic_direct_kernel<<<(unsigned)((mine + 127) / 128), 128, 0, st>>>(
                ctx->basis_dev(), ctx->boys(), ctx->opts.mode, cs->d_bf_covered, 1, ctx->opts.screen, thr, ctx->d_diag, ctx->nelem,
                ctx->opts.rank, ctx->opts.world, ctx->d_dvec, ctx->d_yvec);
            launches += 1;
        }
        // deterministic mode sums the fixed-point partials of all ranks as integers (exact), then converts once
        if (fx_scale != 0.0 && ctx->comm) {
            FOCK_TRY(fock_comm_allreduce(ctx, ctx->d_A, (size_t)ctx->N * ctx->N, true, st));
            launches += 1;
        }
        combine_g_kernel<<<nn_blocks, 256, 0, st>>>(ctx->d_yvec, ctx->d_A, ctx->N, G_dev, fx_scale != 0.0 ? 1.0 / fx_scale : 0.0);
        launches += 1;
        if (fx_scale != 0.0 && ctx->comm) {
            FOCK_CUDA_TRY(ctx, cudaGetLastError());
            FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_F], st));
            ctx->stats[7] = launches;
            return FOCK_OK;
        }
    }
    FOCK_CUDA_TRY(ctx, cudaGetLastError());
    // the one exchange step of the path: partial G of every rank summed over NVLink (no-op without a communicator)
    if (ctx->comm) {
        FOCK_TRY(fock_comm_allreduce(ctx, G_dev, (size_t)ctx->N * ctx->N, false, st));
        launches += 1;
    }
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_F], st));
    ctx->stats[7] = launches;
    return FOCK_OK;
}

int fock_build_g(fock_ctx *ctx, const double *D, double *G) {
    if (!ctx || !D || !G) return FOCK_ERR_BAD_ARG;
    if (!ctx->built) {
        ctx->err = "fock_build_g before fock_eri_build";
        return FOCK_ERR_STATE;
    }
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    FOCK_TRY(ensure_g_buffers(ctx));
    const size_t nn = (size_t)ctx->N * ctx->N;
    cudaStream_t st = ctx->stream;
    // page-locked caller buffers (fock_pinned_alloc, cudaHostRegister, ...) are copied directly; pageable ones are staged
    auto page_locked = [](const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    };
    const bool d_direct = page_locked(D), g_direct = page_locked(G);
    if (!d_direct) std::memcpy(ctx->h_pinned, D, nn * sizeof(double));
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_C], st));
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_D, d_direct ? D : ctx->h_pinned, nn * sizeof(double), cudaMemcpyHostToDevice, st));
    FOCK_TRY(fock_build_g_device(ctx, ctx->d_D, ctx->d_G));
    FOCK_CUDA_TRY(ctx, cudaMemcpyAsync(g_direct ? G : ctx->h_pinned + nn, ctx->d_G, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
    FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_D], st));
    FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (!g_direct) std::memcpy(G, ctx->h_pinned + nn, nn * sizeof(double));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[EV_E], ctx->ev[EV_F]);
    ctx->timings[3] = ms;
    cudaEventElapsedTime(&ms, ctx->ev[EV_C], ctx->ev[EV_D]);
    ctx->timings[4] = ms;
    return FOCK_OK;
}

int fock_pinned_alloc(size_t bytes, void **out) {
    if (!out) return FOCK_ERR_BAD_ARG;
    *out = nullptr;
    const cudaError_t e = cudaMallocHost(out, bytes ? bytes : 8);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? FOCK_ERR_OOM : FOCK_ERR_CUDA;
    }
    return FOCK_OK;
}

int fock_pinned_free(void *p) {
    if (!p) return FOCK_OK;
    return cudaFreeHost(p) == cudaSuccess ? FOCK_OK : FOCK_ERR_CUDA;
}

int fock_one_electron(fock_ctx *ctx, int32_t n_nuc, const double *pos, const double *chg, double *S, double *T, double *V) {
    if (!ctx || n_nuc < 0 || (n_nuc && (!pos || !chg)) || !S || !T || !V) return FOCK_ERR_BAD_ARG;
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int n = ctx->N;
    const size_t nn = (size_t)n * n;
    {
        int lmax = 0;
        for (int i = 0; i < n; ++i) lmax = std::max(lmax, ctx->h_angular[3 * i] + ctx->h_angular[3 * i + 1] + ctx->h_angular[3 * i + 2]);
        if (2 * lmax > 10) {
            ctx->err = "Boys-Function order is higher than max tabulated order 10";
            return FOCK_ERR_BOYS_ORDER;
        }
    }
    double *d_buf = nullptr;
    FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&d_buf, (3 * nn + 4 * (size_t)std::max(n_nuc, 1)) * sizeof(double)));
    double *dS = d_buf, *dT = d_buf + nn, *dV = d_buf + 2 * nn, *dpos = d_buf + 3 * nn, *dchg = dpos + 3 * (size_t)std::max(n_nuc, 1);
    cudaStream_t st = ctx->stream;
    if (n_nuc) {
        cudaMemcpyAsync(dpos, pos, 3 * n_nuc * sizeof(double), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(dchg, chg, n_nuc * sizeof(double), cudaMemcpyHostToDevice, st);
    }
    const size_t np = ctx->npair;
    one_electron_kernel<<<(unsigned)((np + 63) / 64), 64, 0, st>>>(ctx->basis_dev(), ctx->boys(), ctx->opts.mode, n_nuc, dpos, dchg, dS, dT, dV);
    cudaMemcpyAsync(S, dS, nn * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(T, dT, nn * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(V, dV, nn * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d_buf);
    if (e != cudaSuccess) {
        ctx->err = std::string("fock_one_electron: ") + cudaGetErrorString(e);
        return FOCK_ERR_CUDA;
    }
    return FOCK_OK;
}

int fock_supermatrix_device(fock_ctx *ctx, void **dev_ptr, size_t *n_elements) {
    if (!ctx || !dev_ptr || !n_elements) return FOCK_ERR_BAD_ARG;
    if (!ctx->d_P) {
        ctx->err = "no supermatrix (in-core fock_eri_build first)";
        return FOCK_ERR_STATE;
    }
    *dev_ptr = ctx->d_P;
    *n_elements = prow(ctx->npair);  // row-padded layout, see device_math.cuh: prow()
    return FOCK_OK;
}

int fock_set_digest_range(fock_ctx *ctx, size_t begin, size_t end) {
    if (!ctx || begin > end || end > prow(ctx->npair)) return FOCK_ERR_BAD_ARG;
    ctx->digest_begin = begin;
    ctx->digest_end = end;
    return FOCK_OK;
}

int fock_boys_batch(fock_ctx *ctx, const double *t, size_t n, int32_t order, double *values) {
    if (!ctx || (!t && n) || (!values && n)) return FOCK_ERR_BAD_ARG;
    if (order < 0 || order > 10) {
        ctx->err = "Boys-Function order is higher than max tabulated order 10";
        return FOCK_ERR_BOYS_ORDER;
    }
    if (n == 0) return FOCK_OK;
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double *d = nullptr;
    FOCK_CUDA_TRY(ctx, cudaMalloc((void **)&d, 2 * n * sizeof(double)));
    cudaStream_t st = ctx->stream;
    cudaMemcpyAsync(d, t, n * sizeof(double), cudaMemcpyHostToDevice, st);
    ClassState *cs = cls(ctx);
    if (order < kBoysPolyOrders)
        boys_poly_batch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(cs->d_boys_poly, cs->engine.boys_ni, d, n, order, d + n);
    else
        boys_generic_batch_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ctx->boys(), ctx->opts.mode, d, n, order, d + n);
    cudaMemcpyAsync(values, d + n, n * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) {
        ctx->err = std::string("fock_boys_batch: ") + cudaGetErrorString(e);
        return FOCK_ERR_CUDA;
    }
    return FOCK_OK;
}

int fock_measure_fp64_peak(fock_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return FOCK_ERR_BAD_ARG;
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    FOCK_CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    cudaStream_t st = ctx->stream;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_A], st));
        dfma_peak_kernel<<<blocks, 256, 0, st>>>(ctx->d_diag, iters, 1.0 + rep);
        FOCK_CUDA_TRY(ctx, cudaEventRecord(ctx->ev[EV_B], st));
        FOCK_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[EV_A], ctx->ev[EV_B]);
        const double flops = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    *tflops = best;
    return FOCK_OK;
}

int fock_timings(fock_ctx *ctx, double *ms, int32_t n) {
    if (!ctx || !ms || n < 0) return FOCK_ERR_BAD_ARG;
    // the device-pointer G build leaves its events unread until asked
    if (ctx->built && ctx->ev[EV_F] && cudaEventQuery(ctx->ev[EV_F]) == cudaSuccess) {
        float t = 0;
        if (cudaEventElapsedTime(&t, ctx->ev[EV_E], ctx->ev[EV_F]) == cudaSuccess) ctx->timings[3] = t;
    }
    {
        ClassState *cs = cls(ctx);
        float t = 0;
        if (cs->ev_k1 && cudaEventQuery(cs->ev_k1) == cudaSuccess && cudaEventElapsedTime(&t, cs->ev_k0, cs->ev_k1) == cudaSuccess)
            ctx->timings[7] = t;
    }
    cudaGetLastError();
    for (int i = 0; i < n && i < 8; ++i) ms[i] = ctx->timings[i];
    return FOCK_OK;
}

int fock_stats(fock_ctx *ctx, int64_t *out, int32_t n) {
    if (!ctx || !out || n < 0) return FOCK_ERR_BAD_ARG;
    for (int i = 0; i < n && i < 16; ++i) out[i] = ctx->stats[i];
    return FOCK_OK;
}

}  // extern "C"
