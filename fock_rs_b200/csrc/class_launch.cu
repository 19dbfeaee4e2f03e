// class_launch.cu -- host side of the shell-class ERI kernels: diagonal pass, sublist sorting,
// tile tables for the (bra sublist, ket sublist) launches, kernel dispatch.
#include <algorithm>
#include <cmath>

#include "../../include/fock_b200.h"
#include "class_launch.h"
#include "eri_class.cuh"

namespace fock {

namespace {

template <int LA, int LB, int LC, int LD, int KC0, int KCN, bool DIAG, class Sink>
void launch_mode(const ClassLaunch &cl, const Sink &sink, int mode, dim3 grid, size_t smem, cudaStream_t st) {
    const bool acc = mode & FOCK_MODE_BOYS_ACCURATE_BIT, fix = mode & FOCK_MODE_FIX_R_BIT;
    if (!acc && !fix)
        eri_class_kernel<LA, LB, LC, LD, false, false, KC0, KCN, DIAG, Sink><<<grid, kClassThreads, smem, st>>>(cl, sink);
    else if (!acc && fix)
        eri_class_kernel<LA, LB, LC, LD, false, true, KC0, KCN, DIAG, Sink><<<grid, kClassThreads, smem, st>>>(cl, sink);
    else if (acc && !fix)
        eri_class_kernel<LA, LB, LC, LD, true, false, KC0, KCN, DIAG, Sink><<<grid, kClassThreads, smem, st>>>(cl, sink);
    else
        eri_class_kernel<LA, LB, LC, LD, true, true, KC0, KCN, DIAG, Sink><<<grid, kClassThreads, smem, st>>>(cl, sink);
}

// Returns the number of kernel launches issued.
template <bool DIAG, class Sink>
int launch_class(int cb, int ck, const ClassLaunch &cl, const Sink &sink, int mode, dim3 grid, cudaStream_t st) {
    const size_t smem = DIAG ? 0 : (size_t)cl.Kab * pair_fields(cb) * sizeof(double);
    const int key = cb * 3 + ck;
    switch (key) {
        case 0: launch_mode<0, 0, 0, 0, 0, 1, DIAG>(cl, sink, mode, grid, smem, st); return 1;
        case 3: if constexpr (!DIAG) launch_mode<1, 0, 0, 0, 0, 1, DIAG>(cl, sink, mode, grid, smem, st); return 1;
        case 4: launch_mode<1, 0, 1, 0, 0, 3, DIAG>(cl, sink, mode, grid, smem, st); return 1;
        case 6: if constexpr (!DIAG) launch_mode<1, 1, 0, 0, 0, 1, DIAG>(cl, sink, mode, grid, smem, st); return 1;
        case 7: if constexpr (!DIAG) launch_mode<1, 1, 1, 0, 0, 3, DIAG>(cl, sink, mode, grid, smem, st); return 1;
        case 8:  // (pp|pp): 81 accumulators do not fit one thread's registers -> three ket-component slices
            launch_mode<1, 1, 1, 1, 0, 3, DIAG>(cl, sink, mode, grid, smem, st);
            launch_mode<1, 1, 1, 1, 3, 3, DIAG>(cl, sink, mode, grid, smem, st);
            launch_mode<1, 1, 1, 1, 6, 3, DIAG>(cl, sink, mode, grid, smem, st);
            return 3;
        default: return 0;
    }
}

const int kClsLA[3] = {0, 1, 1}, kClsLB[3] = {0, 0, 1};

}  // namespace

int ClassEngine::upload_tables(std::string &err) {
    std::vector<double> bra, ket;
    std::vector<int> meta;
    build_pair_tables(*sd, bra, ket, meta);
    auto re = [&](void **p, size_t bytes) {
        if (*p) cudaFree(*p);
        *p = nullptr;
        return cudaMalloc(p, bytes ? bytes : 8);
    };
    cudaError_t e;
    if ((e = re((void **)&dev->d_bra, bra.size() * 8)) != cudaSuccess ||
        (e = re((void **)&dev->d_ket, ket.size() * 8)) != cudaSuccess ||
        (e = re((void **)&dev->d_meta, meta.size() * 4)) != cudaSuccess) {
        err = std::string("pair tables: ") + cudaGetErrorString(e);
        return 1;
    }
    cudaMemcpyAsync(dev->d_bra, bra.data(), bra.size() * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(dev->d_ket, ket.data(), ket.size() * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(dev->d_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, stream);
    if (!dev->d_shell_bf) {
        std::vector<int> sbf(sd->shells.size() * 6);
        for (size_t s = 0; s < sd->shells.size(); ++s)
            for (int c = 0; c < 6; ++c) sbf[s * 6 + c] = sd->shells[s].bf[c];
        if ((e = cudaMalloc((void **)&dev->d_shell_bf, sbf.size() * 4 + 8)) != cudaSuccess) {
            err = std::string("shell table: ") + cudaGetErrorString(e);
            return 1;
        }
        cudaMemcpyAsync(dev->d_shell_bf, sbf.data(), sbf.size() * 4, cudaMemcpyHostToDevice, stream);
    }
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        err = std::string("pair table upload: ") + cudaGetErrorString(e);
        return 1;
    }
    dev->bra_doubles = bra.size();
    dev->ket_doubles = ket.size();
    dev->meta_ints = meta.size();
    return 0;
}

ClassLaunch ClassEngine::base_launch(const Sublist &sb, const Sublist &sk) const {
    ClassLaunch cl{};
    cl.bra = dev->d_bra + sb.bra_off;
    cl.ket = dev->d_ket + sk.ket_off;
    cl.bra_meta = dev->d_meta + sb.meta_off;
    cl.ket_meta = dev->d_meta + sk.meta_off;
    cl.shell_bf = dev->d_shell_bf;
    cl.boys_lut = boys.lut;
    cl.boys_tab = boys.tab;
    cl.Kab = sb.K;
    cl.Kcd = sk.K;
    cl.nb = (int)sb.pair_ids.size();
    cl.nk = (int)sk.pair_ids.size();
    cl.nk_pad = sk.npad;
    cl.rank = 0;
    cl.world = 1;
    return cl;
}

// Diagonal pass: every class pair's (AB|AB) block; only (xy|xy) kept.
int ClassEngine::run_diagonal(double *d_diag, std::string &err) {
    int launches = 0;
    for (const Sublist &s : sd->sublists) {
        if (s.pair_ids.empty()) continue;
        ClassLaunch cl = base_launch(s, s);
        dim3 grid((unsigned)((cl.nb + kClassThreads - 1) / kClassThreads));
        DiagSink sink{d_diag};
        launches += launch_class<true>(s.cls, s.cls, cl, sink, mode, grid, stream);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("diagonal kernels: ") + cudaGetErrorString(e);
        return -1;
    }
    return launches;
}

// After the diagonal is on the host: per-pair max |(xy|xy)|, sort sublists, rebuild tables.
int ClassEngine::sort_by_diagonal(const std::vector<double> &h_diag, std::string &err) {
    for (ShellPair &p : sd->pairs) {
        const Shell &A = sd->shells[p.A], &B = sd->shells[p.B];
        double m = 0.0;
        for (int ca = 0; ca < ncart(A.L); ++ca)
            for (int cb = 0; cb < ncart(B.L); ++cb) {
                const int x = A.bf[ca], y = B.bf[cb];
                if (x < 0 || y < 0) continue;
                const double d = std::fabs(h_diag[x >= y ? tri(x, y) : tri(y, x)]);
                if (d > m || std::isnan(d)) m = std::isnan(d) ? INFINITY : d;
            }
        p.maxdiag = m;
    }
    sort_sublists(*sd);
    return upload_tables(err);
}

// Number of ket pairs c (sublist sk, sorted descending) that pass the shell-level bound against
// bra pair value `db`: the basis-function rule sqrt(d1*d2) > thr is monotone in both arguments, so
// if the bound on the per-pair maxima fails every component fails.
static int surviving_prefix(double db, const std::vector<double> &dk, int screen, double thr) {
    if (screen == FOCK_SCREEN_OFF) return (int)dk.size();
    int lo = 0, hi = (int)dk.size();  // first index that fails
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (std::sqrt(db * dk[mid]) > thr) lo = mid + 1; else hi = mid;
    }
    return lo;
}

int ClassEngine::run_build(const StoreSink &sink, int screen, double thr, int rank, int world, ClassStats &stats,
                           std::string &err) {
    const int ns = (int)sd->sublists.size();
    std::vector<std::vector<double>> md(ns);
    for (int s = 0; s < ns; ++s) {
        md[s].reserve(sd->sublists[s].pair_ids.size());
        for (int id : sd->sublists[s].pair_ids) md[s].push_back(sd->pairs[id].maxdiag);
    }
    // tile tables for all launches, one upload
    struct Job {
        int sb, sk;
        size_t off;
        long long tiles;
    };
    std::vector<Job> jobs;
    std::vector<int> table;
    stats = ClassStats();
    for (int b = 0; b < ns; ++b) {
        const Sublist &sb = sd->sublists[b];
        const int nb = (int)sb.pair_ids.size();
        if (!nb) continue;
        for (int k = 0; k <= b; ++k) {
            const Sublist &sk = sd->sublists[k];
            const int nk = (int)sk.pair_ids.size();
            if (!nk) continue;
            Job j{b, k, table.size(), 0};
            table.resize(table.size() + 2 * (size_t)nb + 1);
            int *prefix = table.data() + j.off, *ncols = prefix + nb + 1;
            long long tiles = 0, quartets = 0;
            for (int r = 0; r < nb; ++r) {
                int nc = surviving_prefix(md[b][r], md[k], screen, thr);
                if (b == k) nc = std::min(nc, r + 1);
                ncols[r] = nc;
                prefix[r] = (int)tiles;
                tiles += (nc + kClassThreads - 1) / kClassThreads;
                quartets += nc;
            }
            prefix[nb] = (int)tiles;
            if (tiles > 2000000000LL) {
                err = "tile table overflow";
                return -1;
            }
            j.tiles = tiles;
            stats.shell_quartets_total += (b == k) ? (long long)nb * (nb + 1) / 2 : (long long)nb * nk;
            // work this rank will do: tiles are dealt round-robin, quartet counts are split evenly enough
            stats.shell_quartets += quartets;
            stats.prim_quartets += quartets * (long long)sb.K * sk.K;
            stats.model_flops += (double)quartets * sb.K * sk.K *
                                 model_flops_per_prim_quartet(kClsLA[sb.cls], kClsLB[sb.cls], kClsLA[sk.cls], kClsLB[sk.cls]);
            if (tiles > 0) jobs.push_back(j);
        }
    }
    if (table.size() > table_cap) {
        if (d_table) cudaFree(d_table);
        d_table = nullptr;
        table_cap = table.size() + table.size() / 4 + 1024;
        cudaError_t e = cudaMalloc((void **)&d_table, table_cap * sizeof(int));
        if (e != cudaSuccess) {
            err = std::string("tile table: ") + cudaGetErrorString(e);
            table_cap = 0;
            return -1;
        }
    }
    cudaMemcpyAsync(d_table, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice, stream);
    int launches = 0;
    for (const Job &j : jobs) {
        const Sublist &sb = sd->sublists[j.sb], &sk = sd->sublists[j.sk];
        ClassLaunch cl = base_launch(sb, sk);
        cl.tile_prefix = d_table + j.off;
        cl.ncols = cl.tile_prefix + cl.nb + 1;
        cl.rank = rank;
        cl.world = world;
        cl.total_tiles = j.tiles;
        const long long mine = (j.tiles - rank + world - 1) / world;
        if (mine <= 0) continue;
        dim3 grid((unsigned)mine);
        launches += launch_class<false>(sb.cls, sk.cls, cl, sink, mode, grid, stream);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("class kernels: ") + cudaGetErrorString(e);
        return -1;
    }
    // the host tile table must stay alive until the copy has been consumed
    cudaStreamSynchronize(stream);
    return launches;
}

ClassEngine::~ClassEngine() {
    if (d_table) cudaFree(d_table);
}

}  // namespace fock
