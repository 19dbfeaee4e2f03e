// sq_launch.cuh -- host side of the general shell-quartet kernel (eri_shellq.cuh): shell tables,
// the list G of shell pairs that involve a non-class shell (d shells, irregular s/p shells), the
// combined list C of class pairs, screening prefixes and the launch.  Included by fock_api.cu only.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "eri_shellq.cuh"
#include "shells.h"

namespace fock {

// max |(xy|xy)| over the components of each G pair, from the basis-function diagonal
__global__ void sq_pairmax_kernel(const SqShell *__restrict__ shells, const int *__restrict__ meta, int n,
                                  const double *__restrict__ diag, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SqShell A = shells[meta[2 * i]], B = shells[meta[2 * i + 1]];
    double m = 0.0;
    for (int ca = 0; ca < ncart_c(A.L); ++ca)
        for (int cb = 0; cb < ncart_c(B.L); ++cb) {
            const int x = A.bf[ca], y = B.bf[cb];
            if (x < 0 || y < 0) continue;
            const double d = fabs(diag[x >= y ? tri(x, y) : tri(y, x)]);
            if (!(d <= m)) m = (d == d) ? d : __longlong_as_double(0x7ff0000000000000LL);
        }
    out[i] = m;
}

struct SqState {
    bool enabled = false;
    int nG = 0, nC = 0;
    std::vector<int> g_nat;  // [nG][2] shell ids, natural order
    std::vector<int> c_pair_ids;  // indices into ShellData::pairs
    SqShell *d_shells = nullptr;
    SqTables *d_tab = nullptr;
    double *d_exps = nullptr, *d_coefs = nullptr;
    int *d_gmeta_nat = nullptr;
    double *d_gmax = nullptr, *h_gmax = nullptr;
    int *d_meta = nullptr, *h_meta = nullptr;  // sorted: G pairs then C pairs, [.][2]
    long long *d_prefix = nullptr, *h_prefix = nullptr;
    long long quartets = 0, quartets_total = 0;
    SqLaunch launch{};

    void destroy() {
        void *dp[] = {d_shells, d_tab, d_exps, d_coefs, d_gmeta_nat, d_gmax, d_meta, d_prefix};
        for (void *p : dp)
            if (p) cudaFree(p);
        void *hp[] = {h_gmax, h_meta, h_prefix};
        for (void *p : hp)
            if (p) cudaFreeHost(p);
    }

    // Called once from fock_create.  Shells with L <= 2 take part; anything higher stays with the
    // per-function generic kernel.
    int init(const ShellData &sd, std::string &err) {
        const int ns = (int)sd.shells.size();
        std::vector<int> eligible(ns);
        bool any_nonclass = false, any_d = false;
        for (int s = 0; s < ns; ++s) {
            eligible[s] = sd.shells[s].L <= 2;
            if (eligible[s] && !sd.shells[s].class_ok) any_nonclass = true;
            if (sd.shells[s].L == 2) any_d = true;
        }
        // the shell tables also serve the class engine's warp-per-quartet fallback (d classes with L >= 6)
        if (!any_nonclass && !any_d) return 0;
        for (int a = 0; a < ns; ++a)
            for (int b = 0; b <= a; ++b)
                if (eligible[a] && eligible[b] && !(sd.shells[a].class_ok && sd.shells[b].class_ok)) {
                    g_nat.push_back(a);
                    g_nat.push_back(b);
                }
        nG = (int)g_nat.size() / 2;
        nC = (int)sd.pairs.size();
        c_pair_ids.resize(nC);
        for (int i = 0; i < nC; ++i) c_pair_ids[i] = i;
        std::vector<SqShell> hs(ns);
        for (int s = 0; s < ns; ++s) {
            const Shell &sh = sd.shells[s];
            for (int k = 0; k < 3; ++k) hs[s].c[k] = sh.c[k];
            hs[s].L = sh.L;
            hs[s].nprim = sh.nprim;
            hs[s].prim_begin = sh.prim_begin;
            for (int k = 0; k < 6; ++k) {
                hs[s].bf[k] = sh.bf[k];
                hs[s].scale[k] = sh.scale[k];
            }
        }
        SqTables tab{};
        int e = 0;
        for (int deg = 0; deg <= 8; ++deg)
            for (int t = 0; t <= deg; ++t)
                for (int u = 0; u <= deg - t; ++u) {
                    const int v = deg - t - u;
                    tab.t[e] = (unsigned char)t;
                    tab.u[e] = (unsigned char)u;
                    tab.v[e] = (unsigned char)v;
                    tab.idx[t][u][v] = (unsigned char)e;
                    ++e;
                }
        cudaError_t ce;
        const size_t npairs = (size_t)nG + nC;
        if ((ce = cudaMalloc((void **)&d_tab, sizeof(tab))) != cudaSuccess ||
            (ce = cudaMemcpy(d_tab, &tab, sizeof(tab), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (ce = cudaMalloc((void **)&d_shells, hs.size() * sizeof(SqShell))) != cudaSuccess ||
            (ce = cudaMalloc((void **)&d_exps, sd.exps.size() * sizeof(double) + 8)) != cudaSuccess ||
            (ce = cudaMalloc((void **)&d_coefs, sd.coefs.size() * sizeof(double) + 8)) != cudaSuccess ||
            (ce = cudaMalloc((void **)&d_gmeta_nat, g_nat.size() * sizeof(int) + 8)) != cudaSuccess ||
            (ce = cudaMalloc((void **)&d_gmax, nG * sizeof(double) + 8)) != cudaSuccess ||
            (ce = cudaMallocHost((void **)&h_gmax, nG * sizeof(double) + 8)) != cudaSuccess ||
            (ce = cudaMalloc((void **)&d_meta, npairs * 2 * sizeof(int) + 8)) != cudaSuccess ||
            (ce = cudaMallocHost((void **)&h_meta, npairs * 2 * sizeof(int) + 8)) != cudaSuccess ||
            (ce = cudaMalloc((void **)&d_prefix, (2 * (size_t)nG + 2) * sizeof(long long))) != cudaSuccess ||
            (ce = cudaMallocHost((void **)&h_prefix, (2 * (size_t)nG + 2) * sizeof(long long))) != cudaSuccess) {
            err = std::string("shell-quartet tables: ") + cudaGetErrorString(ce);
            return 1;
        }
        cudaMemcpy(d_shells, hs.data(), hs.size() * sizeof(SqShell), cudaMemcpyHostToDevice);
        cudaMemcpy(d_exps, sd.exps.data(), sd.exps.size() * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(d_coefs, sd.coefs.data(), sd.coefs.size() * sizeof(double), cudaMemcpyHostToDevice);
        cudaMemcpy(d_gmeta_nat, g_nat.data(), g_nat.size() * sizeof(int), cudaMemcpyHostToDevice);
        enabled = nG > 0;
        return 0;
    }

    // queue the per-pair maxima of the G pairs (needs the basis-function diagonal on the device)
    void launch_pairmax(const double *d_diag, cudaStream_t st) {
        if (!enabled) return;
        sq_pairmax_kernel<<<(nG + 127) / 128, 128, 0, st>>>(d_shells, d_gmeta_nat, nG, d_diag, d_gmax);
        cudaMemcpyAsync(h_gmax, d_gmax, nG * sizeof(double), cudaMemcpyDeviceToHost, st);
    }

    // after the maxima are on the host (and the class pairs carry theirs): sort, screen, upload
    int prepare(const ShellData &sd, int screen, double thr, BoysTables boys, int mode, int rank, int world, cudaStream_t st) {
        if (!enabled) return 0;
        std::vector<std::pair<double, int>> gk(nG), ck(nC);
        for (int i = 0; i < nG; ++i) gk[i] = std::make_pair(-h_gmax[i], i);
        for (int i = 0; i < nC; ++i) ck[i] = std::make_pair(-sd.pairs[i].maxdiag, i);
        std::sort(gk.begin(), gk.end());
        std::sort(ck.begin(), ck.end());
        for (int i = 0; i < nG; ++i) {
            h_meta[2 * i] = g_nat[2 * gk[i].second];
            h_meta[2 * i + 1] = g_nat[2 * gk[i].second + 1];
        }
        for (int i = 0; i < nC; ++i) {
            h_meta[2 * (nG + i)] = sd.pairs[ck[i].second].A;
            h_meta[2 * (nG + i) + 1] = sd.pairs[ck[i].second].B;
        }
        const double t2 = thr * thr * (1.0 - 1e-12);
        auto passes = [&](double a, double b) { return screen == 1 || a * b > t2; };
        // job 0: G x C (rectangle), job 1: G x G (triangle c <= r)
        long long *p0 = h_prefix, *p1 = h_prefix + nG + 1;
        long long q0 = 0, q1 = 0;
        int nc = nC, ng = nG;
        for (int r = 0; r < nG; ++r) {
            const double dr = -gk[r].first;
            while (nc > 0 && !passes(dr, -ck[nc - 1].first)) --nc;
            while (ng > 0 && !passes(dr, -gk[ng - 1].first)) --ng;
            p0[r] = q0;
            p1[r] = q1;
            q0 += nc;
            q1 += std::min(ng, r + 1);
        }
        p0[nG] = q0;
        p1[nG] = q1;
        quartets = q0 + q1;
        quartets_total = (long long)nG * nC + (long long)nG * (nG + 1) / 2;
        cudaMemcpyAsync(d_meta, h_meta, ((size_t)nG + nC) * 2 * sizeof(int), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_prefix, h_prefix, (2 * (size_t)nG + 2) * sizeof(long long), cudaMemcpyHostToDevice, st);
        launch = SqLaunch{};
        int nj = 0;
        if (q0 > 0) launch.jobs[nj++] = SqJob{0, d_meta, d_meta + 2 * nG, d_prefix, nG, 0};
        if (q1 > 0) {
            launch.jobs[nj] = SqJob{nj ? q0 : 0, d_meta, d_meta, d_prefix + nG + 1, nG, 1};
            ++nj;
        }
        launch.njobs = nj;
        launch.total_quartets = quartets;
        launch.shells = d_shells;
        launch.exps = d_exps;
        launch.coefs = d_coefs;
        launch.boys = boys;
        launch.tab = d_tab;
        launch.mode = mode;
        launch.rank = rank;
        launch.world = world;
        return 0;
    }

    template <class Sink>
    int run(const Sink &sink, int num_sms, cudaStream_t st) {
        if (!enabled || launch.njobs == 0) return 0;
        const long long mine = (launch.total_quartets - launch.rank + launch.world - 1) / launch.world;
        if (mine <= 0) return 0;
        const unsigned grid = (unsigned)std::min<long long>((mine + kSqWarps - 1) / kSqWarps, (long long)num_sms * 16);
        const size_t smem = (size_t)kSqWarps * kSqSmemPerWarp * sizeof(double);
        eri_shellq_kernel<Sink><<<grid, kSqThreads, smem, st>>>(launch, sink);
        return 1;
    }
};

}  // namespace fock
