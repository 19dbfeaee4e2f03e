// eri_shellq.cuh -- general shell-quartet kernel for shells up to d (L <= 2 per shell, total L <= 8).
//
// It covers every shell quartet that is not a pure s/p class quartet (those have the register-resident
// class kernels of eri_class.cuh): anything with a d shell, and s/p shells whose components do not
// share one coefficient list.  One WARP evaluates one contracted shell quartet:
//   * per primitive pair the Hermite coefficient tables E^{ij}_t (i,j <= 2, t <= 4, three dimensions)
//     are built in shared memory with the reference's recursion order (basis/src/utils.rs:33-50);
//   * per primitive quartet the Coulomb table R^0_{tuv}, t+u+v <= L, is built ONCE in shared memory,
//     level by level in the auxiliary index with the lanes working on different entries, using the
//     reference's branch order and its x-branch quirk (basis/src/utils.rs:73-106);
//   * the lanes then share out the n_a*n_b*n_c*n_d Cartesian components and accumulate the reference's
//     six-fold Hermite sum (basis/src/contracted_gaussian.rs:170-204) into shared-memory accumulators.
// The reference evaluates each component separately and re-expands E and R for every term; sharing
// the tables across the components of a shell quartet is where the speed-up over the per-function
// generic kernel comes from (36x fewer R builds for a (dd|ss) quartet, 1296x for (dd|dd)).
#pragma once
#include "device_math.cuh"
#include "eri_class.cuh"

namespace fock {

constexpr int kSqThreads = 64;  // two warps = two quartets per block
constexpr int kSqWarps = kSqThreads / 32;
constexpr int kSqMaxComp = 1296;
constexpr int kSqREntries = 165;  // (t,u,v) with t+u+v <= 8
constexpr int kSqETable = 3 * 3 * 3 * 5;
constexpr int kSqSmemPerWarp = kSqMaxComp + 2 * kSqREntries + 2 * kSqETable;  // doubles
constexpr int kWarpqTile = 32;  // quartets per warp tile of the class engine (ksplit = 1)

struct SqShell {
    double c[3];
    int L, nprim, prim_begin;
    int bf[6];
    double scale[6];
};

struct SqJob {
    long long tile_begin;         // first quartet of this job in the launch-wide numbering
    const int *bra_meta;          // [nb][2] shell ids (sorted by maxdiag, descending)
    const int *ket_meta;          // [nk][2]
    const long long *row_prefix;  // [nb+1] quartets in rows < r
    int nb;
    int same_list;                // bra list == ket list: row == col is the same shell pair
};

struct SqLaunch {
    SqJob jobs[2];
    int njobs;
    long long total_quartets;
    const SqShell *shells;
    const double *exps, *coefs;
    BoysTables boys;
    const SqTables *tab;  // device copy (global memory; staged into shared memory by every block)
    int mode, rank, world;
};

// Hermite index tables: entries (t,u,v) ordered by total degree (so degree <= m is a prefix of
// C(m+3,3) entries) and the inverse map.
struct SqTables {
    unsigned char t[kSqREntries], u[kSqREntries], v[kSqREntries];
    unsigned char idx[9][9][9];
};

// Cartesian powers of component c of a shell with L <= 2, reference order (x outer .. z inner)
__device__ __forceinline__ void sq_cart(int L, int c, int &x, int &y, int &z) {
    // L=0: (000); L=1: (001),(010),(100); L=2: (002),(011),(020),(101),(110),(200)
    const int code = L == 0 ? 0 : (L == 1 ? (c == 0 ? 0x001 : (c == 1 ? 0x010 : 0x100))
                                          : (c == 0 ? 0x002 : c == 1 ? 0x011 : c == 2 ? 0x020 : c == 3 ? 0x101 : c == 4 ? 0x110 : 0x200));
    x = (code >> 8) & 0xf;
    y = (code >> 4) & 0xf;
    z = code & 0xf;
}

// E^{ij}_t for i <= la, j <= lb, one dimension, into E[(i*3+j)*5+t] (zero where t > i+j).
__device__ __forceinline__ void sq_hermite_dim(double *E, int la, int lb, double dist, double a, double b) {
    const double p = a + b, q = a * b / p, half_ip = 1.0 / (2.0 * p);
    const double fa = q * dist / a, fb = q * dist / b;
    for (int k = 0; k < 45; ++k) E[k] = 0.0;
    E[0] = exp(-q * (dist * dist));
    for (int i = 1; i <= la; ++i)
        for (int t = 0; t <= i; ++t)
            E[(i * 3) * 5 + t] = half_ip * (t > 0 ? E[((i - 1) * 3) * 5 + t - 1] : 0.0) - fa * E[((i - 1) * 3) * 5 + t] +
                                 (t + 1 <= i - 1 ? (double)(t + 1) * E[((i - 1) * 3) * 5 + t + 1] : 0.0);
    for (int i = 0; i <= la; ++i)
        for (int j = 1; j <= lb; ++j)
            for (int t = 0; t <= i + j; ++t)
                E[(i * 3 + j) * 5 + t] = half_ip * (t > 0 ? E[(i * 3 + j - 1) * 5 + t - 1] : 0.0) + fb * E[(i * 3 + j - 1) * 5 + t] +
                                         (t + 1 <= i + j - 1 ? (double)(t + 1) * E[(i * 3 + j - 1) * 5 + t + 1] : 0.0);
}

// What one warp needs besides its shared-memory scratch.
struct SqEnv {
    const SqShell *shells;
    const double *exps, *coefs;
    BoysTables boys;
    int mode;
    const SqTables *tab;
};

// per-warp shared-memory scratch: acc[kSqMaxComp], R ping-pong, bra / ket Hermite tables
struct SqScratch {
    double *acc, *Rbuf, *Eb, *Ek;
    __device__ __forceinline__ explicit SqScratch(double *base)
        : acc(base), Rbuf(base + kSqMaxComp), Eb(base + kSqMaxComp + 2 * kSqREntries), Ek(base + kSqMaxComp + 2 * kSqREntries + kSqETable) {}
};

// One contracted shell quartet (AB|CD) by one warp: the integrals are accumulated in sc.acc and the
// unique ones handed to sink.put (same uniqueness rule as emit_quartet of the class kernels).
template <class Sink>
__device__ __forceinline__ void sq_eval_quartet(const SqEnv &env, const SqTables &s_tab, const SqScratch &sc, int lane, int shA,
                                                int shB, int shC, int shD, bool same_pair, const Sink &sink) {
    double *acc = sc.acc, *Rbuf = sc.Rbuf, *Eb = sc.Eb, *Ek = sc.Ek;
    const bool fix_r = (env.mode & FOCK_MODE_FIX_R_BIT) != 0, accurate = (env.mode & FOCK_MODE_BOYS_ACCURATE_BIT) != 0;
    const SqShell A = env.shells[shA], B = env.shells[shB], C = env.shells[shC], D = env.shells[shD];
    const int nA = ncart_c(A.L), nB = ncart_c(B.L), nC = ncart_c(C.L), nD = ncart_c(D.L);
    const int ncomp = nA * nB * nC * nD;
    const int L = A.L + B.L + C.L + D.L;
    for (int c = lane; c < ncomp; c += 32) acc[c] = 0.0;
    const double pref = 2.0 * sqrt(powi_rt(FOCK_PI, 5));
    for (int i = 0; i < A.nprim; ++i)
        for (int jb = 0; jb < B.nprim; ++jb) {
            const double a = env.exps[A.prim_begin + i], b = env.exps[B.prim_begin + jb];
            const double p = a + b;
            const double Px = (a * A.c[0] + b * B.c[0]) / p, Py = (a * A.c[1] + b * B.c[1]) / p, Pz = (a * A.c[2] + b * B.c[2]) / p;
            const double cab = env.coefs[A.prim_begin + i] * env.coefs[B.prim_begin + jb];
            __syncwarp();
            if (lane < 3) sq_hermite_dim(Eb + lane * 45, A.L, B.L, B.c[lane] - A.c[lane], a, b);
            for (int k = 0; k < C.nprim; ++k)
                for (int l = 0; l < D.nprim; ++l) {
                    const double c = env.exps[C.prim_begin + k], d = env.exps[D.prim_begin + l];
                    const double qq = c + d;
                    const double X = (c * C.c[0] + d * D.c[0]) / qq - Px;
                    const double Y = (c * C.c[1] + d * D.c[1]) / qq - Py;
                    const double Z = (c * C.c[2] + d * D.c[2]) / qq - Pz;
                    const double alpha = p * qq / (p + qq);
                    const double T = alpha * (X * X + Y * Y + Z * Z);
                    __syncwarp();
                    if (lane < 3) sq_hermite_dim(Ek + lane * 45, C.L, D.L, D.c[lane] - C.c[lane], c, d);
                    double F[9];
                    if (accurate) boys_accurate(T, L, F, env.boys.tab); else boys_faithful(T, L, F, env.boys.lut);
                    // R table, level by level (level n holds entries of degree <= L - n)
                    for (int n = L; n >= 0; --n) {
                        double *cur = Rbuf + (n & 1) * kSqREntries;
                        const double *prev = Rbuf + ((n + 1) & 1) * kSqREntries;
                        const int deg = L - n;
                        const int cnt = (deg + 1) * (deg + 2) * (deg + 3) / 6;
                        for (int e = lane; e < cnt; e += 32) {
                            const int t = s_tab.t[e], u = s_tab.u[e], v = s_tab.v[e];
                            double r;
                            if (e == 0) {
                                r = powi_rt(-2.0 * alpha, n) * F[n];
                            } else if (t == 0 && u == 0) {
                                r = Z * prev[s_tab.idx[0][0][v - 1]] + (v > 1 ? (double)(v - 1) * prev[s_tab.idx[0][0][v - 2]] : 0.0);
                            } else if (t == 0) {
                                r = Y * prev[s_tab.idx[0][u - 1][v]] + (u > 1 ? (double)(u - 1) * prev[s_tab.idx[0][u - 2][v]] : 0.0);
                            } else {
                                double second = 0.0;
                                if (t > 1) {
                                    second = prev[s_tab.idx[t - 2][u][v]];
                                    if (fix_r) second *= (double)(t - 1);
                                }
                                r = X * prev[s_tab.idx[t - 1][u][v]] + second;
                            }
                            cur[e] = r;
                        }
                        __syncwarp();
                    }
                    const double *R = Rbuf;  // level 0
                    const double w = cab * env.coefs[C.prim_begin + k] * env.coefs[D.prim_begin + l] * (pref * (1.0 / (p * qq * sqrt(p + qq))));
                    for (int comp = lane; comp < ncomp; comp += 32) {
                        int rem = comp;
                        const int cd = rem % nD; rem /= nD;
                        const int cc = rem % nC; rem /= nC;
                        const int cb = rem % nB;
                        const int ca = rem / nB;
                        int ax, ay, az, bx, by, bz, cx, cy, cz, dx, dy, dz;
                        sq_cart(A.L, ca, ax, ay, az);
                        sq_cart(B.L, cb, bx, by, bz);
                        sq_cart(C.L, cc, cx, cy, cz);
                        sq_cart(D.L, cd, dx, dy, dz);
                        const double *ebx = Eb + (ax * 3 + bx) * 5, *eby = Eb + 45 + (ay * 3 + by) * 5, *ebz = Eb + 90 + (az * 3 + bz) * 5;
                        const double *ekx = Ek + (cx * 3 + dx) * 5, *eky = Ek + 45 + (cy * 3 + dy) * 5, *ekz = Ek + 90 + (cz * 3 + dz) * 5;
                        double s = 0.0;
                        for (int t1 = 0; t1 <= ax + bx; ++t1)
                            for (int u1 = 0; u1 <= ay + by; ++u1)
                                for (int v1 = 0; v1 <= az + bz; ++v1) {
                                    const double eb = ebx[t1] * eby[u1] * ebz[v1];
                                    for (int t2 = 0; t2 <= cx + dx; ++t2)
                                        for (int u2 = 0; u2 <= cy + dy; ++u2)
                                            for (int v2 = 0; v2 <= cz + dz; ++v2) {
                                                const double term = eb * ekx[t2] * eky[u2] * ekz[v2] * R[s_tab.idx[t1 + t2][u1 + u2][v1 + v2]];
                                                s += ((t2 + u2 + v2) & 1) ? -term : term;
                                            }
                                }
                        acc[comp] += w * s;
                    }
                }
        }
    __syncwarp();
    // emit: unique integrals only (same rule as emit_quartet of the class kernels)
    const bool ab_same = shA == shB, cd_same = shC == shD;
    for (int comp = lane; comp < ncomp; comp += 32) {
        int rem = comp;
        const int cd = rem % nD; rem /= nD;
        const int cc = rem % nC; rem /= nC;
        const int cb = rem % nB;
        const int ca = rem / nB;
        const int x = A.bf[ca], y = B.bf[cb], z = C.bf[cc], w = D.bf[cd];
        if (x < 0 || y < 0 || z < 0 || w < 0) continue;
        if (ab_same && x < y) continue;
        if (cd_same && z < w) continue;
        if (same_pair) {
            const size_t xy = x >= y ? tri(x, y) : tri(y, x);
            const size_t zw = z >= w ? tri(z, w) : tri(w, z);
            if (xy < zw) continue;
        }
        sink.put(x, y, z, w, acc[comp] * A.scale[ca] * B.scale[cb] * C.scale[cc] * D.scale[cd]);
    }
    __syncwarp();
}

__device__ __forceinline__ void sq_stage_tables(SqTables &s_tab, const SqTables *g_tab, int tid, int nthreads) {
    for (int i = tid; i < (int)(sizeof(SqTables) / sizeof(int)); i += nthreads)
        reinterpret_cast<int *>(&s_tab)[i] = reinterpret_cast<const int *>(g_tab)[i];
}

template <class Sink>
__global__ void __launch_bounds__(kSqThreads) eri_shellq_kernel(SqLaunch sl, Sink sink) {
    extern __shared__ double s_sq[];
    // index tables in shared memory: constant memory serialises lane-divergent lookups
    __shared__ SqTables s_tab;
    sq_stage_tables(s_tab, sl.tab, threadIdx.x, kSqThreads);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SqScratch sc(s_sq + (size_t)warp * kSqSmemPerWarp);
    const SqEnv env{sl.shells, sl.exps, sl.coefs, sl.boys, sl.mode, sl.tab};
    const long long mine = (sl.total_quartets - sl.rank + sl.world - 1) / sl.world;
    for (long long g = (long long)blockIdx.x * kSqWarps + warp; g < mine; g += (long long)gridDim.x * kSqWarps) {
        const long long quartet = g * sl.world + sl.rank;
        const int j = (sl.njobs > 1 && quartet >= sl.jobs[1].tile_begin) ? 1 : 0;
        const SqJob job = sl.jobs[j];
        const long long q = quartet - job.tile_begin;
        const int row = warp_search(job.row_prefix, job.nb, q, lane);
        const int col = (int)(q - job.row_prefix[row]);
        sq_eval_quartet(env, s_tab, sc, lane, job.bra_meta[2 * row], job.bra_meta[2 * row + 1], job.ket_meta[2 * col],
                        job.ket_meta[2 * col + 1], job.same_list && row == col, sink);
    }
}

// The same evaluation driven by the class engine's job / warp-tile tables (class_launch.cu), for the
// shell classes that have no register-resident class kernel (total angular momentum >= 6): a warp walks
// the (up to 32) quartets of its tile one after the other.
template <class Sink>
__global__ void __launch_bounds__(kSqThreads) eri_class_warpq_kernel(ClassLaunch cl, SqEnv env, Sink sink) {
    extern __shared__ double s_sq[];
    __shared__ SqTables s_tab;
    sq_stage_tables(s_tab, env.tab, threadIdx.x, kSqThreads);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SqScratch sc(s_sq + (size_t)warp * kSqSmemPerWarp);
    const long long my_tiles = (cl.total_tiles - cl.rank + cl.world - 1) / cl.world;
    for (long long g = (long long)blockIdx.x * kSqWarps + warp; g < my_tiles; g += (long long)gridDim.x * kSqWarps) {
        const long long tile = g * cl.world + cl.rank;
        int j = 0;
        while (j + 1 < cl.njobs && cl.jobs[j + 1].tile_begin <= tile) ++j;
        const ClassJob job = cl.jobs[j];
        const long long q0 = (tile - job.tile_begin) * kWarpqTile;  // ksplit is 1 for these jobs
        const long long total_q = job_prefix(job, job.nb);
        int row = warp_search(job.row_prefix, job.nb, q0 + job.q_base, lane);
        for (long long q = q0; q < q0 + kWarpqTile && q < total_q; ++q) {
            while (row + 1 < job.nb && job_prefix(job, row + 1) <= q) ++row;
            int brow = row, kcol = (int)(q - job_prefix(job, row));
            if (job.col_idx) {  // screened plan in spatial order: CSR of the surviving quartets (eri_class.cuh: ClassJob)
                kcol = job.col_idx[q];
                brow = job.row_map[row];
            }
            sq_eval_quartet(env, s_tab, sc, lane, job.bra_meta[2 * brow], job.bra_meta[2 * brow + 1], job.ket_meta[2 * kcol],
                            job.ket_meta[2 * kcol + 1], job.same_sublist && brow == kcol, sink);
        }
    }
}

}  // namespace fock
