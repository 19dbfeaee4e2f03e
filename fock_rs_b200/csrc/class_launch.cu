// class_launch.cu -- host side of the shell-class ERI kernels: diagonal pass, sublist sorting
// (device-side permutation of the primitive-pair tables), the job / warp-tile plan, kernel dispatch.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include "../../include/fock_b200.h"
#include "class_launch.h"
#include "eri_class.cuh"
#include "eri_shellq.cuh"

namespace fock {

namespace {

template <int LA, int LB, int LC, int LD, int KC0, int KCN, class Sink>
void launch_mode(const ClassLaunch &cl, const Sink &sink, int mode, dim3 grid, cudaStream_t st) {
    constexpr int NAB = ncart_c(LA) * ncart_c(LB);
    const size_t smem = class_smem_bytes(LA, LB, LA + LB + LC + LD, cl.boys_ni);
    if (cl.ring) st = cl.ring->next();
    if (mode & FOCK_MODE_FIX_R_BIT)
        eri_class_kernel<LA, LB, LC, LD, 1, 0, NAB, KC0, KCN, Sink><<<grid, kClassThreads, smem, st>>>(cl, sink);
    else
        eri_class_kernel<LA, LB, LC, LD, 0, 0, NAB, KC0, KCN, Sink><<<grid, kClassThreads, smem, st>>>(cl, sink);
}

// pair classes: 0 ss, 1 ps, 2 pp (here); 3 ds, 4 dp, 5 dd bras in dclass_*.cu; returns 0 when no
// register-resident kernel exists for the class pair
template <class Sink>
int launch_class(int cb, int ck, const ClassLaunch &cl, const Sink &sink, int mode, dim3 grid, cudaStream_t st) {
    if (cb == 3) return launch_dclass_ds(ck, cl, sink, grid, st);
    if (cb == 4) return launch_dclass_dp(ck, cl, sink, grid, st);
    if (cb == 5) return launch_dclass_dd(ck, cl, sink, grid, st);
    switch (cb * 3 + ck) {
        case 0: launch_mode<0, 0, 0, 0, 0, 1>(cl, sink, mode, grid, st); return 1;
        case 3: launch_mode<1, 0, 0, 0, 0, 1>(cl, sink, mode, grid, st); return 1;
        case 4: launch_mode<1, 0, 1, 0, 0, 3>(cl, sink, mode, grid, st); return 1;
        case 6: launch_mode<1, 1, 0, 0, 0, 1>(cl, sink, mode, grid, st); return 1;
        case 7: launch_mode<1, 1, 1, 0, 0, 3>(cl, sink, mode, grid, st); return 1;
        case 8:  // (pp|pp): 81 accumulators do not fit one thread's registers -> three ket-component slices
            launch_mode<1, 1, 1, 1, 0, 3>(cl, sink, mode, grid, st);
            launch_mode<1, 1, 1, 1, 3, 3>(cl, sink, mode, grid, st);
            launch_mode<1, 1, 1, 1, 6, 3>(cl, sink, mode, grid, st);
            return 3;
        default: return 0;
    }
}

template <int LA, int LB, int KC0, int KCN>
void launch_diag_mode(const DiagLaunch &dl, const DiagSink &sink, int mode, dim3 grid, cudaStream_t st) {
    const size_t smem = boys_smem_bytes(2 * (LA + LB), dl.boys_ni);
    if (mode & FOCK_MODE_FIX_R_BIT)
        eri_diag_kernel<LA, LB, 1, KC0, KCN><<<grid, kClassThreads, smem, st>>>(dl, sink);
    else
        eri_diag_kernel<LA, LB, 0, KC0, KCN><<<grid, kClassThreads, smem, st>>>(dl, sink);
}

// sorted tables from the natural-order master: position i of a sublist takes natural pair perm[i].
// One launch for all sublists; block = one pair.
__global__ void permute_pairs_kernel(const PermuteDesc *__restrict__ descs, int count, const double *__restrict__ bra_nat, const int *__restrict__ meta_nat,
                                     const int *__restrict__ perm, double *__restrict__ bra, double *__restrict__ ket,
                                     int *__restrict__ meta, const unsigned long long *__restrict__ maxdiag_nat, double *__restrict__ md_sorted) {
    const int g = blockIdx.x;
    int s = 0, hi = count - 1;  // largest s with descs[s].first <= g
    while (s < hi) {
        const int mid = (s + hi + 1) >> 1;
        if (descs[mid].first <= g) s = mid; else hi = mid - 1;
    }
    const PermuteDesc d = descs[s];
    const int i = g - d.first;
    if (i >= d.n) return;
    const int src = perm[d.meta_off / 2 + i];
    const int len = d.K * d.nf;
    const double *in = bra_nat + d.bra_off + (size_t)src * len;
    double *ob = bra + d.bra_off + (size_t)i * len, *ok = ket + d.ket_off + i;
    for (int e = threadIdx.x; e < len; e += blockDim.x) {
        const double v = in[e];
        ob[e] = v;
        ok[(size_t)e * d.npad] = v;  // e = ij*nf + f
    }
    if (threadIdx.x < 2) meta[d.meta_off + 2 * i + threadIdx.x] = meta_nat[d.meta_off + 2 * src + threadIdx.x];
    if (threadIdx.x == 2) md_sorted[d.meta_off / 2 + i] = __longlong_as_double((long long)maxdiag_nat[d.meta_off / 2 + src]);
}

// per-pair max |(xy|xy)| of the pairs [first, first+n) of the natural-order meta table, from the
// basis-function diagonal (sublists with a d shell: their diagonal comes from the generic kernel)
__global__ void class_pairmax_kernel(const int *__restrict__ meta_nat, const int *__restrict__ shell_bf, int first, int n,
                                     const double *__restrict__ diag, unsigned long long *__restrict__ maxdiag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int shA = meta_nat[2 * (first + i)], shB = meta_nat[2 * (first + i) + 1];
    double m = 0.0;
    for (int ca = 0; ca < 6; ++ca)
        for (int cb = 0; cb < 6; ++cb) {
            const int x = shell_bf[shA * 6 + ca], y = shell_bf[shB * 6 + cb];
            if (x < 0 || y < 0) continue;
            const double d = fabs(diag[x >= y ? tri(x, y) : tri(y, x)]);
            if (!(d <= m)) m = (d == d) ? d : __longlong_as_double(0x7ff0000000000000LL);
        }
    maxdiag[first + i] = (unsigned long long)__double_as_longlong(m);
}

// sort key of the spatial order: magnitude bucket of the pair's largest diagonal integral (strong first) above the
// Morton code of the pair centre
__global__ void spatial_keys_kernel(const unsigned long long *__restrict__ maxdiag, const unsigned long long *__restrict__ morton, int n,
                                    int bucket_bits, unsigned long long *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = (int)(maxdiag[i] >> 52) & 0x7ff;  // biased exponent of a non-negative double (inf / NaN-as-inf: 2047 -> bucket 0)
    const int b = min(255, max(0, 1030 - e) >> bucket_bits);
    keys[i] = ((unsigned long long)b << 48) | (morton[i] & 0xffffffffffffULL);
}

// ---- CSR of the surviving quartets (spatial pair order) ---------------------------------------------------------------
// One candidate job = (bra sublist, ket sublist) of a class pair.  Its rows are numbered plan-wide from `row_base`; a quartet
// (r, c) survives when the shell-level bound on the per-pair maxima passes (pair_bound_passes below; the exact per-function
// rule is re-applied in the sinks) and, inside one sublist, c <= r.
struct PlanJobDesc {
    long long row_base;
    int md_b, md_k;  // first pair of the bra / ket sublist in the sorted per-pair maxima
    int nb, nk;
    int same, pad;
};

__device__ __forceinline__ int plan_job_of_row(const PlanJobDesc *__restrict__ jobs, int njobs, long long w) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].row_base <= w) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// pass 1: survivors per row (one warp per row)
__global__ void plan_count_kernel(const PlanJobDesc *__restrict__ jobs, int njobs, long long total_rows, const double *__restrict__ md,
                                  double thr2, long long *__restrict__ cnt, long long *__restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total_rows; w += nwarps) {
        const PlanJobDesc j = jobs[plan_job_of_row(jobs, njobs, w)];
        const int r = (int)(w - j.row_base);
        const double db = md[j.md_b + r];
        const int limit = j.same ? r + 1 : j.nk;
        int n = 0;
        for (int c0 = 0; c0 < limit; c0 += 32) {
            const int c = c0 + lane;
            const bool ok = c < limit && db * md[j.md_k + c] > thr2;
            n += __popc(__ballot_sync(0xffffffffu, ok));
        }
        if (lane == 0) {
            cnt[w] = n;
            flag[w] = n > 0;
        }
    }
}

// pass 2 (after the two scans): list the non-empty rows -- their quartet offset and bra pair
__global__ void plan_compact_kernel(const PlanJobDesc *__restrict__ jobs, int njobs, long long total_rows, const long long *__restrict__ cnt,
                                    const long long *__restrict__ gprefix, const long long *__restrict__ cflag,
                                    long long *__restrict__ cprefix, int *__restrict__ crow) {
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w <= total_rows; w += (long long)gridDim.x * blockDim.x) {
        if (w == total_rows) {
            cprefix[cflag[w]] = gprefix[w];  // sentinel: one past the last listed row
        } else if (cnt[w] > 0) {
            const PlanJobDesc j = jobs[plan_job_of_row(jobs, njobs, w)];
            cprefix[cflag[w]] = gprefix[w];
            crow[cflag[w]] = (int)(w - j.row_base);
        }
    }
}

// pass 3: the ket pair of every surviving quartet, in row order (one warp per row)
__global__ void plan_fill_kernel(const PlanJobDesc *__restrict__ jobs, int njobs, long long total_rows, const double *__restrict__ md,
                                 double thr2, const long long *__restrict__ gprefix, int *__restrict__ col) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total_rows; w += nwarps) {
        if (gprefix[w + 1] == gprefix[w]) continue;
        const PlanJobDesc j = jobs[plan_job_of_row(jobs, njobs, w)];
        const int r = (int)(w - j.row_base);
        const double db = md[j.md_b + r];
        const int limit = j.same ? r + 1 : j.nk;
        long long out = gprefix[w];
        for (int c0 = 0; c0 < limit; c0 += 32) {
            const int c = c0 + lane;
            const bool ok = c < limit && db * md[j.md_k + c] > thr2;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) col[out + __popc(m & ((1u << lane) - 1u))] = c;
            out += __popc(m);
        }
    }
}

// quartet / listed-row offsets at the start of every candidate job (+ the totals)
__global__ void plan_jobsum_kernel(const PlanJobDesc *__restrict__ jobs, int njobs, long long total_rows, const long long *__restrict__ gprefix,
                                   const long long *__restrict__ cflag, long long *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > njobs) return;
    const long long w = j < njobs ? jobs[j].row_base : total_rows;
    out[2 * j] = gprefix[w];
    out[2 * j + 1] = cflag[w];
}

// (job, first row) of every warp tile of a class launch: what the kernel would otherwise search for per tile
// (two dependent ballot rounds over the job table, three over the row prefixes); one thread per tile, once per plan.
__global__ void tile_table_kernel(const ClassJob *__restrict__ jobs, int njobs, long long total_tiles, int2 *__restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total_tiles; t += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = njobs - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
        }
        const long long *prefix = jobs[lo].row_prefix;
        const long long q0 = (t - jobs[lo].tile_begin) * (32 / jobs[lo].ksplit) + jobs[lo].q_base;
        int a = 0, b = jobs[lo].nb - 1;
        while (a < b) {  // largest row with prefix[row] <= q0 (skips empty rows, like warp_search)
            const int mid = (a + b + 1) >> 1;
            if (prefix[mid] <= q0) a = mid; else b = mid - 1;
        }
        out[t] = make_int2(lo, a);
    }
}

}  // namespace

int ClassEngine::upload_master(std::string &err) {
    std::vector<double> bra, ket;
    std::vector<int> meta;
    build_pair_tables(*sd, bra, ket, meta);
    nat_ids.clear();
    for (const Sublist &s : sd->sublists) nat_ids.push_back(s.pair_ids);
    cudaError_t e;
    auto alloc = [&](void **p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 8); };
    if ((e = alloc((void **)&d_bra_nat, bra.size() * 8)) != cudaSuccess || (e = alloc((void **)&dev->d_bra, bra.size() * 8)) != cudaSuccess ||
        (e = alloc((void **)&dev->d_ket, ket.size() * 8)) != cudaSuccess || (e = alloc((void **)&d_meta_nat, meta.size() * 4)) != cudaSuccess ||
        (e = alloc((void **)&dev->d_meta, meta.size() * 4)) != cudaSuccess || (e = alloc((void **)&d_perm, meta.size() * 2 + 8)) != cudaSuccess ||
        (e = cudaMallocHost((void **)&h_perm, meta.size() * 2 + 8)) != cudaSuccess ||
        (e = alloc((void **)&d_maxdiag, meta.size() * 4 + 8)) != cudaSuccess ||
        (e = cudaMallocHost((void **)&h_maxdiag, meta.size() * 4 + 8)) != cudaSuccess ||
        (e = alloc((void **)&d_keys_sorted, meta.size() * 4 + 8)) != cudaSuccess ||
        (e = cudaMallocHost((void **)&h_keys_sorted, meta.size() * 4 + 8)) != cudaSuccess ||
        (e = alloc((void **)&d_iota, meta.size() * 2 + 8)) != cudaSuccess ||
        (e = alloc((void **)&d_seg, (2 * sd->sublists.size() + 2) * sizeof(int))) != cudaSuccess ||
        (e = alloc((void **)&d_md_sorted, meta.size() * 4 + 8)) != cudaSuccess || (e = alloc((void **)&d_morton, meta.size() * 4 + 8)) != cudaSuccess ||
        (e = alloc((void **)&d_sortkeys, meta.size() * 4 + 8)) != cudaSuccess) {
        err = std::string("pair tables: ") + cudaGetErrorString(e);
        return 1;
    }
    {
        const char *po = getenv("FOCK_PAIR_ORDER");
        spatial_order = !(po && std::string(po) == "diag");
        // static sort keys of the spatial order: 3 x 16-bit Morton code of the pair centre (the product centre of the two most
        // diffuse primitives: the primitive pairs that decide near / far field), natural order of every sublist
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (const Shell &sh : sd->shells)
            for (int d = 0; d < 3; ++d) {
                lo[d] = std::min(lo[d], sh.c[d]);
                hi[d] = std::max(hi[d], sh.c[d]);
            }
        double ext = 1e-9;
        for (int d = 0; d < 3; ++d) ext = std::max(ext, hi[d] - lo[d]);
        auto min_exp = [&](const Shell &sh) {
            double m = sd->exps[sh.prim_begin];
            for (int i = 1; i < sh.nprim; ++i) m = std::min(m, sd->exps[sh.prim_begin + i]);
            return m;
        };
        auto spread = [](unsigned long long v) {  // 16 bits -> every third bit
            unsigned long long r = 0;
            for (int i = 0; i < 16; ++i) r |= ((v >> i) & 1ULL) << (3 * i);
            return r;
        };
        std::vector<unsigned long long> keys(meta.size() / 2 + 1, 0ULL);
        for (const Sublist &sl : sd->sublists)
            for (size_t i = 0; i < sl.pair_ids.size(); ++i) {
                const ShellPair &pr = sd->pairs[sl.pair_ids[i]];
                const Shell &A = sd->shells[pr.A], &B = sd->shells[pr.B];
                const double a = min_exp(A), b = min_exp(B);
                unsigned long long key = 0;
                for (int d = 0; d < 3; ++d) {
                    const double c = (a * A.c[d] + b * B.c[d]) / (a + b);
                    const double t = std::min(1.0, std::max(0.0, (c - lo[d]) / ext));
                    key |= spread((unsigned long long)(t * 65535.0)) << d;
                }
                keys[sl.meta_off / 2 + i] = key;
            }
        if (const char *bb = getenv("FOCK_ORDER_BUCKET_BITS")) order_bucket_bits = std::max(0, std::min(10, atoi(bb)));
        cudaMemcpy(d_morton, keys.data(), (meta.size() / 2) * sizeof(unsigned long long), cudaMemcpyHostToDevice);
        // pure Morton order is static, so it is fixed here, on the host, with a stable sort: every rank of a multi-GPU run must
        // number the pairs identically (ranks take interleaved warp tiles of the same plan)
        if (spatial_order && order_bucket_bits == 0) {
            std::vector<int> perm(meta.size() / 2 + 1, 0), idx;
            for (const Sublist &sl : sd->sublists) {
                const size_t base = sl.meta_off / 2, n = sl.pair_ids.size();
                idx.resize(n);
                std::iota(idx.begin(), idx.end(), 0);
                std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return keys[base + x] < keys[base + y]; });
                for (size_t i = 0; i < n; ++i) perm[base + i] = idx[i];
            }
            cudaMemcpy(d_perm, perm.data(), (meta.size() / 2) * sizeof(int), cudaMemcpyHostToDevice);
        }
    }
    std::vector<int> sbf(sd->shells.size() * 6);
    for (size_t s = 0; s < sd->shells.size(); ++s)
        for (int c = 0; c < 6; ++c) sbf[s * 6 + c] = sd->shells[s].bf[c];
    std::vector<double> sscale(sd->shells.size() * 6);
    for (size_t s = 0; s < sd->shells.size(); ++s)
        for (int c = 0; c < 6; ++c) sscale[s * 6 + c] = sd->shells[s].scale[c];
    if ((e = cudaMalloc((void **)&dev->d_shell_bf, sbf.size() * 4 + 8)) != cudaSuccess ||
        (e = cudaMalloc((void **)&d_shell_scale, sscale.size() * 8 + 8)) != cudaSuccess) {
        err = std::string("shell table: ") + cudaGetErrorString(e);
        return 1;
    }
    cudaMemcpy(d_shell_scale, sscale.data(), sscale.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpyAsync(d_bra_nat, bra.data(), bra.size() * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(dev->d_bra, bra.data(), bra.size() * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(dev->d_ket, ket.data(), ket.size() * 8, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_meta_nat, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(dev->d_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(dev->d_shell_bf, sbf.data(), sbf.size() * 4, cudaMemcpyHostToDevice, stream);
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        err = std::string("pair table upload: ") + cudaGetErrorString(e);
        return 1;
    }
    dev->bra_doubles = bra.size();
    dev->ket_doubles = ket.size();
    dev->meta_ints = meta.size();
    // device-side segmented sort of the per-pair maxima (one segment per sublist)
    {
        const int npairs = (int)(meta.size() / 2), nseg = (int)sd->sublists.size();
        std::vector<int> iota(npairs), seg(2 * nseg);
        std::vector<PermuteDesc> descs;
        int first = 0;
        for (int si = 0; si < nseg; ++si) {
            const Sublist &s = sd->sublists[si];
            const int n = (int)s.pair_ids.size(), base = (int)(s.meta_off / 2);
            for (int i = 0; i < n; ++i) iota[base + i] = i;
            seg[si] = base;
            seg[nseg + si] = base + n;
            if (n > 0) {
                descs.push_back(PermuteDesc{s.bra_off, s.ket_off, s.meta_off, first, n, s.K, s.nf, s.npad});
                first += n;
            }
        }
        permute_blocks = first;
        permute_count = (int)descs.size();
        if ((e = cudaMalloc((void **)&d_permute_desc, (descs.size() + 1) * sizeof(PermuteDesc))) != cudaSuccess) {
            err = std::string("permutation table: ") + cudaGetErrorString(e);
            return 1;
        }
        cudaMemcpy(d_permute_desc, descs.data(), descs.size() * sizeof(PermuteDesc), cudaMemcpyHostToDevice);
        cudaMemcpy(d_iota, iota.data(), iota.size() * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(d_seg, seg.data(), seg.size() * sizeof(int), cudaMemcpyHostToDevice);
        sort_temp_bytes = 0;
        cudaGetLastError();  // the size query reports any error left pending on this thread as its own
        size_t bytes_stable = 0;
        if ((e = cub::DeviceSegmentedSort::SortPairsDescending(nullptr, sort_temp_bytes, d_maxdiag, d_keys_sorted, d_iota, d_perm, npairs, nseg,
                                                               d_seg, d_seg + nseg, stream)) != cudaSuccess ||
            (e = cub::DeviceSegmentedSort::StableSortPairs(nullptr, bytes_stable, d_sortkeys, d_keys_sorted, d_iota, d_perm, npairs, nseg, d_seg,
                                                           d_seg + nseg, stream)) != cudaSuccess) {
            err = std::string("sort scratch query: ") + cudaGetErrorString(e);
            return 1;
        }
        sort_temp_bytes = std::max(sort_temp_bytes, bytes_stable);
        if ((e = cudaMalloc(&d_sort_temp, sort_temp_bytes + 16)) != cudaSuccess) {
            err = std::string("sort scratch: ") + cudaGetErrorString(e);
            return 1;
        }
    }
    if ((e = cudaMalloc((void **)&d_tile_counters, (size_t)kNumPairClasses * kNumPairClasses * kTileCounterSlots * sizeof(unsigned long long))) != cudaSuccess) {
        err = std::string("tile counters: ") + cudaGetErrorString(e);
        return 1;
    }
    plan_valid = false;
    return 0;
}

// Diagonal pass: every class pair's (AB|AB) block from the natural-order tables; only (xy|xy) kept.
int ClassEngine::run_diagonal(double *d_diag, std::string &err) {
    int launches = 0;
    const size_t npairs = dev->meta_ints / 2;
    cudaMemsetAsync(d_maxdiag, 0, npairs * sizeof(unsigned long long), stream);
    const DiagSink sink{d_diag, d_maxdiag};
    for (size_t si = 0; si < sd->sublists.size(); ++si) {
        const Sublist &s = sd->sublists[si];
        const int n = (int)s.pair_ids.size();
        if (!n) continue;
        DiagLaunch dl{d_bra_nat + s.bra_off, d_meta_nat + s.meta_off, dev->d_shell_bf, boys_poly, boys_ni, s.K, n, (int)(s.meta_off / 2)};
        dim3 grid((unsigned)std::min((n + kClassWarps - 1) / kClassWarps, num_sms * 4));
        if (s.cls == 0) {
            launch_diag_mode<0, 0, 0, 1>(dl, sink, mode, grid, stream);
            launches += 1;
        } else if (s.cls == 1) {
            launch_diag_mode<1, 0, 0, 3>(dl, sink, mode, grid, stream);
            launches += 1;
        } else if (s.cls == 2) {
            launch_diag_mode<1, 1, 0, 3>(dl, sink, mode, grid, stream);
            launch_diag_mode<1, 1, 3, 3>(dl, sink, mode, grid, stream);
            launch_diag_mode<1, 1, 6, 3>(dl, sink, mode, grid, stream);
            launches += 3;
        }  // pairs with a d shell: the generic per-function diagonal kernel (fock_api.cu), then finish_diagonal
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("diagonal kernels: ") + cudaGetErrorString(e);
        return -1;
    }
    return launches;
}

int ClassEngine::finish_diagonal(const double *d_diag, std::string &err) {
    int launches = 0;
    const size_t npairs = dev->meta_ints / 2;
    for (size_t si = 0; si < sd->sublists.size(); ++si) {
        const Sublist &s = sd->sublists[si];
        const int n = (int)s.pair_ids.size();
        if (!n || s.cls < 3) continue;
        class_pairmax_kernel<<<(n + 127) / 128, 128, 0, stream>>>(d_meta_nat, dev->d_shell_bf, (int)(s.meta_off / 2), n, d_diag, d_maxdiag);
        launches += 1;
    }
    // order every sublist by its pairs' maxima (descending) on the device -- non-negative doubles sort
    // like their bit patterns -- and permute the primitive-pair tables accordingly; the host only needs
    // the sorted maxima (for the screening prefixes) and, for the shell-quartet path, the natural ones
    const int nseg = (int)sd->sublists.size();
    // spatial order: (magnitude bucket, Morton code) keys, stable sort (ties keep the natural order, so every rank of a
    // multi-GPU run numbers the pairs identically); pure Morton order was fixed at upload
    cudaError_t se = cudaSuccess;
    if (!spatial_order) {
        se = cub::DeviceSegmentedSort::SortPairsDescending(d_sort_temp, sort_temp_bytes, d_maxdiag, d_keys_sorted, d_iota, d_perm, (int)npairs, nseg,
                                                           d_seg, d_seg + nseg, stream);
    } else if (order_bucket_bits > 0) {
        spatial_keys_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, stream>>>(d_maxdiag, d_morton, (int)npairs, order_bucket_bits, d_sortkeys);
        se = cub::DeviceSegmentedSort::StableSortPairs(d_sort_temp, sort_temp_bytes, d_sortkeys, d_keys_sorted, d_iota, d_perm, (int)npairs, nseg,
                                                       d_seg, d_seg + nseg, stream);
        launches += 1;
    }
    if (se != cudaSuccess) {
        err = std::string("sublist sort: ") + cudaGetErrorString(se);
        return -1;
    }
    if (permute_blocks > 0)
        permute_pairs_kernel<<<permute_blocks, 64, 0, stream>>>(d_permute_desc, permute_count, d_bra_nat, d_meta_nat, d_perm, dev->d_bra, dev->d_ket,
                                                               dev->d_meta, d_maxdiag, d_md_sorted);
    launches += 2;
    cudaMemcpyAsync(h_keys_sorted, d_md_sorted, npairs * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);  // maxima in sorted order
    cudaMemcpyAsync(h_maxdiag, d_maxdiag, npairs * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("diagonal kernels: ") + cudaGetErrorString(e);
        return -1;
    }
    return launches;
}

// With the diagonal on the host: per-pair max |(xy|xy)|, sort each sublist (descending), permute the
// device tables accordingly.
int ClassEngine::sort_by_diagonal(std::string &err) {
    (void)err;
    // the sort and the permutation already ran on the device (run_diagonal); record the natural maxima
    // on the host pairs (the shell-quartet path screens against them)
    const double *maxd = reinterpret_cast<const double *>(h_maxdiag);
    for (size_t si = 0; si < sd->sublists.size(); ++si) {
        const Sublist &s = sd->sublists[si];
        const std::vector<int> &nat = nat_ids[si];
        const double *m = maxd + s.meta_off / 2;
        for (size_t i = 0; i < nat.size(); ++i) sd->pairs[nat[i]].maxdiag = m[i];
    }
    plan_valid = false;
    return 0;
}

// Number of ket pairs (sorted descending by maxdiag) that pass the shell-level bound against a bra
// pair with value db.  The basis-function rule sqrt(d1*d2) > thr is monotone in both arguments, so
// when the bound on the per-pair maxima fails every component fails.
// (The shell-level bound only has to be conservative -- the exact rule is re-applied per integral in
// the sink -- so it compares the product with thr^2 shrunk by a few ulps instead of taking a sqrt.)
static inline bool pair_bound_passes(double db, double dk, double thr2_slack) { return db * dk > thr2_slack; }
static int surviving_prefix(double db, const std::vector<double> &dk, int screen, double thr) {
    if (screen == FOCK_SCREEN_OFF) return (int)dk.size();
    const double t2 = thr * thr * (1.0 - 1e-12);
    int lo = 0, hi = (int)dk.size();  // first index that fails
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (pair_bound_passes(db, dk[mid], t2)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

int ClassEngine::make_plan(int screen, double thr, int rank, int world, std::string &err) {
    if (plan_valid && plan_screen == screen && plan_thr == thr && plan_rank == rank && plan_world == world) return 0;
    const int ns = (int)sd->sublists.size();
    // screened plan in spatial order: the survivors are a CSR matrix built on the device; otherwise (no screening, or the r01
    // order by largest diagonal integral) a dense staircase whose row lengths the host walks down
    const bool csr = spatial_order && screen != FOCK_SCREEN_OFF;
    std::vector<std::vector<double>> md(ns);  // per sublist: maxima in sorted order
    {
        const double *sorted = reinterpret_cast<const double *>(h_keys_sorted);
        for (int s = 0; s < ns; ++s) {
            const size_t base = sd->sublists[s].meta_off / 2, n = sd->sublists[s].pair_ids.size();
            md[s].assign(sorted + base, sorted + base + n);
        }
    }
    cudaError_t e;
    // ---- CSR, phase 1: every candidate job (same nested order as the launch loop below) -> row counts, scans, column lists
    std::vector<PlanJobDesc> cand;
    std::vector<int> cand_of(csr ? (size_t)ns * ns : 0, -1);
    if (csr) {
        long long total_rows = 0;
        for (int cb = 0; cb < kNumPairClasses; ++cb)
            for (int ck = 0; ck <= cb; ++ck)
                for (int b = 0; b < ns; ++b) {
                    const Sublist &sb = sd->sublists[b];
                    if (sb.cls != cb || sb.pair_ids.empty()) continue;
                    for (int k = 0; k <= b; ++k) {
                        const Sublist &sk = sd->sublists[k];
                        if (sk.cls != ck || sk.pair_ids.empty()) continue;
                        cand_of[(size_t)b * ns + k] = (int)cand.size();
                        cand.push_back(PlanJobDesc{total_rows, (int)(sb.meta_off / 2), (int)(sk.meta_off / 2), (int)sb.pair_ids.size(),
                                                   (int)sk.pair_ids.size(), b == k ? 1 : 0, 0});
                        total_rows += (long long)sb.pair_ids.size();
                    }
                }
        const int nc = (int)cand.size();
        if (nc > 0) {
            if ((size_t)nc + 2 > plan_desc_cap) {
                if (d_plan_desc) cudaFree(d_plan_desc);
                if (h_plan_desc) cudaFreeHost(h_plan_desc);
                if (d_jobsum) cudaFree(d_jobsum);
                if (h_jobsum) cudaFreeHost(h_jobsum);
                d_plan_desc = h_plan_desc = nullptr;
                d_jobsum = h_jobsum = nullptr;
                plan_desc_cap = (size_t)nc + 64;
                if ((e = cudaMalloc(&d_plan_desc, plan_desc_cap * sizeof(PlanJobDesc))) != cudaSuccess ||
                    (e = cudaMallocHost(&h_plan_desc, plan_desc_cap * sizeof(PlanJobDesc))) != cudaSuccess ||
                    (e = cudaMalloc((void **)&d_jobsum, plan_desc_cap * 2 * sizeof(long long))) != cudaSuccess ||
                    (e = cudaMallocHost((void **)&h_jobsum, plan_desc_cap * 2 * sizeof(long long))) != cudaSuccess) {
                    err = std::string("plan descriptors: ") + cudaGetErrorString(e);
                    plan_desc_cap = 0;
                    return 1;
                }
            }
            if ((size_t)total_rows + 2 > rows_cap) {
                for (void *q : {(void *)d_rowcnt, (void *)d_gprefix, (void *)d_rowflag, (void *)d_cflag, (void *)d_cprefix, (void *)d_crow, d_scan_temp})
                    if (q) cudaFree(q);
                d_rowcnt = d_gprefix = d_rowflag = d_cflag = d_cprefix = nullptr;
                d_crow = nullptr;
                d_scan_temp = nullptr;
                rows_cap = (size_t)total_rows + (size_t)total_rows / 8 + 1024;
                scan_temp_bytes = 0;
                cudaGetLastError();
                if ((e = cudaMalloc((void **)&d_rowcnt, rows_cap * sizeof(long long))) != cudaSuccess ||
                    (e = cudaMalloc((void **)&d_gprefix, rows_cap * sizeof(long long))) != cudaSuccess ||
                    (e = cudaMalloc((void **)&d_rowflag, rows_cap * sizeof(long long))) != cudaSuccess ||
                    (e = cudaMalloc((void **)&d_cflag, rows_cap * sizeof(long long))) != cudaSuccess ||
                    (e = cudaMalloc((void **)&d_cprefix, rows_cap * sizeof(long long))) != cudaSuccess ||
                    (e = cudaMalloc((void **)&d_crow, rows_cap * sizeof(int))) != cudaSuccess ||
                    (e = cub::DeviceScan::ExclusiveSum(nullptr, scan_temp_bytes, d_rowcnt, d_gprefix, (int)rows_cap, stream)) != cudaSuccess ||
                    (e = cudaMalloc(&d_scan_temp, scan_temp_bytes + 16)) != cudaSuccess) {
                    err = std::string("plan row tables: ") + cudaGetErrorString(e);
                    rows_cap = 0;
                    return 1;
                }
            }
            std::copy(cand.begin(), cand.end(), reinterpret_cast<PlanJobDesc *>(h_plan_desc));
            const PlanJobDesc *dj = reinterpret_cast<const PlanJobDesc *>(d_plan_desc);
            const double thr2 = thr * thr * (1.0 - 1e-12);  // pair_bound_passes
            cudaMemcpyAsync(d_plan_desc, h_plan_desc, (size_t)nc * sizeof(PlanJobDesc), cudaMemcpyHostToDevice, stream);
            cudaMemsetAsync(d_rowcnt + total_rows, 0, sizeof(long long), stream);
            cudaMemsetAsync(d_rowflag + total_rows, 0, sizeof(long long), stream);
            const unsigned wblocks = (unsigned)std::min<long long>((total_rows + 7) / 8, (long long)num_sms * 32);
            plan_count_kernel<<<wblocks, 256, 0, stream>>>(dj, nc, total_rows, d_md_sorted, thr2, d_rowcnt, d_rowflag);
            size_t tb = scan_temp_bytes;
            cub::DeviceScan::ExclusiveSum(d_scan_temp, tb, d_rowcnt, d_gprefix, (int)(total_rows + 1), stream);
            tb = scan_temp_bytes;
            cub::DeviceScan::ExclusiveSum(d_scan_temp, tb, d_rowflag, d_cflag, (int)(total_rows + 1), stream);
            plan_jobsum_kernel<<<(nc + 1 + 127) / 128, 128, 0, stream>>>(dj, nc, total_rows, d_gprefix, d_cflag, d_jobsum);
            cudaMemcpyAsync(h_jobsum, d_jobsum, (size_t)(nc + 1) * 2 * sizeof(long long), cudaMemcpyDeviceToHost, stream);
            if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) {
                err = std::string("plan row counts: ") + cudaGetErrorString(e);
                return 1;
            }
            const long long total_q = h_jobsum[2 * nc];
            if ((size_t)total_q + 1 > colidx_cap) {
                if (d_colidx) cudaFree(d_colidx);
                d_colidx = nullptr;
                colidx_cap = (size_t)total_q + (size_t)total_q / 16 + 1024;
                if ((e = cudaMalloc((void **)&d_colidx, colidx_cap * sizeof(int))) != cudaSuccess) {
                    err = std::string("plan column lists (") + std::to_string(colidx_cap * sizeof(int) >> 20) + " MiB): " + cudaGetErrorString(e);
                    colidx_cap = 0;
                    return 1;
                }
            }
            plan_compact_kernel<<<(unsigned)std::min<long long>((total_rows + 256) / 256, (long long)num_sms * 32), 256, 0, stream>>>(
                dj, nc, total_rows, d_rowcnt, d_gprefix, d_cflag, d_cprefix, d_crow);
            plan_fill_kernel<<<wblocks, 256, 0, stream>>>(dj, nc, total_rows, d_md_sorted, thr2, d_gprefix, d_colidx);
        }
    }
    std::vector<long long> table;
    std::vector<ClassJob> jobs;
    std::vector<size_t> job_table_off;
    plan.clear();
    plan_stats = ClassStats();
    // quartets of the (b, k) job: exact (CSR) or estimated from 64 sampled rows (staircase)
    auto job_quartets_estimate = [&](int b, int k) -> double {
        const int nb = (int)sd->sublists[b].pair_ids.size(), nk = (int)sd->sublists[k].pair_ids.size();
        if (csr) {
            const int ci = cand_of[(size_t)b * ns + k];
            return (double)(h_jobsum[2 * (ci + 1)] - h_jobsum[2 * ci]);
        }
        if (screen == FOCK_SCREEN_OFF) return (b == k) ? 0.5 * nb * (nb + 1.0) : (double)nb * nk;
        const int step = std::max(1, nb / 64);
        int cnt = 0;
        double q = 0.0;
        for (int r = 0; r < nb; r += step, ++cnt) {
            int nc = surviving_prefix(md[b][r], md[k], screen, thr);
            if (b == k) nc = std::min(nc, r + 1);
            q += nc;
        }
        return q * (double)nb / std::max(cnt, 1);
    };
    // class pairs in a fixed order; within one, every (bra sublist >= ket sublist) job
    for (int cb = 0; cb < kNumPairClasses; ++cb)
        for (int ck = 0; ck <= cb; ++ck) {
            ClassPlanLaunch pl;
            pl.cb = cb;
            pl.ck = ck;
            pl.job_begin = (int)jobs.size();
            long long tiles_launch = 0;
            // K-split threshold: one thread's primitive loop (class cost x depth, in instructions) may not
            // exceed a quarter of the average work per resident thread of this launch, nor ~6000
            // instructions when the launch is small -- otherwise it becomes the launch's critical path.
            // (rough instruction counts per primitive quartet; classes with a d shell scale with their flop model)
            static const int kCost[6][6] = {{80, 0, 0, 0, 0, 0},          {110, 200, 0, 0, 0, 0},          {220, 600, 1000, 0, 0, 0},
                                            {200, 400, 1500, 900, 0, 0}, {450, 1400, 4500, 3000, 9000, 0}, {1200, 3600, 12000, 8000, 24000, 60000}};
            const bool reg_kernel = cb < 3 || kPairClassLA[cb] + kPairClassLB[cb] + kPairClassLA[ck] + kPairClassLB[ck] <= 5;
            double launch_work = 0.0;
            for (int b = 0; b < ns; ++b) {
                const Sublist &sb = sd->sublists[b];
                if (sb.cls != cb || sb.pair_ids.empty()) continue;
                for (int k = 0; k <= b; ++k) {
                    const Sublist &sk = sd->sublists[k];
                    if (sk.cls != ck || sk.pair_ids.empty()) continue;
                    launch_work += job_quartets_estimate(b, k) * sb.K * sk.K * kCost[cb][ck];
                }
            }
            // (scaling the limit with 1/world was measured on an emulated 1/8 share: more K-split tiles, 8.9 -> 11.0 ms on config 5 -- not done)
            const double loop_limit = std::max(6000.0, 0.25 * launch_work / ((double)num_sms * 8 * 128));
            for (int b = 0; b < ns; ++b) {
                const Sublist &sb = sd->sublists[b];
                if (sb.cls != cb || sb.pair_ids.empty()) continue;
                for (int k = 0; k <= b; ++k) {
                    const Sublist &sk = sd->sublists[k];
                    if (sk.cls != ck || sk.pair_ids.empty()) continue;
                    const int nb = (int)sb.pair_ids.size(), nk = (int)sk.pair_ids.size();
                    const size_t off = table.size();
                    long long quartets = 0;
                    int rows_listed = nb;
                    const int ci = csr ? cand_of[(size_t)b * ns + k] : -1;
                    if (csr) {
                        quartets = h_jobsum[2 * (ci + 1)] - h_jobsum[2 * ci];
                        rows_listed = (int)(h_jobsum[2 * (ci + 1) + 1] - h_jobsum[2 * ci + 1]);
                    } else {
                        table.resize(off + (size_t)nb + 1);
                        long long *prefix = table.data() + off;
                        const double *mdb = md[b].data(), *mdk = md[k].data();
                        const double thr2_slack = thr * thr * (1.0 - 1e-12);
                        // ncols(r) is non-increasing in r (both lists sorted descending): walk it down
                        int nc = (screen == FOCK_SCREEN_OFF) ? nk : surviving_prefix(md[b][0], md[k], screen, thr);
                        for (int r = 0; r < nb; ++r) {
                            if (screen != FOCK_SCREEN_OFF)
                                while (nc > 0 && !pair_bound_passes(mdb[r], mdk[nc - 1], thr2_slack)) --nc;
                            prefix[r] = quartets;
                            quartets += (b == k) ? std::min(nc, r + 1) : nc;
                        }
                        prefix[nb] = quartets;
                    }
                    // K-split: keep one thread's loop below ~6000 instructions (class cost x depth)
                    static const int kSplitMax = getenv("FOCK_KSPLIT_MAX") ? atoi(getenv("FOCK_KSPLIT_MAX")) : 32;
                    int S = 1;
                    while (reg_kernel && S < kSplitMax && 2 * S <= sk.K && (double)sb.K * sk.K * kCost[cb][ck] / S > loop_limit) S *= 2;
                    const long long per_tile = 32 / S;
                    const long long tiles = (quartets + per_tile - 1) / per_tile;
                    plan_stats.shell_quartets_total += (b == k) ? (long long)nb * (nb + 1) / 2 : (long long)nb * nk;
                    plan_stats.shell_quartets += quartets;
                    plan_stats.prim_quartets += quartets * (long long)sb.K * sk.K;
                    plan_stats.model_flops += (double)quartets * sb.K * sk.K *
                                              model_flops_per_prim_quartet(kPairClassLA[cb], kPairClassLB[cb], kPairClassLA[ck], kPairClassLB[ck]);
                    plan_stats.warp_tiles += tiles;
                    if (tiles == 0) {
                        table.resize(off);
                        continue;
                    }
                    ClassJob j{};
                    j.tile_begin = tiles_launch;
                    j.bra = dev->d_bra + sb.bra_off;
                    j.ket = dev->d_ket + sk.ket_off;
                    j.bra_meta = dev->d_meta + sb.meta_off;
                    j.ket_meta = dev->d_meta + sk.meta_off;
                    j.Kab = sb.K;
                    j.Kcd = sk.K;
                    j.nb = rows_listed;
                    j.nk_pad = sk.npad;
                    j.same_sublist = (b == k) ? 1 : 0;
                    j.ksplit = S;
                    if (csr) {
                        j.q_base = h_jobsum[2 * ci];
                        j.row_prefix = d_cprefix + h_jobsum[2 * ci + 1];
                        j.row_map = d_crow + h_jobsum[2 * ci + 1];
                        j.col_idx = d_colidx + j.q_base;
                    }
                    jobs.push_back(j);
                    job_table_off.push_back(off);
                    tiles_launch += tiles;
                }
            }
            // heaviest tiles first (longest per-lane primitive loop = Kab*Kcd/ksplit): the launch then ends on its cheapest
            // tiles, which is what bounds the tail of a persistent grid -- most of all when the tile range is 1/8 of the job
            {
                static const bool kLpt = !getenv("FOCK_NO_LPT");
                const int j0 = pl.job_begin, nj = (int)jobs.size() - j0;
                if (kLpt && nj > 1) {
                    std::vector<int> order(nj);
                    std::iota(order.begin(), order.end(), 0);
                    std::vector<long long> ntiles(nj);
                    for (int i = 0; i < nj; ++i)
                        ntiles[i] = (i + 1 < nj ? jobs[j0 + i + 1].tile_begin : tiles_launch) - jobs[j0 + i].tile_begin;
                    auto cost = [&](int i) { return (double)jobs[j0 + i].Kab * jobs[j0 + i].Kcd / jobs[j0 + i].ksplit; };
                    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
                    std::vector<ClassJob> sorted_jobs(nj);
                    std::vector<size_t> sorted_off(nj);
                    long long tb = 0;
                    for (int i = 0; i < nj; ++i) {
                        sorted_jobs[i] = jobs[j0 + order[i]];
                        sorted_off[i] = job_table_off[j0 + order[i]];
                        sorted_jobs[i].tile_begin = tb;
                        tb += ntiles[order[i]];
                    }
                    std::copy(sorted_jobs.begin(), sorted_jobs.end(), jobs.begin() + j0);
                    std::copy(sorted_off.begin(), sorted_off.end(), job_table_off.begin() + j0);
                }
            }
            pl.njobs = (int)jobs.size() - pl.job_begin;
            pl.total_tiles = tiles_launch;
            if (pl.njobs > 0) plan.push_back(pl);
        }
    if (table.size() + 1 > table_cap) {
        if (d_table) cudaFree(d_table);
        if (h_table) cudaFreeHost(h_table);
        d_table = nullptr;
        h_table = nullptr;
        table_cap = table.size() + table.size() / 4 + 1024;
        if ((e = cudaMalloc((void **)&d_table, table_cap * sizeof(long long))) != cudaSuccess ||
            (e = cudaMallocHost((void **)&h_table, table_cap * sizeof(long long))) != cudaSuccess) {
            err = std::string("tile table: ") + cudaGetErrorString(e);
            table_cap = 0;
            return 1;
        }
    }
    if (jobs.size() + 1 > jobs_cap) {
        if (d_jobs) cudaFree(d_jobs);
        if (h_jobs) cudaFreeHost(h_jobs);
        d_jobs = nullptr;
        h_jobs = nullptr;
        jobs_cap = jobs.size() + 64;
        if ((e = cudaMalloc((void **)&d_jobs, jobs_cap * sizeof(ClassJob))) != cudaSuccess ||
            (e = cudaMallocHost((void **)&h_jobs, jobs_cap * sizeof(ClassJob))) != cudaSuccess) {
            err = std::string("job table: ") + cudaGetErrorString(e);
            jobs_cap = 0;
            return 1;
        }
    }
    if (!csr)
        for (size_t i = 0; i < jobs.size(); ++i) jobs[i].row_prefix = d_table + job_table_off[i];
    // pinned staging: the copies are truly asynchronous and ordered before the launches on `stream`
    std::copy(table.begin(), table.end(), h_table);
    std::copy(jobs.begin(), jobs.end(), h_jobs);
    cudaMemcpyAsync(d_table, h_table, table.size() * sizeof(long long), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(d_jobs, h_jobs, jobs.size() * sizeof(ClassJob), cudaMemcpyHostToDevice, stream);
    // per-tile descriptors of every launch (one pass on the device; the plan is reused by every G build of a geometry)
    {
        long long all_tiles = 0;
        for (ClassPlanLaunch &pl : plan) {
            pl.tile_base = all_tiles;
            all_tiles += pl.total_tiles;
        }
        if ((size_t)all_tiles + 1 > tile_tab_cap) {
            if (d_tile_tab) cudaFree(d_tile_tab);
            d_tile_tab = nullptr;
            tile_tab_cap = (size_t)all_tiles + (size_t)all_tiles / 8 + 1024;
            if ((e = cudaMalloc((void **)&d_tile_tab, tile_tab_cap * sizeof(int2))) != cudaSuccess) {
                err = std::string("tile descriptor table: ") + cudaGetErrorString(e);
                tile_tab_cap = 0;
                return 1;
            }
        }
        for (const ClassPlanLaunch &pl : plan) {
            if (pl.total_tiles <= 0) continue;
            const unsigned blocks = (unsigned)std::min<long long>((pl.total_tiles + 255) / 256, (long long)num_sms * 16);
            tile_table_kernel<<<blocks, 256, 0, stream>>>(d_jobs + pl.job_begin, pl.njobs, pl.total_tiles, d_tile_tab + pl.tile_base);
        }
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("plan upload: ") + cudaGetErrorString(e);
        return 1;
    }
    plan_valid = true;
    plan_screen = screen;
    plan_thr = thr;
    plan_rank = rank;
    plan_world = world;
    return 0;
}

template <class Sink>
int ClassEngine::run_plan(const Sink &sink, std::string &err) {
    int launches = 0;
    // The class launches are independent (their sinks commute), so they are forked onto auxiliary
    // streams, heaviest class first, and joined back: the long tail of one overlaps the next.
    if (!ev_fork) {
        cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
        for (int i = 0; i < kAuxStreams; ++i) {
            cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming);
        }
    }
    static const int kStreams = getenv("FOCK_STREAMS") ? std::max(1, std::min(kAuxStreams, atoi(getenv("FOCK_STREAMS")))) : kAuxStreams;
    static const int kBlocksPerSm = getenv("FOCK_BLOCKS_PER_SM") ? atoi(getenv("FOCK_BLOCKS_PER_SM")) : 16;
    if (kDynTilesEnabled)
        cudaMemsetAsync(d_tile_counters, 0, (size_t)kNumPairClasses * kNumPairClasses * kTileCounterSlots * sizeof(unsigned long long), stream);
    cudaEventRecord(ev_fork, stream);
    for (int i = 0; i < kAuxStreams; ++i) cudaStreamWaitEvent(aux[i], ev_fork, 0);
    int slot = 0;
    StreamRing ring;
    ring.n = kStreams;
    for (int i = 0; i < kStreams; ++i) ring.s[i] = aux[i];
    static const bool kSpreadSlices = !getenv("FOCK_NO_SLICE_SPREAD");
    for (int pi = (int)plan.size() - 1; pi >= 0; --pi) {  // plan is ordered light -> heavy
        const ClassPlanLaunch &pl = plan[pi];
        ClassLaunch cl{};
        cl.ring = kSpreadSlices ? &ring : nullptr;
        cl.jobs = d_jobs + pl.job_begin;
        cl.njobs = pl.njobs;
        cl.total_tiles = pl.total_tiles;
        cl.shell_bf = dev->d_shell_bf;
        cl.shell_scale = d_shell_scale;
        cl.fix_r = (mode & FOCK_MODE_FIX_R_BIT) ? 1 : 0;
        cl.boys_poly = boys_poly;
        cl.boys_ni = boys_ni;
        cl.tile_tab = d_tile_tab + pl.tile_base;
        cl.tile_counter = d_tile_counters + (size_t)(pl.cb * kNumPairClasses + pl.ck) * kTileCounterSlots;
#ifdef FOCK_NEAR_STATS
        {
            unsigned long long z[8] = {};
            cudaDeviceSynchronize();
            cudaMemcpyToSymbol(g_near_stats, z, sizeof(z));
        }
#endif
        cl.rank = plan_rank;
        cl.world = plan_world;
        const long long mine = (pl.total_tiles - plan_rank + plan_world - 1) / plan_world;  // warp tiles of this rank
        if (mine <= 0) continue;
        // persistent blocks: enough to fill the machine, each strides over the warp tiles
        // every block stages its own copy of the Boys tables (17-34 KB): a launch with few tiles gets fewer blocks, at least
        // kTilesPerWarp tiles per warp once there is a block per SM (FOCK_TILES_PER_WARP; measured 1 / 2 / 4 / 8: 2 is 1-4 % ahead on
        // (H2O)32/6-31G** and on an emulated 1/8 share, 4 and 8 lose -- profiles/r02/variants_call27.log)
        static const long long kTilesPerWarp = getenv("FOCK_TILES_PER_WARP") ? std::max(1, atoi(getenv("FOCK_TILES_PER_WARP"))) : 2;
        const long long full = (long long)num_sms * std::max(1, kBlocksPerSm * 128 / kClassThreads);
        long long nblocks = (mine + kClassWarps - 1) / kClassWarps;
        if (kTilesPerWarp > 1 && nblocks > num_sms) nblocks = std::max<long long>(num_sms, (mine + kClassWarps * kTilesPerWarp - 1) / (kClassWarps * kTilesPerWarp));
        dim3 grid((unsigned)std::min<long long>(nblocks, full));  // kBlocksPerSm counts 128-thread blocks
        // (FOCK_NO_HI_INCORE=1: in-core builds of the L >= 6 class pairs through the warp-per-quartet fallback, as in r01 -- A/B switch)
        const bool kNoHiIncore = getenv("FOCK_NO_HI_INCORE") != nullptr;  // read per build, so one process can compare both
        const int l_total = kPairClassLA[pl.cb] + kPairClassLB[pl.cb] + kPairClassLA[pl.ck] + kPairClassLB[pl.ck];
        int l = (kNoHiIncore && std::is_same<Sink, TensorSink>::value && l_total >= 6)
                    ? 0
                    : launch_class(pl.cb, pl.ck, cl, sink, mode, grid, aux[slot % kStreams]);
        if (l == 0) {
            // no register-resident kernel for this class pair (total angular momentum >= 6): one warp per
            // quartet with shared-memory tables, driven by the same job / tile tables
            if (!sq_shells) {
                err = "class pair without a kernel and no shell-quartet tables";
                return -1;
            }
            const SqEnv env{sq_shells, sq_exps, sq_coefs, sq_boys, mode, sq_tab};
            const unsigned wgrid = (unsigned)std::min<long long>((mine + kSqWarps - 1) / kSqWarps, (long long)num_sms * 16);
            eri_class_warpq_kernel<Sink><<<wgrid, kSqThreads, (size_t)kSqWarps * kSqSmemPerWarp * sizeof(double), aux[slot % kStreams]>>>(cl, env, sink);
            l = 1;
        }
        launches += l;
        ++slot;
#ifdef FOCK_NEAR_STATS
        if (pl.cb < 3) {
            unsigned long long z[8];
            cudaDeviceSynchronize();
            cudaMemcpyFromSymbol(z, g_near_stats, sizeof(z));
            fprintf(stderr, "NEAR_STATS class(%d|%d) slices %d tiles %llu lanes_active %llu lanes_with_near %llu near_prim %llu prim %llu rounds32 %llu | K>=8: lanes_with_near %llu near_prim %llu\n",
                    pl.cb, pl.ck, l, z[0], z[5], z[1], z[2], z[3], z[4], z[6], z[7]);
        }
#endif
    }
    for (int i = 0; i < kAuxStreams; ++i) {
        cudaEventRecord(ev_join[i], aux[i]);
        cudaStreamWaitEvent(stream, ev_join[i], 0);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        err = std::string("class kernels: ") + cudaGetErrorString(e);
        return -1;
    }
    return launches;
}

template int ClassEngine::run_plan<TensorSink>(const TensorSink &, std::string &);
template int ClassEngine::run_plan<BlockDirectSink>(const BlockDirectSink &, std::string &);

ClassEngine::~ClassEngine() {
    for (void *q : {(void *)d_md_sorted, (void *)d_morton, (void *)d_sortkeys, d_plan_desc, (void *)d_rowcnt, (void *)d_gprefix, (void *)d_rowflag, (void *)d_cflag,
                    (void *)d_cprefix, (void *)d_crow, (void *)d_colidx, (void *)d_jobsum, d_scan_temp})
        if (q) cudaFree(q);
    if (h_plan_desc) cudaFreeHost(h_plan_desc);
    if (h_jobsum) cudaFreeHost(h_jobsum);
    if (d_table) cudaFree(d_table);
    if (d_tile_tab) cudaFree(d_tile_tab);
    if (d_tile_counters) cudaFree(d_tile_counters);
    if (d_permute_desc) cudaFree(d_permute_desc);
    if (d_shell_scale) cudaFree(d_shell_scale);
    if (d_jobs) cudaFree(d_jobs);
    if (d_bra_nat) cudaFree(d_bra_nat);
    if (d_meta_nat) cudaFree(d_meta_nat);
    if (d_perm) cudaFree(d_perm);
    if (h_perm) cudaFreeHost(h_perm);
    if (d_maxdiag) cudaFree(d_maxdiag);
    if (h_maxdiag) cudaFreeHost(h_maxdiag);
    if (d_keys_sorted) cudaFree(d_keys_sorted);
    if (h_keys_sorted) cudaFreeHost(h_keys_sorted);
    if (d_iota) cudaFree(d_iota);
    if (d_seg) cudaFree(d_seg);
    if (d_sort_temp) cudaFree(d_sort_temp);
    if (h_table) cudaFreeHost(h_table);
    if (h_jobs) cudaFreeHost(h_jobs);
    if (ev_fork) {
        cudaEventDestroy(ev_fork);
        for (int i = 0; i < kAuxStreams; ++i) {
            cudaStreamDestroy(aux[i]);
            cudaEventDestroy(ev_join[i]);
        }
    }
}

}  // namespace fock
