// class_launch.h -- host driver of the shell-class ERI kernels (see class_launch.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "device_math.cuh"
#include "shells.h"

namespace fock {

struct TensorSink;
struct DirectSink;
struct BlockDirectSink;
struct ClassJob;
struct ClassLaunch;
struct SqShell;
struct SqTables;

// Register-resident class kernels for bra pair classes with a d shell (ds, dp, dd), total angular
// momentum <= 5; one translation unit per (bra class, sink) so they compile in parallel
// (dclass_*.cu).  Return the number of launches, 0 when the class pair has no such kernel.
template <class Sink>
int launch_dclass_ds(int ck, const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st);
template <class Sink>
int launch_dclass_dp(int ck, const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st);
template <class Sink>
int launch_dclass_dd(int ck, const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st);

struct ClassStats {
    long long shell_quartets_total = 0;  // unique class shell quartets before screening
    long long shell_quartets = 0;        // evaluated (all ranks together)
    long long prim_quartets = 0;
    double model_flops = 0;
    long long warp_tiles = 0;
};

// One class-pair launch (all (bra sublist, ket sublist) jobs of one angular class pair).
struct ClassPlanLaunch {
    int cb = 0, ck = 0;       // pair classes 0 ss, 1 ps, 2 pp, 3 ds, 4 dp, 5 dd
    int job_begin = 0, njobs = 0;
    long long total_tiles = 0;
    long long tile_base = 0;  // first entry of this launch in the engine's tile table
};

// one entry per non-empty sublist (class x primitive-pair count); the table lives in device memory, so the number of
// distinct contraction depths of a basis is not limited
struct PermuteDesc {
    size_t bra_off, ket_off, meta_off;
    int first, n, K, nf, npad;
};

struct ClassEngine {
    ShellData *sd = nullptr;
    PairTablesDev *dev = nullptr;
    const double *boys_poly = nullptr;  // device, per-interval polynomial table of the active mode
    int boys_ni = 200;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    int mode = 0;

    // natural-order master tables (built once) and the permutation scratch
    double *d_bra_nat = nullptr;
    int *d_meta_nat = nullptr;
    int *d_perm = nullptr;
    unsigned long long *d_maxdiag = nullptr, *h_maxdiag = nullptr;  // per class pair, natural order (device / pinned)
    std::vector<std::vector<int>> nat_ids;  // per sublist: pair ids in natural order

    // plan (jobs + tile tables) of the current sorted order / screening
    long long *d_table = nullptr;
    size_t table_cap = 0;
    long long *h_table = nullptr;  // pinned staging
    size_t h_table_cap = 0;
    ClassJob *h_jobs = nullptr;    // pinned staging
    int *h_perm = nullptr;         // pinned staging
    unsigned long long *d_keys_sorted = nullptr, *h_keys_sorted = nullptr;  // maxima, sorted per sublist
    int *d_iota = nullptr, *d_seg = nullptr;
    void *d_sort_temp = nullptr;
    size_t sort_temp_bytes = 0;
    PermuteDesc *d_permute_desc = nullptr;
    int permute_count = 0;
    int permute_blocks = 0;
    double host_ms[4] = {0, 0, 0, 0};  // sort, plan
    ClassJob *d_jobs = nullptr;
    size_t jobs_cap = 0;
    // Pair order of the sublists.  Spatial (default): Morton order of the pair centres, so that the 32 ket pairs of a warp tile
    // lie close together and meet the near / far Boys branch together; the quartets that survive the cut are then a CSR matrix
    // built on the device per plan (row counts -> scans -> column lists).  FOCK_PAIR_ORDER=diag restores the r01 order (sorted by
    // the largest diagonal integral, survivors of a row are a prefix, no column lists).
    bool spatial_order = true;
    double *d_md_sorted = nullptr;           // per pair, sorted order: max |(xy|xy)|
    // sort keys of the spatial order: (magnitude bucket of the pair's largest diagonal integral, Morton code of its centre).
    // Pairs that never survive the cut sink to the end of their sublist and the survivors of a row are dense runs of whole
    // buckets (coalesced ket loads) plus one partially surviving bucket; inside a bucket neighbours in the list are neighbours
    // in space.  order_bucket_bits = log2 of the bucket width (FOCK_ORDER_BUCKET_BITS; 0 = pure Morton order, fixed at upload).
    unsigned long long *d_morton = nullptr, *d_sortkeys = nullptr;  // per pair, natural order
    int order_bucket_bits = 4;
    void *d_plan_desc = nullptr;             // PlanJobDesc[candidates]
    void *h_plan_desc = nullptr;
    size_t plan_desc_cap = 0;
    long long *d_rowcnt = nullptr, *d_gprefix = nullptr, *d_rowflag = nullptr, *d_cflag = nullptr;  // [rows + 1]
    size_t rows_cap = 0;
    long long *d_cprefix = nullptr;  // [non-empty rows + 1] quartet offset of every listed row
    int *d_crow = nullptr;           // [non-empty rows] bra pair of every listed row
    int *d_colidx = nullptr;         // [quartets] ket pair of every surviving quartet
    size_t colidx_cap = 0;
    long long *d_jobsum = nullptr, *h_jobsum = nullptr;  // [candidates + 1][2]: quartet / listed-row offsets at job starts
    size_t jobsum_cap = 0;
    void *d_scan_temp = nullptr;
    size_t scan_temp_bytes = 0;
    unsigned long long *d_tile_counters = nullptr;  // [class pairs][kTileCounterSlots], zeroed at the start of every run_plan
    int2 *d_tile_tab = nullptr;  // (job, first row) per warp tile of every launch of the plan
    size_t tile_tab_cap = 0;
    std::vector<ClassPlanLaunch> plan;
    ClassStats plan_stats;
    bool plan_valid = false;
    int plan_screen = -1, plan_rank = -1, plan_world = -1;
    double plan_thr = 0;

    static constexpr int kAuxStreams = 16;
    cudaStream_t aux[kAuxStreams] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kAuxStreams] = {};

    // warp-per-quartet fallback (classes without a register kernel): shell tables owned by SqState
    const SqShell *sq_shells = nullptr;
    const double *sq_exps = nullptr, *sq_coefs = nullptr;
    BoysTables sq_boys{};
    const SqTables *sq_tab = nullptr;
    double *d_shell_scale = nullptr;  // [nshell][6]

    int upload_master(std::string &err);                 // once, after build_pairs
    int run_diagonal(double *d_diag, std::string &err);  // s/p diagonal blocks; returns launches or -1
    // after every (xy|xy) is in d_diag: per-pair maxima of the d sublists, device sort, table permutation
    int finish_diagonal(const double *d_diag, std::string &err);
    int sort_by_diagonal(std::string &err);  // uses h_maxdiag; returns launches or -1
    int make_plan(int screen, double thr, int rank, int world, std::string &err);
    template <class Sink>
    int run_plan(const Sink &sink, std::string &err);  // returns launches or -1
    ~ClassEngine();
};

}  // namespace fock
