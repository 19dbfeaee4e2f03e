// class_launch.h -- host driver of the shell-class ERI kernels (see class_launch.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "device_math.cuh"
#include "shells.h"

namespace fock {

struct StoreSink;
struct ClassLaunch;

struct ClassStats {
    long long shell_quartets_total = 0;  // unique class shell quartets before screening
    long long shell_quartets = 0;        // evaluated (all ranks together)
    long long prim_quartets = 0;
    double model_flops = 0;
};

struct ClassEngine {
    ShellData *sd = nullptr;
    PairTablesDev *dev = nullptr;
    BoysTables boys{};
    cudaStream_t stream = nullptr;
    int mode = 0;
    int *d_table = nullptr;
    size_t table_cap = 0;

    int upload_tables(std::string &err);
    int run_diagonal(double *d_diag, std::string &err);  // returns launches or -1
    int sort_by_diagonal(const std::vector<double> &h_diag, std::string &err);
    int run_build(const StoreSink &sink, int screen, double thr, int rank, int world, ClassStats &stats,
                  std::string &err);  // returns launches or -1
    ClassLaunch base_launch(const Sublist &sb, const Sublist &sk) const;
    ~ClassEngine();
};

}  // namespace fock
