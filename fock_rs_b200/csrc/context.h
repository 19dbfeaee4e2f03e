// context.h -- the opaque fock_ctx behind include/fock_b200.h and small host utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/fock_b200.h"
#include "device_math.cuh"
#include "eri_generic.cuh"
#include "shells.h"

enum { EV_A = 0, EV_B, EV_C, EV_D, EV_E, EV_F, EV_ORB0, EV_ORB1, EV_COUNT };  // EV_ORB*: the orbital path's own pair (fock_timings reads EV_E/EV_F as the last G build)

struct fock_ctx {
    fock_opts opts{};
    int N = 0;
    size_t npair = 0;  // N(N+1)/2
    size_t nelem = 0;  // npair(npair+1)/2
    int nprim = 0;

    // host copy of the basis (SoA)
    std::vector<double> h_center, h_exp, h_coef;
    std::vector<int32_t> h_angular, h_primoff;

    // device basis
    double *d_center = nullptr, *d_exp = nullptr, *d_coef = nullptr;
    int *d_angular = nullptr, *d_primoff = nullptr;
    double *d_boys_lut = nullptr, *d_boys_tab = nullptr;

    // tensors
    double *d_diag = nullptr;  // [npair] (xy|xy)
    double *d_V = nullptr;     // [nelem] raw integrals (in-core, keep_eri)
    double *d_P = nullptr;     // [nelem] supermatrix (ij|kl) - 1/4 (ik|jl) - 1/4 (il|jk)
    double *d_dvec = nullptr;  // [npair] packed density
    double *d_yvec = nullptr;  // [npair] packed result
    double *d_Ds = nullptr, *d_A = nullptr;  // [N*N] symmetrised density / block-digest accumulator (direct mode)
    double *d_D = nullptr, *d_G = nullptr;  // [N*N] staging for the host-buffer entry point
    double *h_pinned = nullptr;              // [2*N*N] pinned staging

    // shell / pair machinery for the class kernels
    fock::ShellData shells;
    fock::PairTablesDev pairs;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[EV_COUNT]{};
    double timings[8]{};
    int64_t stats[16]{};
    size_t digest_begin = 0, digest_end = 0;
    double orbital_ms = 0.0;  // last fock_orbital_voxelize kernel time
    int orbital_functions = 0, orbital_groups = 0;
    bool built = false;
    int device = 0;
    void *comm = nullptr;      // ncclComm_t when fock_comm_init attached one (fock_comm.cu)
    double det_scale = 0.0;    // deterministic mode: fixed-point scale of the last direct G build (a power of two)
    double det_bound_jk = 0.0; // deterministic mode: bound of |A| per unit of max|D|, from the diagonal (fock_eri_build)
    double *d_det = nullptr;   // deterministic mode: [2] device scalars (max |D| as bits, scratch)
    std::string err;

    fock::BasisDev basis_dev() const {
        return fock::BasisDev{N, d_center, d_angular, d_primoff, d_exp, d_coef};
    }
    fock::BoysTables boys() const { return fock::BoysTables{d_boys_lut, d_boys_tab}; }
};

// fock_comm.cu: in-place sum of `buf` over the attached communicator's ranks on `st` (no-op without one);
// `as_int64`: the buffer holds 64-bit fixed-point integers (deterministic mode)
int fock_comm_allreduce(fock_ctx *ctx, void *buf, size_t n, bool as_int64, cudaStream_t st);

#define FOCK_CUDA_TRY(ctx, expr)                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                             \
            return (_e == cudaErrorMemoryAllocation) ? FOCK_ERR_OOM : FOCK_ERR_CUDA;                     \
        }                                                                                                \
    } while (0)

#define FOCK_TRY(expr)                   \
    do {                                 \
        int _s = (expr);                 \
        if (_s != FOCK_OK) return _s;    \
    } while (0)
