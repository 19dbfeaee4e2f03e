// dclass_dispatch.cuh -- launch tables of the register-resident class kernels whose bra pair holds a d
// shell.  Every (bra class, sink) pair is instantiated in its own translation unit (dclass_*.cu) so that
// the slices compile in parallel.  A class whose n_ab x n_cd component block exceeds ~18 accumulators
// is covered by several SLICE launches ([CB0,CB0+CBN) x [KC0,KC0+KCN)), each recomputing the Boys values
// and the R table of its primitive quartets (cheap next to the Hermite contractions of these classes).
// Every class pair up to (dd|dd) has such kernels for both sinks of the class engine (dclass_hi_*.cu / dclass_hit_*.cu hold the
// total-angular-momentum >= 6 ones); class_launch.cu falls back to the warp-per-quartet kernel of eri_shellq.cuh only when a
// launcher returns 0 (FOCK_NO_HI_INCORE, or a sink without instantiations).
#pragma once
#include <type_traits>

#include "class_launch.h"
#include "eri_class.cuh"

namespace fock {

template <int LA, int LB, int LC, int LD, int CB0, int CBN, int KC0, int KCN, class Sink>
inline void launch_slice(const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st) {
    const size_t smem = class_smem_bytes(LA, LB, LA + LB + LC + LD, cl.boys_ni);
    // FIXR = 2: the (t-1) factor of the R recursion is a run-time flag here (halves the instantiations)
    eri_class_kernel<LA, LB, LC, LD, 2, CB0, CBN, KC0, KCN, Sink><<<grid, kClassThreads, smem, cl.ring ? cl.ring->next() : st>>>(cl, sink);
}

// ket component slices [0, N) in steps of STEP, all bra components [CB0, CB0+CBN)
template <int LA, int LB, int LC, int LD, int CB0, int CBN, int STEP, int KC0, class Sink>
inline int launch_ket_slices(const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st) {
    constexpr int NCD = ncart_c(LC) * ncart_c(LD);
    if constexpr (KC0 < NCD) {
        launch_slice<LA, LB, LC, LD, CB0, CBN, KC0, STEP>(cl, sink, grid, st);
        return 1 + launch_ket_slices<LA, LB, LC, LD, CB0, CBN, STEP, KC0 + STEP>(cl, sink, grid, st);
    } else {
        return 0;
    }
}

// ket component slices [KC0, KCEND) one at a time, bra components [CB0, CB0+CBN)
template <int LA, int LB, int LC, int LD, int CB0, int CBN, int KC0, int KCEND, class Sink>
inline int launch_ket_range(const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st) {
    if constexpr (KC0 < KCEND) {
        launch_slice<LA, LB, LC, LD, CB0, CBN, KC0, 1>(cl, sink, grid, st);
        return 1 + launch_ket_range<LA, LB, LC, LD, CB0, CBN, KC0 + 1, KCEND>(cl, sink, grid, st);
    } else {
        return 0;
    }
}

// Class pairs with total angular momentum 6..8 (dclass_hi_*.cu: integral-direct sink, dclass_hit_*.cu: in-core sink): with ONE ket
// component per slice the compiler keeps only the sub-box of the R table that component reaches.  That fits
// the register file up to (dd|pp)/(dd|ds); the (dp|dp), (dd|dp) and (dd|dd) slices carry part of the table per thread in local
// memory (profiles/r01_v22_ptxas_class_kernels.md).
#define FOCK_DECLARE_HI(name)                                                                              \
    int name(const ClassLaunch &cl, const BlockDirectSink &sink, dim3 grid, cudaStream_t st);             \
    int name(const ClassLaunch &cl, const TensorSink &sink, dim3 grid, cudaStream_t st);
FOCK_DECLARE_HI(launch_hi_dpdp)
FOCK_DECLARE_HI(launch_hi_ddpp)
FOCK_DECLARE_HI(launch_hi_ddds)
FOCK_DECLARE_HI(launch_hi_dddp_a)
FOCK_DECLARE_HI(launch_hi_dddp_b)
FOCK_DECLARE_HI(launch_hi_dddd_a)
FOCK_DECLARE_HI(launch_hi_dddd_b)
FOCK_DECLARE_HI(launch_hi_dddd_c)
FOCK_DECLARE_HI(launch_hi_dddd_d)
#undef FOCK_DECLARE_HI
// the class engine runs plans for these two sinks only (class_launch.cu: run_plan)
template <class Sink>
constexpr bool kHasHiKernels = std::is_same<Sink, BlockDirectSink>::value || std::is_same<Sink, TensorSink>::value;

#ifdef FOCK_DCLASS_DS
template <class Sink>
int launch_dclass_ds(int ck, const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st) {
    switch (ck) {
        case 0: return launch_ket_slices<2, 0, 0, 0, 0, 6, 1, 0>(cl, sink, grid, st);  // (ds|ss)  6 x 1
        case 1: return launch_ket_slices<2, 0, 1, 0, 0, 6, 3, 0>(cl, sink, grid, st);  // (ds|ps)  6 x 3
        case 2: return launch_ket_slices<2, 0, 1, 1, 0, 6, 3, 0>(cl, sink, grid, st);  // (ds|pp)  6 x 9 in 3 slices
        case 3: return launch_ket_slices<2, 0, 2, 0, 0, 6, 3, 0>(cl, sink, grid, st);  // (ds|ds)  6 x 6 in 2 slices
        default: return 0;
    }
}
#endif

#ifdef FOCK_DCLASS_DP
template <class Sink>
int launch_dclass_dp(int ck, const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st) {
    switch (ck) {
        case 0: return launch_ket_slices<2, 1, 0, 0, 0, 18, 1, 0>(cl, sink, grid, st);  // (dp|ss) 18 x 1
        case 1: return launch_ket_slices<2, 1, 1, 0, 0, 18, 1, 0>(cl, sink, grid, st);  // (dp|ps) 18 x 3 in 3 slices
        case 2: return launch_ket_slices<2, 1, 1, 1, 0, 18, 1, 0>(cl, sink, grid, st);  // (dp|pp) 18 x 9 in 9 slices
        case 3: return launch_ket_slices<2, 1, 2, 0, 0, 18, 1, 0>(cl, sink, grid, st);  // (dp|ds) 18 x 6 in 6 slices
        case 4:  // (dp|dp): L = 6, 18 x 18 in 18 slices
            if constexpr (kHasHiKernels<Sink>) return launch_hi_dpdp(cl, sink, grid, st);
            return 0;
        default: return 0;
    }
}
#endif

#ifdef FOCK_DCLASS_DD
template <class Sink>
int launch_dclass_dd(int ck, const ClassLaunch &cl, const Sink &sink, dim3 grid, cudaStream_t st) {
    switch (ck) {
        case 0:  // (dd|ss) 36 x 1 in 2 bra slices
            return launch_ket_slices<2, 2, 0, 0, 0, 18, 1, 0>(cl, sink, grid, st) +
                   launch_ket_slices<2, 2, 0, 0, 18, 18, 1, 0>(cl, sink, grid, st);
        case 1:  // (dd|ps) 36 x 3 in 6 slices
            return launch_ket_slices<2, 2, 1, 0, 0, 18, 1, 0>(cl, sink, grid, st) +
                   launch_ket_slices<2, 2, 1, 0, 18, 18, 1, 0>(cl, sink, grid, st);
        default:  // (dd|pp), (dd|ds), (dd|dp), (dd|dd): L = 6, 6, 7, 8
            if constexpr (kHasHiKernels<Sink>) {
                if (ck == 2) return launch_hi_ddpp(cl, sink, grid, st);
                if (ck == 3) return launch_hi_ddds(cl, sink, grid, st);
                if (ck == 4) return launch_hi_dddp_a(cl, sink, grid, st) + launch_hi_dddp_b(cl, sink, grid, st);
                if (ck == 5)
                    return launch_hi_dddd_a(cl, sink, grid, st) + launch_hi_dddd_b(cl, sink, grid, st) +
                           launch_hi_dddd_c(cl, sink, grid, st) + launch_hi_dddd_d(cl, sink, grid, st);
            }
            return 0;
    }
}
#endif

}  // namespace fock
