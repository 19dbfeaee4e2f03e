// dclass_hit_dddp_a.cu -- register-resident slice kernels of a class pair with total angular momentum >= 6
// (in-core sink: TensorSink); see dclass_dispatch.cuh.
#include "dclass_dispatch.cuh"

namespace fock {
int launch_hi_dddp_a(const ClassLaunch &cl, const TensorSink &sink, dim3 grid, cudaStream_t st) {
    return launch_ket_range<2, 2, 2, 1, 0, 18, 0, 18>(cl, sink, grid, st);
}
}  // namespace fock
