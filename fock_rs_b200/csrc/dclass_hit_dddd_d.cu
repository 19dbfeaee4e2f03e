// dclass_hit_dddd_d.cu -- register-resident slice kernels of a class pair with total angular momentum >= 6
// (in-core sink: TensorSink); see dclass_dispatch.cuh.
#include "dclass_dispatch.cuh"

namespace fock {
int launch_hi_dddd_d(const ClassLaunch &cl, const TensorSink &sink, dim3 grid, cudaStream_t st) {
    return launch_ket_range<2, 2, 2, 2, 18, 18, 18, 36>(cl, sink, grid, st);
}
}  // namespace fock
