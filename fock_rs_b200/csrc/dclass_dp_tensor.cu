// dclass_dp_tensor.cu -- (dp|..) class kernels with the TensorSink epilogue; see dclass_dispatch.cuh.
#define FOCK_DCLASS_DP
#include "dclass_dispatch.cuh"

namespace fock {
template int launch_dclass_dp<TensorSink>(int, const ClassLaunch &, const TensorSink &, dim3, cudaStream_t);
}
