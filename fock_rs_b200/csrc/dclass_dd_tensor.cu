// dclass_dd_tensor.cu -- (dd|..) class kernels with the TensorSink epilogue; see dclass_dispatch.cuh.
#define FOCK_DCLASS_DD
#include "dclass_dispatch.cuh"

namespace fock {
template int launch_dclass_dd<TensorSink>(int, const ClassLaunch &, const TensorSink &, dim3, cudaStream_t);
}
