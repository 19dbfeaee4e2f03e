// dclass_ds_tensor.cu -- (ds|..) class kernels with the TensorSink epilogue; see dclass_dispatch.cuh.
#define FOCK_DCLASS_DS
#include "dclass_dispatch.cuh"

namespace fock {
template int launch_dclass_ds<TensorSink>(int, const ClassLaunch &, const TensorSink &, dim3, cudaStream_t);
}
