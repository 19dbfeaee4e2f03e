// dclass_ds_block.cu -- (ds|..) class kernels with the BlockDirectSink epilogue; see dclass_dispatch.cuh.
#define FOCK_DCLASS_DS
#include "dclass_dispatch.cuh"

namespace fock {
template int launch_dclass_ds<BlockDirectSink>(int, const ClassLaunch &, const BlockDirectSink &, dim3, cudaStream_t);
}
