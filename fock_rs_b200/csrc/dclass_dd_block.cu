// dclass_dd_block.cu -- (dd|..) class kernels with the BlockDirectSink epilogue; see dclass_dispatch.cuh.
#define FOCK_DCLASS_DD
#include "dclass_dispatch.cuh"

namespace fock {
template int launch_dclass_dd<BlockDirectSink>(int, const ClassLaunch &, const BlockDirectSink &, dim3, cudaStream_t);
}
