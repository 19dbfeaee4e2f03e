// dclass_dp_block.cu -- (dp|..) class kernels with the BlockDirectSink epilogue; see dclass_dispatch.cuh.
#define FOCK_DCLASS_DP
#include "dclass_dispatch.cuh"

namespace fock {
template int launch_dclass_dp<BlockDirectSink>(int, const ClassLaunch &, const BlockDirectSink &, dim3, cudaStream_t);
}
