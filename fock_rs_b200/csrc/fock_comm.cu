// fock_comm.cu -- the one exchange step of the path behind the C ABI: an all-reduce (sum) of the N x N partial G
// over NCCL (NVLink 5 / NVSwitch), issued on the context's stream right after the kernels that produce the partial.
//
// The reference has no counterpart (single process, CPU); SURVEY.md section 8(b) asks that `fock_build_g` return the
// COMPLETE G -- what scf/src/lib.rs:177-185 computes -- however many GPUs the job is sharded over, so a Rust host
// needs no NCCL binding of its own: rank 0 calls fock_comm_unique_id, ships the 128 bytes to the other ranks by
// whatever means it has (MPI, a socket, a file), every rank calls fock_comm_init.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2) so that the library loads, and single-GPU use works, on a
// machine without NCCL; in a process that already loaded an NCCL (torch bundles one) the loader hands back that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "context.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string err;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("FOCK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
            return;
        }
        auto sym = [&](const char *s) { return dlsym(api.handle, s); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString)
            api.err = "libnccl lacks a required symbol";
    });
    return &api;
}

thread_local std::string g_comm_error;

}  // namespace

// used by fock_build_g[_device]: sums `buf` (n doubles, device) over the ranks in place, on `st`
int fock_comm_allreduce(fock_ctx *ctx, void *buf, size_t n, bool as_int64, cudaStream_t st) {
    if (!ctx->comm) return FOCK_OK;
    NcclApi *api = nccl_api();
    const ncclResult_t r = api->AllReduce(buf, buf, n, as_int64 ? ncclInt64 : ncclDouble, ncclSum, (ncclComm_t)ctx->comm, st);
    if (r != ncclSuccess) {
        ctx->err = std::string("ncclAllReduce: ") + api->GetErrorString(r);
        return FOCK_ERR_CUDA;
    }
    return FOCK_OK;
}

extern "C" {

int fock_comm_unique_id(void *id_out) {
    if (!id_out) return FOCK_ERR_BAD_ARG;
    NcclApi *api = nccl_api();
    if (!api->err.empty()) {
        g_comm_error = api->err;
        return FOCK_ERR_UNSUPPORTED;
    }
    ncclUniqueId id;
    const ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) {
        g_comm_error = std::string("ncclGetUniqueId: ") + api->GetErrorString(r);
        return FOCK_ERR_CUDA;
    }
    static_assert(sizeof(id) == FOCK_COMM_ID_BYTES, "NCCL unique id size");
    memcpy(id_out, &id, sizeof(id));
    return FOCK_OK;
}

int fock_comm_init(fock_ctx *ctx, const void *id_bytes) {
    if (!ctx || !id_bytes) return FOCK_ERR_BAD_ARG;
    if (ctx->comm) {
        ctx->err = "a communicator is already attached";
        return FOCK_ERR_STATE;
    }
    if (ctx->opts.world <= 1) return FOCK_OK;  // nothing to exchange
    NcclApi *api = nccl_api();
    if (!api->err.empty()) {
        ctx->err = api->err;
        return FOCK_ERR_UNSUPPORTED;
    }
    FOCK_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = api->CommInitRank(&comm, ctx->opts.world, id, ctx->opts.rank);
    if (r != ncclSuccess) {
        ctx->err = std::string("ncclCommInitRank: ") + api->GetErrorString(r);
        return FOCK_ERR_CUDA;
    }
    ctx->comm = comm;
    return FOCK_OK;
}

int fock_comm_destroy(fock_ctx *ctx) {
    if (!ctx) return FOCK_ERR_BAD_ARG;
    if (!ctx->comm) return FOCK_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    nccl_api()->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    return FOCK_OK;
}

int fock_comm_attached(fock_ctx *ctx) { return ctx && ctx->comm ? 1 : 0; }

const char *fock_comm_last_error(void) { return g_comm_error.c_str(); }

}  // extern "C"
