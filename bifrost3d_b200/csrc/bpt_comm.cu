// Multi-GPU combine of the accumulation buffers: one NCCL reduce over NVLink (every GPU renders a disjoint
// range of sample indices of the same pixels into its own fp64 sum buffer; sum + count add up, mean = sum / count on resolve).
// The reference is single device (Renderer.cpp:289-291); this is the one collective the sharded path needs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded into the process - e.g. torch's - or the system
// one), so libbpt.so has no link-time dependency on it and single-GPU hosts never load it. The C ABI stays free of NCCL
// types: the 128-byte unique id travels as plain bytes and the host distributes it however it likes (MPI, a socket, a file,
// torch.distributed's store).
#include "bpt_context.h"
#include "../../include/bpt_c_api.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

namespace bpt {
namespace {

// The slice of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 bytes, ncclFloat64 = 8, ncclSum = 0).
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_SUCCESS = 0;

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*CommGetAsyncError)(ncclComm_t, int*) = nullptr; // optional: absent from very old NCCL builds
    std::string error;
};

Nccl* nccl() {
    static Nccl n;
    if (n.handle || !n.error.empty()) return &n;
    const char* override_path = getenv("BPT_NCCL_LIB");
    const char* candidates[] = { override_path, "libnccl.so.2", "libnccl.so" };
    for (const char* name : candidates) {
        if (!name) continue;
        n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) break;
    }
    if (!n.handle) { n.error = std::string("libnccl.so.2 not found (set BPT_NCCL_LIB): ") + dlerror(); return &n; }
    auto sym = [&](const char* name) { void* p = dlsym(n.handle, name); if (!p) n.error = std::string("missing NCCL symbol ") + name; return p; };
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.Reduce = reinterpret_cast<decltype(n.Reduce)>(sym("ncclReduce"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    n.CommGetAsyncError = reinterpret_cast<decltype(n.CommGetAsyncError)>(dlsym(n.handle, "ncclCommGetAsyncError"));
    return &n;
}

int nccl_fail(Context* ctx, const char* what, int status) {
    Nccl* n = nccl();
    return ctx->fail(BPT_ERROR_CUDA, std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(status) : "NCCL error"));
}

} // namespace
} // namespace bpt

using namespace bpt;

extern "C" {

int bpt_comm_unique_id(char out_id[BPT_COMM_ID_BYTES]) {
    Nccl* n = nccl();
    if (!n->error.empty() || !out_id) return BPT_ERROR_NOT_READY;
    NcclUniqueId id;
    if (n->GetUniqueId(&id) != NCCL_SUCCESS) return BPT_ERROR_CUDA;
    static_assert(sizeof(id) == BPT_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(out_id, &id, sizeof(id));
    return BPT_OK;
}

int bpt_comm_init(bpt_ctx* c, const char id[BPT_COMM_ID_BYTES], int rank_count, int rank) {
    Context* ctx = as_context(c);
    if (!id || rank_count < 1 || rank < 0 || rank >= rank_count) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_comm_init: bad arguments");
    Nccl* n = nccl();
    if (!n->error.empty()) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_comm_init: " + n->error);
    if (ctx->comm) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_comm_init: the context already has a communicator (bpt_comm_destroy first)");
    cudaSetDevice(ctx->device);
    NcclUniqueId unique_id;
    memcpy(&unique_id, id, sizeof(unique_id));
    ncclComm_t comm = nullptr;
    int status = n->CommInitRank(&comm, rank_count, unique_id, rank);
    if (status != NCCL_SUCCESS) return nccl_fail(ctx, "ncclCommInitRank", status);
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_rank_count = rank_count;
    return BPT_OK;
}

int bpt_comm_destroy(bpt_ctx* c) {
    Context* ctx = as_context(c);
    if (!ctx->comm) return BPT_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    int status = nccl()->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_rank_count = 1;
    return status == NCCL_SUCCESS ? BPT_OK : nccl_fail(ctx, "ncclCommDestroy", status);
}

int bpt_comm_check(bpt_ctx* c) {
    Context* ctx = as_context(c);
    if (!ctx->comm) return BPT_OK;
    Nccl* n = nccl();
    if (!n->CommGetAsyncError) return BPT_OK;
    int async_status = NCCL_SUCCESS;
    int status = n->CommGetAsyncError(static_cast<ncclComm_t>(ctx->comm), &async_status);
    if (status != NCCL_SUCCESS) return nccl_fail(ctx, "ncclCommGetAsyncError", status);
    if (async_status != NCCL_SUCCESS) return nccl_fail(ctx, "the communicator reports an asynchronous error", async_status);
    return BPT_OK;
}

int bpt_reduce_accumulation(bpt_ctx* c, int root) {
    Context* ctx = as_context(c);
    if (!ctx->comm) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_reduce_accumulation: call bpt_comm_init first");
    if (root >= ctx->comm_rank_count) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_reduce_accumulation: root out of range");
    if (!ctx->accumulation.ptr || ctx->width <= 0) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_reduce_accumulation: nothing rendered");
    cudaSetDevice(ctx->device);
    // In place, on the render stream: ordered after the samples already enqueued, no host synchronisation.
    const size_t count = (size_t)4 * ctx->width * ctx->height;
    Nccl* n = nccl();
    int status = root < 0 ? n->AllReduce(ctx->accumulation.ptr, ctx->accumulation.ptr, count, NCCL_FLOAT64, NCCL_SUM, static_cast<ncclComm_t>(ctx->comm), ctx->stream)
                          : n->Reduce(ctx->accumulation.ptr, ctx->accumulation.ptr, count, NCCL_FLOAT64, NCCL_SUM, root, static_cast<ncclComm_t>(ctx->comm), ctx->stream);
    if (status != NCCL_SUCCESS) return nccl_fail(ctx, root < 0 ? "ncclAllReduce" : "ncclReduce", status);
    return BPT_OK;
}

} // extern "C"
