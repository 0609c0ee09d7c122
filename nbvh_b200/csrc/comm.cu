// comm.cu -- data-parallel plumbing inside the library (nbvh_allreduce_setup(ctx, ncclComm_t); 8(e): one process per
// GPU, scene and weights replicated, gradients summed by NCCL over NVLink / NVSwitch).
//
// The reference is single-GPU: its trainer steps on the gradients of its own batch (ext/tiny-cuda-nn trainer.h:163-190).  Here every
// rank computes the gradients of its share of the global batch, scaled by 1 / global batch (nbvh_training_forward_backward), and the
// library sums them in two buckets on its own communication stream: the MLP matrices as soon as k_train_mlp has flushed them -- that
// allreduce overlaps the hash-grid backward kernels -- and the grid table (+ the loss scalar) when those are done.  The optimiser step
// is replicated, so every rank holds bit-identical parameters afterwards (NCCL's allreduce returns the same bits on every rank).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already carries NCCL -- the Python host through torch -- shares
// that copy, a C++ host gets the system library, and single-GPU users need none.  Only the stable C ABI is used, declared here.
#include "host_context.hpp"

#include <dlfcn.h>

namespace {

struct NcclUniqueId {
    char internal[NBVH_NCCL_UNIQUE_ID_BYTES];
};
typedef void *NcclComm;
enum { kNcclSuccess = 0 };
enum { kNcclSum = 0 };
enum { kNcclInt32 = 2, kNcclUint32 = 3, kNcclUint64 = 5, kNcclFloat32 = 7, kNcclFloat64 = 8 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *)                                                          = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int)                                     = nullptr;
    int (*CommDestroy)(NcclComm)                                                                = nullptr;
    int (*CommCount)(NcclComm, int *)                                                           = nullptr;
    int (*CommUserRank)(NcclComm, int *)                                                        = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t)            = nullptr;
    const char *(*GetErrorString)(int)                                                          = nullptr;
    std::string error;
};

NcclApi &nccl_api()
{
    static NcclApi api;
    if (api.lib || !api.error.empty())
        return api;
    const char *names[] = {getenv("NBVH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n)
            continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.lib)
            break;
    }
    if (!api.lib) {
        api.error = std::string("NCCL is not loadable (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
        return api;
    }
    bool ok = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(api.lib, name);
        ok      = ok && p != nullptr;
        return p;
    };
    api.GetUniqueId    = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank   = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy    = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.CommCount      = reinterpret_cast<decltype(api.CommCount)>(sym("ncclCommCount"));
    api.CommUserRank   = reinterpret_cast<decltype(api.CommUserRank)>(sym("ncclCommUserRank"));
    api.AllReduce      = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
        api.error = "libnccl lacks a required symbol";
        dlclose(api.lib);
        api.lib = nullptr;
    }
    return api;
}

int nccl_fail(nbvh_ctx *ctx, const char *what, int code)
{
    NcclApi &a = nccl_api();
    return fail(ctx, NBVH_ERR_CUDA, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(code) : "NCCL error"));
}

int attach(nbvh_ctx *ctx, NcclComm comm, bool owned, int rank, int world)
{
    ctx->nccl_comm = comm, ctx->nccl_owned = owned, ctx->dp_rank = rank, ctx->dp_world = world;
    if (!ctx->comm_stream)
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&ctx->ev_bucket[0], &ctx->ev_bucket[1], &ctx->ev_reduced})
        if (!*e)
            CU_TRY(ctx, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return NBVH_OK;
}

} // namespace

namespace nbvh_host {

// One gradient bucket: `st` has produced grads[first, first + count) -> the communication stream sums it across the ranks.
int dp_allreduce_bucket(nbvh_ctx *ctx, int bucket, float *ptr, size_t count, cudaStream_t st)
{
    if (!ctx->nccl_comm || ctx->dp_world <= 1 || count == 0)
        return NBVH_OK;
    NcclApi &a = nccl_api();
    CU_TRY(ctx, cudaEventRecord(ctx->ev_bucket[bucket], st));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_bucket[bucket], 0));
    const int rc = a.AllReduce(ptr, ptr, count, kNcclFloat32, kNcclSum, ctx->nccl_comm, ctx->comm_stream);
    if (rc != kNcclSuccess)
        return nccl_fail(ctx, "ncclAllReduce", rc);
    return NBVH_OK;
}

// `st` continues only after every bucket handed to the communication stream has been reduced.
int dp_join(nbvh_ctx *ctx, cudaStream_t st)
{
    if (!ctx->nccl_comm || ctx->dp_world <= 1)
        return NBVH_OK;
    CU_TRY(ctx, cudaEventRecord(ctx->ev_reduced, ctx->comm_stream));
    CU_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_reduced, 0));
    return NBVH_OK;
}

void dp_release(nbvh_ctx *ctx)
{
    if (ctx->nccl_comm && ctx->nccl_owned && nccl_api().CommDestroy)
        nccl_api().CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr, ctx->nccl_owned = false, ctx->dp_rank = 0, ctx->dp_world = 1;
    if (ctx->comm_stream)
        cudaStreamDestroy(ctx->comm_stream);
    ctx->comm_stream = nullptr;
    for (cudaEvent_t *e : {&ctx->ev_bucket[0], &ctx->ev_bucket[1], &ctx->ev_reduced}) {
        if (*e)
            cudaEventDestroy(*e);
        *e = nullptr;
    }
}

} // namespace nbvh_host

extern "C" {

int nbvh_nccl_unique_id(void *id_out)
{
    if (!id_out)
        return NBVH_ERR_INVALID_ARG;
    NcclApi &a = nccl_api();
    if (!a.lib)
        return NBVH_ERR_STATE;
    NcclUniqueId id;
    if (a.GetUniqueId(&id) != kNcclSuccess)
        return NBVH_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return NBVH_OK;
}

int nbvh_allreduce_setup(nbvh_ctx *ctx, void *nccl_comm)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !nccl_comm)
        return NBVH_ERR_INVALID_ARG;
    NcclApi &a = nccl_api();
    if (!a.lib)
        return fail(ctx, NBVH_ERR_STATE, a.error);
    int world = 0, rank = 0, rc;
    if ((rc = a.CommCount(nccl_comm, &world)) != kNcclSuccess || (rc = a.CommUserRank(nccl_comm, &rank)) != kNcclSuccess)
        return nccl_fail(ctx, "ncclCommCount / ncclCommUserRank", rc);
    dp_release(ctx);
    return attach(ctx, nccl_comm, false, rank, world);
}

int nbvh_allreduce_setup_unique_id(nbvh_ctx *ctx, const void *unique_id, int rank, int world)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || !unique_id || world < 1 || rank < 0 || rank >= world)
        return ctx ? fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_allreduce_setup_unique_id: bad rank / world") : NBVH_ERR_INVALID_ARG;
    NcclApi &a = nccl_api();
    if (!a.lib)
        return fail(ctx, NBVH_ERR_STATE, a.error);
    dp_release(ctx);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    NcclComm comm = nullptr;
    const int rc  = a.CommInitRank(&comm, world, id, rank);
    if (rc != kNcclSuccess)
        return nccl_fail(ctx, "ncclCommInitRank", rc);
    return attach(ctx, comm, true, rank, world);
}

int nbvh_allreduce_teardown(nbvh_ctx *ctx)
{
    DeviceGuard device_guard(ctx);
    if (!ctx)
        return NBVH_ERR_INVALID_ARG;
    if (ctx->comm_stream)
        cudaStreamSynchronize(ctx->comm_stream);
    dp_release(ctx);
    return NBVH_OK;
}

void *nbvh_nccl_comm(const nbvh_ctx *ctx) { return ctx ? ctx->nccl_comm : nullptr; }

int nbvh_data_parallel_world(const nbvh_ctx *ctx, int *rank_out)
{
    if (!ctx)
        return 0;
    if (rank_out)
        *rank_out = ctx->dp_rank;
    return ctx->nccl_comm ? ctx->dp_world : 1;
}

int nbvh_allreduce_buffer(nbvh_ctx *ctx, void *device_ptr, uint64_t count, int dtype, void *stream)
{
    DeviceGuard device_guard(ctx);
    if (!ctx || (!device_ptr && count))
        return NBVH_ERR_INVALID_ARG;
    if (!ctx->nccl_comm || ctx->dp_world <= 1 || count == 0)
        return NBVH_OK;
    int t;
    switch (dtype) {
    case NBVH_DTYPE_F32: t = kNcclFloat32; break;
    case NBVH_DTYPE_U32: t = kNcclUint32; break;
    case NBVH_DTYPE_I32: t = kNcclInt32; break;
    case NBVH_DTYPE_U64: t = kNcclUint64; break;
    case NBVH_DTYPE_F64: t = kNcclFloat64; break;
    default: return fail(ctx, NBVH_ERR_INVALID_ARG, "nbvh_allreduce_buffer: unknown dtype");
    }
    const int rc = nccl_api().AllReduce(device_ptr, device_ptr, (size_t)count, t, kNcclSum, ctx->nccl_comm, (cudaStream_t)stream);
    if (rc != kNcclSuccess)
        return nccl_fail(ctx, "ncclAllReduce", rc);
    return NBVH_OK;
}

} // extern "C"
