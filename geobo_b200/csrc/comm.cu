// NCCL plumbing for the multi-GPU path (one process per GPU).  libnccl is resolved lazily with
// dlopen so that the single-GPU library has no load-time dependency on it; the only collective on
// the data path is the all-reduce of the per-shard AkA partial sums.
#include <dlfcn.h>

#include "common.cuh"
#include "comm.h"

namespace {
typedef struct { char internal[128]; } ncclUniqueId_;
typedef int (*fn_getuid)(ncclUniqueId_*);
typedef int (*fn_initrank)(void**, int, ncclUniqueId_, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_getuid get_uid = nullptr;
    fn_initrank init_rank = nullptr;
    fn_allreduce all_reduce = nullptr;
    fn_allgather all_gather = nullptr;
    fn_broadcast broadcast = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
} g_nccl;

int load_nccl(gb_ctx* ctx) {
    if (g_nccl.handle) return GB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return gb_fail(ctx, GB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.get_uid = (fn_getuid)dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_initrank)dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.all_reduce = (fn_allreduce)dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.all_gather = (fn_allgather)dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.broadcast = (fn_broadcast)dlsym(g_nccl.handle, "ncclBroadcast");
    g_nccl.destroy = (fn_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.all_reduce || !g_nccl.destroy)
        return gb_fail(ctx, GB_ERR_NCCL, "libnccl is missing required symbols");
    return GB_OK;
}
}  // namespace

extern "C" int gb_comm_unique_id(gb_ctx* ctx, void* id128) {
    if (!ctx || !id128) return gb_fail(ctx, GB_ERR_ARG, "gb_comm_unique_id: null argument");
    GB_TRY(load_nccl(ctx));
    ncclUniqueId_ id;
    int rc = g_nccl.get_uid(&id);
    if (rc != 0) return gb_fail(ctx, GB_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    memcpy(id128, &id, 128);
    return GB_OK;
}

extern "C" int gb_comm_init(gb_ctx* ctx, const void* id128, int rank, int nranks) {
    if (!ctx || !id128 || rank < 0 || nranks < 1 || rank >= nranks) return gb_fail(ctx, GB_ERR_ARG, "gb_comm_init: bad argument");
    GB_TRY(load_nccl(ctx));
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId_ id;
    memcpy(&id, id128, 128);
    void* comm = nullptr;
    int rc = g_nccl.init_rank(&comm, nranks, id, rank);
    if (rc != 0) return gb_fail(ctx, GB_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return GB_OK;
}

int comm_allreduce_sum_f64(gb_ctx* ctx, double* buf, size_t count) {
    if (ctx->nranks <= 1) return GB_OK;
    if (!ctx->nccl_comm) return gb_fail(ctx, GB_ERR_NCCL, "multi-rank problem without gb_comm_init");
    const int ncclFloat64 = 8, ncclSum = 0;
    int rc = g_nccl.all_reduce(buf, buf, count, ncclFloat64, ncclSum, ctx->nccl_comm, ctx->stream);
    if (rc != 0) return gb_fail(ctx, GB_ERR_NCCL, "ncclAllReduce: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    return GB_OK;
}

// In-place broadcast of `count` doubles from `root` on the context stream (Cholesky panels of the distributed factorisation)
int comm_broadcast_f64(gb_ctx* ctx, double* buf, size_t count, int root) {
    if (ctx->nranks <= 1) return GB_OK;
    if (!ctx->nccl_comm || !g_nccl.broadcast) return gb_fail(ctx, GB_ERR_NCCL, "multi-rank problem without gb_comm_init");
    const int ncclFloat64 = 8;
    int rc = g_nccl.broadcast(buf, buf, count, ncclFloat64, root, ctx->nccl_comm, ctx->stream);
    if (rc != 0) return gb_fail(ctx, GB_ERR_NCCL, "ncclBroadcast: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    return GB_OK;
}

// Gathers `count` doubles per rank (host buffers; staged through device memory, NCCL over NVLink): out[r * count ...] = rank r's data.
// Used for the 6 N result values (mean / variance shards) so that every rank returns whole cubes from Inversion.cubing.
extern "C" int gb_comm_allgather(gb_ctx* ctx, const double* local, int64_t count, double* out) {
    if (!ctx || !local || !out || count < 1) return gb_fail(ctx, GB_ERR_ARG, "gb_comm_allgather: bad argument");
    if (ctx->nranks <= 1) {
        memcpy(out, local, (size_t)count * sizeof(double));
        return GB_OK;
    }
    if (!ctx->nccl_comm || !g_nccl.all_gather) return gb_fail(ctx, GB_ERR_NCCL, "gb_comm_allgather without gb_comm_init");
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    void *send = nullptr, *recv = nullptr;
    int rc_out = GB_OK;
    cudaError_t e = gb_dev_malloc(ctx, &send, (size_t)count * sizeof(double));
    if (e == cudaSuccess) e = gb_dev_malloc(ctx, &recv, (size_t)count * ctx->nranks * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(send, local, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const int ncclFloat64 = 8;
        const int rc = g_nccl.all_gather(send, recv, (size_t)count, ncclFloat64, ctx->nccl_comm, ctx->stream);
        if (rc != 0) rc_out = gb_fail(ctx, GB_ERR_NCCL, "ncclAllGather: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    }
    if (e == cudaSuccess && rc_out == GB_OK)
        e = cudaMemcpyAsync(out, recv, (size_t)count * ctx->nranks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    gb_dev_free(ctx, send);       // both back to the context's cache, also on the error paths
    gb_dev_free(ctx, recv);
    if (e != cudaSuccess) return gb_fail(ctx, e == cudaErrorMemoryAllocation ? GB_ERR_NOMEM : GB_ERR_CUDA, "gb_comm_allgather: %s", cudaGetErrorString(e));
    return rc_out;
}

void comm_destroy(gb_ctx* ctx) {
    if (ctx->nccl_comm && g_nccl.destroy) g_nccl.destroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}
