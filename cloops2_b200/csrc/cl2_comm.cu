// cl2_comm.cu -- the path's only collective: a sum of int64 counters (PETs / candidates / loops per rank) over the
// GPUs of a job, NCCL over NVLink.  The reference has nothing to reduce (one process, joblib workers return their
// lists to it, callCisLoops.py:134-143); with one process per GPU the totals it logs and writes (petMeta.json's
// "Unique PETs", the candidate and loop counts) are sums over ranks.
//
// libnccl is opened at run time (dlopen), not linked: libcl2.so must load on hosts without NCCL, and inside a PyTorch
// process it has to share the copy PyTorch already loaded.  CL2_NCCL_LIB names another library file.
#include "cl2_common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <new>

struct cl2_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

namespace {
struct NcclApi {
    void *h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    char why[256] = {0};
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load() {
    const char *names[] = {getenv("CL2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !nm[0]) continue;
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
        snprintf(g_nccl.why, sizeof(g_nccl.why), "%s", dlerror());
    }
    if (!g_nccl.h) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        snprintf(g_nccl.why, sizeof(g_nccl.why), "libnccl lacks a required symbol");
        g_nccl.h = nullptr;
    }
}
bool nccl_ready() {
    std::call_once(g_nccl_once, nccl_load);
    return g_nccl.h != nullptr;
}
int nccl_fail(cl2_ctx *ctx, const char *what, ncclResult_t r) {
    if (ctx)
        snprintf(ctx->err, sizeof(ctx->err), "%s failed: %s", what,
                 g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return CL2_ERR_CUDA;
}
}  // namespace

extern "C" int cl2_comm_unique_id(uint8_t *id_out) {
    if (!id_out) return CL2_ERR_ARG;
    if (!nccl_ready()) return CL2_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == CL2_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return CL2_ERR_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return CL2_OK;
}

extern "C" int cl2_comm_init(cl2_ctx *ctx, const uint8_t *id, int32_t rank, int32_t world, cl2_comm **out) {
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_comm_init: bad argument");
    *out = nullptr;
    if (!nccl_ready()) return cl2_fail(ctx, CL2_ERR_CUDA, "cl2_comm_init: cannot load libnccl (%s)", g_nccl.why);
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    cl2_comm *c = new (std::nothrow) cl2_comm();
    if (!c) return CL2_ERR_NOMEM;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) {
        delete c;
        return nccl_fail(ctx, "ncclCommInitRank", r);
    }
    c->rank = rank;
    c->world = world;
    *out = c;
    return CL2_OK;
}

extern "C" int cl2_stats_allreduce(cl2_ctx *ctx, cl2_comm *comm, int64_t *vals, int32_t k) {
    if (!ctx || !comm || k < 0 || (k > 0 && !vals)) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_stats_allreduce: bad argument");
    if (k == 0 || comm->world == 1) return CL2_OK;
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    int64_t *d;
    CL2_TRY(sc.get(&d, (size_t)k));
    CL2_CUDA(ctx, cudaMemcpyAsync(d, vals, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = g_nccl.AllReduce(d, d, (size_t)k, ncclInt64, ncclSum, comm->comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce", r);
    CL2_CUDA(ctx, cudaMemcpyAsync(vals, d, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CL2_OK;
}

extern "C" void cl2_comm_destroy(cl2_comm *comm) {
    if (!comm) return;
    if (comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->comm);
    delete comm;
}
