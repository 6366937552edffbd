// nccl.cu -- final aggregate across GPUs: one ncclAllReduce(sum) of the partial
// aggregate state, the analogue of CoalescePartitionsExec + AggregateExec(Final).  NCCL is resolved with
// dlopen at first use so that the library has no link-time dependency on a particular libnccl (the process
// may already have torch's bundled NCCL loaded; RTLD_NOLOAD picks that one up first).
#include <mutex>
#include <dlfcn.h>

#include <cstring>

#include "internal.h"

namespace exon {

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt64 = 4, ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
    std::lock_guard<std::mutex> g(g_nccl_mu);
    if (g_nccl.handle) return EXON_GPU_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;
    if (!h)
        for (const char *n : names)
            if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return fail(EXON_GPU_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.GroupStart ||
        !g_nccl.GroupEnd || !g_nccl.GetErrorString)
        return fail(EXON_GPU_ERR_NCCL, "libnccl is missing a required symbol");
    g_nccl.handle = h;
    return EXON_GPU_OK;
}

#define NCCL_TRY(expr)                                                                                  \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != 0) return fail(EXON_GPU_ERR_NCCL, "%s: %s", #expr, g_nccl.GetErrorString(_r));        \
    } while (0)
#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)
}  // namespace

void nccl_teardown(Ctx *c) {
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
}

int nccl_allreduce_i64(Ctx *c, int64_t *device_buf, size_t n) {
    if (!c->nccl_comm) return fail(EXON_GPU_ERR_STATE, "all-reduce: exon_gpu_nccl_init has not been called");
    NCCL_TRY(g_nccl.AllReduce(device_buf, device_buf, n, ncclInt64, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_nccl_unique_id(uint8_t id[EXON_GPU_NCCL_ID_BYTES]) {
    if (!id) return fail(EXON_GPU_ERR_ARG, "nccl_unique_id: id is NULL");
    if (int rc = load_nccl()) return rc;
    ncclUniqueId u;
    NCCL_TRY(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(u) == EXON_GPU_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, sizeof(u));
    return EXON_GPU_OK;
}

int exon_gpu_nccl_init(exon_gpu_ctx *c, const uint8_t id[EXON_GPU_NCCL_ID_BYTES], int n_ranks, int rank) {
    if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(EXON_GPU_ERR_ARG, "nccl_init: bad argument");
    if (int rc = load_nccl()) return rc;
    CUDA_TRY(cudaSetDevice(c->device));
    nccl_teardown(c);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    NCCL_TRY(g_nccl.CommInitRank(&comm, n_ranks, u, rank));
    c->nccl_comm = comm;
    c->nccl_ranks = n_ranks;
    return EXON_GPU_OK;
}

// AggregateExec(Final): count and integer sum are reduced as int64 (bit-exact), the float sum as float64
// (order-dependent across ranks: 1e-6 relative, north_star).
int exon_gpu_allreduce_partial(exon_gpu_ctx *c, exon_gpu_partial *inout) {
    if (!c || !inout) return fail(EXON_GPU_ERR_ARG, "allreduce_partial: NULL argument");
    if (!c->nccl_comm) return fail(EXON_GPU_ERR_STATE, "allreduce_partial: exon_gpu_nccl_init has not been called");
    CUDA_TRY(cudaSetDevice(c->device));
    if (int rc = c->ensure_scratch(64, 64)) return rc;
    static_assert(sizeof(exon_gpu_partial) == 24, "partial is {i64, i64, f64}");
    memcpy(c->h_scratch, inout, sizeof(*inout));
    CUDA_TRY(cudaMemcpyAsync(c->scratch, c->h_scratch, sizeof(*inout), cudaMemcpyHostToDevice, c->stream));
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    NCCL_TRY(g_nccl.GroupStart());
    NCCL_TRY(g_nccl.AllReduce(c->scratch, c->scratch, 2, ncclInt64, ncclSum, comm, c->stream));
    NCCL_TRY(g_nccl.AllReduce((char *)c->scratch + 16, (char *)c->scratch + 16, 1, ncclFloat64, ncclSum, comm, c->stream));
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaMemcpyAsync(c->h_scratch, c->scratch, sizeof(*inout), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(inout, c->h_scratch, sizeof(*inout));
    return EXON_GPU_OK;
}

// AggregateExec(Final) of a GROUP BY with a small, rank-independent group set (per-reference counts): one
// ncclAllReduce(sum, int64, n) of the count vector.
int exon_gpu_allreduce_counts(exon_gpu_ctx *c, int64_t *inout, int32_t n) {
    if (!c || (!inout && n) || n < 0) return fail(EXON_GPU_ERR_ARG, "allreduce_counts: bad argument");
    if (!c->nccl_comm) return fail(EXON_GPU_ERR_STATE, "allreduce_counts: exon_gpu_nccl_init has not been called");
    if (n == 0) return EXON_GPU_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    std::lock_guard<std::mutex> work(c->work_mu);
    const size_t bytes = (size_t)n * sizeof(int64_t);
    if (int rc = c->ensure_scratch(bytes, bytes)) return rc;
    memcpy(c->h_scratch, inout, bytes);
    CUDA_TRY(cudaMemcpyAsync(c->scratch, c->h_scratch, bytes, cudaMemcpyHostToDevice, c->stream));
    if (int rc = nccl_allreduce_i64(c, (int64_t *)c->scratch, (size_t)n)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_scratch, c->scratch, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(inout, c->h_scratch, bytes);
    return EXON_GPU_OK;
}

}  // extern "C"
