// nccl.cu -- final aggregate across GPUs: one ncclAllReduce(sum) of the partial
// aggregate state, the analogue of CoalescePartitionsExec + AggregateExec(Final).  NCCL is resolved with
// dlopen at first use so that the library has no link-time dependency on a particular libnccl (the process
// may already have torch's bundled NCCL loaded; RTLD_NOLOAD picks that one up first).
#include <mutex>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"

namespace exon {

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclInt32 = 2, ncclInt64 = 4, ncclFloat64 = 8, ncclSum = 0, ncclMin = 3 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;  // optional (peer exchange setup)
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
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
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

// ---- scalar all-reduce over peer memory ------------------------------------------------------------------------------
// The final aggregate of the headline query is ONE int64 per rank.  A separate NCCL launch for it costs ~90 us per step at
// 8 GPUs (measured: step 0.689 ms against a 0.571 ms scan), a separate 32-thread exchange kernel + read-back still ~40 us;
// here the LAST CTA of the scan kernel itself (vcf_scan.cu: scan_finalize) stores the rank's partial straight into a slot of
// every peer's exchange buffer over NVLink (CUDA IPC mappings set up once at exon_gpu_nccl_init), publishes it with a
// sequence number behind a system-scope fence, sums the slots of its own buffer as they arrive and writes the result to
// mapped host memory.  This file owns the buffers and the sequence numbers.
//   slot layout (per rank): [2 parities][n ranks] x {value, seq}; parity = seq & 1 so that a rank that runs one step ahead
//   never overwrites values a peer is still reading (it cannot run two ahead: every step needs everybody's contribution)
// NCCL stays the fallback (setup failure on any rank, EXON_GPU_PEER_XCHG=0) and the path for vectors / float state.
struct PeerXchg {
    int n = 0, rank = 0;
    unsigned long long *local = nullptr;       // this rank's slots (cudaMalloc, IPC-exported)
    unsigned long long **d_peers = nullptr;    // device array: peer p's slots as mapped here (d_peers[rank] == local)
    std::vector<void *> opened;                // IPC mappings to close
    unsigned long long seq = 0;
};

static void peer_xchg_teardown(Ctx *c) {
    auto *x = static_cast<PeerXchg *>(c->peer_xchg);
    if (!x) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (void *p : x->opened) cudaIpcCloseMemHandle(p);
    cudaFree(x->d_peers);
    cudaFree(x->local);
    delete x;
    c->peer_xchg = nullptr;
}

// Collective over the communicator: every rank calls it; the exchange is enabled only when it worked everywhere.
static void peer_xchg_setup(Ctx *c, int n, int rank) {
    if (n < 2) return;  // every rank sees the same n: nobody waits in the collectives below
    const char *e = getenv("EXON_GPU_PEER_XCHG");
    bool want = !(e && atoi(e) == 0) && n <= 32 && g_nccl.AllGather;
    auto *x = new PeerXchg();
    x->n = n;
    x->rank = rank;
    int ok = want ? 1 : 0;
    uint8_t *d_h = nullptr;
    std::vector<uint8_t> all((size_t)n * sizeof(cudaIpcMemHandle_t));
    if (ok && (cudaMalloc((void **)&x->local, (size_t)2 * n * 16) != cudaSuccess || cudaMemset(x->local, 0, (size_t)2 * n * 16) != cudaSuccess)) ok = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) ok = 0;  // the slots are zero before anybody can learn their address
    if (cudaMalloc((void **)&d_h, all.size() + 64) != cudaSuccess) ok = 0, d_h = nullptr;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, x->local) != cudaSuccess) ok = 0;
    // handles travel with ncclAllGather (every rank takes part, whatever its own state, so that nobody is left waiting)
    if (d_h) {
        cudaMemcpyAsync(d_h + all.size(), &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream);
        if (g_nccl.AllGather && g_nccl.AllGather(d_h + all.size(), d_h, sizeof(mine), ncclUint8, (ncclComm_t)c->nccl_comm, c->stream) != 0) ok = 0;
        cudaMemcpyAsync(all.data(), d_h, all.size(), cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) ok = 0;
    }
    std::vector<unsigned long long *> peers((size_t)n, nullptr);
    if (ok) {
        for (int p = 0; p < n && ok; ++p) {
            if (p == rank) {
                peers[(size_t)p] = x->local;
                continue;
            }
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + (size_t)p * sizeof(h), sizeof(h));
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = 0;
                break;
            }
            x->opened.push_back(ptr);
            peers[(size_t)p] = (unsigned long long *)ptr;
        }
        if (ok && (cudaMalloc((void **)&x->d_peers, (size_t)n * 8) != cudaSuccess ||
                   cudaMemcpy(x->d_peers, peers.data(), (size_t)n * 8, cudaMemcpyHostToDevice) != cudaSuccess))
            ok = 0;
    }
    cudaGetLastError();  // a failed IPC call must not poison later launches
    // agree: min over ranks of `ok` (also the barrier that orders every rank's memset before anybody's first store)
    int agreed = 0;
    if (d_h) {
        cudaMemcpyAsync(d_h, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream);
        if (g_nccl.AllReduce(d_h, d_h, 1, ncclInt32, ncclMin, (ncclComm_t)c->nccl_comm, c->stream) == 0) {
            cudaMemcpyAsync(&agreed, d_h, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) agreed = 0;
        }
        cudaFree(d_h);
    }
    c->peer_xchg = x;
    if (!agreed) peer_xchg_teardown(c);
}

// Fills the peer fields of a scan tail (vcf_scan.cu: scan_finalize runs the exchange inside the scan's last CTA) and
// consumes one sequence number.  Every rank must launch exactly one tail per armed exchange, in the same order.
bool peer_xchg_arm(Ctx *c, ScanTail *tail) {
    auto *x = static_cast<PeerXchg *>(c->peer_xchg);
    if (!x) return false;
    std::lock_guard<std::mutex> g(c->mu);
    tail->n_ranks = x->n;
    tail->rank = x->rank;
    tail->peers = x->d_peers;
    tail->xseq = ++x->seq;
    return true;
}

void nccl_teardown(Ctx *c) {
    peer_xchg_teardown(c);
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
    peer_xchg_setup(c, n_ranks, rank);
    return EXON_GPU_OK;
}

// AggregateExec(Final): count and integer sum are reduced as int64 (bit-exact), the float sum as float64
// (order-dependent across ranks: 1e-6 relative, north_star).
int exon_gpu_allreduce_partial(exon_gpu_ctx *c, exon_gpu_partial *inout) {
    if (!c || !inout) return fail(EXON_GPU_ERR_ARG, "allreduce_partial: NULL argument");
    if (!c->nccl_comm) return fail(EXON_GPU_ERR_STATE, "allreduce_partial: exon_gpu_nccl_init has not been called");
    CUDA_TRY(cudaSetDevice(c->device));
    std::lock_guard<std::recursive_mutex> work(c->work_mu);
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
    std::lock_guard<std::recursive_mutex> work(c->work_mu);
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
