// abi.cu -- the C ABI declared in include/exon_gpu.h: contexts, partition streams, the device arena that
// holds fed bytes, and the glue that launches the kernels.  No compute happens on the host here: every
// result comes from a CUDA kernel, and every entry point fails loudly without a usable sm_100 device.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/exon_gpu.h"
#include "internal.h"

namespace exon {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// Arena: device blocks that hold fed body bytes.  Blocks are recycled through the context so that a
// steady-state query does no cudaMalloc.
// ---------------------------------------------------------------------------------------------------
int Ctx::get_block(size_t min_bytes, DevBlock *out) {
    std::lock_guard<std::mutex> g(mu);
    for (size_t i = 0; i < free_blocks.size(); ++i) {
        if (free_blocks[i].cap >= min_bytes) {
            *out = free_blocks[i];
            free_blocks.erase(free_blocks.begin() + (long)i);
            return EXON_GPU_OK;
        }
    }
    size_t cap = std::max(min_bytes, kArenaBlock);
    void *p = nullptr;
    // +256: tail padding so that the 16-byte granule holding the last byte is always readable
    CUDA_TRY(cudaMalloc(&p, cap + 256));
    out->ptr = (uint8_t *)p;
    out->cap = cap;
    out->used = 0;
    return EXON_GPU_OK;
}

void Ctx::put_block(DevBlock b) {
    std::lock_guard<std::mutex> g(mu);
    b.used = 0;
    free_blocks.push_back(b);
}

static int ensure_device(Ctx *c) {
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;


extern "C" {

const char *exon_gpu_last_error(void) { return g_last_error.c_str(); }
const char *exon_gpu_version(void) { return "exon_gpu 0.1.0 sm_100a"; }

int exon_gpu_ctx_create(int device, void *cuda_stream, exon_gpu_ctx **out) {
    if (!out) return fail(EXON_GPU_ERR_ARG, "exon_gpu_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(EXON_GPU_ERR_CUDA, "no CUDA device available (%s); exon_gpu has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(EXON_GPU_ERR_ARG, "device %d out of range [0, %d)", device, n);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(EXON_GPU_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    CUDA_TRY(cudaSetDevice(device));
    auto *c = new exon_gpu_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cuda_stream) {
        c->stream = (cudaStream_t)cuda_stream;
        c->own_stream = false;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return fail(EXON_GPU_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    {  // keep freed column stores cached in the stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    for (int i = 0; i < Ctx::kEvRing; ++i) {
        cudaEventCreate(&c->ev[i][0]);
        cudaEventCreate(&c->ev[i][1]);
    }
    *out = c;
    return EXON_GPU_OK;
}

int exon_gpu_ctx_destroy(exon_gpu_ctx *c) {
    if (!c) return EXON_GPU_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &b : c->free_blocks) cudaFree(b.ptr);
    if (c->scratch) cudaFree(c->scratch);
    if (c->scratch_b) cudaFree(c->scratch_b);
    if (c->inf_tokens) cudaFree(c->inf_tokens);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    nccl_teardown(c);
    for (int i = 0; i < Ctx::kEvRing; ++i) {
        cudaEventDestroy(c->ev[i][0]);
        cudaEventDestroy(c->ev[i][1]);
    }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return EXON_GPU_OK;
}

int exon_gpu_ctx_launch_count(exon_gpu_ctx *c, int64_t *out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "launch_count: NULL argument");
    *out = c->launches.load();
    return EXON_GPU_OK;
}

int exon_gpu_ctx_last_kernel_ms(exon_gpu_ctx *c, float *out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "last_kernel_ms: NULL argument");
    if (int rc = ensure_device(c)) return rc;
    int32_t n = 0;
    if (int rc = exon_gpu_ctx_kernel_ms_history(c, out, 1, &n)) return rc;
    if (n == 0) return fail(EXON_GPU_ERR_STATE, "no fused-scan kernel has been launched on this context");
    return EXON_GPU_OK;
}

// Durations of the most recent fused-scan launches, oldest first (at most 64 are kept).  Benches read a whole timed
// region's launches afterwards, so that no step pays for an event synchronisation.
int exon_gpu_ctx_kernel_ms_history(exon_gpu_ctx *c, float *out, int32_t cap, int32_t *out_n) {
    if (!c || !out_n || (!out && cap > 0) || cap < 0) return fail(EXON_GPU_ERR_ARG, "kernel_ms_history: bad argument");
    if (int rc = ensure_device(c)) return rc;
    int64_t count;
    {
        std::lock_guard<std::mutex> g(c->mu);
        count = c->ev_count;
    }
    int64_t n = std::min<int64_t>({count, (int64_t)Ctx::kEvRing, (int64_t)cap});
    for (int64_t i = 0; i < n; ++i) {
        const int slot = (int)((count - n + i) % Ctx::kEvRing);
        CUDA_TRY(cudaEventSynchronize(c->ev[slot][1]));
        CUDA_TRY(cudaEventElapsedTime(out + i, c->ev[slot][0], c->ev[slot][1]));
    }
    *out_n = (int32_t)n;
    return EXON_GPU_OK;
}

int exon_gpu_ctx_synchronize(exon_gpu_ctx *c) {
    if (!c) return fail(EXON_GPU_ERR_ARG, "synchronize: ctx is NULL");
    if (int rc = ensure_device(c)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return EXON_GPU_OK;
}

int exon_gpu_host_alloc(exon_gpu_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "host_alloc: NULL argument");
    if (int rc = ensure_device(c)) return rc;
    CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return EXON_GPU_OK;
}
int exon_gpu_host_free(exon_gpu_ctx *c, void *p) {
    if (!c) return fail(EXON_GPU_ERR_ARG, "host_free: ctx is NULL");
    if (p) CUDA_TRY(cudaFreeHost(p));
    return EXON_GPU_OK;
}
int exon_gpu_device_alloc(exon_gpu_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "device_alloc: NULL argument");
    if (int rc = ensure_device(c)) return rc;
    CUDA_TRY(cudaMalloc(out, (bytes ? bytes : 1) + 256));
    return EXON_GPU_OK;
}
int exon_gpu_device_free(exon_gpu_ctx *c, void *p) {
    if (!c) return fail(EXON_GPU_ERR_ARG, "device_free: ctx is NULL");
    if (int rc = ensure_device(c)) return rc;
    if (p) CUDA_TRY(cudaFree(p));
    return EXON_GPU_OK;
}
int exon_gpu_memcpy_h2d(exon_gpu_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || (!dst && bytes) || (!src && bytes)) return fail(EXON_GPU_ERR_ARG, "memcpy_h2d: NULL argument");
    if (int rc = ensure_device(c)) return rc;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return EXON_GPU_OK;
}

int exon_gpu_memcpy_h2d_async(exon_gpu_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || (!dst && bytes) || (!src && bytes)) return fail(EXON_GPU_ERR_ARG, "memcpy_h2d_async: NULL argument");
    if (int rc = ensure_device(c)) return rc;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return EXON_GPU_OK;
}

int exon_gpu_memset(exon_gpu_ctx *c, void *p, int value, size_t bytes) {
    if (!c || (!p && bytes)) return fail(EXON_GPU_ERR_ARG, "memset: NULL argument");
    if (int rc = ensure_device(c)) return rc;
    CUDA_TRY(cudaMemsetAsync(p, value, bytes, c->stream));
    return EXON_GPU_OK;
}

// ---- VCF partition stream ---------------------------------------------------------------------------

int exon_gpu_vcf_open(exon_gpu_ctx *c, const exon_gpu_vcf_opts *o, exon_gpu_stream **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "vcf_open: NULL argument");
    *out = nullptr;
    if (int rc = ensure_device(c)) return rc;
    auto *s = new exon_gpu_stream();
    s->ctx = c;
    s->batch_rows = 8192;
    if (o) {
        if (o->batch_rows < 0 || o->n_projection < 0 || (o->n_projection > 0 && !o->projection)) {
            delete s;
            return fail(EXON_GPU_ERR_ARG, "vcf_open: bad batch_rows / projection");
        }
        if (o->batch_rows > 0) s->batch_rows = o->batch_rows;
        for (int i = 0; i < o->n_projection; ++i) {
            if (o->projection[i] < 0 || o->projection[i] > 8) {
                delete s;
                return fail(EXON_GPU_ERR_ARG, "vcf_open: projection index %d is not a VCF file-schema column",
                            o->projection[i]);
            }
            for (int j = 0; j < i; ++j)
                if (o->projection[j] == o->projection[i]) {
                    delete s;
                    return fail(EXON_GPU_ERR_ARG, "vcf_open: column %d is projected twice", o->projection[i]);
                }
            s->projection.push_back(o->projection[i]);
        }
        s->columns_on_device = o->columns_on_device != 0;
        s->strict = o->strict != 0;
        s->variant = o->kernel_variant;
        if (o->pushdown) {
            if (int rc = s->pushdown.assign(o->pushdown)) {
                delete s;
                return rc;
            }
            s->has_pushdown = true;
        }
    }
    // device accumulators + the MAPPED pinned record the scan tail publishes into (layout: internal.h)
    cudaError_t e = cudaMalloc((void **)&s->d_res, 16 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&s->h_res, 16 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) {
        memset(s->h_res, 0, 16 * sizeof(unsigned long long));
        e = cudaHostGetDevicePointer((void **)&s->h_res_dev, s->h_res, 0);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_res, 0, 16 * sizeof(unsigned long long), c->stream);
    if (e != cudaSuccess) {
        if (s->d_res) cudaFree(s->d_res);
        delete s;
        return fail(EXON_GPU_ERR_CUDA, "vcf_open: %s", cudaGetErrorString(e));
    }
    *out = s;
    return EXON_GPU_OK;
}

int exon_gpu_vcf_close(exon_gpu_stream *s) {
    if (!s) return EXON_GPU_OK;
    Ctx *c = s->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    s->release_all();
    if (s->d_res) cudaFree(s->d_res);
    if (s->h_res) cudaFreeHost(s->h_res);
    if (s->d_segs) cudaFree(s->d_segs);
    if (s->d_tiles) cudaFree(s->d_tiles);
    s->gz_teardown();
    if (s->d_bam) cudaFree(s->d_bam);
    delete s;
    return EXON_GPU_OK;
}

int exon_gpu_vcf_reset(exon_gpu_stream *s) {
    if (!s) return fail(EXON_GPU_ERR_ARG, "vcf_reset: stream is NULL");
    if (int rc = ensure_device(s->ctx)) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    s->release_all();
    CUDA_TRY(cudaMemsetAsync(s->d_res, 0, 16 * sizeof(unsigned long long), s->ctx->stream));
    return EXON_GPU_OK;
}

// The header's ##INFO definitions, as the reference's builder gets them from the noodles Header the opener parsed
// (VCFOpener::open, exon/exon-core/src/datasources/vcf/file_opener/unindex_file_opener.rs:74-88; LazyVCFArrayBuilder::create,
// exon/exon-vcf/src/array_builder/lazy_array_builder.rs:70-75).  Only ID and Type are needed.
int exon_gpu_vcf_set_header(exon_gpu_stream *s, const char *text, size_t len) {
    if (!s || (!text && len)) return fail(EXON_GPU_ERR_ARG, "vcf_set_header: NULL argument");
    InfoDefs defs[2];  // ##INFO, ##FORMAT
    const char *p = text, *end = text + len;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        int which = -1;
        size_t skip = 0;
        if (le - p > 8 && memcmp(p, "##INFO=<", 8) == 0) which = 0, skip = 8;
        else if (le - p > 10 && memcmp(p, "##FORMAT=<", 10) == 0) which = 1, skip = 10;
        if (which >= 0) {
            auto field = [&](const char *key, std::string *out) {
                const size_t kl = strlen(key);
                for (const char *q = p + skip; q + kl <= le; ++q) {
                    if ((q == p + skip || q[-1] == ',') && memcmp(q, key, kl) == 0) {
                        const char *v = q + kl, *ve = v;
                        while (ve < le && *ve != ',' && *ve != '>') ++ve;
                        out->assign(v, (size_t)(ve - v));
                        return true;
                    }
                    if (*q == '"') {  // quoted text (Description) may hold commas
                        ++q;
                        while (q < le && *q != '"') ++q;
                    }
                }
                return false;
            };
            std::string id, type, number;
            if (!field("ID=", &id) || !field("Type=", &type) || id.empty())
                return fail(EXON_GPU_ERR_PARSE, "vcf_set_header: an ##INFO / ##FORMAT line without ID / Type");
            field("Number=", &number);
            int t = type == "Integer" ? 0 : type == "Float" ? 1 : type == "Flag" ? 2 : type == "Character" ? 3 : type == "String" ? 4 : -1;
            if (t < 0 || (which == 1 && t == 2)) return fail(EXON_GPU_ERR_PARSE, "vcf_set_header: unknown %s type '%s'", which ? "FORMAT" : "INFO", type.c_str());
            defs[which].ids.push_back(id);
            defs[which].types.push_back((uint8_t)t);
            defs[which].single.push_back(number == "1" ? 1 : 0);
        }
        if (!nl) break;
        p = nl + 1;
    }
    defs[0].set = defs[1].set = true;
    s->info_defs = std::move(defs[0]);
    s->format_defs = std::move(defs[1]);
    return EXON_GPU_OK;
}

int exon_gpu_vcf_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last) {
    if (!s || (!text && len)) return fail(EXON_GPU_ERR_ARG, "vcf_feed: NULL argument");
    if (int rc = ensure_device(s->ctx)) return rc;
    if (s->drained) return fail(EXON_GPU_ERR_STATE, "vcf_feed: the stream has already produced batches");
    if (int rc = s->flush_gz()) return rc;  // compressed files fed earlier come first
    return is_device_ptr ? s->feed_device(text, len, is_last != 0) : s->feed_host(text, len, is_last != 0);
}

int exon_gpu_vcf_filter_count_async(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *device_out) {
    if (!s || !device_out) return fail(EXON_GPU_ERR_ARG, "vcf_filter_count_async: NULL argument");
    if (int rc = ensure_device(s->ctx)) return rc;
    return s->filter_count(region, device_out, nullptr);
}

int exon_gpu_vcf_filter_count(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *out_count) {
    if (!s || !out_count) return fail(EXON_GPU_ERR_ARG, "vcf_filter_count: NULL argument");
    if (int rc = ensure_device(s->ctx)) return rc;
    return s->filter_count(region, nullptr, out_count);
}

int exon_gpu_vcf_filter_count_global(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *out_local,
                                     int64_t *out_global) {
    if (!s || (!out_local && !out_global)) return fail(EXON_GPU_ERR_ARG, "vcf_filter_count_global: NULL argument");
    if (int rc = ensure_device(s->ctx)) return rc;
    if (!s->ctx->nccl_comm) return fail(EXON_GPU_ERR_STATE, "vcf_filter_count_global: exon_gpu_nccl_init has not been called");
    return s->filter_count_global(region, out_local, out_global);
}

int exon_gpu_vcf_rows(exon_gpu_stream *s, int64_t *out_rows) {
    if (!s || !out_rows) return fail(EXON_GPU_ERR_ARG, "vcf_rows: NULL argument");
    if (int rc = ensure_device(s->ctx)) return rc;
    return s->filter_count(nullptr, nullptr, out_rows);
}

int exon_gpu_vcf_body_bytes(exon_gpu_stream *s, int64_t *out_bytes) {
    if (!s || !out_bytes) return fail(EXON_GPU_ERR_ARG, "vcf_body_bytes: NULL argument");
    if (int rc = ensure_device(s->ctx)) return rc;
    if (int rc = s->flush_gz()) return rc;
    *out_bytes = s->body_bytes;
    return EXON_GPU_OK;
}

}  // extern "C"
