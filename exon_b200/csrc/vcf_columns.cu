// vcf_columns.cu -- K2: VCF text -> Arrow columns {chrom: utf8, pos: int64} in reference-sized batches.
//
// Replaces AsyncBatchStream::read_batch (exon/exon-vcf/src/async_batch_stream.rs:80-109),
// LazyVCFArrayBuilder::{append, finish} for the projected columns 0 and 1
// (exon/exon-vcf/src/array_builder/lazy_array_builder.rs:157-168, 451-484) and
// ExonArrayBuilder::try_into_record_batch (exon/exon-common/src/array_builder.rs:25-36).
//
// The reference builds one batch at a time, row by row.  Here the whole resident partition is converted in
// five data-parallel passes and then handed out as zero-copy slices of batch_rows rows:
//   1. count_lines     line starts per 16 KiB block                       (reads the text once)
//   2. exclusive scan  block -> first row index                            (cub)
//   3. index_lines     row -> pointer of its first byte                    (reads the text a second time)
//   4. parse_rows      row -> (chrom length, POS as int64), validation     (touches ~12 bytes per row)
//   5. exclusive scan of chrom lengths + gather_chrom + batch_offsets      (Arrow utf8 layout per batch:
//                      int32 offsets that restart at 0 in every batch, as arrow-rs' StringBuilder emits)
// Batches own nothing: they are views into the stream's column store, kept alive by a reference count that
// the Arrow release callbacks decrement.
#include <cub/device/device_scan.cuh>

#include <atomic>
#include <cstring>
#include <new>

#include "common.cuh"
#include "internal.h"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

constexpr int kBlockBytes = 16384;  // text bytes per CTA in passes 1 and 3
constexpr int kThreads = 256;

struct RunDesc {
    const uint8_t *base;  // first valid byte
    int64_t len;
    int64_t block0;  // index of the run's first block in the launch-wide block numbering
    int64_t row0;    // filled after pass 2: first row of the run
};

// Newline flags (0x80 per byte) of the 16-byte chunk at aligned address `p`, restricted to positions
// [0, len - 1) relative to `base` (a '\n' that is the run's last byte starts no line).
__device__ __forceinline__ void chunk_flags(const uint8_t *p, const uint8_t *base, int64_t len, uint32_t f[4]) {
    const uint4 w = *reinterpret_cast<const uint4 *>(p);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    const int64_t rel = p - base;  // position of byte 0 of the chunk
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t x = zero_bytes_exact(ws[k] ^ kNL4);
        const int64_t r = rel + 4 * k;
        if (r < 0 || r + 4 > len - 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (r + j < 0 || r + j >= len - 1) x &= ~(0x80u << (8 * j));
        }
        f[k] = x;
    }
}

__device__ __forceinline__ int find_run(const RunDesc *runs, int n_runs, int64_t block) {
    int lo = 0, hi = n_runs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (runs[mid].block0 <= block) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// pass 1: block_rows[b] = number of lines that START in block b
__global__ void __launch_bounds__(kThreads) count_lines(const RunDesc *runs, int n_runs, int64_t n_blocks,
                                                       unsigned long long *block_rows) {
    __shared__ uint32_t warp_sums[kThreads / 32];
    for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const int r = find_run(runs, n_runs, b);
        const uint8_t *base = runs[r].base;
        const int64_t len = runs[r].len;
        const uint8_t *abase = base - ((uintptr_t)base & 15);
        const int64_t bi = b - runs[r].block0;
        const uint8_t *p0 = abase + bi * kBlockBytes;
        const int64_t span = (base + len) - p0;  // bytes from block start to run end
        uint32_t c = 0;
        for (int i = threadIdx.x; i < kBlockBytes / 16; i += kThreads) {
            if ((int64_t)i * 16 < span) {
                uint32_t f[4];
                chunk_flags(p0 + i * 16, base, len, f);
                c += __popc(f[0]) + __popc(f[1]) + __popc(f[2]) + __popc(f[3]);
            }
        }
        c = warp_sum(c);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kThreads / 32; ++w) t += warp_sums[w];
            if (bi == 0 && len > 0) t += 1;  // the run's first line has no '\n' before it
            block_rows[b] = t;
        }
        __syncthreads();
    }
}

// pass 3: line_ptr[row] = address of the row's first byte, rows numbered in byte order
__global__ void __launch_bounds__(kThreads) index_lines(const RunDesc *runs, int n_runs, int64_t n_blocks,
                                                       const unsigned long long *block_row0,
                                                       const uint8_t **line_ptr) {
    __shared__ uint32_t warp_sums[kThreads / 32];
    __shared__ uint32_t round_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const int r = find_run(runs, n_runs, b);
        const uint8_t *base = runs[r].base;
        const int64_t len = runs[r].len;
        const uint8_t *abase = base - ((uintptr_t)base & 15);
        const int64_t bi = b - runs[r].block0;
        const uint8_t *p0 = abase + bi * kBlockBytes;
        const int64_t span = (base + len) - p0;
        unsigned long long row = block_row0[b];
        if (bi == 0 && len > 0) {
            if (threadIdx.x == 0) line_ptr[row] = base;
            row += 1;
        }
        if (threadIdx.x == 0) round_base = 0;
        __syncthreads();
        for (int i0 = 0; i0 < kBlockBytes / 16; i0 += kThreads) {
            const int i = i0 + threadIdx.x;
            uint32_t f[4] = {0, 0, 0, 0};
            if ((int64_t)i * 16 < span) chunk_flags(p0 + i * 16, base, len, f);
            const uint32_t c = __popc(f[0]) + __popc(f[1]) + __popc(f[2]) + __popc(f[3]);
            // exclusive prefix of c over the block, in thread (= byte) order
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            uint32_t before = round_base;
            for (int w = 0; w < warp; ++w) before += warp_sums[w];
            uint32_t k = before + incl - c;
            const uint8_t *cp = p0 + i * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t x = f[q];
                while (x) {
                    const int j = (__ffs(x) - 1) >> 3;
                    x &= x - 1;
                    line_ptr[row + k++] = cp + 4 * q + j + 1;
                }
            }
            __syncthreads();
            if (threadIdx.x == kThreads - 1) round_base = before + incl;
            __syncthreads();
        }
    }
}

// pass 4: one thread per row.  CHROM = bytes up to the first '\t'; POS = decimal usize (Rust from_str rules),
// non-zero, <= i64::MAX.  Reads stop at the run's end.
__global__ void __launch_bounds__(kThreads) parse_rows(const RunDesc *runs, int n_runs, int64_t n_rows,
                                                      const uint8_t *const *line_ptr, int want_chrom, int want_pos,
                                                      int32_t *chrom_len, int64_t *pos, uint32_t *flags,
                                                      unsigned long long *first_bad_row) {
    const int64_t row = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (row >= n_rows) return;
    int lo = 0, hi = n_runs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (runs[mid].row0 <= row) lo = mid;
        else hi = mid - 1;
    }
    const uint8_t *end = runs[lo].base + runs[lo].len;
    const uint8_t *p = line_ptr[row];
    uint32_t err = 0;
    const uint8_t *q = p;
    while (q < end && *q != '\t' && *q != '\n') ++q;
    if (q >= end || *q != '\t' || q == p) err |= kErrShortLine;
    if (want_chrom) chrom_len[row] = err ? 0 : (int32_t)(q - p);
    if (want_pos) {
        long long out = 0;
        if (!err) {
            ++q;
            if (q < end && *q == '+') ++q;
            unsigned long long v = 0;
            int sig = 0, nd = 0;
            bool ovf = false;
            while (q < end && (uint32_t)(*q - '0') <= 9u) {
                const uint32_t dg = *q - '0';
                if (v | dg) ++sig;
                if (sig > 19) ovf = true;
                v = v * 10ull + dg;
                ++nd;
                ++q;
            }
            if (nd == 0 || q >= end || *q != '\t') err |= (q >= end || *q == '\n') ? kErrShortLine : kErrBadPos;
            else if (ovf || v == 0ull || v > 0x7FFFFFFFFFFFFFFFull) err |= kErrBadPos;
            else out = (long long)v;
        }
        pos[row] = out;
    }
    if (err) {
        atomicOr(flags, err);
        atomicMin(first_bad_row, (unsigned long long)row);
    }
}

// Row index of every file end: first row of the mark's run whose line starts at or after the mark.
struct MarkDesc {
    int32_t run;  // index into the RunDesc table
    int32_t pad_;
    int64_t len;  // run bytes that belong to files up to and including the marked one
};
__global__ void mark_rows(const RunDesc *runs, int n_runs, const MarkDesc *marks, int n_marks,
                          const uint8_t *const *line_ptr, int64_t n_rows, long long *out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_marks) return;
    const int r = marks[m].run;
    int64_t lo = runs[r].row0, hi = (r + 1 < n_runs) ? runs[r + 1].row0 : n_rows;
    const uint8_t *lim = runs[r].base + marks[m].len;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (line_ptr[mid] < lim) lo = mid + 1;
        else hi = mid;
    }
    out[m] = lo;
}

// out[i] = src[idx[i]]
__global__ void gather_i64(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

// pass 5b: values[voff[row] .. ) = CHROM bytes; offsets restart at 0 in every batch.  Batches restart at every
// file: file f covers rows [file_row0[f], file_row0[f + 1]) and owns batches file_batch0[f] .. in steps of
// batch_rows rows.
__global__ void __launch_bounds__(kThreads) gather_chrom(int64_t n_rows, int batch_rows, int n_files,
                                                        const long long *file_row0, const long long *file_batch0,
                                                        const uint8_t *const *line_ptr, const int32_t *chrom_len,
                                                        const long long *voff, uint8_t *values, int32_t *offsets) {
    const int64_t row = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (row >= n_rows) return;
    int lo = 0, hi = n_files - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (file_row0[mid] <= row) lo = mid;
        else hi = mid - 1;
    }
    const int64_t f0 = file_row0[lo], f1 = file_row0[lo + 1];
    const int64_t k = (row - f0) / batch_rows;
    const int64_t batch = file_batch0[lo] + k, b0 = f0 + k * batch_rows;
    const int64_t in_batch = row - b0;
    const long long v0 = voff[b0];
    const long long v = voff[row];
    const int32_t n = chrom_len[row];
    int32_t *o = offsets + batch * (batch_rows + 1);
    o[in_batch] = (int32_t)(v - v0);
    if (row + 1 == f1 || in_batch + 1 == batch_rows) o[in_batch + 1] = (int32_t)(v + n - v0);
    const uint8_t *src = line_ptr[row];
    for (int32_t j = 0; j < n; ++j) values[v + j] = src[j];
}

}  // namespace

// Column store of one stream; batches are views into it.
struct Columns {
    std::atomic<int> refs{1};  // the stream holds one reference
    bool on_device = false;
    int device = 0;
    int64_t n_rows = 0, n_batches = 0, next = 0;
    int batch_rows = 8192;
    bool want_chrom = false, want_pos = false;
    // device store
    int64_t *d_pos = nullptr;
    int32_t *d_offsets = nullptr;
    uint8_t *d_values = nullptr;
    long long *d_voff = nullptr;
    // host mirrors (pinned) when !on_device
    int64_t *h_pos = nullptr;
    int32_t *h_offsets = nullptr;
    uint8_t *h_values = nullptr;
    std::vector<long long> batch_row0;  // first row of each batch (+ n_rows at the end); batches never span files
    std::vector<long long> batch_v0;    // values offset of each batch's first row (+ total at the end)
    std::vector<int> projection;

    void unref() {
        if (refs.fetch_sub(1) == 1) {
            cudaSetDevice(device);
            cudaFree(d_pos);
            cudaFree(d_offsets);
            cudaFree(d_values);
            cudaFree(d_voff);
            cudaFreeHost(h_pos);
            cudaFreeHost(h_offsets);
            cudaFreeHost(h_values);
            delete this;
        }
    }
};

void columns_free(VcfStream *s) {
    if (s->cols) {
        s->cols->unref();
        s->cols = nullptr;
    }
}

namespace {

struct ChildPriv {
    const void *buffers[3];
};
struct BatchPriv {
    Columns *cols;
    int n_children;
    ArrowArray children[2];
    ArrowArray *child_ptrs[2];
    ChildPriv child_priv[2];
    const void *struct_buffers[1];
};

void release_child(ArrowArray *a) { a->release = nullptr; }
void release_batch(ArrowArray *a) {
    auto *p = static_cast<BatchPriv *>(a->private_data);
    for (int i = 0; i < p->n_children; ++i)
        if (p->children[i].release) p->children[i].release(&p->children[i]);
    p->cols->unref();
    delete p;
    a->release = nullptr;
}

struct SchemaPriv {
    int n_children;
    ArrowSchema children[2];
    ArrowSchema *child_ptrs[2];
};
void release_schema_child(ArrowSchema *s) { s->release = nullptr; }
void release_schema(ArrowSchema *s) {
    auto *p = static_cast<SchemaPriv *>(s->private_data);
    for (int i = 0; i < p->n_children; ++i)
        if (p->children[i].release) p->children[i].release(&p->children[i]);
    delete p;
    s->release = nullptr;
}

// VCFSchemaBuilder (exon/exon-core/src/datasources/vcf/schema_builder.rs:85-129): chrom Utf8 !null, pos Int64 !null
void fill_schema(const std::vector<int> &projection, ArrowSchema *out) {
    auto *p = new SchemaPriv();
    p->n_children = (int)projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        ArrowSchema &c = p->children[i];
        memset(&c, 0, sizeof(c));
        c.format = projection[(size_t)i] == 0 ? "u" : "l";
        c.name = projection[(size_t)i] == 0 ? "chrom" : "pos";
        c.flags = 0;  // non-nullable
        c.release = release_schema_child;
        p->child_ptrs[i] = &c;
    }
    memset(out, 0, sizeof(*out));
    out->format = "+s";
    out->name = "";
    out->flags = 0;
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = release_schema;
    out->private_data = p;
}

int build_columns(VcfStream *s) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto *c = new (std::nothrow) Columns();
    if (!c) return fail(EXON_GPU_ERR_OOM, "next_batch: out of host memory");
    s->cols = c;
    c->device = ctx->device;
    c->on_device = s->columns_on_device;
    c->batch_rows = s->batch_rows;
    c->projection = s->projection;
    for (int p : s->projection) {
        if (p == 0) c->want_chrom = true;
        if (p == 1) c->want_pos = true;
    }
    c->batch_row0.assign(1, 0);
    // run table (empty runs dropped; run_map: stream run index -> table index of the first non-empty run at or after it)
    std::vector<RunDesc> h_runs;
    std::vector<int> run_map(s->runs.size() + 1, 0);
    int64_t n_blocks = 0;
    for (size_t i = 0; i < s->runs.size(); ++i) {
        const Run &r = s->runs[i];
        run_map[i] = (int)h_runs.size();
        if (r.len <= 0) continue;
        RunDesc d;
        d.base = r.base;
        d.len = r.len;
        d.block0 = n_blocks;
        d.row0 = 0;
        n_blocks += (int64_t)(((uintptr_t)r.base & 15) + (uint64_t)r.len + kBlockBytes - 1) / kBlockBytes;
        h_runs.push_back(d);
    }
    if (h_runs.empty()) return EXON_GPU_OK;  // zero rows
    const int n_runs = (int)h_runs.size();
    // file ends that fall inside a non-empty run (an empty file adds no rows and no batch)
    std::vector<MarkDesc> h_marks;
    for (const auto &m : s->file_marks) {
        if (m.run >= s->runs.size() || s->runs[m.run].len <= 0 || m.len <= 0) continue;
        MarkDesc d;
        d.run = run_map[m.run];
        d.pad_ = 0;
        d.len = std::min<int64_t>(m.len, s->runs[m.run].len);
        h_marks.push_back(d);
    }
    const int n_marks = (int)h_marks.size();

    RunDesc *d_runs = nullptr;
    MarkDesc *d_marks = nullptr;
    unsigned long long *d_block_rows = nullptr, *d_block_row0 = nullptr, *d_misc = nullptr;
    long long *d_small = nullptr;  // mark rows | file_row0 | file_batch0 | batch_row0 | batch_v0
    const uint8_t **d_line_ptr = nullptr;
    int32_t *d_chrom_len = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&]() {
        cudaFree(d_runs);
        cudaFree(d_marks);
        cudaFree(d_block_rows);
        cudaFree(d_block_row0);
        cudaFree(d_misc);
        cudaFree(d_small);
        cudaFree((void *)d_line_ptr);
        cudaFree(d_chrom_len);
        cudaFree(d_tmp);
    };
#define TRY_OR_CLEAN(expr)                                                                                    \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            cleanup();                                                                                        \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", #expr, \
                        cudaGetErrorString(_e));                                                              \
        }                                                                                                     \
    } while (0)

    TRY_OR_CLEAN(cudaMalloc((void **)&d_runs, sizeof(RunDesc) * (size_t)n_runs));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_block_rows, sizeof(unsigned long long) * (size_t)(n_blocks + 1)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_block_row0, sizeof(unsigned long long) * (size_t)(n_blocks + 1)));
    TRY_OR_CLEAN(cudaMalloc((void **)&d_misc, 4 * sizeof(unsigned long long)));
    TRY_OR_CLEAN(cudaMemcpyAsync(d_runs, h_runs.data(), sizeof(RunDesc) * (size_t)n_runs, cudaMemcpyHostToDevice, st));
    TRY_OR_CLEAN(cudaMemsetAsync(d_block_rows + n_blocks, 0, sizeof(unsigned long long), st));

    const int grid_blocks = (int)std::min<int64_t>(n_blocks, (int64_t)ctx->sm_count * 8);
    count_lines<<<grid_blocks, kThreads, 0, st>>>(d_runs, n_runs, n_blocks, d_block_rows);
    ctx->launches.fetch_add(1);
    TRY_OR_CLEAN(cudaGetLastError());
    TRY_OR_CLEAN(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_block_rows, d_block_row0, (int)(n_blocks + 1), st));
    size_t tmp_cap = tmp_bytes;
    TRY_OR_CLEAN(cudaMalloc(&d_tmp, tmp_cap));
    TRY_OR_CLEAN(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_block_rows, d_block_row0, (int)(n_blocks + 1), st));
    ctx->launches.fetch_add(1);

    // rows per run + total rows come back to the host (planning data, 8 bytes per run)
    std::vector<unsigned long long> h_row0((size_t)n_runs + 1);
    for (int r = 0; r < n_runs; ++r)
        TRY_OR_CLEAN(cudaMemcpyAsync(&h_row0[(size_t)r], d_block_row0 + h_runs[(size_t)r].block0, sizeof(unsigned long long),
                                     cudaMemcpyDeviceToHost, st));
    TRY_OR_CLEAN(cudaMemcpyAsync(&h_row0[(size_t)n_runs], d_block_row0 + n_blocks, sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, st));
    TRY_OR_CLEAN(cudaStreamSynchronize(st));
    const int64_t n_rows = (int64_t)h_row0[(size_t)n_runs];
    for (int r = 0; r < n_runs; ++r) h_runs[(size_t)r].row0 = (int64_t)h_row0[(size_t)r];
    TRY_OR_CLEAN(cudaMemcpyAsync(d_runs, h_runs.data(), sizeof(RunDesc) * (size_t)n_runs, cudaMemcpyHostToDevice, st));
    c->n_rows = n_rows;
    if (n_rows == 0) {
        cleanup();
        return EXON_GPU_OK;
    }

    TRY_OR_CLEAN(cudaMalloc((void **)&d_line_ptr, sizeof(uint8_t *) * (size_t)n_rows));
    index_lines<<<grid_blocks, kThreads, 0, st>>>(d_runs, n_runs, n_blocks, d_block_row0, d_line_ptr);
    ctx->launches.fetch_add(1);
    TRY_OR_CLEAN(cudaGetLastError());

    // ---- file table -> batch table (batches restart at every file) ----
    std::vector<long long> file_row0{0};
    if (n_marks) {
        TRY_OR_CLEAN(cudaMalloc((void **)&d_marks, sizeof(MarkDesc) * (size_t)n_marks));
        TRY_OR_CLEAN(cudaMalloc((void **)&d_small, sizeof(long long) * (size_t)n_marks));
        TRY_OR_CLEAN(cudaMemcpyAsync(d_marks, h_marks.data(), sizeof(MarkDesc) * (size_t)n_marks, cudaMemcpyHostToDevice, st));
        mark_rows<<<(n_marks + 127) / 128, 128, 0, st>>>(d_runs, n_runs, d_marks, n_marks, d_line_ptr, n_rows, d_small);
        ctx->launches.fetch_add(1);
        TRY_OR_CLEAN(cudaGetLastError());
        std::vector<long long> h_mark_rows((size_t)n_marks);
        TRY_OR_CLEAN(cudaMemcpyAsync(h_mark_rows.data(), d_small, sizeof(long long) * (size_t)n_marks, cudaMemcpyDeviceToHost, st));
        TRY_OR_CLEAN(cudaStreamSynchronize(st));
        TRY_OR_CLEAN(cudaFree(d_small));
        d_small = nullptr;
        for (long long r : h_mark_rows)
            if (r > file_row0.back() && r < n_rows) file_row0.push_back(r);
    }
    file_row0.push_back(n_rows);
    const int n_files = (int)file_row0.size() - 1;
    std::vector<long long> file_batch0((size_t)n_files + 1, 0);
    c->batch_row0.clear();
    for (int f = 0; f < n_files; ++f) {
        file_batch0[(size_t)f] = (long long)c->batch_row0.size();
        for (long long r = file_row0[(size_t)f]; r < file_row0[(size_t)f + 1]; r += c->batch_rows) c->batch_row0.push_back(r);
    }
    file_batch0[(size_t)n_files] = (long long)c->batch_row0.size();
    c->n_batches = (int64_t)c->batch_row0.size();
    c->batch_row0.push_back(n_rows);

    if (c->want_chrom) TRY_OR_CLEAN(cudaMalloc((void **)&d_chrom_len, sizeof(int32_t) * (size_t)(n_rows + 1)));
    if (c->want_pos) TRY_OR_CLEAN(cudaMalloc((void **)&c->d_pos, sizeof(int64_t) * (size_t)n_rows));
    const unsigned long long init_misc[4] = {0ull, ~0ull, 0ull, 0ull};
    TRY_OR_CLEAN(cudaMemcpyAsync(d_misc, init_misc, sizeof(init_misc), cudaMemcpyHostToDevice, st));
    const unsigned row_grid = (unsigned)((n_rows + kThreads - 1) / kThreads);
    parse_rows<<<row_grid, kThreads, 0, st>>>(d_runs, n_runs, n_rows, d_line_ptr, c->want_chrom, c->want_pos, d_chrom_len,
                                              c->d_pos, reinterpret_cast<uint32_t *>(d_misc), d_misc + 1);
    ctx->launches.fetch_add(1);
    TRY_OR_CLEAN(cudaGetLastError());

    long long total_values = 0;
    long long *d_file_row0 = nullptr, *d_file_batch0 = nullptr, *d_batch_row0 = nullptr, *d_batch_v0 = nullptr;
    if (c->want_chrom) {
        TRY_OR_CLEAN(cudaMemsetAsync(d_chrom_len + n_rows, 0, sizeof(int32_t), st));
        TRY_OR_CLEAN(cudaMalloc((void **)&c->d_voff, sizeof(long long) * (size_t)(n_rows + 1)));
        TRY_OR_CLEAN(cub::DeviceScan::ExclusiveScan(nullptr, tmp_bytes, d_chrom_len, c->d_voff, cub::Sum(), 0ll, (int)(n_rows + 1), st));
        if (tmp_bytes > tmp_cap) {
            TRY_OR_CLEAN(cudaFree(d_tmp));
            d_tmp = nullptr;
            tmp_cap = tmp_bytes;
            TRY_OR_CLEAN(cudaMalloc(&d_tmp, tmp_cap));
        }
        TRY_OR_CLEAN(cub::DeviceScan::ExclusiveScan(d_tmp, tmp_bytes, d_chrom_len, c->d_voff, cub::Sum(), 0ll, (int)(n_rows + 1), st));
        ctx->launches.fetch_add(1);
        // planning tables on the device: file_row0 | file_batch0 | batch_row0 | batch_v0
        const size_t nf = (size_t)n_files + 1, nb = (size_t)c->n_batches + 1;
        TRY_OR_CLEAN(cudaMalloc((void **)&d_small, sizeof(long long) * (2 * nf + 2 * nb)));
        d_file_row0 = d_small;
        d_file_batch0 = d_small + nf;
        d_batch_row0 = d_small + 2 * nf;
        d_batch_v0 = d_small + 2 * nf + nb;
        TRY_OR_CLEAN(cudaMemcpyAsync(d_file_row0, file_row0.data(), sizeof(long long) * nf, cudaMemcpyHostToDevice, st));
        TRY_OR_CLEAN(cudaMemcpyAsync(d_file_batch0, file_batch0.data(), sizeof(long long) * nf, cudaMemcpyHostToDevice, st));
        TRY_OR_CLEAN(cudaMemcpyAsync(d_batch_row0, c->batch_row0.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, st));
        gather_i64<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(c->d_voff, d_batch_row0, (int64_t)nb, d_batch_v0);
        ctx->launches.fetch_add(1);
        TRY_OR_CLEAN(cudaGetLastError());
        c->batch_v0.resize(nb);
        TRY_OR_CLEAN(cudaMemcpyAsync(c->batch_v0.data(), d_batch_v0, sizeof(long long) * nb, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_misc[4];
    TRY_OR_CLEAN(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    TRY_OR_CLEAN(cudaStreamSynchronize(st));
    if ((uint32_t)h_misc[0]) {
        cleanup();
        return fail(EXON_GPU_ERR_PARSE, "malformed VCF record at row %llu:%s%s", h_misc[1],
                    ((uint32_t)h_misc[0] & kErrBadPos) ? " POS is not a positive decimal integer;" : "",
                    ((uint32_t)h_misc[0] & kErrShortLine) ? " line ended before the field being read;" : "");
    }
    if (c->want_chrom) {
        total_values = c->batch_v0[(size_t)c->n_batches];
        for (int64_t b = 0; b < c->n_batches; ++b) {
            const long long nb = c->batch_v0[(size_t)b + 1] - c->batch_v0[(size_t)b];
            if (nb > 0x7FFFFFFFll) {
                cleanup();
                return fail(EXON_GPU_ERR_UNSUPPORTED, "chrom bytes of batch %lld overflow int32 offsets", (long long)b);
            }
        }
        TRY_OR_CLEAN(cudaMalloc((void **)&c->d_values, (size_t)std::max<long long>(total_values, 1)));
        TRY_OR_CLEAN(cudaMalloc((void **)&c->d_offsets, sizeof(int32_t) * (size_t)(c->n_batches * (c->batch_rows + 1))));
        gather_chrom<<<row_grid, kThreads, 0, st>>>(n_rows, c->batch_rows, n_files, d_file_row0, d_file_batch0, d_line_ptr,
                                                    d_chrom_len, c->d_voff, c->d_values, c->d_offsets);
        ctx->launches.fetch_add(1);
        TRY_OR_CLEAN(cudaGetLastError());
    }
    if (!c->on_device) {
        if (c->want_pos) {
            TRY_OR_CLEAN(cudaHostAlloc((void **)&c->h_pos, sizeof(int64_t) * (size_t)n_rows, cudaHostAllocDefault));
            TRY_OR_CLEAN(cudaMemcpyAsync(c->h_pos, c->d_pos, sizeof(int64_t) * (size_t)n_rows, cudaMemcpyDeviceToHost, st));
        }
        if (c->want_chrom) {
            const size_t ob = sizeof(int32_t) * (size_t)(c->n_batches * (c->batch_rows + 1));
            TRY_OR_CLEAN(cudaHostAlloc((void **)&c->h_offsets, ob, cudaHostAllocDefault));
            TRY_OR_CLEAN(cudaHostAlloc((void **)&c->h_values, (size_t)std::max<long long>(total_values, 1), cudaHostAllocDefault));
            TRY_OR_CLEAN(cudaMemcpyAsync(c->h_offsets, c->d_offsets, ob, cudaMemcpyDeviceToHost, st));
            TRY_OR_CLEAN(cudaMemcpyAsync(c->h_values, c->d_values, (size_t)total_values, cudaMemcpyDeviceToHost, st));
        }
    }
    TRY_OR_CLEAN(cudaStreamSynchronize(st));
    cleanup();
    return EXON_GPU_OK;
#undef TRY_OR_CLEAN
}

}  // namespace

int columns_next_batch(VcfStream *s, ArrowArray *out, ArrowSchema *out_schema) {
    if (!s->cols) {
        if (s->file_open && s->tail_len > 0)
            return fail(EXON_GPU_ERR_STATE, "next_batch: the current file ends mid-line; finish it with is_last first");
        if (int rc = build_columns(s)) {
            columns_free(s);
            return rc;
        }
        s->drained = true;
    }
    Columns *c = s->cols;
    if (out_schema) fill_schema(s->projection, out_schema);
    memset(out, 0, sizeof(*out));
    if (c->next >= c->n_batches) return EXON_GPU_OK;  // end of stream: release == NULL
    const int64_t b = c->next++;
    const int64_t row0 = c->batch_row0[(size_t)b];
    const int64_t rows = c->batch_row0[(size_t)b + 1] - row0;
    auto *p = new BatchPriv();
    p->cols = c;
    c->refs.fetch_add(1);
    p->n_children = (int)s->projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        ArrowArray &a = p->children[i];
        memset(&a, 0, sizeof(a));
        a.length = rows;
        a.null_count = 0;
        a.offset = 0;
        ChildPriv &cp = p->child_priv[i];
        cp.buffers[0] = nullptr;  // no validity bitmap: both columns are non-nullable
        if (s->projection[(size_t)i] == 0) {
            const int32_t *off = (c->on_device ? c->d_offsets : c->h_offsets) + b * (c->batch_rows + 1);
            const uint8_t *val = (c->on_device ? c->d_values : c->h_values) + c->batch_v0[(size_t)b];
            cp.buffers[1] = off;
            cp.buffers[2] = val;
            a.n_buffers = 3;
        } else {
            cp.buffers[1] = (c->on_device ? c->d_pos : c->h_pos) + row0;
            a.n_buffers = 2;
        }
        a.buffers = cp.buffers;
        a.release = release_child;
        p->child_ptrs[i] = &a;
    }
    p->struct_buffers[0] = nullptr;
    out->length = rows;
    out->null_count = 0;
    out->offset = 0;
    out->n_buffers = 1;
    out->buffers = p->struct_buffers;
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = release_batch;
    out->private_data = p;
    return EXON_GPU_OK;
}

}  // namespace exon

extern "C" int exon_gpu_vcf_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema) {
    if (!s || !out) return exon::fail(EXON_GPU_ERR_ARG, "vcf_next_batch: NULL argument");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return exon::fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return exon::columns_next_batch(s, out, out_schema);
}
