// fasta_columns.cu -- FASTA text -> Arrow columns {id, description, sequence} in reference-sized batches
// (exon_gpu_fasta_next_batch).
//
// Replaces BatchReader::{read_record, read_batch} (exon/exon-fasta/src/batch_reader.rs:52-103: noodles-fasta
// `read_definition` + `read_sequence` per record) and FASTAArrayBuilder::{append, finish} for the Utf8 sequence type
// (exon/exon-fasta/src/array_builder.rs:108-160): the definition line is parsed by noodles `Definition::from_str` -- '>',
// name up to the first ASCII whitespace (required), description = the rest, trimmed, None when there is no whitespace --
// and the sequence is every line up to the next definition with its line terminator ("\n" or "\r\n") removed.
// Schema: exon/exon-fasta/src/config.rs:162-226 (id !null, description nullable, sequence !null).
//
// Line-parallel on the partition's line index (build_line_index, fastq_scan.cu), so a chromosome-sized record is copied by
// as many threads as it has lines:
//   1. measure  one thread per line: definition or sequence line, lengths of name / description / sequence bytes
//   2. 4 exclusive scans over lines: record index, byte offsets of the three columns
//   3. records  definition line of every record (scatter), per-file record counts -> batch table (batches restart at every file)
//   4. emit     one thread per line: a sequence line copies its bytes to their final place; a definition line writes its
//               record's name / description bytes, the three batch-relative offsets and the description validity bit
// A file whose first line is not a definition, a definition without a name, and a record without any sequence line are
// errors, as in the reference ("invalid definition" / "invalid sequence").
// The column store is the FASTQ one (FqColumns: up to four utf8 columns); fastq_next_batch exports the batches.
#include <algorithm>
#include <cstring>
#include <new>

#include "common.cuh"
#include "internal.h"
#include "scan_i64.cuh"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

constexpr uint32_t kFaErrFirst = 1u;  // a file does not begin with a definition line
constexpr uint32_t kFaErrName = 2u;   // '>' followed by whitespace or nothing
constexpr uint32_t kFaErrSeq = 4u;    // a definition without any sequence line

enum { kFaRec = 0, kFaName = 1, kFaDesc = 2, kFaSeq = 3, kFaN = 4 };

struct FaColArgs {
    int64_t n_lines;
    const uint8_t *const *line_start;
    const uint8_t *const *line_end;
    int32_t *cnt[kFaN];
    const long long *pre[kFaN];
    uint8_t *lflags;            // bit0 definition line, bit1 description present
    long long *rec_line;        // n_records + 1: definition line of every record, then n_lines
    const long long *file_line0;  // n_files + 1
    int32_t n_files;
    const long long *brow;      // n_batches + 1 (records)
    int64_t n_batches, n_records;
    int32_t batch_rows, wpb;
    int32_t *off[3];            // id, description, sequence: n_batches * (batch_rows + 1)
    uint8_t *val[3];
    uint32_t *desc_valid;
    uint32_t *flags;
};

// Rust char::is_ascii_whitespace: space, \t, \n, \x0C, \r
__device__ __forceinline__ bool ascii_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == 0x0C || c == '\r'; }
// str::trim on ASCII input: the above plus \x0B
__device__ __forceinline__ bool trim_ws(uint8_t c) { return ascii_ws(c) || c == 0x0B; }

struct FaLine {
    const uint8_t *s, *e;  // line without its terminator
    bool def;
    const uint8_t *name_e;         // definition: end of the name
    const uint8_t *desc_s, *desc_e;  // definition: trimmed description (desc_s == nullptr: none)
};
__device__ __forceinline__ FaLine fa_line(const uint8_t *s, const uint8_t *e) {
    FaLine L;
    if (e > s && e[-1] == '\r') --e;
    L.s = s;
    L.e = e;
    L.def = e > s && s[0] == '>';
    L.name_e = L.desc_s = L.desc_e = nullptr;
    if (L.def) {
        const uint8_t *p = s + 1;
        while (p < e && !ascii_ws(*p)) ++p;
        L.name_e = p;
        if (p < e) {
            const uint8_t *a = p + 1, *b = e;
            while (a < b && trim_ws(*a)) ++a;
            while (b > a && trim_ws(b[-1])) --b;
            L.desc_s = a;
            L.desc_e = b;
        }
    }
    return L;
}

__global__ void __launch_bounds__(256) fa_measure_kernel(const __grid_constant__ FaColArgs a) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_lines) return;
    const FaLine L = fa_line(a.line_start[i], a.line_end[i]);
    int32_t c[kFaN] = {0, 0, 0, 0};
    uint8_t f = 0;
    if (L.def) {
        f = 1;
        c[kFaRec] = 1;
        c[kFaName] = (int32_t)(L.name_e - (L.s + 1));
        if (c[kFaName] == 0) atomicOr(a.flags, kFaErrName);
        if (L.desc_s) {
            f |= 2;
            c[kFaDesc] = (int32_t)(L.desc_e - L.desc_s);
        }
    } else {
        const long long n = L.e - L.s;
        c[kFaSeq] = (int32_t)(n > 0x7FFFFFFFll ? 0x7FFFFFFFll : n);
    }
#pragma unroll
    for (int k = 0; k < kFaN; ++k) a.cnt[k][i] = c[k];
    a.lflags[i] = f;
}

// rec_line[record] = its definition line; the first line of every file must be one, and every definition needs a sequence line
__global__ void __launch_bounds__(256) fa_records_kernel(const __grid_constant__ FaColArgs a) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_lines) return;
    int lo = 0, hi = a.n_files - 1;  // file of the line
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.file_line0[mid] <= i) lo = mid;
        else hi = mid - 1;
    }
    const bool def = a.lflags[i] & 1u;
    if (i == a.file_line0[lo] && !def) atomicOr(a.flags, kFaErrFirst);
    if (!def) return;
    a.rec_line[a.pre[kFaRec][i]] = i;
    if (i + 1 >= a.file_line0[lo + 1] || (a.lflags[i + 1] & 1u)) atomicOr(a.flags, kFaErrSeq);
    if (i == 0) a.rec_line[a.n_records] = a.n_lines;  // sentinel (line 0 of a non-empty partition is a definition or an error)
}

__global__ void __launch_bounds__(256) fa_emit_kernel(const __grid_constant__ FaColArgs a) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_lines) return;
    const FaLine L = fa_line(a.line_start[i], a.line_end[i]);
    if (!L.def) {
        if (a.val[2]) {
            uint8_t *dst = a.val[2] + a.pre[kFaSeq][i];
            const long long n = L.e - L.s;
            for (long long j = 0; j < n; ++j) dst[j] = L.s[j];
        }
        return;
    }
    const long long rec = a.pre[kFaRec][i];
    int64_t lo = 0, hi = a.n_batches;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (a.brow[mid] <= rec) lo = mid;
        else hi = mid;
    }
    const int64_t b = lo, r0 = a.brow[b];
    const int in_batch = (int)(rec - r0);
    const bool last = rec + 1 == a.brow[b + 1];
    const long long l0 = a.rec_line[r0];                         // definition line of the batch's first record
    const long long l1 = last ? a.rec_line[rec + 1] : 0;         // ... of the record after this one (n_lines at the end)
    const int64_t row = b * (int64_t)(a.batch_rows + 1) + in_batch;
    const int pk[3] = {kFaName, kFaDesc, kFaSeq};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!a.off[k]) continue;
        const long long *P = a.pre[pk[k]];
        a.off[k][row] = (int32_t)(P[i] - P[l0]);
        if (last) a.off[k][row + 1] = (int32_t)(P[l1] - P[l0]);
    }
    if (a.val[0]) {
        uint8_t *dst = a.val[0] + a.pre[kFaName][i];
        const int n = (int)(L.name_e - (L.s + 1));
        for (int j = 0; j < n; ++j) dst[j] = L.s[1 + j];
    }
    if (a.val[1] && L.desc_s) {
        uint8_t *dst = a.val[1] + a.pre[kFaDesc][i];
        const int n = (int)(L.desc_e - L.desc_s);
        for (int j = 0; j < n; ++j) dst[j] = L.desc_s[j];
        atomicOr(a.desc_valid + b * a.wpb + (in_batch >> 5), 1u << (in_batch & 31));
    }
}

// out[i] = src[idx[i]]
__global__ void fa_gather(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
// out[i] = src[rec_line[idx[i]]]
__global__ void fa_gather2(const long long *src, const long long *rec_line, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[rec_line[idx[i]]];
}

size_t fal256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// The caller (fastq_next_batch) holds ctx->work_mu and owns the error path (fq_columns_free).
int fasta_build_columns(VcfStream *s) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto *c = new (std::nothrow) FqColumns();
    if (!c) return fail(EXON_GPU_ERR_OOM, "fasta_next_batch: out of host memory");
    s->fq_cols = c;
    c->device = ctx->device;
    c->on_device = s->columns_on_device;
    c->batch_rows = s->batch_rows;
    c->words_per_batch = (s->batch_rows + 31) / 32;
    c->projection = s->projection;
    bool want[3] = {false, false, false};
    for (int p : s->projection) want[p] = true;
    c->batch_row0.assign(1, 0);

    const size_t per_line = kFaN * 4 + kFaN * 8 + 1 + 8;  // counts | prefixes | flags | rec_line (records <= lines)
    LineIndex li;
    if (int rc = build_line_index(s, per_line, (2 * kFaN + 8) * 256 + (1 << 20), &li)) return rc;
    const int64_t n_lines = li.n_lines;
    if (n_lines == 0) return EXON_GPU_OK;
    const size_t nl1 = (size_t)n_lines + 1;
    const int n_files = (int)li.file_line0.size() - 1;
    size_t cub_bytes = 0;
    CUDA_TRY(exclusive_sum_i32_i64(nullptr, cub_bytes, (const int32_t *)nullptr, (long long *)nullptr, (int)nl1, st));
    uint8_t *x = li.extra;
    auto take = [&](size_t bytes) {
        uint8_t *p = x;
        x += fal256(bytes);
        return p;
    };
    FaColArgs a;
    memset(&a, 0, sizeof(a));
    a.n_lines = n_lines;
    a.line_start = li.line_start;
    a.line_end = li.line_end;
    long long *pre[kFaN];
    for (int k = 0; k < kFaN; ++k) {
        a.cnt[k] = (int32_t *)take(nl1 * 4);
        pre[k] = (long long *)take(nl1 * 8);
        a.pre[k] = pre[k];
        CUDA_TRY(cudaMemsetAsync(a.cnt[k] + n_lines, 0, 4, st));
    }
    a.lflags = take(nl1);
    a.rec_line = (long long *)take(nl1 * 8);
    // scan scratch grows with the row count (64-bit tile states): from the pool, not from the fixed part of scratch_b
    uint8_t *cub_tmp = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&cub_tmp, cub_bytes + 256, st));
    struct CubFree {
        void *p;
        cudaStream_t st;
        ~CubFree() { cudaFreeAsync(p, st); }
    } cub_guard{cub_tmp, st};
    uint32_t *d_flags = (uint32_t *)take(64);
    CUDA_TRY(cudaMemsetAsync(d_flags, 0, 64, st));
    CUDA_TRY(cudaMemsetAsync(a.lflags + n_lines, 0, 1, st));
    a.flags = d_flags;
    // small tables from the pool: file_line0 | per-file record prefix | batch table | 3 batch bases
    const size_t nf1 = (size_t)n_files + 1;
    long long *d_small = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&d_small, 2 * fal256(nf1 * 8), st));
    struct PoolFree {
        void *p;
        cudaStream_t st;
        ~PoolFree() {
            if (p) cudaFreeAsync(p, st);
        }
    } g0{d_small, st}, g1{nullptr, st};
    long long *d_fl0 = d_small, *d_frec = reinterpret_cast<long long *>(reinterpret_cast<uint8_t *>(d_small) + fal256(nf1 * 8));
    CUDA_TRY(cudaMemcpyAsync(d_fl0, li.file_line0.data(), nf1 * 8, cudaMemcpyHostToDevice, st));
    a.file_line0 = d_fl0;
    a.n_files = n_files;

    const unsigned grid = (unsigned)((n_lines + 255) / 256);
    fa_measure_kernel<<<grid, 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    for (int k = 0; k < kFaN; ++k) {
        size_t tb = cub_bytes;
        CUDA_TRY(exclusive_sum_i32_i64(cub_tmp, tb, (const int32_t *)a.cnt[k], pre[k], (int)nl1, st));
    }
    fa_gather<<<(unsigned)((nf1 + 255) / 256), 256, 0, st>>>(pre[kFaRec], d_fl0, (int64_t)nf1, d_frec);
    ctx->launches.fetch_add(6);
    std::vector<long long> frec(nf1);
    CUDA_TRY(cudaMemcpyAsync(frec.data(), d_frec, nf1 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int64_t n_records = frec[(size_t)n_files];
    c->n_rows = n_records;
    a.n_records = n_records;
    // batches restart at every file
    c->batch_row0.clear();
    for (int f = 0; f < n_files; ++f)
        for (long long r = frec[(size_t)f]; r < frec[(size_t)f + 1]; r += c->batch_rows) c->batch_row0.push_back(r);
    c->n_batches = (int64_t)c->batch_row0.size();
    c->batch_row0.push_back(n_records);
    const size_t nb1 = (size_t)c->n_batches + 1;
    long long *d_b = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&d_b, 4 * fal256(nb1 * 8), st));
    g1.p = d_b;
    auto d_bx = [&](int k) { return reinterpret_cast<long long *>(reinterpret_cast<uint8_t *>(d_b) + (size_t)k * fal256(nb1 * 8)); };
    CUDA_TRY(cudaMemcpyAsync(d_bx(0), c->batch_row0.data(), nb1 * 8, cudaMemcpyHostToDevice, st));
    a.brow = d_bx(0);
    a.n_batches = c->n_batches;
    a.batch_rows = c->batch_rows;
    a.wpb = c->words_per_batch;
    fa_records_kernel<<<grid, 256, 0, st>>>(a);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    uint32_t h_flags = 0;
    CUDA_TRY(cudaMemcpyAsync(&h_flags, d_flags, 4, cudaMemcpyDeviceToHost, st));
    const int pk[3] = {kFaName, kFaDesc, kFaSeq};
    const int col_slot[3] = {0, 1, 2};  // FqColumns slot of id / description / sequence
    for (int k = 0; k < 3; ++k) {
        if (!want[k] || n_records == 0) continue;
        fa_gather2<<<(unsigned)((nb1 + 255) / 256), 256, 0, st>>>(pre[pk[k]], a.rec_line, d_bx(0), (int64_t)nb1, d_bx(1 + k));
        c->batch_v0[col_slot[k]].resize(nb1);
        CUDA_TRY(cudaMemcpyAsync(c->batch_v0[col_slot[k]].data(), d_bx(1 + k), nb1 * 8, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (h_flags)
        return fail(EXON_GPU_ERR_PARSE, "malformed FASTA:%s%s%s", (h_flags & kFaErrFirst) ? " a file does not begin with a definition line ('>');" : "",
                    (h_flags & kFaErrName) ? " a definition without a name;" : "", (h_flags & kFaErrSeq) ? " a definition without a sequence;" : "");
    if (n_records == 0) return EXON_GPU_OK;
    const size_t off_elems = (size_t)c->n_batches * (size_t)(c->batch_rows + 1);
    size_t total[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        if (!want[k]) continue;
        const int slot = col_slot[k];
        total[k] = (size_t)c->batch_v0[slot][(size_t)c->n_batches];
        for (int64_t b = 0; b < c->n_batches; ++b)
            if (c->batch_v0[slot][(size_t)b + 1] - c->batch_v0[slot][(size_t)b] > 0x7FFFFFFFll)
                return fail(EXON_GPU_ERR_UNSUPPORTED, "fasta: the %s bytes of batch %lld overflow the int32 offsets of a Utf8 column (the reference offers LargeUtf8 for this)",
                            k == 2 ? "sequence" : "definition", (long long)b);
        CUDA_TRY(cudaMallocAsync((void **)&c->d_values[slot], std::max<size_t>(total[k], 1), st));
        CUDA_TRY(cudaMallocAsync((void **)&c->d_offsets[slot], off_elems * 4, st));
        a.val[k] = c->d_values[slot];
        a.off[k] = c->d_offsets[slot];
    }
    const size_t valid_bytes = (size_t)c->n_batches * (size_t)c->words_per_batch * 4;
    if (want[1]) {
        CUDA_TRY(cudaMallocAsync((void **)&c->d_valid, valid_bytes, st));
        CUDA_TRY(cudaMemsetAsync(c->d_valid, 0, valid_bytes, st));
        a.desc_valid = c->d_valid;
    }
    fa_emit_kernel<<<grid, 256, 0, st>>>(a);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    if (!c->on_device) {
        for (int k = 0; k < 3; ++k) {
            if (!want[k]) continue;
            const int slot = col_slot[k];
            CUDA_TRY(cudaHostAlloc((void **)&c->h_values[slot], std::max<size_t>(total[k], 1), cudaHostAllocDefault));
            CUDA_TRY(cudaHostAlloc((void **)&c->h_offsets[slot], off_elems * 4, cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(c->h_values[slot], c->d_values[slot], total[k], cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(c->h_offsets[slot], c->d_offsets[slot], off_elems * 4, cudaMemcpyDeviceToHost, st));
        }
        if (want[1]) {
            CUDA_TRY(cudaHostAlloc((void **)&c->h_valid, valid_bytes, cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(c->h_valid, c->d_valid, valid_bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_fasta_open_columns(exon_gpu_ctx *c, const exon_gpu_fastq_opts *o, exon_gpu_stream **out) {
    if (!c || !o || !out) return fail(EXON_GPU_ERR_ARG, "fasta_open_columns: NULL argument");
    if (o->batch_rows < 0 || o->n_projection < 0 || o->n_projection > 3 || (o->n_projection > 0 && !o->projection))
        return fail(EXON_GPU_ERR_ARG, "fasta_open_columns: bad batch_rows / projection");
    for (int i = 0; i < o->n_projection; ++i) {
        if (o->projection[i] < 0 || o->projection[i] > 2) return fail(EXON_GPU_ERR_ARG, "fasta_open_columns: projection index %d is not a FASTA file-schema column", o->projection[i]);
        for (int j = 0; j < i; ++j)
            if (o->projection[j] == o->projection[i]) return fail(EXON_GPU_ERR_ARG, "fasta_open_columns: column %d is projected twice", o->projection[i]);
    }
    if (int rc = exon_gpu_fasta_open(c, out)) return rc;
    if (o->batch_rows > 0) (*out)->batch_rows = o->batch_rows;
    (*out)->projection.assign(o->projection, o->projection + o->n_projection);
    (*out)->columns_on_device = o->columns_on_device != 0;
    return EXON_GPU_OK;
}

int exon_gpu_fasta_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema) {
    if (!s || !out || s->fmt != kFmtFasta) return fail(EXON_GPU_ERR_ARG, "fasta_next_batch: not a FASTA stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return fastq_next_batch(s, out, out_schema);
}

}  // extern "C"
