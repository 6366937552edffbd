// gff_columns.cu -- GFF text -> Arrow columns 0..7 {seqname, source, type, start, end, score, strand, phase}
// (exon_gpu_gff_next_batch).
//
// Replaces BatchReader::{read_line, read_batch} (exon/exon-gff/src/batch_reader.rs:56-130) and GFFArrayBuilder::{append,
// finish} (exon/exon-gff/src/array_builder.rs:84-200) over noodles-gff lazy records; schema exon/exon-gff/src/config.rs:81-108.
// Reference behaviour that is kept on purpose:
//   * read_batch has no row limit: ONE batch per file
//   * score "." -> NULL, else Rust f32::from_str; phase "." -> NULL, else "0" | "1" | "2"
//   * strand "+" | "-" as text; "." and "?" are appended as NULL into a column the schema declares NON-nullable, so the
//     reference's batch construction fails for such a file when strand is projected -- reported as EXON_GPU_ERR_PARSE here
//   * "##" directives and "#" comments are not records; an empty line and a line with fewer than 9 fields are errors
//   * attributes (column 8, Map<Utf8, List<Utf8>>): noodles-gff 0.41 lazy attributes -- `key=value` fields split at ';', a
//     value that holds ',' is an array (split at ','), keys and values percent-decoded -- appended exactly as
//     GFFArrayBuilder::append does (array_builder.rs:142-160), INCLUDING its off-by-one: for a plain string value the builder
//     closes the value list BEFORE it appends the string (`values().append(true)` precedes `values().values().append_value`),
//     so every string value lands in the list of the NEXT entry of the batch (the first list is empty, the last string
//     dangles behind the last offset); array values are appended before their list is closed and so stay -- together with a
//     string left pending by the entry before them.  What a DataFusion query sees through the reference is what is built here.
//
// Row-parallel on the partition's line index (build_line_index, fastq_scan.cu):
//   1. measure  one thread per line: record or not, the eight tab offsets (aligned 8-byte SWAR), byte lengths of the five
//               string columns, start / end parsed, validity flags; scores that are not short unsigned integers are left to
//               the exact parser's own kernel (gff_score_kernel), as in vcf_wide.cu
//   2. 6 exclusive scans over lines: row index and the five byte offsets
//   3. emit     one thread per record line: offsets relative to the file's first row, bytes, values, validity bits
#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>

#include "common.cuh"
#include "f32_parse.cuh"
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

constexpr uint32_t kGErrFields = 1u, kGErrEmptyLine = 2u, kGErrPos = 4u, kGErrScore = 8u, kGErrStrand = 16u, kGErrPhase = 32u,
                   kGErrStrandNull = 64u, kGErrSeqname = 128u, kGErrAttr = 256u;

// scan slots: rows, the five string columns' bytes, and for attributes: entries, key bytes, value strings, value bytes
enum { kGRow = 0, kGSeq = 1, kGSrc = 2, kGType = 3, kGStrand = 4, kGPhase = 5, kGAttrN = 6, kGKeyB = 7, kGStrN = 8, kGStrB = 9, kGN = 10 };
constexpr int kGStrCols = 6;  // slots [kGSeq, kGStrCols) are plain string columns

struct GffColArgs {
    int64_t n_lines;
    const uint8_t *const *line_start;
    const uint8_t *const *line_end;
    int32_t want[9];
    int32_t *cnt[kGN];
    const long long *pre[kGN];
    uint8_t *lflags;  // bit0 record, bit1 score valid, bit2 phase valid, bit3 score pending (exact parser)
    const long long *file_line0;  // n_files + 1
    int32_t n_files;
    const long long *brow;  // n_batches + 1: first row of every batch (= file)
    const long long *bline;  // n_batches + 1: first line of every batch's file
    int64_t n_batches;
    const long long *bit0;  // n_batches + 1: first validity WORD of every batch
    long long *start, *end;
    float *score;
    int32_t *off[kGN];  // [kGSeq..kGPhase]: n_rows + n_batches entries, batch b's offsets start at brow[b] + b
    uint8_t *val[kGN];
    uint32_t *score_valid, *phase_valid;
    // attributes: map offsets (index row + batch), key offsets / list offsets (index entry + batch), string offsets (index string + batch)
    int32_t *map_off, *key_off, *list_off, *str_off;
    uint8_t *key_val, *str_val;
    uint32_t *flags;
    unsigned long long *misc;  // [1] first bad line, [2] scores left to the exact parser
};

__device__ __forceinline__ int gff_hex(uint32_t c) {
    if (c - '0' <= 9u) return (int)(c - '0');
    c |= 0x20u;
    return c - 'a' <= 5u ? (int)(c - 'a' + 10) : -1;
}
// percent-decoded copy of [p, p + n) (noodles percent_decode); returns the decoded length, writes when dst is set
__device__ __forceinline__ int32_t gff_pct(const uint8_t *p, int32_t n, uint8_t *dst) {
    int32_t o = 0;
    for (int32_t q = 0; q < n; ++q) {
        uint32_t c = __ldg(p + q);
        if (c == '%' && q + 2 < n) {
            const int h = gff_hex(__ldg(p + q + 1)), l = gff_hex(__ldg(p + q + 2));
            if (h >= 0 && l >= 0) {
                c = (uint32_t)(h * 16 + l);
                q += 2;
            }
        }
        if (dst) dst[o] = (uint8_t)c;
        ++o;
    }
    return o;
}

// where the emit pass writes one record's attributes (NULL: measure pass)
struct GffAttrOut {
    int32_t *key_off, *list_off, *str_off;  // entries of the record's first key / first string
    uint8_t *key_val, *str_val;             // the batch's byte bases
    int32_t key0, str0;                     // byte offsets (batch-relative) of the record's first key / string
    int32_t strn0;                          // strings of the batch before this record
};
// Walks the attributes field [f, f + n): c[] = entries, key bytes, strings, string bytes.  false: a field without '='.
__device__ bool gff_attr_walk(const uint8_t *f, int32_t n, int32_t c[4], const GffAttrOut *o) {
    int32_t ne = 0, kb = 0, ns = 0, sb = 0;
    if (!(n == 1 && __ldg(f) == '.')) {
        int32_t i = 0;
        while (i < n) {
            int32_t e = i, eq = -1;
            while (e < n && __ldg(f + e) != ';') {
                if (eq < 0 && __ldg(f + e) == '=') eq = e;
                ++e;
            }
            if (eq < 0) return false;
            const int32_t klen = gff_pct(f + i, eq - i, o ? o->key_val + o->key0 + kb : nullptr);
            if (o) o->key_off[ne] = o->key0 + kb;
            kb += klen;
            bool is_array = false;
            for (int32_t q = eq + 1; q < e && !is_array; ++q) is_array = __ldg(f + q) == ',';
            // the builder closes a STRING value's list before it appends the string (see the file header)
            if (o && !is_array) o->list_off[ne + 1] = o->strn0 + ns;
            int32_t v = eq + 1;
            while (true) {
                int32_t ve = v;
                while (ve < e && __ldg(f + ve) != ',') ++ve;
                const int32_t m = gff_pct(f + v, ve - v, o ? o->str_val + o->str0 + sb : nullptr);
                if (o) o->str_off[ns] = o->str0 + sb;
                sb += m;
                ++ns;
                if (ve >= e) break;
                v = ve + 1;
            }
            if (o && is_array) o->list_off[ne + 1] = o->strn0 + ns;
            ++ne;
            i = e + 1;
        }
    }
    c[0] = ne, c[1] = kb, c[2] = ns, c[3] = sb;
    return true;
}

// offsets (from ls) of the first N tabs of [ls, le); false when there are fewer
template <int N>
__device__ __forceinline__ bool find_tabs_n(const uint8_t *ls, const uint8_t *le, int32_t t[N]) {
    const uint8_t *wp = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(ls) & ~(uintptr_t)7);
    int32_t off = (int32_t)(wp - ls);
    int nt = 0;
    while (wp < le) {
        const unsigned long long w = __ldg(reinterpret_cast<const unsigned long long *>(wp));
        const unsigned long long x = w ^ 0x0909090909090909ull;
        const unsigned long long y = (x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full;
        unsigned long long m = ~(y | x | 0x7F7F7F7F7F7F7F7Full);
        if (off < 0) m &= ~0ull << (8 * -off);
        const long long rem = le - wp;
        if (rem < 8) m &= (1ull << (8 * rem)) - 1ull;
        while (m) {
            const int32_t p = off + ((__ffsll((long long)m) - 1) >> 3);
#pragma unroll
            for (int k = 0; k < N; ++k)
                if (nt == k) t[k] = p;
            if (++nt == N) return true;
            m &= m - 1ull;
        }
        wp += 8;
        off += 8;
    }
    return false;
}

// decimal usize > 0 (noodles Position), at most 18 digits
__device__ __forceinline__ bool parse_pos(const uint8_t *f, int n, long long *out) {
    if (n < 1 || n > 18) return false;
    unsigned long long v = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t d = (uint32_t)__ldg(f + i) - '0';
        if (d > 9u) return false;
        v = v * 10ull + d;
    }
    *out = (long long)v;
    return v != 0ull;
}

__global__ void __launch_bounds__(256) gff_measure_kernel(const __grid_constant__ GffColArgs a) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_lines) return;
    const uint8_t *ls = a.line_start[i], *le = a.line_end[i];
    int32_t c[kGN] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t f = 0;
    uint32_t err = 0;
    long long start = 0, end = 0;
    float score = 0.0f;
    if (le == ls) {
        err = kGErrEmptyLine;
    } else if (__ldg(ls) != '#') {
        int32_t t[8];
        if (le - ls > 0x7FFFFFF0ll || !find_tabs_n<8>(ls, le, t)) {
            err = kGErrFields;
        } else {
            f = 1;
            c[kGRow] = 1;
            c[kGSeq] = t[0];
            c[kGSrc] = t[1] - t[0] - 1;
            c[kGType] = t[2] - t[1] - 1;
            if (t[0] == 0) err |= kGErrSeqname;
            if ((a.want[3] && !parse_pos(ls + t[2] + 1, t[3] - t[2] - 1, &start)) || (a.want[4] && !parse_pos(ls + t[3] + 1, t[4] - t[3] - 1, &end))) err |= kGErrPos;
            if (a.want[5]) {
                const uint8_t *s = ls + t[4] + 1;
                const int n = t[5] - t[4] - 1;
                if (!(n == 1 && __ldg(s) == '.')) {
                    uint32_t v = 0;
                    bool plain = n >= 1 && n <= 7;
                    for (int k = 0; plain && k < n; ++k) {
                        const uint32_t d = (uint32_t)__ldg(s + k) - '0';
                        plain = d <= 9u;
                        v = v * 10u + d;
                    }
                    if (plain) {
                        score = (float)v;
                        f |= 2;
                    } else {
                        f |= 8;
                    }
                }
            }
            if (a.want[6]) {
                const int n = t[6] - t[5] - 1;
                const uint8_t ch = n == 1 ? __ldg(ls + t[5] + 1) : 0;
                if (ch == '+' || ch == '-') c[kGStrand] = 1;
                else if (ch == '.' || ch == '?') err |= kGErrStrandNull;
                else err |= kGErrStrand;
            }
            if (a.want[7]) {
                const int n = t[7] - t[6] - 1;
                const uint8_t ch = n == 1 ? __ldg(ls + t[6] + 1) : 0;
                if (ch == '0' || ch == '1' || ch == '2') {
                    c[kGPhase] = 1;
                    f |= 4;
                } else if (ch != '.') {
                    err |= kGErrPhase;
                }
            }
            if (a.want[8] && !gff_attr_walk(ls + t[7] + 1, (int32_t)(le - ls) - t[7] - 1, c + kGAttrN, nullptr)) err |= kGErrAttr;
        }
    }
#pragma unroll
    for (int k = 0; k < kGN; ++k)
        if (a.cnt[k]) a.cnt[k][i] = c[k];
    a.lflags[i] = f;
    if (a.start) a.start[i] = start;  // per LINE here; compacted to rows by the emit pass
    if (a.end) a.end[i] = end;
    if (a.score) a.score[i] = score;
    const uint32_t pend = __ballot_sync(__activemask(), (f & 8u) != 0u);
    if (pend && (threadIdx.x & 31) == __ffs((int)pend) - 1) atomicAdd(a.misc + 2, (unsigned long long)__popc(pend));
    if (err) {
        atomicOr(a.flags, err);
        atomicMin(a.misc + 1, (unsigned long long)i);
    }
}

__global__ void __launch_bounds__(256) gff_score_kernel(const __grid_constant__ GffColArgs a) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_lines) return;
    const uint8_t f = a.lflags[i];
    if (!(f & 8u)) return;
    const uint8_t *ls = a.line_start[i], *le = a.line_end[i];
    int32_t t[8];
    if (!find_tabs_n<8>(ls, le, t)) return;
    float v = 0.0f;
    const int n = t[5] - t[4] - 1;
    const int rc = n > 4096 ? kF32Malformed : parse_f32_rust(ls + t[4] + 1, n, &v);
    if (rc == kF32Ok) {
        a.score[i] = v;
        a.lflags[i] = (uint8_t)((f & ~8u) | 2u);
    } else {
        atomicOr(a.flags, kGErrScore);
        atomicMin(a.misc + 1, (unsigned long long)i);
    }
}

__global__ void __launch_bounds__(256) gff_emit_kernel(const __grid_constant__ GffColArgs a, long long *out_start, long long *out_end, float *out_score) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= a.n_lines) return;
    const uint8_t f = a.lflags[i];
    if (!(f & 1u)) return;
    int64_t lo = 0, hi = a.n_batches;  // batch (= file) of the line
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (a.bline[mid] <= i) lo = mid;
        else hi = mid;
    }
    const int64_t b = lo, r0 = a.brow[b], l0 = a.bline[b];
    const long long row = a.pre[kGRow][i];
    const int64_t in_batch = row - r0;
    const bool last = row + 1 == a.brow[b + 1];
    if (out_start) out_start[row] = a.start[i];
    if (out_end) out_end[row] = a.end[i];
    if (out_score) out_score[row] = a.score[i];
    const long long word = a.bit0[b] + (in_batch >> 5);
    const uint32_t bit = 1u << (in_batch & 31);
    if (a.score_valid && (f & 2u)) atomicOr(a.score_valid + word, bit);
    if (a.phase_valid && (f & 4u)) atomicOr(a.phase_valid + word, bit);
    const uint8_t *ls = a.line_start[i], *le = a.line_end[i];
    int32_t t[8];
    if (!find_tabs_n<8>(ls, le, t)) return;
    const long long lend = a.bline[b + 1];  // first line of the next file: prefix there = end of this batch
    const int fs[kGStrCols] = {0, 0, t[0] + 1, t[1] + 1, t[5] + 1, t[6] + 1};
    if (a.map_off) {
        const long long *PE = a.pre[kGAttrN], *PK = a.pre[kGKeyB], *PS = a.pre[kGStrN], *PB = a.pre[kGStrB];
        int32_t *mo = a.map_off + r0 + b;
        mo[in_batch] = (int32_t)(PE[i] - PE[l0]);
        GffAttrOut o;
        o.key_off = a.key_off + PE[i] + b;
        o.list_off = a.list_off + PE[i] + b;
        o.str_off = a.str_off + PS[i] + b;
        o.key_val = a.key_val + PK[l0];
        o.str_val = a.str_val + PB[l0];
        o.key0 = (int32_t)(PK[i] - PK[l0]);
        o.str0 = (int32_t)(PB[i] - PB[l0]);
        o.strn0 = (int32_t)(PS[i] - PS[l0]);
        if (in_batch == 0) a.list_off[PE[l0] + b] = 0;  // offsets[0] of the batch's value lists
        if (last) {
            mo[in_batch + 1] = (int32_t)(PE[lend] - PE[l0]);
            a.key_off[PE[lend] + b] = (int32_t)(PK[lend] - PK[l0]);
            a.str_off[PS[lend] + b] = (int32_t)(PB[lend] - PB[l0]);
        }
        int32_t cc[4];
        gff_attr_walk(ls + t[7] + 1, (int32_t)(le - ls) - t[7] - 1, cc, &o);
    }
#pragma unroll
    for (int k = kGSeq; k < kGStrCols; ++k) {
        if (!a.off[k]) continue;
        const long long *P = a.pre[k];
        int32_t *o = a.off[k] + r0 + b;
        o[in_batch] = (int32_t)(P[i] - P[l0]);
        if (last) o[in_batch + 1] = (int32_t)(P[lend] - P[l0]);
        const int n = (int)(P[i + 1] - P[i]);
        uint8_t *dst = a.val[k] + P[i];
        for (int j = 0; j < n; ++j) dst[j] = __ldg(ls + fs[k] + j);
    }
}

__global__ void gff_gather(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

size_t gal256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

struct GffBuf {
    void *d = nullptr, *h = nullptr;
    size_t bytes = 0;
};

struct GffColumns {
    std::atomic<int> refs{1};
    int device = 0;
    bool on_device = false;
    int64_t n_rows = 0, n_batches = 0, next = 0;
    std::vector<int> projection;
    GffBuf off[kGN], val[kGN], start, end, score, score_valid, phase_valid, map_off, key_off, list_off, str_off, key_val, str_val;
    std::vector<long long> batch_row0, bit0, base[kGN];
    template <class T>
    const T *p(const GffBuf &b) const { return static_cast<const T *>(on_device ? b.d : b.h); }
    template <class F>
    void each(F fn) {
        for (int k = 0; k < kGN; ++k) fn(off[k]), fn(val[k]);
        fn(start), fn(end), fn(score), fn(score_valid), fn(phase_valid);
        fn(map_off), fn(key_off), fn(list_off), fn(str_off), fn(key_val), fn(str_val);
    }
    void unref() {
        if (refs.fetch_sub(1) == 1) {
            cudaSetDevice(device);
            each([](GffBuf &b) {
                cudaFree(b.d);
                cudaFreeHost(b.h);
            });
            delete this;
        }
    }
};

void gff_columns_free(VcfStream *s) {
    if (s->gff_cols) {
        s->gff_cols->unref();
        s->gff_cols = nullptr;
    }
}

namespace {

constexpr int kColSlot[8] = {kGSeq, kGSrc, kGType, -1, -1, -1, kGStrand, kGPhase};

int gff_build_columns(VcfStream *s) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto *c = new (std::nothrow) GffColumns();
    if (!c) return fail(EXON_GPU_ERR_OOM, "gff_next_batch: out of host memory");
    s->gff_cols = c;
    c->device = ctx->device;
    c->on_device = s->columns_on_device;
    c->projection = s->projection;
    bool want[9] = {false, false, false, false, false, false, false, false, false};
    for (int p : s->projection) want[p] = true;
    c->batch_row0.assign(1, 0);

    const size_t per_line = kGN * 4 + kGN * 8 + 1 + 8 + 8 + 4;  // counts | prefixes | flags | start | end | score (per line)
    LineIndex li;
    if (int rc = build_line_index(s, per_line, (2 * kGN + 8) * 256 + (1 << 20), &li)) return rc;
    const int64_t n_lines = li.n_lines;
    if (n_lines == 0) return EXON_GPU_OK;
    const size_t nl1 = (size_t)n_lines + 1;
    const int n_files = (int)li.file_line0.size() - 1;
    size_t cub_bytes = 0;
    CUDA_TRY(exclusive_sum_i32_i64(nullptr, cub_bytes, (const int32_t *)nullptr, (long long *)nullptr, (int)nl1, st));
    uint8_t *x = li.extra;
    auto take = [&](size_t bytes) {
        uint8_t *p = x;
        x += gal256(bytes);
        return p;
    };
    GffColArgs a;
    memset(&a, 0, sizeof(a));
    a.n_lines = n_lines;
    a.line_start = li.line_start;
    a.line_end = li.line_end;
    for (int k = 0; k < 9; ++k) a.want[k] = want[k];
    bool need[kGN] = {true, want[0], want[1], want[2], want[6], want[7], want[8], want[8], want[8], want[8]};
    long long *pre[kGN];
    for (int k = 0; k < kGN; ++k) {
        pre[k] = nullptr;
        if (!need[k]) continue;
        a.cnt[k] = (int32_t *)take(nl1 * 4);
        pre[k] = (long long *)take(nl1 * 8);
        a.pre[k] = pre[k];
        CUDA_TRY(cudaMemsetAsync(a.cnt[k] + n_lines, 0, 4, st));
    }
    a.lflags = take(nl1);
    if (want[3]) a.start = (long long *)take(nl1 * 8);
    if (want[4]) a.end = (long long *)take(nl1 * 8);
    if (want[5]) a.score = (float *)take(nl1 * 4);
    // scan scratch grows with the row count (64-bit tile states): from the pool, not from the fixed part of scratch_b
    uint8_t *cub_tmp = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&cub_tmp, cub_bytes + 256, st));
    struct CubFree {
        void *p;
        cudaStream_t st;
        ~CubFree() { cudaFreeAsync(p, st); }
    } cub_guard{cub_tmp, st};
    unsigned long long *d_misc = (unsigned long long *)take(64);
    const unsigned long long init_misc[3] = {0ull, ~0ull, 0ull};
    CUDA_TRY(cudaMemcpyAsync(d_misc, init_misc, sizeof(init_misc), cudaMemcpyHostToDevice, st));
    a.flags = reinterpret_cast<uint32_t *>(d_misc);
    a.misc = d_misc;
    const size_t nf1 = (size_t)n_files + 1;
    long long *d_small = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&d_small, (4 + kGN) * gal256(nf1 * 8), st));
    struct PoolFree {
        void *p;
        cudaStream_t st;
        ~PoolFree() { cudaFreeAsync(p, st); }
    } g0{d_small, st};
    auto d_tab = [&](int k) { return reinterpret_cast<long long *>(reinterpret_cast<uint8_t *>(d_small) + (size_t)k * gal256(nf1 * 8)); };
    CUDA_TRY(cudaMemcpyAsync(d_tab(0), li.file_line0.data(), nf1 * 8, cudaMemcpyHostToDevice, st));
    a.file_line0 = d_tab(0);
    a.n_files = n_files;

    const unsigned grid = (unsigned)((n_lines + 255) / 256);
    gff_measure_kernel<<<grid, 256, 0, st>>>(a);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    std::vector<long long> h_base[kGN];
    for (int k = 0; k < kGN; ++k) {
        if (!need[k]) continue;
        size_t tb = cub_bytes;
        CUDA_TRY(exclusive_sum_i32_i64(cub_tmp, tb, (const int32_t *)a.cnt[k], pre[k], (int)nl1, st));
        gff_gather<<<(unsigned)((nf1 + 255) / 256), 256, 0, st>>>(pre[k], d_tab(0), (int64_t)nf1, d_tab(4 + k));
        ctx->launches.fetch_add(2);
        h_base[k].resize(nf1);
        CUDA_TRY(cudaMemcpyAsync(h_base[k].data(), d_tab(4 + k), nf1 * 8, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_misc[3];
    CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (want[5] && h_misc[2] && !h_misc[0]) {
        gff_score_kernel<<<grid, 256, 0, st>>>(a);
        ctx->launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (const uint32_t e = (uint32_t)h_misc[0])
        return fail(EXON_GPU_ERR_PARSE, "malformed GFF at line %llu:%s%s%s%s%s%s%s%s%s", h_misc[1], (e & kGErrFields) ? " fewer than 9 tab-separated fields;" : "",
                    (e & kGErrEmptyLine) ? " empty line;" : "", (e & kGErrPos) ? " start / end is not a positive decimal integer;" : "",
                    (e & kGErrScore) ? " score is not a float literal;" : "", (e & kGErrStrand) ? " invalid strand;" : "", (e & kGErrPhase) ? " invalid phase;" : "",
                    (e & kGErrStrandNull) ? " strand '.' / '?' is NULL in the reference's non-nullable strand column;" : "", (e & kGErrSeqname) ? " empty seqname;" : "",
                    (e & kGErrAttr) ? " an attribute without '=';" : "");
    // one batch per non-empty file (read_batch has no row limit)
    std::vector<long long> brow, bline;
    for (int f = 0; f < n_files; ++f)
        if (h_base[kGRow][(size_t)f + 1] > h_base[kGRow][(size_t)f]) {
            brow.push_back(h_base[kGRow][(size_t)f]);
            bline.push_back(li.file_line0[(size_t)f]);
        }
    const int64_t n_rows = h_base[kGRow][(size_t)n_files];
    c->n_rows = n_rows;
    c->n_batches = (int64_t)brow.size();
    brow.push_back(n_rows);
    bline.push_back(n_lines);
    c->batch_row0 = brow;
    if (n_rows == 0) return EXON_GPU_OK;
    const size_t nb1 = (size_t)c->n_batches + 1;
    c->bit0.resize(nb1);
    long long words = 0;
    for (size_t b = 0; b < nb1; ++b) {
        c->bit0[b] = words;
        if (b + 1 < nb1) words += ((brow[b + 1] - brow[b] + 63) / 64) * 2;
    }
    // per-batch bases of the string columns = prefix at the batch's first line (a file boundary: already on the host)
    for (int k = kGSeq; k < kGN; ++k) {
        if (!need[k]) continue;
        c->base[k].resize(nb1);
        size_t bi = 0;
        for (int f = 0; f < n_files; ++f)
            if (h_base[kGRow][(size_t)f + 1] > h_base[kGRow][(size_t)f]) c->base[k][bi++] = h_base[k][(size_t)f];
        c->base[k][bi] = h_base[k][(size_t)n_files];
        for (size_t b = 0; b + 1 < nb1; ++b)
            if (c->base[k][b + 1] - c->base[k][b] > 0x7FFFFFFFll) return fail(EXON_GPU_ERR_UNSUPPORTED, "gff_next_batch: file %zu overflows int32 offsets", b);
    }
    CUDA_TRY(cudaMemcpyAsync(d_tab(1), brow.data(), nb1 * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_tab(2), bline.data(), nb1 * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_tab(3), c->bit0.data(), nb1 * 8, cudaMemcpyHostToDevice, st));
    a.brow = d_tab(1);
    a.bline = d_tab(2);
    a.bit0 = d_tab(3);
    a.n_batches = c->n_batches;
    auto dev_alloc = [&](GffBuf &b, size_t bytes, bool zero) -> int {
        b.bytes = std::max<size_t>(bytes, 8);
        CUDA_TRY(cudaMallocAsync(&b.d, b.bytes, st));
        if (zero) CUDA_TRY(cudaMemsetAsync(b.d, 0, b.bytes, st));
        return EXON_GPU_OK;
    };
    for (int k = kGSeq; k < kGStrCols; ++k) {
        if (!need[k]) continue;
        if (int rc = dev_alloc(c->off[k], ((size_t)n_rows + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->val[k], (size_t)c->base[k][nb1 - 1], false)) return rc;
        a.off[k] = (int32_t *)c->off[k].d;
        a.val[k] = (uint8_t *)c->val[k].d;
    }
    if (want[8]) {
        const size_t n_ent = (size_t)c->base[kGAttrN][nb1 - 1], n_str = (size_t)c->base[kGStrN][nb1 - 1];
        if (int rc = dev_alloc(c->map_off, ((size_t)n_rows + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->key_off, (n_ent + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->list_off, (n_ent + nb1) * 4, true)) return rc;
        if (int rc = dev_alloc(c->str_off, (n_str + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->key_val, (size_t)c->base[kGKeyB][nb1 - 1], false)) return rc;
        if (int rc = dev_alloc(c->str_val, (size_t)c->base[kGStrB][nb1 - 1], false)) return rc;
        a.map_off = (int32_t *)c->map_off.d, a.key_off = (int32_t *)c->key_off.d, a.list_off = (int32_t *)c->list_off.d, a.str_off = (int32_t *)c->str_off.d;
        a.key_val = (uint8_t *)c->key_val.d, a.str_val = (uint8_t *)c->str_val.d;
    }
    long long *o_start = nullptr, *o_end = nullptr;
    float *o_score = nullptr;
    if (want[3]) {
        if (int rc = dev_alloc(c->start, (size_t)n_rows * 8, false)) return rc;
        o_start = (long long *)c->start.d;
    }
    if (want[4]) {
        if (int rc = dev_alloc(c->end, (size_t)n_rows * 8, false)) return rc;
        o_end = (long long *)c->end.d;
    }
    if (want[5]) {
        if (int rc = dev_alloc(c->score, (size_t)n_rows * 4, false)) return rc;
        if (int rc = dev_alloc(c->score_valid, (size_t)words * 4, true)) return rc;
        o_score = (float *)c->score.d;
        a.score_valid = (uint32_t *)c->score_valid.d;
    }
    if (want[7]) {
        if (int rc = dev_alloc(c->phase_valid, (size_t)words * 4, true)) return rc;
        a.phase_valid = (uint32_t *)c->phase_valid.d;
    }
    gff_emit_kernel<<<grid, 256, 0, st>>>(a, o_start, o_end, o_score);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    if (!c->on_device) {
        int rc = EXON_GPU_OK;
        c->each([&](GffBuf &b) {
            if (!b.d || rc) return;
            if (cudaHostAlloc(&b.h, b.bytes, cudaHostAllocDefault) != cudaSuccess || cudaMemcpyAsync(b.h, b.d, b.bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                rc = fail(EXON_GPU_ERR_OOM, "gff_next_batch: host copy of the columns failed");
        });
        if (rc) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return EXON_GPU_OK;
}

struct GffBatchPriv {
    GffColumns *cols;
    int n_children;
    ArrowArray children[9];
    ArrowArray *child_ptrs[9];
    const void *bufs[9][3];
    const void *struct_buffers[1];
    // attributes: map -> entries struct -> {keys utf8, values list -> item utf8}
    ArrowArray entries, keys, values, items;
    ArrowArray *entries_ptr, *kv_ptrs[2], *items_ptr;
    const void *entries_bufs[1], *keys_bufs[3], *values_bufs[2], *items_bufs[3];
};
void gff_release_child(ArrowArray *a) { a->release = nullptr; }
void gff_release_batch(ArrowArray *a) {
    auto *p = static_cast<GffBatchPriv *>(a->private_data);
    p->cols->unref();
    delete p;
    a->release = nullptr;
}
struct GffSchemaPriv {
    int n_children;
    ArrowSchema children[9];
    ArrowSchema *child_ptrs[9];
    ArrowSchema entries, keys, values, items;
    ArrowSchema *entries_ptr, *kv_ptrs[2], *items_ptr;
};
void gff_release_schema_child(ArrowSchema *s) { s->release = nullptr; }
void gff_release_schema(ArrowSchema *s) {
    delete static_cast<GffSchemaPriv *>(s->private_data);
    s->release = nullptr;
}
// new_gff_schema_builder, exon/exon-gff/src/config.rs:81-108
void gff_fill_schema(const std::vector<int> &projection, ArrowSchema *out) {
    static const char *names[9] = {"seqname", "source", "type", "start", "end", "score", "strand", "phase", "attributes"};
    static const char *formats[9] = {"u", "u", "u", "l", "l", "f", "u", "u", "+m"};
    static const bool nullable[9] = {false, true, false, false, false, true, false, true, true};
    auto *p = new GffSchemaPriv();
    p->n_children = (int)projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        const int col = projection[(size_t)i];
        ArrowSchema &c = p->children[i];
        memset(&c, 0, sizeof(c));
        c.format = formats[col];
        c.name = names[col];
        c.flags = nullable[col] ? ARROW_FLAG_NULLABLE : 0;
        c.release = gff_release_schema_child;
        if (col == 8) {
            // Field::new_map("attributes", "entries", keys: Utf8 !null, values: List<item: Utf8>, sorted = false, nullable = true)
            auto init = [](ArrowSchema &x, const char *fmt, const char *name, bool nullable_) {
                memset(&x, 0, sizeof(x));
                x.format = fmt;
                x.name = name;
                x.flags = nullable_ ? ARROW_FLAG_NULLABLE : 0;
                x.release = gff_release_schema_child;
            };
            init(p->entries, "+s", "entries", false);
            init(p->keys, "u", "keys", false);
            init(p->values, "+l", "values", true);
            init(p->items, "u", "item", true);
            p->items_ptr = &p->items;
            p->values.n_children = 1;
            p->values.children = &p->items_ptr;
            p->kv_ptrs[0] = &p->keys;
            p->kv_ptrs[1] = &p->values;
            p->entries.n_children = 2;
            p->entries.children = p->kv_ptrs;
            p->entries_ptr = &p->entries;
            c.n_children = 1;
            c.children = &p->entries_ptr;
        }
        p->child_ptrs[i] = &c;
    }
    memset(out, 0, sizeof(*out));
    out->format = "+s";
    out->name = "";
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = gff_release_schema;
    out->private_data = p;
}

}  // namespace

void gff_stream_schema(VcfStream *s, ArrowSchema *out) { gff_fill_schema(s->projection, out); }

int gff_next_batch(VcfStream *s, ArrowArray *out, ArrowSchema *out_schema) {
    if (!s->gff_cols) {
        if (int rc = s->flush_gz()) return rc;
        std::lock_guard<std::recursive_mutex> work(s->ctx->work_mu);
        if (int rc = gff_build_columns(s)) {
            gff_columns_free(s);
            return rc;
        }
        s->drained = true;
    }
    GffColumns *c = s->gff_cols;
    if (out_schema) gff_fill_schema(s->projection, out_schema);
    memset(out, 0, sizeof(*out));
    if (c->next >= c->n_batches) return EXON_GPU_OK;
    const int64_t b = c->next++;
    const int64_t row0 = c->batch_row0[(size_t)b], rows = c->batch_row0[(size_t)b + 1] - row0;
    auto *p = new GffBatchPriv();
    memset(static_cast<void *>(p), 0, sizeof(*p));
    p->cols = c;
    c->refs.fetch_add(1);
    p->n_children = (int)s->projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        const int col = s->projection[(size_t)i];
        ArrowArray &a = p->children[i];
        a.length = rows;
        a.buffers = p->bufs[i];
        a.release = gff_release_child;
        p->bufs[i][0] = nullptr;
        if (col == 8) {
            const long long e0 = c->base[kGAttrN][(size_t)b], n_ent = c->base[kGAttrN][(size_t)b + 1] - e0;
            const long long s0 = c->base[kGStrN][(size_t)b], n_str = c->base[kGStrN][(size_t)b + 1] - s0;
            a.n_buffers = 2;
            p->bufs[i][1] = c->p<int32_t>(c->map_off) + row0 + b;
            auto init = [](ArrowArray &x, int64_t len, int nb, const void **bufs) {
                memset(&x, 0, sizeof(x));
                x.length = len;
                x.n_buffers = nb;
                x.buffers = bufs;
                x.release = gff_release_child;
            };
            init(p->entries, n_ent, 1, p->entries_bufs);
            init(p->keys, n_ent, 3, p->keys_bufs);
            init(p->values, n_ent, 2, p->values_bufs);
            init(p->items, n_str, 3, p->items_bufs);  // every string of the batch, the dangling last one included
            p->entries_bufs[0] = nullptr;
            p->keys_bufs[0] = nullptr;
            p->keys_bufs[1] = c->p<int32_t>(c->key_off) + e0 + b;
            p->keys_bufs[2] = c->p<uint8_t>(c->key_val) + c->base[kGKeyB][(size_t)b];
            p->values_bufs[0] = nullptr;
            p->values_bufs[1] = c->p<int32_t>(c->list_off) + e0 + b;
            p->items_bufs[0] = nullptr;
            p->items_bufs[1] = c->p<int32_t>(c->str_off) + s0 + b;
            p->items_bufs[2] = c->p<uint8_t>(c->str_val) + c->base[kGStrB][(size_t)b];
            p->items_ptr = &p->items;
            p->values.n_children = 1;
            p->values.children = &p->items_ptr;
            p->kv_ptrs[0] = &p->keys;
            p->kv_ptrs[1] = &p->values;
            p->entries.n_children = 2;
            p->entries.children = p->kv_ptrs;
            p->entries_ptr = &p->entries;
            a.n_children = 1;
            a.children = &p->entries_ptr;
        } else if (col == 3 || col == 4) {
            a.n_buffers = 2;
            p->bufs[i][1] = c->p<long long>(col == 3 ? c->start : c->end) + row0;
        } else if (col == 5) {
            a.n_buffers = 2;
            a.null_count = -1;
            p->bufs[i][0] = c->p<uint32_t>(c->score_valid) + c->bit0[(size_t)b];
            p->bufs[i][1] = c->p<float>(c->score) + row0;
        } else {
            const int k = kColSlot[col];
            a.n_buffers = 3;
            if (col == 7) {
                a.null_count = -1;
                p->bufs[i][0] = c->p<uint32_t>(c->phase_valid) + c->bit0[(size_t)b];
            }
            p->bufs[i][1] = c->p<int32_t>(c->off[k]) + row0 + b;
            p->bufs[i][2] = c->p<uint8_t>(c->val[k]) + c->base[k][(size_t)b];
        }
        p->child_ptrs[i] = &a;
    }
    p->struct_buffers[0] = nullptr;
    out->length = rows;
    out->n_buffers = 1;
    out->buffers = p->struct_buffers;
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = gff_release_batch;
    out->private_data = p;
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_gff_open_columns(exon_gpu_ctx *c, const exon_gpu_fastq_opts *o, exon_gpu_stream **out) {
    if (!c || !o || !out) return fail(EXON_GPU_ERR_ARG, "gff_open_columns: NULL argument");
    if (o->n_projection < 0 || o->n_projection > 9 || (o->n_projection > 0 && !o->projection)) return fail(EXON_GPU_ERR_ARG, "gff_open_columns: bad projection");
    for (int i = 0; i < o->n_projection; ++i) {
        if (o->projection[i] < 0 || o->projection[i] > 8) return fail(EXON_GPU_ERR_ARG, "gff_open_columns: projection index %d is not a GFF file-schema column", o->projection[i]);
        for (int j = 0; j < i; ++j)
            if (o->projection[j] == o->projection[i]) return fail(EXON_GPU_ERR_ARG, "gff_open_columns: column %d is projected twice", o->projection[i]);
    }
    if (int rc = exon_gpu_gff_open(c, out)) return rc;
    (*out)->projection.assign(o->projection, o->projection + o->n_projection);
    (*out)->columns_on_device = o->columns_on_device != 0;
    return EXON_GPU_OK;
}

int exon_gpu_gff_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema) {
    if (!s || !out || s->fmt != kFmtGff) return fail(EXON_GPU_ERR_ARG, "gff_next_batch: not a GFF stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return gff_next_batch(s, out, out_schema);
}

}  // extern "C"
