// filter_agg.cu -- K3: columnar filter + partial aggregate over Arrow record batches, and the evaluated region
// UDFs (region_match / chrom_match / interval_match).
//
// Replaces DataFusion's FilterExec (arrow-ord `eq` Utf8-vs-scalar, `gt_eq` / `lt_eq` Int64, arrow-arith
// `and_kleene`, arrow-select `filter_record_batch`) followed by AggregateExec(Partial) `count(*)`, `count(x)`,
// `sum(x)`, `avg(x)` (datafusion-physical-plan / datafusion-functions-aggregate 44.0.0, third party; predicate
// shape as in exon/exon-core/src/physical_plan/pos_interval_physical_expr.rs:79-98 and
// region_physical_expr.rs:220-240), and the row-at-a-time UDFs of exon/exon-core/src/udfs/vcf/mod.rs:65-274.
// Nothing is compacted: the selection mask lives in registers and feeds the accumulators directly, so each
// column byte is read once and 24 bytes come back.
//
// Null semantics of the filter: a NULL operand makes a comparison NULL, `and_kleene` keeps NULL unless the other
// side is false, and FilterExec drops rows whose predicate is NULL or false -- i.e. a row is selected iff every
// operand is valid and every comparison true.  count(x) / sum(x) / avg(x) skip NULL x.
#include <cstring>

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

struct FilterAggArgs {
    int64_t n_rows;
    // chrom: utf8
    const uint8_t *chrom_valid;  // may be NULL
    const int32_t *chrom_offsets;
    const uint8_t *chrom_values;
    int64_t chrom_off;  // logical offset of row 0
    int32_t has_chrom;
    int32_t lit_len;
    uint8_t lit[kMaxChrom + 1];
    // pos: int64
    const uint8_t *pos_valid;
    const int64_t *pos;
    int64_t pos_off;
    int32_t has_pos;
    int64_t lo, hi;
    // aggregated value
    const uint8_t *val_valid;
    const void *val;
    int64_t val_off;
    int32_t val_type;
    int32_t agg_kind;
    // out: [0] count (u64) [1] sum_i64 (as u64, two's complement) [2] sum_f64
    unsigned long long *out;
    // region UDFs
    int32_t udf_kind;
    uint8_t *mask_values, *mask_valid;
    uint32_t *mask_err;
};

__device__ __forceinline__ bool bit_set(const uint8_t *bits, int64_t i) {
    return bits == nullptr || ((bits[i >> 3] >> (i & 7)) & 1);
}

__device__ __forceinline__ bool chrom_equals(const FilterAggArgs &a, int64_t r) {
    const int32_t s = a.chrom_offsets[r], e = a.chrom_offsets[r + 1];
    bool eq = (e - s) == a.lit_len;
    for (int32_t j = 0; eq && j < a.lit_len; ++j) eq = a.chrom_values[s + j] == a.lit[j];
    return eq;
}

constexpr int kFaThreads = 256;

__global__ void __launch_bounds__(kFaThreads) filter_agg_kernel(const __grid_constant__ FilterAggArgs a) {
    unsigned long long cnt = 0;
    long long si = 0;
    double sf = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kFaThreads;
    for (int64_t i = (int64_t)blockIdx.x * kFaThreads + threadIdx.x; i < a.n_rows; i += stride) {
        bool sel = true;
        if (a.has_pos) {  // the cheap coalesced test first; AND is commutative under Kleene logic for selection
            const int64_t r = i + a.pos_off;
            sel = bit_set(a.pos_valid, r);
            if (sel) {
                const int64_t v = a.pos[r];
                sel = (v >= a.lo) & (v <= a.hi);
            }
        }
        if (sel && a.has_chrom) {
            const int64_t r = i + a.chrom_off;
            sel = bit_set(a.chrom_valid, r) && chrom_equals(a, r);
        }
        if (!sel) continue;
        if (a.agg_kind == EXON_GPU_AGG_COUNT_STAR) {
            ++cnt;
        } else {
            const int64_t r = i + a.val_off;
            if (!bit_set(a.val_valid, r)) continue;
            ++cnt;
            if (a.agg_kind != EXON_GPU_AGG_COUNT) {
                if (a.val_type == kValI64) si += static_cast<const int64_t *>(a.val)[r];
                else if (a.val_type == kValI32) si += static_cast<const int32_t *>(a.val)[r];
                else if (a.val_type == kValF64) sf += static_cast<const double *>(a.val)[r];
                else if (a.val_type == kValF32) sf += (double)static_cast<const float *>(a.val)[r];
            }
        }
    }
    // warp shuffle reduction, then one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        si += __shfl_xor_sync(0xFFFFFFFFu, si, d);
        sf += __shfl_xor_sync(0xFFFFFFFFu, sf, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(a.out, cnt);
        if (si) atomicAdd(a.out + 1, (unsigned long long)si);
        if (sf != 0.0) atomicAdd(reinterpret_cast<double *>(a.out + 2), sf);
    }
}

// The same operators over MANY device-resident batches in one launch (a partition's whole column store): work is
// cut into units of kUnitRows rows of one batch, units are dealt round-robin to a persistent grid, and every
// thread evaluates kRowsPerThread rows whose loads are issued together (coalesced 8-byte / 4-byte accesses, four
// independent rows in flight per thread).  The CHROM bytes are fetched only for rows whose POS and CHROM length
// already match.
constexpr int kRowsPerThread = 8;
constexpr int kUnitRows = kFaThreads * kRowsPerThread;

template <bool NULLS>
__global__ void __launch_bounds__(kFaThreads) filter_agg_multi_kernel(const __grid_constant__ FilterAggArgs a, const FaBatchDesc *descs,
                                                                      int n_batches, int units_per_batch) {
    unsigned long long cnt = 0;
    long long si = 0;
    double sf = 0.0;
    uint32_t cnt32 = 0;
    const int lane = threadIdx.x & 31;
    const int64_t n_units = (int64_t)n_batches * units_per_batch;
    // lo <= v <= hi as one unsigned compare: (v - lo) <= (hi - lo); an empty interval never matches
    const unsigned long long span = a.hi >= a.lo ? (unsigned long long)a.hi - (unsigned long long)a.lo : 0ull;
    const bool never = a.has_pos && a.hi < a.lo;
    for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int b = (int)(unit / units_per_batch);
        const int base = (int)(unit - (int64_t)b * units_per_batch) * kUnitRows;
        const FaBatchDesc *d = descs + b;
        const int64_t n64 = __ldg(&d->n_rows);
        if (base >= n64) continue;
        const int n = (int)(n64 - base < kUnitRows ? n64 - base : kUnitRows);  // rows of this unit
        const int64_t pos_off = __ldg(&d->pos_off), chrom_off = __ldg(&d->chrom_off);
        const int64_t *pos = a.has_pos ? reinterpret_cast<const int64_t *>(__ldg(reinterpret_cast<const unsigned long long *>(&d->pos))) + pos_off + base : nullptr;
        const int32_t *off = a.has_chrom ? reinterpret_cast<const int32_t *>(__ldg(reinterpret_cast<const unsigned long long *>(&d->chrom_offsets))) + chrom_off + base : nullptr;
        int64_t pv[kRowsPerThread];
        int32_t o0[kRowsPerThread], o1[kRowsPerThread];
        // all loads of the unit are issued before anything is consumed (8 x 8 B + 8 x 4 B in flight per thread)
#pragma unroll
        for (int k = 0; k < kRowsPerThread; ++k) {
            const int i = k * kFaThreads + (int)threadIdx.x;
            pv[k] = 0;
            o0[k] = 0;
            if (i < n) {
                if (a.has_pos) pv[k] = __ldg(pos + i);
                if (a.has_chrom) o0[k] = __ldg(off + i);
            }
        }
        if (a.has_chrom) {
            // end offset of row i = start offset of row i + 1: taken from the next lane, loaded only by lane 31
#pragma unroll
            for (int k = 0; k < kRowsPerThread; ++k) {
                const int i = k * kFaThreads + (int)threadIdx.x;
                o1[k] = __shfl_down_sync(0xFFFFFFFFu, o0[k], 1);
                if ((lane == 31 || i + 1 >= n) && i < n) o1[k] = __ldg(off + i + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < kRowsPerThread; ++k) {
            const int i = k * kFaThreads + (int)threadIdx.x;
            bool sel = i < n && !never;
            if (a.has_pos) sel &= ((unsigned long long)pv[k] - (unsigned long long)a.lo) <= span;
            if (a.has_chrom) sel &= (o1[k] - o0[k]) == a.lit_len;
            if (!sel) continue;
            // rare from here on (rows whose POS and CHROM length already match)
            const int64_t row = (int64_t)base + i;
            if (NULLS && a.has_pos && !bit_set(d->pos_valid, row + pos_off)) continue;
            if (a.has_chrom) {
                if (NULLS && !bit_set(d->chrom_valid, row + chrom_off)) continue;
                const uint8_t *cv = d->chrom_values + o0[k];
                bool eq = true;
                for (int32_t j = 0; eq && j < a.lit_len; ++j) eq = cv[j] == a.lit[j];
                if (!eq) continue;
            }
            if (a.agg_kind == EXON_GPU_AGG_COUNT_STAR) {
                ++cnt32;
            } else {
                const int64_t r = row + d->val_off;
                if (NULLS && !bit_set(d->val_valid, r)) continue;
                ++cnt32;
                if (a.agg_kind != EXON_GPU_AGG_COUNT) {
                    if (a.val_type == kValI64) si += static_cast<const int64_t *>(d->val)[r];
                    else if (a.val_type == kValI32) si += static_cast<const int32_t *>(d->val)[r];
                    else if (a.val_type == kValF64) sf += static_cast<const double *>(d->val)[r];
                    else if (a.val_type == kValF32) sf += (double)static_cast<const float *>(d->val)[r];
                }
            }
        }
        cnt += cnt32;
        cnt32 = 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        si += __shfl_xor_sync(0xFFFFFFFFu, si, d);
        sf += __shfl_xor_sync(0xFFFFFFFFu, sf, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(a.out, cnt);
        if (si) atomicAdd(a.out + 1, (unsigned long long)si);
        if (sf != 0.0) atomicAdd(reinterpret_cast<double *>(a.out + 2), sf);
    }
}

// region_match (udf_kind 0): chrom and pos must be non-NULL and pos >= 1 (else the query fails), value =
// name == chrom && lo <= pos <= hi, never NULL.  chrom_match (1): NULL in -> NULL out.  interval_match (2):
// NULL pos -> false; pos < 1 fails.  exon/exon-core/src/udfs/vcf/mod.rs:65-131, 167-196, 232-274.
constexpr uint32_t kUdfErrNull = 1u, kUdfErrPos = 2u;
__global__ void __launch_bounds__(kFaThreads) region_udf_kernel(const __grid_constant__ FilterAggArgs a) {
    const int64_t i = (int64_t)blockIdx.x * kFaThreads + threadIdx.x;
    if (i >= a.n_rows) return;
    uint8_t value = 0, valid = 1;
    uint32_t err = 0;
    if (a.udf_kind == 0) {
        const bool cv = bit_set(a.chrom_valid, i + a.chrom_off), pv = bit_set(a.pos_valid, i + a.pos_off);
        if (!cv || !pv) err = kUdfErrNull;
        else {
            const int64_t v = a.pos[i + a.pos_off];
            if (v < 1) err = kUdfErrPos;
            else value = chrom_equals(a, i + a.chrom_off) && v >= a.lo && v <= a.hi;
        }
    } else if (a.udf_kind == 1) {
        if (!bit_set(a.chrom_valid, i + a.chrom_off)) valid = 0;
        else value = chrom_equals(a, i + a.chrom_off);
    } else {
        if (bit_set(a.pos_valid, i + a.pos_off)) {
            const int64_t v = a.pos[i + a.pos_off];
            if (v < 1) err = kUdfErrPos;
            else value = v >= a.lo && v <= a.hi;
        }
    }
    a.mask_values[i] = value;
    a.mask_valid[i] = valid;
    if (err) atomicOr(a.mask_err, err);
}

// Validates the batch against the request, fills `a`, and stages host buffers into the context's scratch area
// (device buffers are used in place).  `front` bytes at the start of the scratch area are left to the caller.
int prepare(Ctx *c, const ArrowArray *batch, const ArrowSchema *schema, int buffers_on_device, const exon_gpu_pred *pred,
            const exon_gpu_agg *agg, size_t front, size_t host_bytes, FilterAggArgs &a) {
    if (!schema->format || strcmp(schema->format, "+s") != 0 || schema->n_children != batch->n_children)
        return fail(EXON_GPU_ERR_ARG, "batch must be a struct array (\"+s\") matching its schema");
    const int64_t n = batch->length;
    const int nc = (int)batch->n_children;
    auto child_ok = [&](int idx) { return idx >= 0 && idx < nc && batch->children[idx] && schema->children[idx]; };
    memset(&a, 0, sizeof(a));
    a.n_rows = n;

    struct Piece { const void *src; size_t bytes; const void **dst; };
    std::vector<Piece> pieces;
    const void *p_chrom_valid = nullptr, *p_chrom_off = nullptr, *p_chrom_val = nullptr, *p_pos_valid = nullptr,
               *p_pos = nullptr, *p_val_valid = nullptr, *p_val = nullptr;

    if (pred && pred->chrom_col >= 0 && pred->region.has_chrom) {
        if (!child_ok(pred->chrom_col)) return fail(EXON_GPU_ERR_ARG, "chrom_col out of range");
        const ArrowArray *ch = batch->children[pred->chrom_col];
        if (strcmp(schema->children[pred->chrom_col]->format, "u") != 0)
            return fail(EXON_GPU_ERR_UNSUPPORTED, "chrom column must be utf8 (\"u\")");
        if (ch->n_buffers != 3 || (n > 0 && !ch->buffers[1])) return fail(EXON_GPU_ERR_ARG, "malformed utf8 array");
        if (pred->region.chrom_len < 0 || pred->region.chrom_len > kMaxChrom || (!pred->region.chrom && pred->region.chrom_len))
            return fail(EXON_GPU_ERR_ARG, "bad chrom literal");
        a.has_chrom = 1;
        a.lit_len = pred->region.chrom_len;
        memcpy(a.lit, pred->region.chrom, (size_t)a.lit_len);
        a.chrom_off = batch->offset + ch->offset;
        const int64_t rows_end = a.chrom_off + n;
        if (ch->buffers[0]) pieces.push_back({ch->buffers[0], (size_t)((rows_end + 7) / 8), &p_chrom_valid});
        pieces.push_back({ch->buffers[1], sizeof(int32_t) * (size_t)(rows_end + 1), &p_chrom_off});
        if (!buffers_on_device) {  // device buffers are used in place: their extent is never needed
            const size_t vbytes = n > 0 ? (size_t)((const int32_t *)ch->buffers[1])[rows_end] : 0;
            pieces.push_back({ch->buffers[2], vbytes, &p_chrom_val});
        } else {
            pieces.push_back({ch->buffers[2], 0, &p_chrom_val});
        }
    }
    if (pred && pred->pos_col >= 0 && pred->region.has_interval) {
        if (!child_ok(pred->pos_col)) return fail(EXON_GPU_ERR_ARG, "pos_col out of range");
        const ArrowArray *ps = batch->children[pred->pos_col];
        if (strcmp(schema->children[pred->pos_col]->format, "l") != 0)
            return fail(EXON_GPU_ERR_UNSUPPORTED, "pos column must be int64 (\"l\")");
        if (ps->n_buffers != 2) return fail(EXON_GPU_ERR_ARG, "malformed int64 array");
        a.has_pos = 1;
        a.lo = pred->region.lo;
        a.hi = pred->region.hi;
        a.pos_off = batch->offset + ps->offset;
        const int64_t rows_end = a.pos_off + n;
        if (ps->buffers[0]) pieces.push_back({ps->buffers[0], (size_t)((rows_end + 7) / 8), &p_pos_valid});
        pieces.push_back({ps->buffers[1], sizeof(int64_t) * (size_t)rows_end, &p_pos});
    }
    if (agg && agg->kind != EXON_GPU_AGG_COUNT_STAR) {
        if (!child_ok(agg->value_col)) return fail(EXON_GPU_ERR_ARG, "value_col out of range");
        const ArrowArray *v = batch->children[agg->value_col];
        const char *f = schema->children[agg->value_col]->format;
        size_t width = 0;
        if (!strcmp(f, "l")) { a.val_type = kValI64; width = 8; }
        else if (!strcmp(f, "g")) { a.val_type = kValF64; width = 8; }
        else if (!strcmp(f, "f")) { a.val_type = kValF32; width = 4; }
        else if (!strcmp(f, "i")) { a.val_type = kValI32; width = 4; }
        else if (agg->kind == EXON_GPU_AGG_COUNT) { a.val_type = kValNone; }
        else return fail(EXON_GPU_ERR_UNSUPPORTED, "cannot aggregate a column of format \"%s\"", f);
        a.val_off = batch->offset + v->offset;
        const int64_t rows_end = a.val_off + n;
        if (v->n_buffers >= 1 && v->buffers[0]) pieces.push_back({v->buffers[0], (size_t)((rows_end + 7) / 8), &p_val_valid});
        if (width) {
            if (v->n_buffers != 2) return fail(EXON_GPU_ERR_ARG, "malformed primitive array");
            pieces.push_back({v->buffers[1], width * (size_t)rows_end, &p_val});
        }
    }
    if (agg) a.agg_kind = agg->kind;

    size_t total = (front + 255) & ~(size_t)255;
    std::vector<size_t> offs(pieces.size());
    for (size_t i = 0; i < pieces.size(); ++i) {
        offs[i] = total;
        if (!buffers_on_device) total += (pieces[i].bytes + 255) & ~(size_t)255;
    }
    if (int rc = c->ensure_scratch(total, host_bytes)) return rc;
    uint8_t *scratch = (uint8_t *)c->scratch;
    for (size_t i = 0; i < pieces.size(); ++i) {
        if (buffers_on_device) {
            *pieces[i].dst = pieces[i].src;
        } else {
            if (pieces[i].bytes)
                CUDA_TRY(cudaMemcpyAsync(scratch + offs[i], pieces[i].src, pieces[i].bytes, cudaMemcpyHostToDevice, c->stream));
            *pieces[i].dst = scratch + offs[i];
        }
    }
    a.chrom_valid = (const uint8_t *)p_chrom_valid;
    a.chrom_offsets = (const int32_t *)p_chrom_off;
    a.chrom_values = (const uint8_t *)p_chrom_val;
    a.pos_valid = (const uint8_t *)p_pos_valid;
    a.pos = (const int64_t *)p_pos;
    a.val_valid = (const uint8_t *)p_val_valid;
    a.val = p_val;
    return EXON_GPU_OK;
}

int launch_agg(Ctx *c, const FilterAggArgs &a) {
    if (a.n_rows > 0) {
        int grid = (int)std::min<int64_t>((a.n_rows + kFaThreads - 1) / kFaThreads, (int64_t)c->sm_count * 8);
        filter_agg_kernel<<<grid, kFaThreads, 0, c->stream>>>(a);
        c->launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    return EXON_GPU_OK;
}

}  // namespace

// Launches the multi-batch kernel over a descriptor table that already lives in device memory.
int filter_agg_multi_launch(Ctx *c, const FaBatchDesc *d_descs, int n_batches, int64_t max_rows, const FaCommon &k,
                            unsigned long long *d_out, bool timed) {
    if (n_batches <= 0 || max_rows <= 0) return EXON_GPU_OK;
    FilterAggArgs a;
    memset(&a, 0, sizeof(a));
    a.has_chrom = k.has_chrom;
    a.lit_len = k.lit_len;
    memcpy(a.lit, k.lit, sizeof(a.lit));
    a.has_pos = k.has_pos;
    a.lo = k.lo;
    a.hi = k.hi;
    a.val_type = k.val_type;
    a.agg_kind = k.agg_kind;
    a.out = d_out;
    const int upb = (int)((max_rows + kUnitRows - 1) / kUnitRows);
    const int64_t n_units = (int64_t)n_batches * upb;
    static int occ = 0;  // persistent grid = exactly one resident wave
    if (!occ) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_agg_multi_kernel<true>, kFaThreads, 0));
        if (occ < 1) occ = 1;
    }
    const int grid = (int)std::min<int64_t>(n_units, (int64_t)c->sm_count * occ);
    if (timed) CUDA_TRY(c->timed_begin(c->stream));
    if (k.has_nulls) filter_agg_multi_kernel<true><<<grid, kFaThreads, 0, c->stream>>>(a, d_descs, n_batches, upb);
    else filter_agg_multi_kernel<false><<<grid, kFaThreads, 0, c->stream>>>(a, d_descs, n_batches, upb);
    c->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    if (timed) {
        CUDA_TRY(c->timed_end(c->stream));
    }
    return EXON_GPU_OK;
}

int fa_common_from(const exon_gpu_pred *pred, const exon_gpu_agg *agg, FaCommon &k) {
    memset(&k, 0, sizeof(k));
    if (pred && pred->chrom_col >= 0 && pred->region.has_chrom) {
        if (pred->region.chrom_len < 0 || pred->region.chrom_len > kMaxChrom || (!pred->region.chrom && pred->region.chrom_len))
            return fail(EXON_GPU_ERR_ARG, "bad chrom literal");
        k.has_chrom = 1;
        k.lit_len = pred->region.chrom_len;
        memcpy(k.lit, pred->region.chrom, (size_t)k.lit_len);
    }
    if (pred && pred->pos_col >= 0 && pred->region.has_interval) {
        k.has_pos = 1;
        k.lo = pred->region.lo;
        k.hi = pred->region.hi;
    }
    k.agg_kind = agg ? agg->kind : EXON_GPU_AGG_COUNT_STAR;
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_filter_agg(exon_gpu_ctx *c, const struct ArrowArray *batch, const struct ArrowSchema *schema,
                        int buffers_on_device, const exon_gpu_pred *pred, const exon_gpu_agg *agg, exon_gpu_partial *out) {
    if (!c || !batch || !schema || !agg || !out) return fail(EXON_GPU_ERR_ARG, "filter_agg: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (agg->kind < EXON_GPU_AGG_COUNT_STAR || agg->kind > EXON_GPU_AGG_AVG)
        return fail(EXON_GPU_ERR_ARG, "filter_agg: unknown aggregate kind %d", agg->kind);
    FilterAggArgs a;
    if (int rc = prepare(c, batch, schema, buffers_on_device, pred, agg, 64, 64, a)) return rc;
    a.out = (unsigned long long *)c->scratch;
    CUDA_TRY(cudaMemsetAsync(c->scratch, 0, 64, c->stream));
    if (int rc = launch_agg(c, a)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_scratch, c->scratch, 24, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const unsigned long long *r = (const unsigned long long *)c->h_scratch;
    out->count = (int64_t)r[0];
    out->sum_i64 = (int64_t)r[1];
    memcpy(&out->sum_f64, &r[2], sizeof(double));
    if (a.val_type == kValI64 || a.val_type == kValI32) out->sum_f64 = (double)out->sum_i64;
    return EXON_GPU_OK;
}

int exon_gpu_filter_agg_accumulate(exon_gpu_ctx *c, const struct ArrowArray *batch, const struct ArrowSchema *schema,
                                   const exon_gpu_pred *pred, const exon_gpu_agg *agg, exon_gpu_partial *device_acc) {
    if (!c || !batch || !schema || !agg || !device_acc) return fail(EXON_GPU_ERR_ARG, "filter_agg_accumulate: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (agg->kind < EXON_GPU_AGG_COUNT_STAR || agg->kind > EXON_GPU_AGG_AVG)
        return fail(EXON_GPU_ERR_ARG, "filter_agg_accumulate: unknown aggregate kind %d", agg->kind);
    FilterAggArgs a;
    if (int rc = prepare(c, batch, schema, 1, pred, agg, 0, 0, a)) return rc;
    a.out = reinterpret_cast<unsigned long long *>(device_acc);
    return launch_agg(c, a);
}

int exon_gpu_filter_agg_batches(exon_gpu_ctx *c, const struct ArrowArray *const *batches, int32_t n_batches,
                                const struct ArrowSchema *schema, const exon_gpu_pred *pred, const exon_gpu_agg *agg,
                                exon_gpu_partial *out) {
    if (!c || (!batches && n_batches) || n_batches < 0 || !schema || !agg || !out)
        return fail(EXON_GPU_ERR_ARG, "filter_agg_batches: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (agg->kind < EXON_GPU_AGG_COUNT_STAR || agg->kind > EXON_GPU_AGG_AVG)
        return fail(EXON_GPU_ERR_ARG, "filter_agg_batches: unknown aggregate kind %d", agg->kind);
    std::lock_guard<std::recursive_mutex> work(c->work_mu);
    std::vector<FaBatchDesc> descs((size_t)n_batches);
    FaCommon k;
    if (int rc = fa_common_from(pred, agg, k)) return rc;
    int64_t max_rows = 0;
    for (int32_t b = 0; b < n_batches; ++b) {
        if (!batches[b]) return fail(EXON_GPU_ERR_ARG, "filter_agg_batches: batch %d is NULL", b);
        FilterAggArgs a;
        if (int rc = prepare(c, batches[b], schema, 1, pred, agg, 0, 0, a)) return rc;
        FaBatchDesc &d = descs[(size_t)b];
        d.chrom_valid = a.chrom_valid;
        d.chrom_offsets = a.chrom_offsets;
        d.chrom_values = a.chrom_values;
        d.chrom_off = a.chrom_off;
        d.pos_valid = a.pos_valid;
        d.pos = a.pos;
        d.pos_off = a.pos_off;
        d.val_valid = a.val_valid;
        d.val = a.val;
        d.val_off = a.val_off;
        d.n_rows = a.n_rows;
        k.val_type = a.val_type;
        if (a.chrom_valid || a.pos_valid || a.val_valid) k.has_nulls = 1;
        max_rows = std::max(max_rows, a.n_rows);
    }
    const size_t table = ((sizeof(FaBatchDesc) * (size_t)n_batches + 255) & ~(size_t)255);
    if (int rc = c->ensure_scratch(table + 64, 64)) return rc;
    uint8_t *scr = (uint8_t *)c->scratch;
    unsigned long long *d_out = (unsigned long long *)(scr + table);
    if (n_batches) CUDA_TRY(cudaMemcpyAsync(scr, descs.data(), sizeof(FaBatchDesc) * (size_t)n_batches, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, 64, c->stream));
    if (int rc = filter_agg_multi_launch(c, (const FaBatchDesc *)scr, n_batches, max_rows, k, d_out, true)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_scratch, d_out, 24, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const unsigned long long *r = (const unsigned long long *)c->h_scratch;
    out->count = (int64_t)r[0];
    out->sum_i64 = (int64_t)r[1];
    memcpy(&out->sum_f64, &r[2], sizeof(double));
    if (k.val_type == kValI64 || k.val_type == kValI32) out->sum_f64 = (double)out->sum_i64;
    return EXON_GPU_OK;
}

int exon_gpu_partial_read(exon_gpu_ctx *c, const exon_gpu_partial *device_acc, int sum_is_integer, exon_gpu_partial *out) {
    if (!c || !device_acc || !out) return fail(EXON_GPU_ERR_ARG, "partial_read: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (int rc = c->ensure_scratch(0, 64)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_scratch, device_acc, sizeof(*out), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out, c->h_scratch, sizeof(*out));
    if (sum_is_integer) out->sum_f64 = (double)out->sum_i64;
    return EXON_GPU_OK;
}

int exon_gpu_region_udf(exon_gpu_ctx *c, int kind, const struct ArrowArray *batch, const struct ArrowSchema *schema,
                        int buffers_on_device, const exon_gpu_pred *pred, uint8_t *out_values, uint8_t *out_valid) {
    if (!c || !batch || !schema || !pred || !out_values) return fail(EXON_GPU_ERR_ARG, "region_udf: NULL argument");
    if (kind < EXON_GPU_UDF_REGION_MATCH || kind > EXON_GPU_UDF_INTERVAL_MATCH) return fail(EXON_GPU_ERR_ARG, "region_udf: unknown kind %d", kind);
    CUDA_TRY(cudaSetDevice(c->device));
    exon_gpu_pred p = *pred;
    // the UDFs always read the columns their signature names, whatever the literal leaves open
    if (kind != EXON_GPU_UDF_INTERVAL_MATCH) {
        if (p.chrom_col < 0 || !p.region.has_chrom) return fail(EXON_GPU_ERR_ARG, "region_udf: a chrom column and a name are required");
    } else {
        p.chrom_col = -1;
    }
    if (kind != EXON_GPU_UDF_CHROM_MATCH) {
        if (p.pos_col < 0) return fail(EXON_GPU_ERR_ARG, "region_udf: a pos column is required");
        if (!p.region.has_interval) { p.region.has_interval = 1; p.region.lo = 1; p.region.hi = INT64_MAX; }
    } else {
        p.pos_col = -1;
    }
    const int64_t n = batch->length;
    const size_t front = 256 + 2 * (size_t)((n + 255) & ~(int64_t)255);
    FilterAggArgs a;
    if (int rc = prepare(c, batch, schema, buffers_on_device, &p, nullptr, front, front, a)) return rc;
    uint8_t *scratch = (uint8_t *)c->scratch;
    a.udf_kind = kind;
    a.mask_err = reinterpret_cast<uint32_t *>(scratch);
    a.mask_values = scratch + 256;
    a.mask_valid = scratch + 256 + ((n + 255) & ~(int64_t)255);
    CUDA_TRY(cudaMemsetAsync(scratch, 0, 256, c->stream));
    if (n > 0) {
        region_udf_kernel<<<(unsigned)((n + kFaThreads - 1) / kFaThreads), kFaThreads, 0, c->stream>>>(a);
        c->launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(c->h_scratch, scratch, front, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const uint8_t *h = (const uint8_t *)c->h_scratch;
    const uint32_t err = *reinterpret_cast<const uint32_t *>(h);
    if (err & kUdfErrNull) return fail(EXON_GPU_ERR_PARSE, "Failed to get %s", "chrom or pos (NULL operand)");
    if (err & kUdfErrPos) return fail(EXON_GPU_ERR_PARSE, "Failed to convert pos: a position must be >= 1");
    memcpy(out_values, h + 256, (size_t)n);
    if (out_valid) memcpy(out_valid, h + 256 + ((n + 255) & ~(int64_t)255), (size_t)n);
    return EXON_GPU_OK;
}

}  // extern "C"
