// gff_scan.cu -- GFF3 text -> region predicate -> COUNT, fused.
//
// Replaces, for `SELECT COUNT(*) FROM gff [WHERE gff_region_filter('name[:lo-hi]', seqname[, start])]`, the chain
//   BatchReader::read_line / read_batch   exon/exon-gff/src/batch_reader.rs:56-130 (noodles-gff Line: "##" directive,
//                                         "#" comment, otherwise a record of 9 tab-separated fields)
//   BatchReader::filter                   exon/exon-gff/src/batch_reader.rs:70-96 (seqname == region name, and the
//                                         region's interval, if any, contains the record's START)
//   GFFArrayBuilder::append + AggregateExec count(*)
// One pass of the warp-private TMA tile pipeline (tile_ring.cuh).  GFF lines are long (~100-300 bytes), so there are
// only a few line starts per 512-byte row: every lane handles the line starts of its own 16-byte chunk with a
// byte-exact field walk (first byte '#' -> not a record; seqname compare; two tabs skipped; START parsed as a decimal).
// Like K1's default mode, only the fields the predicate reads are validated (seqname non-empty; START a positive
// decimal when an interval is asked for); a well-formed file gives the reference's count bit-exactly.
#include <cstring>

#include "common.cuh"
#include "internal.h"
#include "tile_ring.cuh"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

using GffRing = TileRing<4096, 2, 8, 16, 112, 16>;
constexpr int kGffU = GffRing::TILE / 512;
constexpr uint32_t kGffErrEmptyLine = 1u, kGffErrBadStart = 2u, kGffErrShort = 4u;

struct GffArgs {
    const ScanSeg *segs;
    int64_t n_tiles;
    int32_t has_name, has_interval, name_len;
    uint8_t name[kMaxChrom + 1];
    int64_t lo, hi;
    unsigned long long *out;  // [0] selected records [1] records
    uint32_t *flags;
};

// the record line that starts at tile index ls: 1 if it satisfies the predicate
template <class V>
__device__ __noinline__ uint32_t gff_line(const V &v, int ls, const GffArgs *ap, uint32_t *err) {
    const GffArgs &a = *ap;
    if (!a.has_name && !a.has_interval) return 1;
    int p = ls;
    bool name_ok = true;
    if (a.has_name) {
        for (int j = 0; j < a.name_len; ++j)
            if (view_byte(v, p + j) != a.name[j]) { name_ok = false; break; }
        if (name_ok && view_byte(v, p + a.name_len) != '\t') name_ok = false;
        if (!name_ok) return 0;
        p += a.name_len + 1;
    } else {
        uint32_t c;
        while ((c = view_byte(v, p)) != '\t') {
            if (c == '\n') { *err |= kGffErrShort; return 0; }
            ++p;
        }
        if (p == ls) { *err |= kGffErrShort; return 0; }
        ++p;
    }
    if (!a.has_interval) return 1;
    for (int f = 0; f < 2; ++f) {  // source, type
        uint32_t c;
        while ((c = view_byte(v, p)) != '\t') {
            if (c == '\n') { *err |= kGffErrShort; return 0; }
            ++p;
        }
        ++p;
    }
    unsigned long long val = 0;
    int nd = 0;
    uint32_t c;
    while ((c = view_byte(v, p)) - '0' <= 9u) {
        if (nd < 19) val = val * 10ull + (c - '0');
        ++nd;
        ++p;
    }
    if (nd == 0 || nd > 18 || c != '\t' || val == 0ull) { *err |= (c == '\n' && nd == 0) ? kGffErrShort : kGffErrBadStart; return 0; }
    return ((long long)val >= a.lo) & ((long long)val <= a.hi);
}

__global__ void __launch_bounds__(GffRing::WARPS * 32, 3) gff_scan_kernel(const __grid_constant__ GffArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    GffRing ring;
    ring.init(smem_raw, a.segs, a.n_tiles);
    const int lane = ring.lane;
    uint32_t cnt = 0, rows = 0, err = 0;
#pragma unroll 1
    for (int64_t T = ring.first_tile(); T < a.n_tiles; T += ring.nw) {
        const GffRing::View v = ring.acquire();
        auto on_line = [&](int ls) {
            const uint32_t c = view_byte(v, ls);
            if (c == '#') return;            // directive or comment
            if (c == '\n') { err |= kGffErrEmptyLine; return; }
            rows += 1;
            cnt += gff_line(v, ls, &a, &err);
        };
        if (v.first && v.hi > v.seg_lo && lane == 0) on_line(v.seg_lo);
#pragma unroll 1
        for (int u = 0; u < kGffU; ++u) {
            const int c0 = (u * 32 + lane) * 16;
            uint32_t m = 0;
            if (v.interior || c0 < v.sm_hi) m = newline_mask16(lds128(v.sa + (uint32_t)c0));
            if (!v.interior) m = clip_mask16(m, c0, v.seg_lo, v.hi);
            while (m) {
                on_line(c0 + __ffs(m));
                m &= m - 1;
            }
        }
        ring.release(T);
    }
    cnt = warp_sum(cnt);
    rows = warp_sum(rows);
    err = __reduce_or_sync(0xFFFFFFFFu, err);
    if (lane == 0) {
        if (cnt) atomicAdd(a.out, (unsigned long long)cnt);
        if (rows) atomicAdd(a.out + 1, (unsigned long long)rows);
        if (err) atomicOr(a.flags, err);
    }
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int gff_filter_count(VcfStream *s, const exon_gpu_region *region, int64_t *out_count, int64_t *out_rows) {
    if (int rc = s->flush_gz()) return rc;
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    std::lock_guard<std::recursive_mutex> work(ctx->work_mu);
    if (out_count) *out_count = 0;
    if (out_rows) *out_rows = 0;
    OwnedRegion r;
    if (int rc = r.assign(region)) return rc;
    if (r.has_chrom && r.chrom.size() > (size_t)kMaxChrom) return EXON_GPU_OK;  // no seqname can equal it
    std::vector<Piece> pieces;
    s->cut_pieces(pieces);
    if (pieces.empty()) return EXON_GPU_OK;
    std::vector<ScanSeg> h_segs;
    int64_t n_tiles = 0;
    for (const Piece &p : pieces) {
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)p.base & 15);
        sg.base = p.base - sg.skip;
        sg.len = p.len;
        sg.tile0 = n_tiles;
        sg.pad_ = 0;
        n_tiles += (sg.skip + p.len + GffRing::TILE - 1) / GffRing::TILE;
        h_segs.push_back(sg);
    }
    ScanSeg sentinel;
    memset(&sentinel, 0, sizeof(sentinel));
    sentinel.tile0 = n_tiles;
    h_segs.push_back(sentinel);
    const size_t o_out = al256(h_segs.size() * sizeof(ScanSeg));
    if (int rc = ctx->ensure_scratch(o_out + 256, 64)) return rc;
    uint8_t *scr = (uint8_t *)ctx->scratch;
    CUDA_TRY(cudaMemcpyAsync(scr, h_segs.data(), h_segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(scr + o_out, 0, 64, st));
    GffArgs a;
    memset(&a, 0, sizeof(a));
    a.segs = (const ScanSeg *)scr;
    a.n_tiles = n_tiles;
    a.has_name = r.has_chrom;
    a.has_interval = r.has_interval;
    a.name_len = (int32_t)r.chrom.size();
    memcpy(a.name, r.chrom.data(), r.chrom.size());
    a.lo = r.lo;
    a.hi = r.hi;
    a.out = (unsigned long long *)(scr + o_out);
    a.flags = (uint32_t *)(scr + o_out + 16);
    static int occ = 0;
    if (!occ) {
        CUDA_TRY(cudaFuncSetAttribute(gff_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GffRing::smem_bytes));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gff_scan_kernel, GffRing::WARPS * 32, GffRing::smem_bytes));
        if (occ < 1) occ = 1;
    }
    int64_t grid = std::min<int64_t>((int64_t)occ * ctx->sm_count, (n_tiles + GffRing::WARPS - 1) / GffRing::WARPS);
    if (grid < 1) grid = 1;
    CUDA_TRY(ctx->timed_begin(st));
    gff_scan_kernel<<<(unsigned)grid, GffRing::WARPS * 32, GffRing::smem_bytes, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(ctx->timed_end(st));
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaMemcpyAsync(s->h_res, a.out, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint32_t flags = (uint32_t)s->h_res[2];
    if (flags)
        return fail(EXON_GPU_ERR_PARSE, "malformed GFF record:%s%s%s", (flags & kGffErrEmptyLine) ? " empty line;" : "",
                    (flags & kGffErrBadStart) ? " start is not a positive decimal integer;" : "", (flags & kGffErrShort) ? " line ended before the field being read;" : "");
    if (out_count) *out_count = (int64_t)s->h_res[0];
    if (out_rows) *out_rows = (int64_t)s->h_res[1];
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_gff_open(exon_gpu_ctx *c, exon_gpu_stream **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "gff_open: NULL argument");
    exon_gpu_vcf_opts vo;
    memset(&vo, 0, sizeof(vo));
    if (int rc = exon_gpu_vcf_open(c, &vo, out)) return rc;
    (*out)->fmt = kFmtGff;
    (*out)->hdr = VcfStream::kBody;  // directives and comments are lines like any other: the kernel skips them
    return EXON_GPU_OK;
}

int exon_gpu_gff_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last) {
    if (!s || s->fmt != kFmtGff) return fail(EXON_GPU_ERR_ARG, "gff_feed: not a GFF stream");
    return exon_gpu_vcf_feed(s, text, len, is_device_ptr, is_last);
}

int exon_gpu_gff_filter_count(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *out_count) {
    if (!s || !out_count) return fail(EXON_GPU_ERR_ARG, "gff_filter_count: NULL argument");
    if (s->fmt != kFmtGff) return fail(EXON_GPU_ERR_ARG, "gff_filter_count: not a GFF stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return gff_filter_count(s, region, out_count, nullptr);
}

}  // extern "C"
