// fastq_scan.cu -- FASTQ text -> mean-quality predicate -> COUNT, fused (BASELINE.json configs[1]).
//
// Replaces, for `SELECT COUNT(*) FROM fastq WHERE mean(quality) > T`, the chain
//   BatchReader::read_batch / read_record   exon/exon-fastq/src/batch_reader.rs:56-82   (noodles-fastq 0.16 reader)
//   FASTQArrayBuilder::append col 3          exon/exon-fastq/src/array_builder.rs:68-102
//   quality_scores_to_list (byte - 33)       exon/exon-core/src/udfs/sequence/quality_score_string_to_list.rs:80-93
//   FilterExec + AggregateExec(Partial) count                                          (DataFusion 44, third party)
// without materialising columns.
//
// A FASTQ record is exactly four lines, and '@' / '+' may also begin a quality line, so a line's role is known
// only from its index in the file modulo 4.  Two passes of the warp-private TMA tile pipeline (tile_ring.cuh):
//   A. lines    per 4 KiB tile: line starts in the tile                              (reads the text once)
//      scan     exclusive prefix over tiles (cub); a tiny kernel turns it into the file-relative line index of
//               every segment's first line and checks that no file ends inside a record
//   B. filter   every warp re-reads its tiles: '\n' masks per 16-byte chunk, rank of every line start (warp
//               ballots), role = index mod 4; definition and '+' lines are validated ('@' / '+' first byte);
//               quality-line starts go to a per-warp queue that is drained one line per lane: the lane sums the
//               bytes of its line 16 at a time with IDP.4A (sum of 4 bytes per instruction) until the '\n', and
//               evaluates  sum - 33 * len  >  T * len  over integers
// A line belongs to the tile that holds the '\n' before it.  Lines longer than the staged halo continue from
// global memory (byte loop), so read length is not limited.
#include <cub/device/device_scan.cuh>

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

constexpr int kFqQueue = 256;
using FqRing = TileRing<4096, 2, 8, 16, 368, kFqQueue * 2>;
constexpr int kFqU = FqRing::TILE / 512;

constexpr uint32_t kFqErrPrefix = 1u;     // a definition line without '@' or a third line without '+'
constexpr uint32_t kFqErrTruncated = 2u;  // a file ends after the first or second line of a record

struct FqArgs {
    const ScanSeg *segs;
    int32_t n_segs;
    int64_t n_tiles;
    unsigned long long *tile_lines;         // pass A out
    const unsigned long long *tile_prefix;  // pass B in: lines that start in earlier tiles
    const unsigned long long *seg_line0;    // pass B in: per segment, prefix value at the first tile of its FILE
    int32_t has_pred, phred_offset;
    long long num, den;                     // mean > num / den
    unsigned long long *out;                // [0] selected records [1] records
    uint32_t *flags;
};

// ---- pass A -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FqRing::WARPS * 32, 3) fq_lines_kernel(const __grid_constant__ FqArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FqRing ring;
    ring.init(smem_raw, a.segs, a.n_tiles);
    const int lane = ring.lane;
#pragma unroll 1
    for (int64_t T = ring.first_tile(); T < a.n_tiles; T += ring.nw) {
        const FqRing::View v = ring.acquire();
        uint32_t n = 0;
        if (v.interior) {
            uint32_t acc = 0;
#pragma unroll
            for (int u = 0; u < kFqU; ++u) {
                const uint4 w = lds128(v.sa + (uint32_t)((u * 32 + lane) * 16));
                acc = __dp4a(zero_bytes_exact(w.x ^ kNL4), 0x01010101u, acc);
                acc = __dp4a(zero_bytes_exact(w.y ^ kNL4), 0x01010101u, acc);
                acc = __dp4a(zero_bytes_exact(w.z ^ kNL4), 0x01010101u, acc);
                acc = __dp4a(zero_bytes_exact(w.w ^ kNL4), 0x01010101u, acc);
            }
            n = acc >> 7;
        } else {
#pragma unroll 1
            for (int u = 0; u < kFqU; ++u) {
                const int c0 = (u * 32 + lane) * 16;
                if (c0 < v.sm_hi) n += __popc(clip_mask16(newline_mask16(lds128(v.sa + (uint32_t)c0)), c0, v.seg_lo, v.hi));
            }
        }
        n = warp_sum(n);
        if (lane == 0) a.tile_lines[T] = (unsigned long long)n + ((v.first && v.hi > v.seg_lo) ? 1u : 0u);
        ring.release(T);
    }
}

// One thread per segment: file-relative line base of the segment, and the truncation check once per file.
// file_first[s] = index of the first segment of s's file; file_next[s] = first segment of the next file (n_segs at the end).
__global__ void fq_segment_table(const ScanSeg *segs, int n_segs, const unsigned long long *prefix, const int32_t *file_first,
                                 const int32_t *file_next, unsigned long long *seg_line0, uint32_t *flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const unsigned long long l0 = prefix[segs[file_first[s]].tile0];
    seg_line0[s] = l0;
    if (file_first[s] == s) {
        const unsigned long long lines = prefix[segs[file_next[s]].tile0] - l0;  // segs[n_segs] is the sentinel (tile0 = n_tiles)
        const unsigned r = (unsigned)(lines & 3ull);
        if (r == 1u || r == 2u) atomicOr(flags, kFqErrTruncated);
    }
}

// ---- pass B -------------------------------------------------------------------------------------------------
// Sum and length of the line that starts at tile index ls (ends at '\n' or at the segment end).
__device__ __forceinline__ void fq_line_sum(const FqRing::View &v, int ls, uint32_t &sum, uint32_t &len) {
    sum = 0;
    len = 0;
    int lim = v.hi < v.sm_hi ? v.hi : v.sm_hi;  // staged segment bytes: [.., lim)
    lim &= ~15;
    int c = ls & ~15;
    bool done = false;
    if (c < lim) {
        // first chunk: bytes before the line start are cleared (0 is neither '\n' nor adds to the sum)
        uint4 w = lds128(v.sa + (uint32_t)c);
        const int off = ls - c;
        uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = off - 4 * q;  // bytes of word q to clear
            if (k >= 4) ws[q] = 0;
            else if (k > 0) ws[q] &= 0xFFFFFFFFu << (8 * k);
        }
        int valid = 16 - off;
#pragma unroll 1
        while (true) {
            const uint32_t f0 = zero_bytes_exact(ws[0] ^ kNL4), f1 = zero_bytes_exact(ws[1] ^ kNL4), f2 = zero_bytes_exact(ws[2] ^ kNL4),
                           f3 = zero_bytes_exact(ws[3] ^ kNL4);
            if ((f0 | f1 | f2 | f3) != 0u) {
                // the line ends in this chunk: keep the bytes before the first '\n'
                const uint32_t m = pack_flags16(f0, f1, f2, f3);
                const int e = __ffs(m) - 1;  // chunk index of the '\n'
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = e - 4 * q;  // bytes of word q to keep
                    uint32_t x = ws[q];
                    if (k <= 0) x = 0;
                    else if (k < 4) x &= (1u << (8 * k)) - 1u;
                    sum = __dp4a(x, 0x01010101u, sum);
                }
                len += (uint32_t)(e - (16 - valid));
                done = true;
                break;
            }
            sum = __dp4a(ws[0], 0x01010101u, sum);
            sum = __dp4a(ws[1], 0x01010101u, sum);
            sum = __dp4a(ws[2], 0x01010101u, sum);
            sum = __dp4a(ws[3], 0x01010101u, sum);
            len += (uint32_t)valid;
            c += 16;
            if (c >= lim) break;
            w = lds128(v.sa + (uint32_t)c);
            ws[0] = w.x;
            ws[1] = w.y;
            ws[2] = w.z;
            ws[3] = w.w;
            valid = 16;
        }
    }
    if (!done) {
        // the rest of the line lies beyond the staged window (or the tile is a short boundary tile): byte loop
        int p = c > ls ? c : ls;
        uint32_t b;
        while ((b = view_byte(v, p)) != '\n') {
            sum += b;
            ++len;
            ++p;
        }
    }
}

__global__ void __launch_bounds__(FqRing::WARPS * 32, 3) fq_filter_kernel(const __grid_constant__ FqArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FqRing ring;
    ring.init(smem_raw, a.segs, a.n_tiles);
    const int lane = ring.lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t queue_sa = smem_u32(ring.extra);
    uint32_t cnt = 0, rows = 0, err = 0;
#pragma unroll 1
    for (int64_t T = ring.first_tile(); T < a.n_tiles; T += ring.nw) {
        const FqRing::View v = ring.acquire();
        // file-relative index of the first line that starts in this tile
        uint32_t line = (uint32_t)(__ldg(&a.tile_prefix[T]) - __ldg(&a.seg_line0[v.seg]));
        int qn = 0;

        auto drain = [&]() {
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < qn; i += 32) {
                const int ls = (int)lds16(queue_sa + 2u * (uint32_t)i);
                uint32_t sum, len;
                fq_line_sum(v, ls, sum, len);
                // mean(byte - off) > num / den  <=>  (sum - off * len) * den > num * len, len > 0
                const long long q = (long long)sum - (long long)a.phred_offset * (long long)len;
                cnt += (len > 0u) && (q * a.den > a.num * (long long)len);
            }
            __syncwarp();
            qn = 0;
        };
        // role of the line with file-relative index `idx` that starts at tile index `ls`
        auto on_line = [&](uint32_t idx, int ls, bool enqueue_now, int slot) {
            const uint32_t role = idx & 3u;
            if (role == 0u) {
                rows += 1;
                if (view_byte(v, ls) != '@') err |= kFqErrPrefix;
            } else if (role == 2u) {
                if (view_byte(v, ls) != '+') err |= kFqErrPrefix;
            } else if (role == 3u && a.has_pred && enqueue_now) {
                sts16(queue_sa + 2u * (uint32_t)slot, (uint32_t)ls);
            }
        };

        if (v.first && v.hi > v.seg_lo) {  // the segment's first line has no '\n' before it
            const bool q = (line & 3u) == 3u && a.has_pred;
            if (lane == 0) on_line(line, v.seg_lo, true, 0);
            qn = q ? 1 : 0;
            line += 1;
        }
#pragma unroll 1
        for (int u = 0; u < kFqU; ++u) {
            const int c0 = (u * 32 + lane) * 16;
            uint32_t m = 0;
            if (v.interior || c0 < v.sm_hi) m = newline_mask16(lds128(v.sa + (uint32_t)c0));
            if (!v.interior) m = clip_mask16(m, c0, v.seg_lo, v.hi);
            const uint32_t n = (uint32_t)__popc(m);
            const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, n >= 1u), b2 = __ballot_sync(0xFFFFFFFFu, n >= 2u);
            if (b1 == 0u) continue;
            uint32_t before, total;
            if (__ballot_sync(0xFFFFFFFFu, n >= 3u) == 0u) {
                before = (uint32_t)(__popc(b1 & lt_mask) + __popc(b2 & lt_mask));
                total = (uint32_t)(__popc(b1) + __popc(b2));
            } else {
                uint32_t incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += x;
                }
                before = incl - n;
                total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            // quality lines among this lane's line starts: at most one per lane unless lines are tiny
            uint32_t mm = m, idx = line + before;
            int my_q = -1, extra_q = 0;
            while (mm) {
                const int ls = c0 + __ffs(mm);
                mm &= mm - 1;
                const uint32_t role = idx & 3u;
                if (role == 3u && a.has_pred) {
                    if (my_q < 0) my_q = ls;
                    else ++extra_q;
                } else {
                    on_line(idx, ls, false, 0);
                }
                ++idx;
            }
            const uint32_t bq = __ballot_sync(0xFFFFFFFFu, my_q >= 0);
            if (my_q >= 0) sts16(queue_sa + 2u * (uint32_t)(qn + __popc(bq & lt_mask)), (uint32_t)my_q);
            qn += __popc(bq);
            if (__ballot_sync(0xFFFFFFFFu, extra_q > 0) != 0u) {
                // >= 2 quality lines start inside one 16-byte chunk (records shorter than 16 bytes): those beyond the
                // first are summed right here by their lane
                uint32_t m2 = m, i2 = line + before;
                bool seen = false;
                while (m2) {
                    const int ls = c0 + __ffs(m2);
                    m2 &= m2 - 1;
                    if ((i2 & 3u) == 3u) {
                        if (seen) {
                            uint32_t sum, len;
                            fq_line_sum(v, ls, sum, len);
                            const long long q = (long long)sum - (long long)a.phred_offset * (long long)len;
                            cnt += (len > 0u) && (q * a.den > a.num * (long long)len);
                        }
                        seen = true;
                    }
                    ++i2;
                }
            }
            line += total;
            if (qn > kFqQueue - 32) drain();
        }
        if (qn) drain();
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

template <class K>
cudaError_t fq_launch(K kern, const FqArgs &a, int sm_count, cudaStream_t st, int *occ_cache) {
    constexpr size_t smem = FqRing::smem_bytes;
    if (!*occ_cache) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_cache, kern, FqRing::WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        if (*occ_cache < 1) return cudaErrorLaunchOutOfResources;
    }
    int64_t grid = (int64_t)*occ_cache * sm_count;
    const int64_t need = (a.n_tiles + FqRing::WARPS - 1) / FqRing::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, FqRing::WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// FASTQ streams reuse the partition-stream machinery (arena, runs, file marks); this is their fused query.
int fastq_filter_count(VcfStream *s, const exon_gpu_fastq_pred *pred, int64_t *out_count, int64_t *out_rows) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    if (int rc = s->flush_gz()) return rc;
    if (pred && (pred->min_mean_den <= 0)) return fail(EXON_GPU_ERR_ARG, "fastq_filter_count: min_mean_den must be positive");
    std::lock_guard<std::mutex> work(ctx->work_mu);
    std::vector<Piece> pieces;
    s->cut_pieces(pieces);
    if (out_count) *out_count = 0;
    if (out_rows) *out_rows = 0;
    if (pieces.empty()) return EXON_GPU_OK;
    std::vector<ScanSeg> h_segs;
    std::vector<int32_t> file_first, file_next;
    int64_t n_tiles = 0;
    for (const Piece &p : pieces) {
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)p.base & 15);
        sg.base = p.base - sg.skip;
        sg.len = p.len;
        sg.tile0 = n_tiles;
        sg.pad_ = 0;
        n_tiles += (sg.skip + p.len + FqRing::TILE - 1) / FqRing::TILE;
        file_first.push_back(p.starts_file || h_segs.empty() ? (int32_t)h_segs.size() : file_first.back());
        h_segs.push_back(sg);
    }
    const int n_segs = (int)h_segs.size();
    file_next.assign((size_t)n_segs, n_segs);
    for (int i = n_segs - 2; i >= 0; --i) file_next[(size_t)i] = file_first[(size_t)i + 1] != file_first[(size_t)i] ? i + 1 : file_next[(size_t)i + 1];
    ScanSeg sentinel;
    sentinel.base = nullptr;
    sentinel.len = 0;
    sentinel.tile0 = n_tiles;
    sentinel.skip = 0;
    sentinel.pad_ = 0;
    h_segs.push_back(sentinel);

    size_t cub_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)(n_tiles + 1), st));
    const size_t o_segs = 0;
    const size_t o_lines = o_segs + al256(h_segs.size() * sizeof(ScanSeg));
    const size_t o_prefix = o_lines + al256((size_t)(n_tiles + 1) * 8);
    const size_t o_cub = o_prefix + al256((size_t)(n_tiles + 1) * 8);
    const size_t o_ff = o_cub + al256(cub_bytes);
    const size_t o_fn = o_ff + al256((size_t)n_segs * 4);
    const size_t o_l0 = o_fn + al256((size_t)n_segs * 4);
    const size_t o_out = o_l0 + al256((size_t)n_segs * 8);
    if (int rc = ctx->ensure_scratch(o_out + 256, 64)) return rc;
    uint8_t *scr = (uint8_t *)ctx->scratch;
    CUDA_TRY(cudaMemcpyAsync(scr + o_segs, h_segs.data(), h_segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_ff, file_first.data(), (size_t)n_segs * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_fn, file_next.data(), (size_t)n_segs * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(scr + o_out, 0, 64, st));
    CUDA_TRY(cudaMemsetAsync(scr + o_lines + (size_t)n_tiles * 8, 0, 8, st));

    FqArgs a;
    memset(&a, 0, sizeof(a));
    a.segs = (const ScanSeg *)(scr + o_segs);
    a.n_segs = n_segs;
    a.n_tiles = n_tiles;
    a.tile_lines = (unsigned long long *)(scr + o_lines);
    a.tile_prefix = (const unsigned long long *)(scr + o_prefix);
    a.seg_line0 = (const unsigned long long *)(scr + o_l0);
    a.has_pred = pred != nullptr;
    a.phred_offset = pred ? pred->phred_offset : 33;
    a.num = pred ? pred->min_mean_num : 0;
    a.den = pred ? pred->min_mean_den : 1;
    a.out = (unsigned long long *)(scr + o_out);
    a.flags = (uint32_t *)(scr + o_out + 16);

    static int occ_a = 0, occ_b = 0;
    CUDA_TRY(cudaEventRecord(ctx->ev0, st));
    CUDA_TRY(fq_launch(fq_lines_kernel, a, ctx->sm_count, st, &occ_a));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(scr + o_cub, cub_bytes, a.tile_lines, (unsigned long long *)(scr + o_prefix), (int)(n_tiles + 1), st));
    fq_segment_table<<<(n_segs + 127) / 128, 128, 0, st>>>(a.segs, n_segs, a.tile_prefix, (const int32_t *)(scr + o_ff),
                                                           (const int32_t *)(scr + o_fn), (unsigned long long *)(scr + o_l0), a.flags);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(fq_launch(fq_filter_kernel, a, ctx->sm_count, st, &occ_b));
    CUDA_TRY(cudaEventRecord(ctx->ev1, st));
    ctx->timed = true;
    ctx->launches.fetch_add(4);
    CUDA_TRY(cudaMemcpyAsync(s->h_res, a.out, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint32_t flags = (uint32_t)s->h_res[2];
    if (flags)
        return fail(EXON_GPU_ERR_PARSE, "malformed FASTQ record:%s%s", (flags & kFqErrPrefix) ? " invalid name prefix or missing '+' line;" : "",
                    (flags & kFqErrTruncated) ? " unexpected end of file inside a record;" : "");
    if (out_count) *out_count = pred ? (int64_t)s->h_res[0] : (int64_t)s->h_res[1];
    if (out_rows) *out_rows = (int64_t)s->h_res[1];
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_fastq_open(exon_gpu_ctx *c, const exon_gpu_fastq_opts *o, exon_gpu_stream **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "fastq_open: NULL argument");
    *out = nullptr;
    if (o && o->n_projection > 0)
        return fail(EXON_GPU_ERR_UNSUPPORTED, "fastq_open: column batches are not built on the GPU yet; open with n_projection = 0 "
                                              "and use exon_gpu_fastq_filter_count");
    exon_gpu_vcf_opts vo;
    memset(&vo, 0, sizeof(vo));
    vo.batch_rows = o ? o->batch_rows : 0;
    if (int rc = exon_gpu_vcf_open(c, &vo, out)) return rc;
    (*out)->fmt = kFmtFastq;
    (*out)->hdr = VcfStream::kBody;  // no header: every byte is record data
    return EXON_GPU_OK;
}

int exon_gpu_fastq_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last) {
    if (!s || s->fmt != kFmtFastq) return fail(EXON_GPU_ERR_ARG, "fastq_feed: not a FASTQ stream");
    return exon_gpu_vcf_feed(s, text, len, is_device_ptr, is_last);
}

int exon_gpu_fastq_filter_count(exon_gpu_stream *s, const exon_gpu_fastq_pred *pred, int64_t *out_count) {
    if (!s || !out_count) return fail(EXON_GPU_ERR_ARG, "fastq_filter_count: NULL argument");
    if (s->fmt != kFmtFastq) return fail(EXON_GPU_ERR_ARG, "fastq_filter_count: not a FASTQ stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return fastq_filter_count(s, pred, out_count, nullptr);
}

int exon_gpu_fastq_rows(exon_gpu_stream *s, int64_t *out_rows) {
    if (!s || !out_rows) return fail(EXON_GPU_ERR_ARG, "fastq_rows: NULL argument");
    if (s->fmt != kFmtFastq) return fail(EXON_GPU_ERR_ARG, "fastq_rows: not a FASTQ stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return fastq_filter_count(s, nullptr, nullptr, out_rows);
}

int exon_gpu_stream_close(exon_gpu_stream *s) { return exon_gpu_vcf_close(s); }
int exon_gpu_stream_reset(exon_gpu_stream *s) { return exon_gpu_vcf_reset(s); }
int exon_gpu_stream_body_bytes(exon_gpu_stream *s, int64_t *out_bytes) { return exon_gpu_vcf_body_bytes(s, out_bytes); }

}  // extern "C"
