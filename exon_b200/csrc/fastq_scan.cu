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

#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>

#include "common.cuh"
#include "internal.h"
#include "scan_i64.cuh"
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
#ifndef EXON_FQ_TILE
#define EXON_FQ_TILE 4096
#endif
#ifndef EXON_FQ_WARPS
#define EXON_FQ_WARPS 8
#endif
#ifndef EXON_FQ_MINB
#define EXON_FQ_MINB 3
#endif
using FqRing = TileRing<EXON_FQ_TILE, 2, EXON_FQ_WARPS, 16, 368, kFqQueue * 2>;
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
__global__ void __launch_bounds__(FqRing::WARPS * 32, EXON_FQ_MINB) fq_lines_kernel(const __grid_constant__ FqArgs a) {
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

__global__ void __launch_bounds__(FqRing::WARPS * 32, EXON_FQ_MINB) fq_filter_kernel(const __grid_constant__ FqArgs a) {
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
        // 32 bytes per lane and step: the ballots, the rank arithmetic and the (divergent) per-line-start work below are paid
        // once per KiB instead of once per 512 bytes
#pragma unroll 1
        for (int u = 0; u < kFqU / 2; ++u) {
            const int c0 = (u * 32 + lane) * 32;
            uint32_t m = 0;
            if (v.interior || c0 < v.sm_hi) m = newline_mask16(lds128(v.sa + (uint32_t)c0));
            if (v.interior || c0 + 16 < v.sm_hi) m |= newline_mask16(lds128(v.sa + (uint32_t)c0 + 16u)) << 16;
            if (!v.interior) {  // '\n' at tile index p starts a line iff p >= seg_lo and p + 1 < hi
                const int j_lo = v.seg_lo - c0 > 0 ? v.seg_lo - c0 : 0;
                const int j_hi = v.hi - 1 - c0 < 32 ? v.hi - 1 - c0 : 32;
                m = (j_hi > j_lo) ? (m & (j_hi >= 32 ? 0xFFFFFFFFu : (1u << j_hi) - 1u) & ~((1u << j_lo) - 1u)) : 0u;
            }
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
                // >= 2 quality lines start inside one 32-byte chunk (records shorter than 32 bytes): those beyond the
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
    std::lock_guard<std::recursive_mutex> work(ctx->work_mu);
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
    CUDA_TRY(ctx->timed_begin(st));
    CUDA_TRY(fq_launch(fq_lines_kernel, a, ctx->sm_count, st, &occ_a));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(scr + o_cub, cub_bytes, a.tile_lines, (unsigned long long *)(scr + o_prefix), (int)(n_tiles + 1), st));
    fq_segment_table<<<(n_segs + 127) / 128, 128, 0, st>>>(a.segs, n_segs, a.tile_prefix, (const int32_t *)(scr + o_ff),
                                                           (const int32_t *)(scr + o_fn), (unsigned long long *)(scr + o_l0), a.flags);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(fq_launch(fq_filter_kernel, a, ctx->sm_count, st, &occ_b));
    CUDA_TRY(ctx->timed_end(st));
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

// =====================================================================================================================
// FASTQ text -> Arrow columns {name, description, sequence, quality_scores} in reference-sized batches.
//
// Replaces BatchReader::read_batch (exon/exon-fastq/src/batch_reader.rs:63-82), FASTQArrayBuilder::{append, finish}
// (exon/exon-fastq/src/array_builder.rs:68-118: four GenericStringBuilder<i32>, description NULL when empty) and
// ExonArrayBuilder::try_into_record_batch (exon/exon-common/src/array_builder.rs:25-36).
//   1. lines  + scan          (as in the fused query) -> first line index of every tile / file
//   2. index                  every '\n' knows its rank: line_start[i], line_end[i] for all lines (TMA tile pipeline)
//   3. fields                 one thread per record: the four (offset, length) pairs, '@' / '+' validation
//   4. 4 exclusive scans      value offsets per column (cub)
//   5. gather                 one warp per record copies its four byte ranges to their final place and writes the
//                             batch-relative int32 offsets and the description validity bit
// Batches restart at every file and are zero-copy views of the column store (reference counted, Arrow release).
// =====================================================================================================================
namespace {

struct FqIndexArgs {
    const ScanSeg *segs;
    int64_t n_tiles;
    const unsigned long long *tile_prefix;
    const uint8_t **line_start;
    const uint8_t **line_end;
};

__global__ void __launch_bounds__(FqRing::WARPS * 32, EXON_FQ_MINB) fq_index_kernel(const __grid_constant__ FqIndexArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FqRing ring;
    ring.init(smem_raw, a.segs, a.n_tiles);
    const int lane = ring.lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll 1
    for (int64_t T = ring.first_tile(); T < a.n_tiles; T += ring.nw) {
        const FqRing::View v = ring.acquire();
        // index of the line ended by the first '\n' of this tile
        unsigned long long next_end = __ldg(&a.tile_prefix[T]) - 1ull;
        if (v.first && v.hi > v.seg_lo) {
            if (lane == 0) a.line_start[next_end + 1ull] = v.g + v.seg_lo;
            next_end += 1ull;
        }
#pragma unroll 1
        for (int u = 0; u < kFqU; ++u) {
            const int c0 = (u * 32 + lane) * 16;
            uint32_t m = 0;
            if (v.interior || c0 < v.sm_hi) m = newline_mask16(lds128(v.sa + (uint32_t)c0));
            if (!v.interior) {  // every '\n' inside the segment, the last byte included
                const int j_lo = v.seg_lo - c0 > 0 ? v.seg_lo - c0 : 0;
                const int j_hi = v.hi - c0 < 16 ? v.hi - c0 : 16;
                m = (j_hi > j_lo) ? (m & ((1u << j_hi) - 1u) & ~((1u << j_lo) - 1u)) : 0u;
            }
            const uint32_t n = (uint32_t)__popc(m);
            const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, n >= 1u);
            if (b1 == 0u) continue;
            const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, n >= 2u);
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
            unsigned long long e = next_end + before;
            uint32_t mm = m;
            while (mm) {
                const int p = c0 + __ffs(mm) - 1;
                mm &= mm - 1;
                a.line_end[e] = v.g + p;
                if (p + 1 < v.hi) a.line_start[e + 1ull] = v.g + p + 1;
                ++e;
            }
            next_end += total;
        }
        // a segment whose last byte is not '\n': its last line ends where the segment ends
        if (v.hi <= FqRing::TILE && v.hi > v.seg_lo && lane == 0 && view_byte(v, v.hi - 1) != '\n') a.line_end[next_end] = v.g + v.hi;
        ring.release(T);
    }
}

struct FqFileTab {
    long long rec0;    // first record of the file (global numbering)
    long long line0;   // first line of the file (global numbering)
    long long lines;   // lines in the file
    long long batch0;  // first batch of the file
};

struct FqColArgs {
    const FqFileTab *files;  // n_files + 1 entries (the last one carries rec0 = n_records)
    int32_t n_files;
    int64_t n_records;
    int32_t batch_rows;
    const uint8_t *const *line_start;
    const uint8_t *const *line_end;
    int32_t *lens[4];          // pass 3 out (n_records + 1 each, the last entry 0); NULL for unprojected columns
    const long long *voff[4];  // pass 5 in
    uint8_t *values[4];
    int32_t *offsets[4];       // n_batches * (batch_rows + 1)
    uint32_t *desc_valid;      // n_batches * words_per_batch, zeroed
    int32_t words_per_batch;
    uint32_t *flags;
};

__device__ __forceinline__ int fq_find_file(const FqFileTab *files, int n_files, long long r) {
    int lo = 0, hi = n_files - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (files[mid].rec0 <= r) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// the four fields of record r: pointers and lengths (name, description, sequence, quality)
__device__ __forceinline__ uint32_t fq_record_fields(const FqColArgs &a, long long r, int f, const uint8_t *ptr[4], int32_t len[4]) {
    const long long local = r - a.files[f].rec0;
    const long long l0 = a.files[f].line0 + 4 * local;
    const int avail = (int)((a.files[f].lines - 4 * local) < 4 ? (a.files[f].lines - 4 * local) : 4);
    uint32_t err = 0;
    const uint8_t *s = a.line_start[l0], *e = a.line_end[l0];
    if (e - s < 1 || *s != '@') err |= kFqErrPrefix;
    const uint8_t *sp = s + 1;
    while (sp < e && *sp != ' ') ++sp;
    ptr[0] = s + 1;
    len[0] = (int32_t)((sp < e ? sp : e) - (s + 1));
    if (len[0] < 0) len[0] = 0;
    ptr[1] = sp < e ? sp + 1 : e;
    len[1] = sp < e ? (int32_t)(e - (sp + 1)) : 0;
    ptr[2] = ptr[3] = e;
    len[2] = len[3] = 0;
    if (avail >= 2) {
        ptr[2] = a.line_start[l0 + 1];
        len[2] = (int32_t)(a.line_end[l0 + 1] - ptr[2]);
    }
    if (avail >= 3) {
        const uint8_t *p = a.line_start[l0 + 2];
        if (a.line_end[l0 + 2] - p < 1 || *p != '+') err |= kFqErrPrefix;
    } else {
        err |= kFqErrTruncated;
    }
    if (avail >= 4) {
        ptr[3] = a.line_start[l0 + 3];
        len[3] = (int32_t)(a.line_end[l0 + 3] - ptr[3]);
    }
    return err;
}

__global__ void __launch_bounds__(256) fq_fields_kernel(const __grid_constant__ FqColArgs a) {
    const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
    if (r >= a.n_records) return;
    const int f = fq_find_file(a.files, a.n_files, r);
    const uint8_t *ptr[4];
    int32_t len[4];
    const uint32_t err = fq_record_fields(a, r, f, ptr, len);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (a.lens[k]) a.lens[k][r] = len[k];
    if (err) atomicOr(a.flags, err);
}

__global__ void __launch_bounds__(256) fq_gather_kernel(const __grid_constant__ FqColArgs a) {
    const long long r = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n_records) return;
    const int f = fq_find_file(a.files, a.n_files, r);
    const uint8_t *ptr[4];
    int32_t len[4];
    fq_record_fields(a, r, f, ptr, len);
    const long long local = r - a.files[f].rec0;
    const long long b = a.files[f].batch0 + local / a.batch_rows;
    const int in_batch = (int)(local % a.batch_rows);
    const long long first = r - in_batch;
    const bool last_of_batch = in_batch + 1 == a.batch_rows || r + 1 == a.files[f + 1].rec0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!a.values[k]) continue;
        const long long v = a.voff[k][r];
        uint8_t *dst = a.values[k] + v;
        for (int j = lane; j < len[k]; j += 32) dst[j] = ptr[k][j];
        if (lane == 0) {
            const long long v0 = a.voff[k][first];
            int32_t *o = a.offsets[k] + b * (a.batch_rows + 1);
            o[in_batch] = (int32_t)(v - v0);
            if (last_of_batch) o[in_batch + 1] = (int32_t)(v + len[k] - v0);
        }
    }
    if (lane == 0 && a.desc_valid && len[1] > 0) atomicOr(a.desc_valid + b * a.words_per_batch + (in_batch >> 5), 1u << (in_batch & 31));
}

__global__ void fq_gather_i64(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

__global__ void fq_gather_u64(const unsigned long long *src, const long long *idx, int n, unsigned long long *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

}  // namespace

void fq_columns_free(VcfStream *s) {
    if (s->fq_cols) {
        s->fq_cols->unref();
        s->fq_cols = nullptr;
    }
}

namespace {

struct FqBatchPriv {
    FqColumns *cols;
    int n_children;
    ArrowArray children[4];
    ArrowArray *child_ptrs[4];
    const void *child_buffers[4][3];
    const void *struct_buffers[1];
};
void fq_release_child(ArrowArray *a) { a->release = nullptr; }
void fq_release_batch(ArrowArray *a) {
    auto *p = static_cast<FqBatchPriv *>(a->private_data);
    for (int i = 0; i < p->n_children; ++i)
        if (p->children[i].release) p->children[i].release(&p->children[i]);
    p->cols->unref();
    delete p;
    a->release = nullptr;
}
struct FqSchemaPriv {
    int n_children;
    ArrowSchema children[4];
    ArrowSchema *child_ptrs[4];
};
void fq_release_schema_child(ArrowSchema *s) { s->release = nullptr; }
void fq_release_schema(ArrowSchema *s) {
    auto *p = static_cast<FqSchemaPriv *>(s->private_data);
    for (int i = 0; i < p->n_children; ++i)
        if (p->children[i].release) p->children[i].release(&p->children[i]);
    delete p;
    s->release = nullptr;
}
// exon/exon-fastq/src/config.rs:79-88: name !null, description nullable, sequence !null, quality_scores !null, all Utf8
void fq_fill_schema(const std::vector<int> &projection, ArrowSchema *out, bool fasta) {
    // FASTA (exon/exon-fasta/src/config.rs:162-226): id !null, description nullable, sequence !null
    static const char *fq_names[4] = {"name", "description", "sequence", "quality_scores"};
    static const char *fa_names[4] = {"id", "description", "sequence", ""};
    const char *const *names = fasta ? fa_names : fq_names;
    auto *p = new FqSchemaPriv();
    p->n_children = (int)projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        ArrowSchema &c = p->children[i];
        memset(&c, 0, sizeof(c));
        c.format = "u";
        c.name = names[projection[(size_t)i]];
        c.flags = projection[(size_t)i] == 1 ? ARROW_FLAG_NULLABLE : 0;
        c.release = fq_release_schema_child;
        p->child_ptrs[i] = &c;
    }
    memset(out, 0, sizeof(*out));
    out->format = "+s";
    out->name = "";
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = fq_release_schema;
    out->private_data = p;
}

int fq_build_columns(VcfStream *s) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto *c = new (std::nothrow) FqColumns();
    if (!c) return fail(EXON_GPU_ERR_OOM, "fastq_next_batch: out of host memory");
    s->fq_cols = c;
    c->device = ctx->device;
    c->on_device = s->columns_on_device;
    c->batch_rows = s->batch_rows;
    c->words_per_batch = (s->batch_rows + 31) / 32;
    c->projection = s->projection;
    bool want[4] = {false, false, false, false};
    for (int p : s->projection) want[p] = true;
    c->batch_row0.assign(1, 0);

    std::vector<Piece> pieces;
    s->cut_pieces(pieces);
    if (pieces.empty()) return EXON_GPU_OK;
    std::vector<ScanSeg> h_segs;
    std::vector<int32_t> file_first, file_next;
    std::vector<long long> file_tiles;
    int64_t n_tiles = 0;
    for (const Piece &p : pieces) {
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)p.base & 15);
        sg.base = p.base - sg.skip;
        sg.len = p.len;
        sg.tile0 = n_tiles;
        sg.pad_ = 0;
        const bool starts = p.starts_file || h_segs.empty();
        if (starts) file_tiles.push_back(n_tiles);
        n_tiles += (sg.skip + p.len + FqRing::TILE - 1) / FqRing::TILE;
        file_first.push_back(starts ? (int32_t)h_segs.size() : file_first.back());
        h_segs.push_back(sg);
    }
    const int n_segs = (int)h_segs.size();
    file_next.assign((size_t)n_segs, n_segs);
    for (int i = n_segs - 2; i >= 0; --i) file_next[(size_t)i] = file_first[(size_t)i + 1] != file_first[(size_t)i] ? i + 1 : file_next[(size_t)i + 1];
    ScanSeg sentinel;
    memset(&sentinel, 0, sizeof(sentinel));
    sentinel.tile0 = n_tiles;
    h_segs.push_back(sentinel);
    file_tiles.push_back(n_tiles);
    const int n_files = (int)file_tiles.size() - 1;

    size_t cub_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)(n_tiles + 1), st));
    const size_t o_segs = 0, o_lines = o_segs + al256(h_segs.size() * sizeof(ScanSeg)), o_prefix = o_lines + al256((size_t)(n_tiles + 1) * 8),
                 o_cub = o_prefix + al256((size_t)(n_tiles + 1) * 8), o_ff = o_cub + al256(cub_bytes), o_fn = o_ff + al256((size_t)n_segs * 4),
                 o_l0 = o_fn + al256((size_t)n_segs * 4), o_ft = o_l0 + al256((size_t)n_segs * 8), o_fp = o_ft + al256(file_tiles.size() * 8),
                 o_ftab = o_fp + al256(file_tiles.size() * 8), o_out = o_ftab + al256(((size_t)n_files + 1) * sizeof(FqFileTab));
    if (int rc = ctx->ensure_scratch(o_out + 256, file_tiles.size() * 8 + 64)) return rc;
    uint8_t *scr = (uint8_t *)ctx->scratch;
    CUDA_TRY(cudaMemcpyAsync(scr + o_segs, h_segs.data(), h_segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_ff, file_first.data(), (size_t)n_segs * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_fn, file_next.data(), (size_t)n_segs * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_ft, file_tiles.data(), file_tiles.size() * 8, cudaMemcpyHostToDevice, st));
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
    a.out = (unsigned long long *)(scr + o_out);
    a.flags = (uint32_t *)(scr + o_out + 16);
    static int occ_a = 0, occ_i = 0;
    CUDA_TRY(fq_launch(fq_lines_kernel, a, ctx->sm_count, st, &occ_a));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(scr + o_cub, cub_bytes, a.tile_lines, (unsigned long long *)(scr + o_prefix), (int)(n_tiles + 1), st));
    fq_segment_table<<<(n_segs + 127) / 128, 128, 0, st>>>(a.segs, n_segs, a.tile_prefix, (const int32_t *)(scr + o_ff),
                                                           (const int32_t *)(scr + o_fn), (unsigned long long *)(scr + o_l0), a.flags);
    fq_gather_u64<<<(unsigned)((file_tiles.size() + 127) / 128), 128, 0, st>>>(a.tile_prefix, (const long long *)(scr + o_ft), (int)file_tiles.size(),
                                                                               (unsigned long long *)(scr + o_fp));
    ctx->launches.fetch_add(4);
    CUDA_TRY(cudaGetLastError());
    unsigned long long *h_fp = (unsigned long long *)ctx->h_scratch;
    CUDA_TRY(cudaMemcpyAsync(h_fp, scr + o_fp, file_tiles.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(s->h_res, a.out, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((uint32_t)s->h_res[2] & kFqErrTruncated) return fail(EXON_GPU_ERR_PARSE, "malformed FASTQ record: unexpected end of file inside a record;");
    // files -> records -> batches (batches restart at every file; a file with 4k + 3 lines ends in a record without quality)
    std::vector<FqFileTab> ftab((size_t)n_files + 1);
    long long n_records = 0;
    c->batch_row0.clear();
    for (int f = 0; f < n_files; ++f) {
        const long long lines = (long long)(h_fp[f + 1] - h_fp[f]);
        const long long recs = (lines + 3) / 4;
        ftab[(size_t)f] = FqFileTab{n_records, (long long)h_fp[f], lines, (long long)c->batch_row0.size()};
        for (long long r = 0; r < recs; r += c->batch_rows) c->batch_row0.push_back(n_records + r);
        n_records += recs;
    }
    ftab[(size_t)n_files] = FqFileTab{n_records, (long long)h_fp[n_files], 0, (long long)c->batch_row0.size()};
    const long long n_lines = (long long)h_fp[n_files];
    c->n_rows = n_records;
    c->n_batches = (int64_t)c->batch_row0.size();
    c->batch_row0.push_back(n_records);
    if (n_records == 0) return EXON_GPU_OK;
    CUDA_TRY(cudaMemcpyAsync(scr + o_ftab, ftab.data(), ftab.size() * sizeof(FqFileTab), cudaMemcpyHostToDevice, st));

    // scratch B: line tables | lens x4 | voff x4 | cub temp | batch tables
    size_t cub2 = 0;
    CUDA_TRY(exclusive_sum_i32_i64(nullptr, cub2, (const int32_t *)nullptr, (long long *)nullptr, (int)(n_records + 1), st));
    const size_t nb1 = (size_t)c->n_batches + 1;
    size_t ob = 0;
    const size_t o_ls = ob; ob += al256((size_t)(n_lines + 1) * 8);
    const size_t o_le = ob; ob += al256((size_t)(n_lines + 1) * 8);
    size_t o_len[4], o_voff[4];
    for (int k = 0; k < 4; ++k) { o_len[k] = ob; ob += al256((size_t)(n_records + 1) * 4); }
    for (int k = 0; k < 4; ++k) { o_voff[k] = ob; ob += al256((size_t)(n_records + 1) * 8); }
    const size_t o_cub2 = ob; ob += al256(cub2);
    const size_t o_brow = ob; ob += al256(nb1 * 8);
    const size_t o_bv0 = ob; ob += al256(nb1 * 8);
    if (int rc = ctx->ensure_scratch_b(ob + 256)) return rc;
    uint8_t *scb = (uint8_t *)ctx->scratch_b;
    FqIndexArgs ia;
    ia.segs = a.segs;
    ia.n_tiles = n_tiles;
    ia.tile_prefix = a.tile_prefix;
    ia.line_start = (const uint8_t **)(scb + o_ls);
    ia.line_end = (const uint8_t **)(scb + o_le);
    {
        constexpr size_t smem = FqRing::smem_bytes;
        if (!occ_i) {
            CUDA_TRY(cudaFuncSetAttribute(fq_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, fq_index_kernel, FqRing::WARPS * 32, smem));
            if (occ_i < 1) occ_i = 1;
        }
        int64_t grid = std::min<int64_t>((int64_t)occ_i * ctx->sm_count, (n_tiles + FqRing::WARPS - 1) / FqRing::WARPS);
        if (grid < 1) grid = 1;
        fq_index_kernel<<<(unsigned)grid, FqRing::WARPS * 32, smem, st>>>(ia);
        CUDA_TRY(cudaGetLastError());
    }
    FqColArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.files = (const FqFileTab *)(scr + o_ftab);
    ca.n_files = n_files;
    ca.n_records = n_records;
    ca.batch_rows = c->batch_rows;
    ca.line_start = ia.line_start;
    ca.line_end = ia.line_end;
    ca.words_per_batch = c->words_per_batch;
    ca.flags = a.flags;
    for (int k = 0; k < 4; ++k) {
        if (!want[k]) continue;
        ca.lens[k] = (int32_t *)(scb + o_len[k]);
        ca.voff[k] = (const long long *)(scb + o_voff[k]);
        CUDA_TRY(cudaMemsetAsync(scb + o_len[k] + (size_t)n_records * 4, 0, 4, st));
    }
    const unsigned rec_grid = (unsigned)((n_records + 255) / 256);
    fq_fields_kernel<<<rec_grid, 256, 0, st>>>(ca);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(scb + o_brow, c->batch_row0.data(), nb1 * 8, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < 4; ++k) {
        if (!want[k]) continue;
        size_t tb = cub2;
        CUDA_TRY(exclusive_sum_i32_i64(scb + o_cub2, tb, (const int32_t *)(scb + o_len[k]), (long long *)(scb + o_voff[k]), (int)(n_records + 1), st));
        fq_gather_i64<<<(unsigned)((nb1 + 255) / 256), 256, 0, st>>>((const long long *)(scb + o_voff[k]), (const long long *)(scb + o_brow), (int64_t)nb1,
                                                                    (long long *)(scb + o_bv0));
        c->batch_v0[k].resize(nb1);
        CUDA_TRY(cudaMemcpyAsync(c->batch_v0[k].data(), scb + o_bv0, nb1 * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));  // o_bv0 is reused by the next column
    }
    CUDA_TRY(cudaMemcpyAsync(s->h_res, a.out, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((uint32_t)s->h_res[2])
        return fail(EXON_GPU_ERR_PARSE, "malformed FASTQ record:%s%s", ((uint32_t)s->h_res[2] & kFqErrPrefix) ? " invalid name prefix or missing '+' line;" : "",
                    ((uint32_t)s->h_res[2] & kFqErrTruncated) ? " unexpected end of file inside a record;" : "");
    size_t total[4] = {0, 0, 0, 0};
    const size_t off_elems = (size_t)c->n_batches * (size_t)(c->batch_rows + 1);
    for (int k = 0; k < 4; ++k) {
        if (!want[k]) continue;
        total[k] = (size_t)c->batch_v0[k][(size_t)c->n_batches];
        for (int64_t b = 0; b < c->n_batches; ++b)
            if (c->batch_v0[k][(size_t)b + 1] - c->batch_v0[k][(size_t)b] > 0x7FFFFFFFll)
                return fail(EXON_GPU_ERR_UNSUPPORTED, "fastq: the bytes of batch %lld overflow int32 offsets", (long long)b);
        CUDA_TRY(cudaMallocAsync((void **)&c->d_values[k], std::max<size_t>(total[k], 1), st));
        CUDA_TRY(cudaMallocAsync((void **)&c->d_offsets[k], off_elems * 4, st));
        ca.values[k] = c->d_values[k];
        ca.offsets[k] = c->d_offsets[k];
    }
    const size_t valid_bytes = (size_t)c->n_batches * (size_t)c->words_per_batch * 4;
    if (want[1]) {
        CUDA_TRY(cudaMallocAsync((void **)&c->d_valid, valid_bytes, st));
        CUDA_TRY(cudaMemsetAsync(c->d_valid, 0, valid_bytes, st));
        ca.desc_valid = c->d_valid;
    }
    fq_gather_kernel<<<(unsigned)((n_records * 32 + 255) / 256), 256, 0, st>>>(ca);
    ctx->launches.fetch_add(4);
    CUDA_TRY(cudaGetLastError());
    if (!c->on_device) {
        for (int k = 0; k < 4; ++k) {
            if (!want[k]) continue;
            CUDA_TRY(cudaHostAlloc((void **)&c->h_values[k], std::max<size_t>(total[k], 1), cudaHostAllocDefault));
            CUDA_TRY(cudaHostAlloc((void **)&c->h_offsets[k], off_elems * 4, cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(c->h_values[k], c->d_values[k], total[k], cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(c->h_offsets[k], c->d_offsets[k], off_elems * 4, cudaMemcpyDeviceToHost, st));
        }
        if (want[1]) {
            CUDA_TRY(cudaHostAlloc((void **)&c->h_valid, valid_bytes, cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(c->h_valid, c->d_valid, valid_bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return EXON_GPU_OK;
}

}  // namespace

// Line index of a whole resident partition (steps 1 and 2 above, shared with the wide VCF column build in vcf_wide.cu):
// line_start[i] / line_end[i] for every line in feed order (line_end points at the '\n', or at the end of a file whose last
// line has none), the first line of every file, and `extra_per_line * (n_lines + 1) + extra_fixed` bytes of scratch_b
// behind the two tables for the caller's per-row temporaries.  The caller holds ctx->work_mu.
int build_line_index(VcfStream *s, size_t extra_per_line, size_t extra_fixed, LineIndex *out) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    out->n_lines = 0;
    out->file_line0.assign(1, 0);
    std::vector<Piece> pieces;
    s->cut_pieces(pieces);
    if (pieces.empty()) return EXON_GPU_OK;
    std::vector<ScanSeg> h_segs;
    std::vector<long long> file_tiles;
    int64_t n_tiles = 0;
    for (const Piece &p : pieces) {
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)p.base & 15);
        sg.base = p.base - sg.skip;
        sg.len = p.len;
        sg.tile0 = n_tiles;
        sg.pad_ = 0;
        if (p.starts_file || h_segs.empty()) file_tiles.push_back(n_tiles);
        n_tiles += (sg.skip + p.len + FqRing::TILE - 1) / FqRing::TILE;
        h_segs.push_back(sg);
    }
    ScanSeg sentinel;
    memset(&sentinel, 0, sizeof(sentinel));
    sentinel.tile0 = n_tiles;
    h_segs.push_back(sentinel);
    file_tiles.push_back(n_tiles);
    size_t cub_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)(n_tiles + 1), st));
    const size_t o_segs = 0, o_lines = o_segs + al256(h_segs.size() * sizeof(ScanSeg)), o_prefix = o_lines + al256((size_t)(n_tiles + 1) * 8),
                 o_cub = o_prefix + al256((size_t)(n_tiles + 1) * 8), o_ft = o_cub + al256(cub_bytes), o_fp = o_ft + al256(file_tiles.size() * 8),
                 o_end = o_fp + al256(file_tiles.size() * 8);
    if (int rc = ctx->ensure_scratch(o_end + 256, file_tiles.size() * 8 + 64)) return rc;
    uint8_t *scr = (uint8_t *)ctx->scratch;
    CUDA_TRY(cudaMemcpyAsync(scr + o_segs, h_segs.data(), h_segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_ft, file_tiles.data(), file_tiles.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(scr + o_lines + (size_t)n_tiles * 8, 0, 8, st));
    FqArgs a;
    memset(&a, 0, sizeof(a));
    a.segs = (const ScanSeg *)(scr + o_segs);
    a.n_segs = (int32_t)h_segs.size() - 1;
    a.n_tiles = n_tiles;
    a.tile_lines = (unsigned long long *)(scr + o_lines);
    a.tile_prefix = (const unsigned long long *)(scr + o_prefix);
    static int occ_a = 0, occ_i = 0;
    CUDA_TRY(fq_launch(fq_lines_kernel, a, ctx->sm_count, st, &occ_a));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(scr + o_cub, cub_bytes, a.tile_lines, (unsigned long long *)(scr + o_prefix), (int)(n_tiles + 1), st));
    fq_gather_u64<<<(unsigned)((file_tiles.size() + 127) / 128), 128, 0, st>>>(a.tile_prefix, (const long long *)(scr + o_ft), (int)file_tiles.size(),
                                                                               (unsigned long long *)(scr + o_fp));
    ctx->launches.fetch_add(3);
    CUDA_TRY(cudaGetLastError());
    unsigned long long *h_fp = (unsigned long long *)ctx->h_scratch;
    CUDA_TRY(cudaMemcpyAsync(h_fp, scr + o_fp, file_tiles.size() * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    out->file_line0.clear();
    for (size_t f = 0; f < file_tiles.size(); ++f) out->file_line0.push_back((long long)h_fp[f]);
    const long long n_lines = out->file_line0.back();
    out->n_lines = n_lines;
    if (n_lines == 0) return EXON_GPU_OK;
    const size_t tab = al256((size_t)(n_lines + 1) * 8);
    if (int rc = ctx->ensure_scratch_b(2 * tab + extra_per_line * (size_t)(n_lines + 1) + extra_fixed + 256)) return rc;
    uint8_t *scb = (uint8_t *)ctx->scratch_b;
    FqIndexArgs ia;
    ia.segs = a.segs;
    ia.n_tiles = n_tiles;
    ia.tile_prefix = a.tile_prefix;
    ia.line_start = (const uint8_t **)scb;
    ia.line_end = (const uint8_t **)(scb + tab);
    constexpr size_t smem = FqRing::smem_bytes;
    if (!occ_i) {
        CUDA_TRY(cudaFuncSetAttribute(fq_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, fq_index_kernel, FqRing::WARPS * 32, smem));
        if (occ_i < 1) occ_i = 1;
    }
    int64_t grid = std::min<int64_t>((int64_t)occ_i * ctx->sm_count, (n_tiles + FqRing::WARPS - 1) / FqRing::WARPS);
    if (grid < 1) grid = 1;
    fq_index_kernel<<<(unsigned)grid, FqRing::WARPS * 32, smem, st>>>(ia);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    out->line_start = ia.line_start;
    out->line_end = ia.line_end;
    out->extra = scb + 2 * tab;
    return EXON_GPU_OK;
}

void fastq_stream_schema(VcfStream *s, ArrowSchema *out) { fq_fill_schema(s->projection, out, s->fmt == kFmtFasta); }

int fastq_next_batch(VcfStream *s, ArrowArray *out, ArrowSchema *out_schema) {
    if (!s->fq_cols) {
        if (int rc = s->flush_gz()) return rc;
        std::lock_guard<std::recursive_mutex> work(s->ctx->work_mu);
        if (int rc = s->fmt == kFmtFasta ? fasta_build_columns(s) : fq_build_columns(s)) {
            fq_columns_free(s);
            return rc;
        }
        s->drained = true;
    }
    FqColumns *c = s->fq_cols;
    if (out_schema) fq_fill_schema(s->projection, out_schema, s->fmt == kFmtFasta);
    memset(out, 0, sizeof(*out));
    if (c->next >= c->n_batches) return EXON_GPU_OK;  // end of stream: release == NULL
    const int64_t b = c->next++;
    const int64_t rows = c->batch_row0[(size_t)b + 1] - c->batch_row0[(size_t)b];
    auto *p = new FqBatchPriv();
    p->cols = c;
    c->refs.fetch_add(1);
    p->n_children = (int)s->projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        const int k = s->projection[(size_t)i];
        ArrowArray &a = p->children[i];
        memset(&a, 0, sizeof(a));
        a.length = rows;
        a.null_count = k == 1 ? -1 : 0;  // description: not counted (the bitmap is authoritative)
        a.n_buffers = 3;
        p->child_buffers[i][0] = k == 1 ? (const void *)((c->on_device ? c->d_valid : c->h_valid) + b * c->words_per_batch) : nullptr;
        p->child_buffers[i][1] = (c->on_device ? c->d_offsets[k] : c->h_offsets[k]) + b * (c->batch_rows + 1);
        p->child_buffers[i][2] = (c->on_device ? c->d_values[k] : c->h_values[k]) + c->batch_v0[k][(size_t)b];
        a.buffers = p->child_buffers[i];
        a.release = fq_release_child;
        p->child_ptrs[i] = &a;
    }
    p->struct_buffers[0] = nullptr;
    out->length = rows;
    out->n_buffers = 1;
    out->buffers = p->struct_buffers;
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = fq_release_batch;
    out->private_data = p;
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_fastq_open(exon_gpu_ctx *c, const exon_gpu_fastq_opts *o, exon_gpu_stream **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "fastq_open: NULL argument");
    *out = nullptr;
    if (o && (o->n_projection < 0 || o->n_projection > 4 || (o->n_projection > 0 && !o->projection)))
        return fail(EXON_GPU_ERR_ARG, "fastq_open: bad projection");
    if (o)
        for (int i = 0; i < o->n_projection; ++i)
            if (o->projection[i] < 0 || o->projection[i] > 3) return fail(EXON_GPU_ERR_ARG, "fastq_open: projection index %d is not a FASTQ column", o->projection[i]);
    exon_gpu_vcf_opts vo;
    memset(&vo, 0, sizeof(vo));
    vo.batch_rows = o ? o->batch_rows : 0;
    vo.columns_on_device = o ? o->columns_on_device : 0;
    if (int rc = exon_gpu_vcf_open(c, &vo, out)) return rc;
    if (o)
        for (int i = 0; i < o->n_projection; ++i) (*out)->projection.push_back(o->projection[i]);
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

int exon_gpu_fastq_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema) {
    if (!s || !out) return fail(EXON_GPU_ERR_ARG, "fastq_next_batch: NULL argument");
    if (s->fmt != kFmtFastq) return fail(EXON_GPU_ERR_ARG, "fastq_next_batch: not a FASTQ stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return fastq_next_batch(s, out, out_schema);
}

int exon_gpu_stream_close(exon_gpu_stream *s) { return exon_gpu_vcf_close(s); }
int exon_gpu_stream_reset(exon_gpu_stream *s) { return exon_gpu_vcf_reset(s); }
int exon_gpu_stream_body_bytes(exon_gpu_stream *s, int64_t *out_bytes) { return exon_gpu_vcf_body_bytes(s, out_bytes); }

}  // extern "C"
