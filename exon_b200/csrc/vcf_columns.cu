// vcf_columns.cu -- K2: VCF text -> Arrow columns {chrom: utf8, pos: int64} in reference-sized batches.
//
// Replaces AsyncBatchStream::read_batch (exon/exon-vcf/src/async_batch_stream.rs:80-109),
// LazyVCFArrayBuilder::{append, finish} for the projected columns 0 and 1
// (exon/exon-vcf/src/array_builder/lazy_array_builder.rs:157-168, 451-484) and
// ExonArrayBuilder::try_into_record_batch (exon/exon-common/src/array_builder.rs:25-36).
//
// The reference builds one batch at a time, row by row.  Here the whole resident partition is converted by two
// passes of the same warp-private TMA tile pipeline K1 uses (vcf_tile.cuh), so the text is read from HBM twice
// and nothing per-row is staged in between:
//   A. measure   per 4 KiB tile: rows that start in it and the CHROM bytes of those rows   (16 B per tile out)
//      scan      exclusive prefix over tiles (cub) -> first row index / first value offset of every tile;
//                the prefix at the first tile of every file gives the file's first row (batches restart there)
//   B. emit      every warp re-parses its tiles with the SWAR line parser and writes pos[row] (int64), the CHROM
//                bytes at their final place in the values buffer, and the absolute value offset of the row (u32)
//   C. offsets   per batch: absolute offsets -> int32 offsets that restart at 0 in every batch (the layout
//                arrow-rs' StringBuilder emits), batch_rows + 1 entries per batch
// A row belongs to the tile that holds the '\n' before it (the first row of a segment to the segment's first
// tile), exactly as in K1.  Segments are cut at file ends, so no batch spans two files (the reference opens one
// AsyncBatchStream per file).  Batches own nothing: they are views into the stream's column store, kept alive by
// a reference count that the Arrow release callbacks decrement.
#include <cub/device/device_scan.cuh>

#include <atomic>
#include <cstring>
#include <new>

#include "common.cuh"
#include "internal.h"
#include "vcf_tile.cuh"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

struct TileSum {
    unsigned long long rows, bytes;
};
struct TileSumAdd {
    __host__ __device__ __forceinline__ TileSum operator()(const TileSum &x, const TileSum &y) const {
        return TileSum{x.rows + y.rows, x.bytes + y.bytes};
    }
};

struct ColArgs {
    const ScanSeg *segs;  // n_segs + 1 entries (sentinel carries tile0 = n_tiles)
    int32_t n_segs;
    int64_t n_tiles;
    int32_t want_chrom, want_pos;
    TileSum *tile_stats;         // pass A out, one per tile
    const TileSum *tile_prefix;  // pass B in: exclusive prefix of tile_stats
    int64_t *pos;                // pass B out
    uint32_t *off32;             // pass B out: low 32 bits of the row's absolute offset into `values`
    uint8_t *values;             // pass B out
    uint32_t *flags;
    unsigned long long *first_bad_row;
};

// Byte-exact field reader: CHROM = bytes up to the first '\t', non-empty; POS = Rust `usize::from_str` (optional
// '+', >= 1 digit) terminated by '\t', non-zero (noodles maps 0 to None and the column is non-nullable,
// lazy_array_builder.rs:163-168), <= i64::MAX.  Returns clen | err << 32; *out_pos = 0 on error.
__device__ __noinline__ unsigned long long line_fields_exact(const uint8_t *sm, const uint8_t *g, int lo, int hi, int sm_lo,
                                                             int sm_hi, int ls, int want_pos, long long *out_pos) {
    const TileView t{sm, g, lo, hi, sm_lo, sm_hi};
    uint32_t err = 0;
    int q = ls;
    uint32_t c;
    while ((c = ld_byte(t, q)) != '\t' && c != '\n') ++q;
    if (c != '\t' || q == ls) err |= kErrShortLine;
    const uint32_t clen = err ? 0u : (uint32_t)(q - ls);
    long long out = 0;
    if (want_pos && !err) {
        c = ld_byte(t, ++q);
        if (c == '+') c = ld_byte(t, ++q);
        unsigned long long v = 0;
        int sig = 0, nd = 0;
        bool ovf = false;
        while (c - '0' <= 9u) {
            const uint32_t dg = c - '0';
            if (v | dg) ++sig;
            if (sig > 19) ovf = true;
            v = v * 10ull + dg;
            ++nd;
            c = ld_byte(t, ++q);
        }
        if (nd == 0 || c != '\t') err |= (c == '\n') ? kErrShortLine : kErrBadPos;
        else if (ovf || v == 0ull || v > 0x7FFFFFFFFFFFFFFFull) err |= kErrBadPos;
        else out = (long long)v;
    }
    *out_pos = out;
    return (unsigned long long)clen | ((unsigned long long)err << 32);
}

// SWAR field reader for interior tiles (same window technique as K1's line_swar).  Sets `slow` when the line
// needs the byte-exact routine (field outside the 16-byte window, '+', > 12 digits, POS 0, anything malformed).
template <bool EMIT>
__device__ __forceinline__ void line_fields_swar(uint32_t sa, int ls, int want_pos, uint32_t &clen, long long &pos, bool &slow) {
    const uint32_t la = sa + (uint32_t)ls;
    const uint32_t a0 = la & ~3u;
    const uint32_t sh = (la & 3u) << 3;
    const uint32_t w0 = lds32(a0), w1 = lds32(a0 + 4), w2 = lds32(a0 + 8), w3 = lds32(a0 + 12), w4 = lds32(a0 + 16);
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh),
                   v3 = __funnelshift_r(w3, w4, sh);
    const uint32_t m = pack16(zero_bytes_exact((v0 & 0xFCFCFCFCu) ^ 0x08080808u), zero_bytes_exact((v1 & 0xFCFCFCFCu) ^ 0x08080808u),
                              zero_bytes_exact((v2 & 0xFCFCFCFCu) ^ 0x08080808u), zero_bytes_exact((v3 & 0xFCFCFCFCu) ^ 0x08080808u));
    const int s1 = __ffs(m) - 1;
    if (s1 < 1 || lds8(la + s1) != '\t') {
        slow = true;
        return;
    }
    clen = (uint32_t)s1;
    if (!EMIT || !want_pos) return;
    const uint32_t m2 = m & (m - 1);
    const int s2 = __ffs(m2) - 1;
    const int n = s2 - s1 - 1;
    if (m2 == 0 || n < 1 || n > 12 || lds8(la + s2) != '\t') {
        slow = true;
        return;
    }
    const uint32_t b = la + (uint32_t)s2 - 12u;
    const uint32_t b0 = b & ~3u;
    const uint32_t sh2 = (b & 3u) << 3;
    const uint32_t x0 = lds32(b0), x1 = lds32(b0 + 4), x2 = lds32(b0 + 8), x3 = lds32(b0 + 12);
    const uint32_t s = (uint32_t)(12 - n) << 3;
    const uint32_t d0 = (__funnelshift_r(x0, x1, sh2) ^ 0x30303030u) & shl_clamp(0xFFFFFFFFu, s);
    const uint32_t d1 = (__funnelshift_r(x1, x2, sh2) ^ 0x30303030u) & shl_clamp(0xFFFFFFFFu, s > 32u ? s - 32u : 0u);
    const uint32_t d2 = (__funnelshift_r(x2, x3, sh2) ^ 0x30303030u) & shl_clamp(0xFFFFFFFFu, s > 64u ? s - 64u : 0u);
    const uint32_t bad = ((d0 + 0x76767676u) | d0 | (d1 + 0x76767676u) | d1 | (d2 + 0x76767676u) | d2) & 0x80808080u;
    if (bad || (d0 | d1 | d2) == 0u) {
        slow = true;
        return;
    }
    const uint32_t q0 = __dp4a(d0, 0x00010A64u, 0u) * 10u + (d0 >> 24);
    const uint32_t q1 = __dp4a(d1, 0x00010A64u, 0u) * 10u + (d1 >> 24);
    const uint32_t q2 = __dp4a(d2, 0x00010A64u, 0u) * 10u + (d2 >> 24);
    pos = (long long)((unsigned long long)(q0 * 10000u + q1) * 10000ull + q2);
}

template <bool EMIT, int U, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, ctas_per_sm<U, S, WARPS>()) vcf_cols_kernel(const __grid_constant__ ColArgs a) {
    using L = SmemLayout<U, S, WARPS>;
    constexpr int TILE = L::TILE, STAGE = L::STAGE;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t *ring = smem_raw + L::ring + (size_t)warp * (S * STAGE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + L::bars) + warp * S;
    StageMeta *meta = reinterpret_cast<StageMeta *>(smem_raw + L::meta) + warp * S;
    const uint32_t ring_sa = smem_u32(ring);
    const uint32_t queue_sa = smem_u32(smem_raw + L::queue) + (uint32_t)(warp * kQueue * sizeof(uint16_t));

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int64_t nw = (int64_t)gridDim.x * WARPS;
    const int64_t wg = (int64_t)blockIdx.x * WARPS + warp;

    // ---- producer (lane 0): one cursor over the segment table, S tiles ahead of the consumer ----
    int pc = 0;
    int64_t p_tile0 = 0, p_next0 = 0;
    if (lane == 0) {
        p_tile0 = __ldg(&a.segs[0].tile0);
        p_next0 = __ldg(&a.segs[1].tile0);
    }
    auto issue = [&](int64_t T, int s) {  // lane 0 only
        while (T >= p_next0) {
            ++pc;
            p_tile0 = p_next0;
            p_next0 = __ldg(&a.segs[pc + 1].tile0);
        }
        const uint8_t *base = a.segs[pc].base;
        const int skip = __ldg(&a.segs[pc].skip);
        const int64_t off = (T - p_tile0) * TILE;
        const int64_t rem = skip + __ldg(&a.segs[pc].len) - off;
        const int pre = off ? kPre : 0;
        const int64_t body = (rem + 15) & ~(int64_t)15;
        const uint32_t bytes = (uint32_t)(body < TILE + kHalo ? body : TILE + kHalo) + pre;
        meta[s].g = base + off;
        meta[s].lo = off ? -kPre : skip;
        meta[s].hi = rem > (1 << 30) ? (1 << 30) : (int)rem;
        mbar_arrive_expect_tx(&bars[s], bytes);
        bulk_g2s(ring + s * STAGE + (kPre - pre), base + off - pre, bytes, &bars[s]);
    };
    if (lane == 0) {
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
            const int64_t T = wg + s * nw;
            if (T < a.n_tiles) issue(T, s);
        }
    }
    __syncwarp();

    const bool need_fields = EMIT || a.want_chrom;  // pass A of a pos-only projection just counts lines
    uint32_t err = 0;
    unsigned long long bad_row = ~0ull;
    uint32_t parity = 0;
    int s = 0;
#pragma unroll 1
    for (int64_t T = wg; T < a.n_tiles; T += nw) {
        const uint8_t *sm = ring + s * STAGE + kPre;
        mbar_wait(&bars[s], parity);
        const uint8_t *g = meta[s].g;
        const int lo = meta[s].lo, hi = meta[s].hi;
        const bool first = lo >= 0;
        const int seg_lo = first ? lo : -(1 << 30);
        const int sm_lo = first ? 0 : -kPre;
        const int sm_hi = hi < TILE + kHalo ? ((hi + 15) & ~15) : TILE + kHalo;
        const bool interior = hi >= TILE + kHalo && (lo <= 0);
        const uint32_t sa = ring_sa + (uint32_t)(s * STAGE + kPre);

        unsigned long long row0 = 0, vb = 0;
        if (EMIT) {
            row0 = __ldg(&a.tile_prefix[T].rows);
            vb = __ldg(&a.tile_prefix[T].bytes);
        }
        uint32_t t_rows = 0;           // rows of this tile already drained (warp-uniform)
        unsigned long long t_bytes = 0;  // EMIT: CHROM bytes already placed (warp-uniform); pass A: this lane's partial sum
        uint32_t lane_rows = 0;        // pass A without CHROM: this lane's line count
        int qn = 0;

        auto drain = [&]() {
            __syncwarp();
#pragma unroll 1
            for (int i0 = 0; i0 < qn; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < qn;
                uint32_t clen = 0, e = 0;
                long long pv = 0;
                int ls = 0;
                if (act) {
                    ls = (int)lds16(queue_sa + 2u * (uint32_t)i);
                    bool slow = !interior;
                    if (!slow) line_fields_swar<EMIT>(sa, ls, a.want_pos, clen, pv, slow);
                    if (slow) {
                        const unsigned long long r = line_fields_exact(sm, g, seg_lo, hi, sm_lo, sm_hi, ls, EMIT ? a.want_pos : 0, &pv);
                        clen = (uint32_t)r;
                        e = (uint32_t)(r >> 32);
                    }
                }
                if (EMIT) {
                    uint32_t incl = clen;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= d) incl += v;
                    }
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    if (act) {
                        const unsigned long long row = row0 + t_rows + (uint32_t)i;
                        if (a.want_pos) a.pos[row] = pv;
                        if (a.want_chrom) {
                            const unsigned long long v = vb + t_bytes + (incl - clen);
                            a.off32[row] = (uint32_t)v;
                            const TileView t{sm, g, seg_lo, hi, sm_lo, sm_hi};
                            for (uint32_t j = 0; j < clen; ++j) a.values[v + j] = (uint8_t)ld_byte(t, ls + (int)j);
                        }
                        if (e) {
                            err |= e;
                            if (row < bad_row) bad_row = row;
                        }
                    }
                    t_bytes += total;
                } else {
                    t_bytes += clen;
                }
            }
            __syncwarp();
            t_rows += (uint32_t)qn;
            qn = 0;
        };

        // the segment's first line has no '\n' before it
        if (first && hi > lo) {
            if (need_fields) {
                if (lane == 0) sts16(queue_sa, (uint32_t)lo);
                qn = 1;
            } else if (lane == 0) {
                lane_rows += 1;
            }
        }
#pragma unroll 2
        for (int u = 0; u < U; ++u) {
            const int c0 = (u * 32 + lane) * 16;
            uint32_t m = 0;
            if (interior || c0 < sm_hi) {
                const uint4 w = lds128(sa + (uint32_t)c0);
                m = pack16(zero_bytes_exact(w.x ^ kNL4), zero_bytes_exact(w.y ^ kNL4), zero_bytes_exact(w.z ^ kNL4),
                           zero_bytes_exact(w.w ^ kNL4));
            }
            if (!interior) {
                // a '\n' at tile index p starts a line iff p >= seg_lo and p + 1 < hi
                const int j_lo = seg_lo - c0 > 0 ? seg_lo - c0 : 0;
                const int j_hi = hi - 1 - c0 < 16 ? hi - 1 - c0 : 16;
                m = (j_hi > j_lo) ? (m & ((1u << j_hi) - 1u) & ~((1u << j_lo) - 1u)) : 0u;
            }
            const uint32_t cnt = (uint32_t)__popc(m);
            if (!need_fields) {
                lane_rows += cnt;
                continue;
            }
            const uint32_t b_any = __ballot_sync(0xFFFFFFFFu, m != 0);
            const uint32_t b_multi = __ballot_sync(0xFFFFFFFFu, cnt > 1);
            if (b_any == 0) continue;
            if (b_multi == 0) {
                if (m) sts16(queue_sa + 2u * (uint32_t)(qn + __popc(b_any & lt_mask)), (uint32_t)(c0 + __ffs(m)));
                qn += __popc(b_any);
            } else {
                // several line starts inside one 16-byte chunk (lines shorter than 16 bytes): rank them exactly
                drain();
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= d) incl += v;
                }
                const int total = (int)__shfl_sync(0xFFFFFFFFu, incl, 31);
#pragma unroll 1
                for (int r0 = 0; r0 < total; r0 += kQueue) {
                    uint32_t mm = m;
                    int idx = (int)(incl - cnt);
                    while (mm) {
                        if (idx >= r0 && idx < r0 + kQueue) sts16(queue_sa + 2u * (uint32_t)(idx - r0), (uint32_t)(c0 + __ffs(mm)));
                        mm &= mm - 1;
                        ++idx;
                    }
                    qn = total - r0 < kQueue ? total - r0 : kQueue;
                    drain();
                }
            }
            if (qn > kQueue - 32) drain();
        }
        if (qn) drain();

        if (!EMIT) {
            unsigned long long rows = (unsigned long long)t_rows + warp_sum(lane_rows);
            unsigned long long bytes = t_bytes;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) bytes += __shfl_xor_sync(0xFFFFFFFFu, bytes, d);
            if (lane == 0) a.tile_stats[T] = TileSum{rows, bytes};
        }
        __syncwarp();
        if (lane == 0) {
            const int64_t Tn = T + (int64_t)S * nw;
            if (Tn < a.n_tiles) issue(Tn, s);
        }
        if (++s == S) {
            s = 0;
            parity ^= 1;
        }
    }
    if (EMIT && err) {
        atomicOr(a.flags, err);
        atomicMin(a.first_bad_row, bad_row);
    }
}

constexpr int kColU = 8, kColS = 2, kColW = 8;  // 4 KiB tiles, the geometry K1 settled on

template <bool EMIT>
cudaError_t launch_cols(const ColArgs &args, int sm_count, cudaStream_t stream) {
    constexpr size_t smem = SmemLayout<kColU, kColS, kColW>::total;
    auto kern = vcf_cols_kernel<EMIT, kColU, kColS, kColW>;
    static int occ = 0;
    if (!occ) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kColW * 32, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
    }
    int64_t grid = (int64_t)occ * sm_count;
    const int64_t need = (args.n_tiles + kColW - 1) / kColW;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kColW * 32, smem, stream>>>(args);
    return cudaGetLastError();
}

// out[i] = prefix[idx[i]]
__global__ void gather_prefix(const TileSum *prefix, const long long *idx, int n, TileSum *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = prefix[idx[i]];
}

// Full 64-bit value offset of the first row of every batch (and the grand total at index n_batches): the tile
// that holds the row is found by binary search over the tile prefix; the row's u32 offset supplies the low bits.
__global__ void batch_value_offsets(const TileSum *prefix, int64_t n_tiles, const long long *batch_row0, int64_t n_batches,
                                    int64_t n_rows, unsigned long long total, const uint32_t *off32, long long *batch_v0) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_batches) return;
    const unsigned long long row = (unsigned long long)batch_row0[b];
    if ((int64_t)row >= n_rows) {
        batch_v0[b] = (long long)total;
        return;
    }
    int64_t lo = 0, hi = n_tiles;  // last tile t with prefix[t].rows <= row
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (prefix[mid].rows <= row) lo = mid;
        else hi = mid;
    }
    const unsigned long long base = prefix[lo].bytes;
    batch_v0[b] = (long long)(base + (uint32_t)(off32[row] - (uint32_t)base));
}

// pass C: offsets[b * (batch_rows + 1) + i] = value offset of row (batch_row0[b] + i) relative to the batch's first row
__global__ void __launch_bounds__(256) batch_offsets(const long long *batch_row0, const long long *batch_v0, int64_t n_rows,
                                                    unsigned long long total, int batch_rows, const uint32_t *off32,
                                                    int32_t *offsets) {
    const int64_t b = blockIdx.x;
    const long long r0 = batch_row0[b];
    const int n = (int)(batch_row0[b + 1] - r0);
    const uint32_t base = (uint32_t)batch_v0[b];
    int32_t *o = offsets + b * (int64_t)(batch_rows + 1);
    for (int i = threadIdx.x; i <= n; i += blockDim.x) {
        const long long r = r0 + i;
        const uint32_t v = r >= n_rows ? (uint32_t)total : off32[r];
        o[i] = (int32_t)(v - base);
    }
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
    FaBatchDesc *d_descs = nullptr;  // K3 descriptor table of every batch (+ 64 bytes of result slots), built on first use
    // host mirrors (pinned) when !on_device
    int64_t *h_pos = nullptr;
    int32_t *h_offsets = nullptr;
    uint8_t *h_values = nullptr;
    std::vector<long long> batch_row0;  // first row of each batch (+ n_rows at the end); batches never span files
    std::vector<long long> batch_v0;    // values offset of each batch's first row (+ total at the end)
    std::vector<int> projection;
    WideStore *wide = nullptr;  // columns 2..6 (vcf_wide.cu)

    void unref() {
        if (refs.fetch_sub(1) == 1) {
            cudaSetDevice(device);
            wide_free(wide);
            cudaFree(d_pos);
            cudaFree(d_offsets);
            cudaFree(d_values);
            cudaFree(d_descs);
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
constexpr int kMaxCols = 9;
struct BatchPriv {
    Columns *cols;
    int n_children;
    ArrowArray children[kMaxCols];
    ArrowArray *child_ptrs[kMaxCols];
    ChildPriv child_priv[kMaxCols];
    WideChildSlot wide_slot[kMaxCols];
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
    ArrowSchema children[kMaxCols];
    ArrowSchema *child_ptrs[kMaxCols];
    ArrowSchema items[kMaxCols];  // the "item" child of list columns
    ArrowSchema *item_ptrs[kMaxCols];
};
void release_schema_child(ArrowSchema *s) { s->release = nullptr; }
void release_schema(ArrowSchema *s) {
    auto *p = static_cast<SchemaPriv *>(s->private_data);
    for (int i = 0; i < p->n_children; ++i)
        if (p->children[i].release) p->children[i].release(&p->children[i]);
    delete p;
    s->release = nullptr;
}

// VCFSchemaBuilder (exon/exon-core/src/datasources/vcf/schema_builder.rs:85-129): chrom Utf8 !null, pos Int64 !null,
// id List<item: Utf8>, ref Utf8 !null, alt List<item: Utf8>, qual Float32, filter List<item: Utf8>
void fill_schema(const std::vector<int> &projection, ArrowSchema *out) {
    static const char *names[9] = {"chrom", "pos", "id", "ref", "alt", "qual", "filter", "info", "formats"};
    static const char *formats[9] = {"u", "l", "+l", "u", "+l", "f", "+l", "u", "u"};
    static const bool nullable[9] = {false, false, true, false, true, true, true, true, true};
    auto *p = new SchemaPriv();
    p->n_children = (int)projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        const int col = projection[(size_t)i];
        ArrowSchema &c = p->children[i];
        memset(&c, 0, sizeof(c));
        c.format = formats[col];
        c.name = names[col];
        c.flags = nullable[col] ? ARROW_FLAG_NULLABLE : 0;
        c.release = release_schema_child;
        if (formats[col][0] == '+') {
            ArrowSchema &it = p->items[i];
            memset(&it, 0, sizeof(it));
            it.format = "u";
            it.name = "item";
            it.flags = ARROW_FLAG_NULLABLE;
            it.release = release_schema_child;
            p->item_ptrs[i] = &it;
            c.n_children = 1;
            c.children = &p->item_ptrs[i];
        }
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

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

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

    std::vector<Piece> pieces;
    s->cut_pieces(pieces);
    if (pieces.empty()) return EXON_GPU_OK;  // zero rows
    if (!c->want_chrom && !c->want_pos && wide_wanted(c->projection)) {
        // only columns 2..6: the line index of the wide build also yields the batch table
        int64_t n = -1;
        if (int rc = wide_build(s, &c->batch_row0, &n, &c->wide)) return rc;
        c->n_rows = n;
        c->n_batches = (int64_t)c->batch_row0.size() - 1;
        return EXON_GPU_OK;
    }
    constexpr int kTile = 512 * kColU;
    std::vector<ScanSeg> h_segs;
    std::vector<long long> file_tiles;  // first tile of every piece that starts a file
    int64_t n_tiles = 0;
    for (const Piece &p : pieces) {
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)p.base & 15);
        sg.base = p.base - sg.skip;
        sg.len = p.len;
        sg.tile0 = n_tiles;
        sg.pad_ = 0;
        if (p.starts_file) file_tiles.push_back(n_tiles);
        n_tiles += (sg.skip + p.len + kTile - 1) / kTile;
        h_segs.push_back(sg);
    }
    ScanSeg sentinel;
    sentinel.base = nullptr;
    sentinel.len = 0;
    sentinel.tile0 = n_tiles;
    sentinel.skip = 0;
    sentinel.pad_ = 0;
    h_segs.push_back(sentinel);
    file_tiles.push_back(n_tiles);  // grand totals
    const int n_files_max = (int)file_tiles.size() - 1;

    // ---- scratch A (persists in the context): segs | tile stats | tile prefix | cub temp | gather in/out | misc ----
    size_t cub_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, cub_bytes, (TileSum *)nullptr, (TileSum *)nullptr, TileSumAdd(), TileSum{0, 0},
                                            (int)(n_tiles + 1), st));
    const size_t o_segs = 0;
    const size_t o_stats = o_segs + align256(h_segs.size() * sizeof(ScanSeg));
    const size_t o_prefix = o_stats + align256((size_t)(n_tiles + 1) * sizeof(TileSum));
    const size_t o_cub = o_prefix + align256((size_t)(n_tiles + 1) * sizeof(TileSum));
    const size_t o_gidx = o_cub + align256(cub_bytes);
    const size_t o_gout = o_gidx + align256(file_tiles.size() * sizeof(long long));
    const size_t o_misc = o_gout + align256(file_tiles.size() * sizeof(TileSum));
    const size_t scratch_bytes = o_misc + 256;
    if (int rc = ctx->ensure_scratch(scratch_bytes, file_tiles.size() * sizeof(TileSum) + 64)) return rc;
    uint8_t *scr = (uint8_t *)ctx->scratch;
    ScanSeg *d_segs = (ScanSeg *)(scr + o_segs);
    TileSum *d_stats = (TileSum *)(scr + o_stats), *d_prefix = (TileSum *)(scr + o_prefix);
    long long *d_gidx = (long long *)(scr + o_gidx);
    TileSum *d_gout = (TileSum *)(scr + o_gout);
    unsigned long long *d_misc = (unsigned long long *)(scr + o_misc);

    CUDA_TRY(cudaMemcpyAsync(d_segs, h_segs.data(), h_segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_gidx, file_tiles.data(), file_tiles.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_stats + n_tiles, 0, sizeof(TileSum), st));
    const unsigned long long init_misc[4] = {0ull, ~0ull, 0ull, 0ull};
    CUDA_TRY(cudaMemcpyAsync(d_misc, init_misc, sizeof(init_misc), cudaMemcpyHostToDevice, st));

    ColArgs a;
    memset(&a, 0, sizeof(a));
    a.segs = d_segs;
    a.n_segs = (int32_t)h_segs.size() - 1;
    a.n_tiles = n_tiles;
    a.want_chrom = c->want_chrom;
    a.want_pos = c->want_pos;
    a.tile_stats = d_stats;
    a.tile_prefix = d_prefix;
    a.flags = reinterpret_cast<uint32_t *>(d_misc);
    a.first_bad_row = d_misc + 1;

    // ---- pass A + scan ----
    CUDA_TRY(launch_cols<false>(a, ctx->sm_count, st));
    ctx->launches.fetch_add(1);
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(scr + o_cub, cub_bytes, d_stats, d_prefix, TileSumAdd(), TileSum{0, 0}, (int)(n_tiles + 1), st));
    ctx->launches.fetch_add(1);
    gather_prefix<<<(unsigned)((file_tiles.size() + 127) / 128), 128, 0, st>>>(d_prefix, d_gidx, (int)file_tiles.size(), d_gout);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    TileSum *h_g = (TileSum *)ctx->h_scratch;
    CUDA_TRY(cudaMemcpyAsync(h_g, d_gout, file_tiles.size() * sizeof(TileSum), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int64_t n_rows = (int64_t)h_g[n_files_max].rows;
    const unsigned long long total_values = h_g[n_files_max].bytes;
    c->n_rows = n_rows;
    if (n_rows == 0) return EXON_GPU_OK;

    // ---- file table -> batch table (batches restart at every file; an empty file adds none) ----
    std::vector<long long> file_row0;
    for (int f = 0; f < n_files_max; ++f) {
        const long long r = (long long)h_g[f].rows;
        if (file_row0.empty() || r > file_row0.back()) file_row0.push_back(r);
    }
    if (file_row0.empty() || file_row0.front() != 0) file_row0.insert(file_row0.begin(), 0);
    while (!file_row0.empty() && file_row0.back() >= n_rows) file_row0.pop_back();
    file_row0.push_back(n_rows);
    c->batch_row0.clear();
    for (size_t f = 0; f + 1 < file_row0.size(); ++f)
        for (long long r = file_row0[f]; r < file_row0[f + 1]; r += c->batch_rows) c->batch_row0.push_back(r);
    c->n_batches = (int64_t)c->batch_row0.size();
    c->batch_row0.push_back(n_rows);
    if (!c->want_chrom && !c->want_pos) return EXON_GPU_OK;  // empty projection: row counts only

    // ---- outputs + scratch B: absolute u32 offsets | batch_row0 | batch_v0 ----
    const size_t nb1 = (size_t)c->n_batches + 1;
    const size_t ob_off32 = 0;
    const size_t ob_brow = ob_off32 + align256(c->want_chrom ? sizeof(uint32_t) * (size_t)(n_rows + 1) : 0);
    const size_t ob_bv0 = ob_brow + align256(nb1 * sizeof(long long));
    if (int rc = ctx->ensure_scratch_b(ob_bv0 + align256(nb1 * sizeof(long long)))) return rc;
    uint8_t *scb = (uint8_t *)ctx->scratch_b;
    uint32_t *d_off32 = (uint32_t *)(scb + ob_off32);
    long long *d_brow = (long long *)(scb + ob_brow), *d_bv0 = (long long *)(scb + ob_bv0);
    // outputs come from the device's stream-ordered pool (release threshold raised in ctx_create): a steady-state
    // query re-uses the memory the previous query's batches released instead of paying cudaMalloc
    if (c->want_pos) CUDA_TRY(cudaMallocAsync((void **)&c->d_pos, sizeof(int64_t) * (size_t)n_rows, st));
    if (c->want_chrom) {
        CUDA_TRY(cudaMallocAsync((void **)&c->d_values, (size_t)std::max<unsigned long long>(total_values, 1), st));
        CUDA_TRY(cudaMallocAsync((void **)&c->d_offsets, sizeof(int32_t) * (size_t)(c->n_batches * (c->batch_rows + 1)), st));
        CUDA_TRY(cudaMemcpyAsync(d_brow, c->batch_row0.data(), nb1 * sizeof(long long), cudaMemcpyHostToDevice, st));
    }
    a.pos = c->d_pos;
    a.off32 = d_off32;
    a.values = c->d_values;

    // ---- pass B (+ C) ----
    CUDA_TRY(launch_cols<true>(a, ctx->sm_count, st));
    ctx->launches.fetch_add(1);
    if (c->want_chrom) {
        batch_value_offsets<<<(unsigned)((nb1 + 127) / 128), 128, 0, st>>>(d_prefix, n_tiles, d_brow, c->n_batches, n_rows, total_values,
                                                                          d_off32, d_bv0);
        batch_offsets<<<(unsigned)c->n_batches, 256, 0, st>>>(d_brow, d_bv0, n_rows, total_values, c->batch_rows, d_off32, c->d_offsets);
        ctx->launches.fetch_add(2);
        CUDA_TRY(cudaGetLastError());
        c->batch_v0.resize(nb1);
        CUDA_TRY(cudaMemcpyAsync(c->batch_v0.data(), d_bv0, nb1 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_misc[4];
    CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((uint32_t)h_misc[0])
        return fail(EXON_GPU_ERR_PARSE, "malformed VCF record at row %llu:%s%s", h_misc[1],
                    ((uint32_t)h_misc[0] & kErrBadPos) ? " POS is not a positive decimal integer;" : "",
                    ((uint32_t)h_misc[0] & kErrShortLine) ? " line ended before the field being read;" : "");
    if (c->want_chrom) {
        for (int64_t b = 0; b < c->n_batches; ++b)
            if (c->batch_v0[(size_t)b + 1] - c->batch_v0[(size_t)b] > 0x7FFFFFFFll)
                return fail(EXON_GPU_ERR_UNSUPPORTED, "chrom bytes of batch %lld overflow int32 offsets", (long long)b);
    }
    if (!c->on_device) {
        if (c->want_pos) {
            CUDA_TRY(cudaHostAlloc((void **)&c->h_pos, sizeof(int64_t) * (size_t)n_rows, cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(c->h_pos, c->d_pos, sizeof(int64_t) * (size_t)n_rows, cudaMemcpyDeviceToHost, st));
        }
        if (c->want_chrom) {
            const size_t ob = sizeof(int32_t) * (size_t)(c->n_batches * (c->batch_rows + 1));
            CUDA_TRY(cudaHostAlloc((void **)&c->h_offsets, ob, cudaHostAllocDefault));
            CUDA_TRY(cudaHostAlloc((void **)&c->h_values, (size_t)std::max<unsigned long long>(total_values, 1), cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(c->h_offsets, c->d_offsets, ob, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(c->h_values, c->d_values, (size_t)total_values, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (wide_wanted(c->projection)) {
        int64_t n = n_rows;
        return wide_build(s, &c->batch_row0, &n, &c->wide);
    }
    return EXON_GPU_OK;
}

}  // namespace

static int ensure_columns(VcfStream *s) {
    if (s->cols) return EXON_GPU_OK;
    if (int rc = s->flush_gz()) return rc;
    if (s->file_open && s->tail_len > 0)
        return fail(EXON_GPU_ERR_STATE, "next_batch: the current file ends mid-line; finish it with is_last first");
    std::lock_guard<std::recursive_mutex> work(s->ctx->work_mu);
    if (int rc = build_columns(s)) {
        columns_free(s);
        return rc;
    }
    s->drained = true;
    return EXON_GPU_OK;
}

// FilterExec + AggregateExec(Partial) over every batch of the stream's column store in ONE launch of the
// multi-batch K3 kernel: the columns K2 built never leave HBM and 24 bytes come back.
int columns_filter_agg(VcfStream *s, const exon_gpu_pred *pred, const exon_gpu_agg *agg, exon_gpu_partial *out) {
    if (int rc = ensure_columns(s)) return rc;
    Ctx *ctx = s->ctx;
    Columns *c = s->cols;
    auto col_of = [&](int child) { return (child >= 0 && child < (int)s->projection.size()) ? s->projection[(size_t)child] : -1; };
    FaCommon k;
    if (int rc = fa_common_from(pred, agg, k)) return rc;
    if (k.has_chrom && col_of(pred->chrom_col) != 0) return fail(EXON_GPU_ERR_ARG, "vcf_filter_agg: chrom_col is not the projected chrom column");
    if (k.has_pos && col_of(pred->pos_col) != 1) return fail(EXON_GPU_ERR_ARG, "vcf_filter_agg: pos_col is not the projected pos column");
    int val_file_col = -1;
    if (agg->kind != EXON_GPU_AGG_COUNT_STAR) {
        val_file_col = col_of(agg->value_col);
        if (val_file_col < 0) return fail(EXON_GPU_ERR_ARG, "vcf_filter_agg: value_col out of range");
        if (val_file_col == 1) k.val_type = kValI64;
        else if (agg->kind == EXON_GPU_AGG_COUNT) k.val_type = kValNone;
        else return fail(EXON_GPU_ERR_UNSUPPORTED, "vcf_filter_agg: cannot sum the chrom column");
    }
    memset(out, 0, sizeof(*out));
    if (c->n_batches == 0) return EXON_GPU_OK;
    if ((k.has_chrom && !c->d_offsets) || ((k.has_pos || val_file_col == 1) && !c->d_pos))
        return fail(EXON_GPU_ERR_ARG, "vcf_filter_agg: the predicate reads a column outside the projection");
    const size_t table = (sizeof(FaBatchDesc) * (size_t)c->n_batches + 255) & ~(size_t)255;
    if (!c->d_descs) {
        std::vector<FaBatchDesc> h((size_t)c->n_batches);
        for (int64_t b = 0; b < c->n_batches; ++b) {
            FaBatchDesc &d = h[(size_t)b];
            memset(&d, 0, sizeof(d));
            const int64_t row0 = c->batch_row0[(size_t)b];
            if (c->d_offsets) {
                d.chrom_offsets = c->d_offsets + b * (c->batch_rows + 1);
                d.chrom_values = c->d_values + c->batch_v0[(size_t)b];
            }
            if (c->d_pos) {
                d.pos = c->d_pos + row0;
                d.val = c->d_pos + row0;
            }
            d.n_rows = c->batch_row0[(size_t)b + 1] - row0;
        }
        CUDA_TRY(cudaMalloc((void **)&c->d_descs, table + 64));
        CUDA_TRY(cudaMemcpyAsync(c->d_descs, h.data(), sizeof(FaBatchDesc) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    unsigned long long *d_out = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(c->d_descs) + table);
    CUDA_TRY(cudaMemsetAsync(d_out, 0, 64, ctx->stream));
    if (int rc = filter_agg_multi_launch(ctx, c->d_descs, (int)c->n_batches, c->batch_rows, k, d_out, true)) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->h_res + 8, d_out, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    out->count = (int64_t)s->h_res[8];
    out->sum_i64 = (int64_t)s->h_res[9];
    memcpy(&out->sum_f64, &s->h_res[10], sizeof(double));
    if (k.val_type == kValI64) out->sum_f64 = (double)out->sum_i64;
    return EXON_GPU_OK;
}

void vcf_stream_schema(VcfStream *s, ArrowSchema *out) { fill_schema(s->projection, out); }

int columns_next_batch(VcfStream *s, ArrowArray *out, ArrowSchema *out_schema) {
    if (int rc = ensure_columns(s)) return rc;
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
        if (s->projection[(size_t)i] >= 2) {
            wide_export(c->wide, s->projection[(size_t)i], b, rows, &a, &p->wide_slot[i]);
            if (p->wide_slot[i].item_ptr) p->wide_slot[i].item.release = release_child;
            a.release = release_child;
            p->child_ptrs[i] = &a;
            continue;
        }
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

extern "C" int exon_gpu_vcf_filter_agg(exon_gpu_stream *s, const exon_gpu_pred *pred, const exon_gpu_agg *agg,
                                       exon_gpu_partial *out) {
    if (!s || !agg || !out) return exon::fail(EXON_GPU_ERR_ARG, "vcf_filter_agg: NULL argument");
    if (agg->kind < EXON_GPU_AGG_COUNT_STAR || agg->kind > EXON_GPU_AGG_AVG)
        return exon::fail(EXON_GPU_ERR_ARG, "vcf_filter_agg: unknown aggregate kind %d", agg->kind);
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return exon::fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return exon::columns_filter_agg(s, pred, agg, out);
}

extern "C" int exon_gpu_vcf_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema) {
    if (!s || !out) return exon::fail(EXON_GPU_ERR_ARG, "vcf_next_batch: NULL argument");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return exon::fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return exon::columns_next_batch(s, out, out_schema);
}
