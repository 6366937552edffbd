// vcf_columns.cu -- K2: VCF text -> Arrow columns {chrom: utf8, pos: int64} (+ columns 2..6) in reference-sized batches.
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
// Columns 2..6 (id, ref, alt, qual, filter; lazy_array_builder.rs:169-216) ride in the same two passes (WIDE instantiations of the
// kernel): per-tile sums of their list entries / bytes in pass A, ranks inside a tile from warp scans in pass B, values at
// their final place plus absolute 32-bit offsets, and an element-wise pass C per batch for the layout arrow-rs emits; the
// per-line field positions come from tab / newline / semicolon bitmaps built once per tile (wide_fields_bitmap below).
// A row belongs to the tile that holds the '\n' before it (the first row of a segment to the segment's first
// tile), exactly as in K1.  Segments are cut at file ends, so no batch spans two files (the reference opens one
// AsyncBatchStream per file).  Batches own nothing: they are views into the stream's column store, kept alive by
// a reference count that the Arrow release callbacks decrement.
#include <cub/device/device_scan.cuh>

#include <atomic>
#include <cstring>
#include <type_traits>
#include <new>

#include "common.cuh"
#include "internal.h"
#include "vcf_tile.cuh"
#include "vcf_wide.cuh"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

// What a tile contributes: rows that start in it, CHROM bytes of those rows and, for the wide columns, list entries / bytes of
// id and filter and the bytes of ref.
struct TileSum {
    unsigned long long rows, bytes, idE, idB, refB, fiE, fiB, pad_;
};
struct TileSumAdd {
    __host__ __device__ __forceinline__ TileSum operator()(const TileSum &x, const TileSum &y) const {
        return TileSum{x.rows + y.rows, x.bytes + y.bytes, x.idE + y.idE, x.idB + y.idB, x.refB + y.refB, x.fiE + y.fiE, x.fiB + y.fiB, 0ull};
    }
};
// the two fields a projection of chrom / pos alone needs (a quarter of the scan's traffic)
struct TileSum2 {
    unsigned long long rows, bytes;
};
struct TileSum2Add {
    __host__ __device__ __forceinline__ TileSum2 operator()(const TileSum2 &x, const TileSum2 &y) const { return TileSum2{x.rows + y.rows, x.bytes + y.bytes}; }
};

struct ColArgs {
    const ScanSeg *segs;  // n_segs + 1 entries (sentinel carries tile0 = n_tiles)
    int32_t n_segs;
    int64_t n_tiles;
    int32_t want_chrom, want_pos;
    void *tile_stats;         // pass A out, one per tile: TileSum (WIDE) or TileSum2
    const void *tile_prefix;  // pass B in: exclusive prefix of tile_stats
    int64_t *pos;                // pass B out
    uint32_t *off32;             // pass B out: low 32 bits of the row's absolute offset into `values`
    uint8_t *values;             // pass B out
    uint32_t *flags;
    unsigned long long *first_bad_row;
    // ---- columns 2..6 inside the same two passes (WIDE instantiations) ----
    int32_t want_id, want_ref, want_alt, want_qual, want_filter;
    // pass B out, ABSOLUTE numbering (low 32 bits; pass C turns them into per-batch offsets): first list entry of a row, byte
    // offset of an entry / of a row's REF
    uint32_t *id_eabs, *id_vabs, *ref_vabs, *fi_eabs, *fi_vabs;
    uint8_t *id_val, *ref_val, *fi_val;
    float *qual;                                              // [row]
    uint32_t *id_valid_abs, *alt_valid_abs, *qual_valid_abs;  // one bit per absolute row (zeroed before the launch)
    QualSlow *qual_list;                                      // rows whose QUAL needs the exact parser
    unsigned long long *qual_list_n;                          // pass A counts them, pass B appends (zeroed in between)
    unsigned long long qual_list_cap;
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

// ---- columns 2..6 of one line (LazyVCFArrayBuilder::append, lazy_array_builder.rs:169-216; vcf_wide.cu has the row-parallel
// twin of this walk).  Readers: RdFast takes bytes straight from the staged tile and gives up at the window's end (the line is
// then walked again through RdView, which falls back to global memory and knows where the segment ends).
struct RdFast {
    uint32_t sa;
    int lim;
    bool over;
    __device__ __forceinline__ uint32_t operator()(int i) {
        if (i >= lim) {
            over = true;
            return '\n';
        }
        return lds8(sa + (uint32_t)i);
    }
};
struct RdView {
    TileView t;
    __device__ __forceinline__ uint32_t operator()(int i) { return ld_byte(t, i); }
};
struct WideRow {
    int t1, t2, t3, t4, t5, t6;  // tile-relative indices of the tabs that end POS, ID, REF, ALT, QUAL, FILTER
    int32_t idE, idB, fiE, fiB;  // list entries / bytes (a missing field: 0 / 0)
    uint32_t rf;                 // 1 id valid | 2 alt valid | 4 qual valid (q holds it) | 8 qual left to the exact parser
    float q;
    uint32_t err;
};
template <class Rd>
__device__ __forceinline__ void wide_fields(Rd &rd, int t0, WideRow &R) {
    R.idE = R.idB = R.fiE = R.fiB = 0;
    R.rf = 0;
    R.q = 0.0f;
    R.err = 0;
    R.t1 = R.t2 = R.t3 = R.t4 = R.t5 = R.t6 = t0;
    int i = t0 + 1;
    uint32_t c;
    auto field_end = [&](int &semi) -> bool {  // i -> the tab that ends the field; false when the line ends first
        semi = 0;
        while ((c = rd(i)) != '\t') {
            if (c == '\n') return false;
            semi += c == ';';
            ++i;
        }
        return true;
    };
    int semi;
    if (!field_end(semi)) goto bad;  // POS
    R.t1 = i;
    {
        const int f0 = ++i;
        if (!field_end(semi)) goto bad;  // ID
        R.t2 = i;
        const int n = i - f0;
        if (!(n == 0 || (n == 1 && rd(f0) == '.'))) {
            R.rf |= 1u;
            R.idE = semi + 1;
            R.idB = n - semi;
        }
    }
    ++i;
    if (!field_end(semi)) goto bad;  // REF
    R.t3 = i;
    {
        const int f0 = ++i;
        if (!field_end(semi)) goto bad;  // ALT
        R.t4 = i;
        const int n = i - f0;
        if (!(n == 0 || (n == 1 && rd(f0) == '.'))) R.rf |= 2u;
    }
    {
        const int f0 = ++i;
        uint32_t v = 0;
        bool plain = true;
        while ((c = rd(i)) != '\t') {  // QUAL: the usual one is a short unsigned integer, exact in f32 below 2^24
            if (c == '\n') goto bad;
            const uint32_t d = c - '0';
            plain = plain && d <= 9u;
            v = v * 10u + d;
            ++i;
        }
        R.t5 = i;
        const int n = i - f0;
        if (!(n == 1 && rd(f0) == '.')) {
            if (plain && n >= 1 && n <= 7) {
                R.q = (float)v;
                R.rf |= 4u;
            } else {
                R.rf |= 8u;
            }
        }
    }
    {
        const int f0 = ++i;
        if (!field_end(semi)) goto bad;  // FILTER (an 8th field must follow: the tab is required)
        R.t6 = i;
        const int n = i - f0;
        if (!(n == 0 || (n == 1 && rd(f0) == '.'))) {
            R.fiE = semi + 1;
            R.fiB = n - semi;
        }
    }
    return;
bad:
    R.err = kWErrFields;
}
// The same walk for the lines of an interior tile, from three BITMAPS of the staged bytes (tab, newline, semicolon: one bit per
// byte, built once per tile by all lanes, 16 bytes each -- a per-line SWAR window looked at every byte 2.4 times and paid
// the unaligned fetch per line): 64 bits of each from the line's first byte on, the first seven tabs by find-first-set, the
// list entries by population counts; single bytes are read only to tell "." from a one-character value and for QUAL.
// false: the line ends before its eighth field or the seven tabs are not within 64 staged bytes -- the caller walks the line
// byte by byte (and reports what is wrong with it).
constexpr int kBmU = 8;                                        // == kColU (asserted at the launch)
constexpr int kBmChunks = (512 * kBmU + kHalo) / 16;           // 16-byte chunks of the staged tile (kBmU = the kernel's U)
constexpr int kBmWords = ((kBmChunks + 1) / 2 + 2 + 3) & ~3;   // + two zero words behind the last one (64-bit reads), 16-byte multiple
constexpr int kBmBytes = 3 * kBmWords * 4;                     // per warp
__device__ __forceinline__ unsigned long long bits64(uint32_t bm_sa, int bit) {
    const uint32_t wa = bm_sa + 4u * (uint32_t)(bit >> 5), sh = (uint32_t)bit & 31u;
    const uint32_t w0 = lds32(wa), w1 = lds32(wa + 4), w2 = lds32(wa + 8);
    return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}
__device__ __forceinline__ unsigned long long bit_range(int a, int b) {  // bits [a, b), 0 <= a <= b <= 64, b - a < 64
    return b <= a ? 0ull : ((~0ull >> (64 - (b - a))) << a);
}
__device__ __forceinline__ bool wide_fields_bitmap(uint32_t bm_sa, uint32_t sa, int ls, WideRow &R) {
    const unsigned long long T = bits64(bm_sa, ls), N = bits64(bm_sa + kBmWords * 4, ls), S = bits64(bm_sa + 2 * kBmWords * 4, ls);
    const int nl = N ? __ffsll((long long)N) - 1 : 64;
    uint32_t lo = (uint32_t)T, hi = (uint32_t)(T >> 32);
    int t[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        if (lo) {
            t[k] = __ffs((int)lo) - 1;
            lo &= lo - 1u;
        } else if (hi) {
            t[k] = 31 + __ffs((int)hi);
            hi &= hi - 1u;
        } else {
            return false;
        }
    }
    if (t[6] >= nl) return false;
    const uint32_t la = sa + (uint32_t)ls;
    R.t1 = ls + t[1], R.t2 = ls + t[2], R.t3 = ls + t[3], R.t4 = ls + t[4], R.t5 = ls + t[5], R.t6 = ls + t[6];
    R.idE = R.idB = R.fiE = R.fiB = 0;
    R.rf = 0;
    R.q = 0.0f;
    R.err = 0;
    auto list_field = [&](int f0, int f1, int32_t &ne, int32_t &nb) -> bool {  // false: missing
        const int n = f1 - f0;
        if (n == 0 || (n == 1 && lds8(la + (uint32_t)f0) == '.')) return false;
        const int semi = __popcll(S & bit_range(f0, f1));
        ne = semi + 1;
        nb = n - semi;
        return true;
    };
    if (list_field(t[1] + 1, t[2], R.idE, R.idB)) R.rf |= 1u;
    {
        const int n = t[4] - t[3] - 1;
        if (!(n == 0 || (n == 1 && lds8(la + (uint32_t)(t[3] + 1)) == '.'))) R.rf |= 2u;
    }
    {
        const int f0 = t[4] + 1, n = t[5] - f0;
        if (!(n == 1 && lds8(la + (uint32_t)f0) == '.')) {
            uint32_t v = 0;
            bool plain = n >= 1 && n <= 7;
            for (int i = 0; plain && i < n; ++i) {
                const uint32_t d = lds8(la + (uint32_t)(f0 + i)) - '0';
                plain = d <= 9u;
                v = v * 10u + d;
            }
            if (plain) {
                R.q = (float)v;
                R.rf |= 4u;
            } else {
                R.rf |= 8u;
            }
        }
    }
    list_field(t[5] + 1, t[6], R.fiE, R.fiB);
    return true;
}
// items of a list cell [f0, f1): byte offsets of the items (absolute) and the bytes without the ';'
template <class Rd>
__device__ __forceinline__ void wide_items(Rd &rd, int f0, int f1, uint32_t *vabs, uint8_t *val, unsigned long long e, unsigned long long v) {
    vabs[e] = (uint32_t)v;
    for (int i = f0; i < f1; ++i) {
        const uint32_t c = rd(i);
        if (c == ';') vabs[++e] = (uint32_t)v;
        else val[v++] = (uint8_t)c;
    }
}

template <bool EMIT, bool WIDE, int U, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WIDE ? 2 : ctas_per_sm<U, S, WARPS>()) vcf_cols_kernel(const __grid_constant__ ColArgs a) {
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
    const uint32_t bm_sa = smem_u32(smem_raw + ((L::total + 15) & ~(size_t)15)) + (uint32_t)(warp * kBmBytes);  // WIDE: the warp's three bitmaps

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

    const bool need_fields = EMIT || WIDE || a.want_chrom;  // pass A of a pos-only projection just counts lines
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
        if (WIDE && interior) {
            // the tile's tab / newline / semicolon bitmaps (halo included): every staged byte is looked at once, by one lane
            static_assert(U == kBmU, "bitmap sizes are fixed for this U");
#pragma unroll 1
            for (int u = 0; u <= U; ++u) {
                const int ch = u * 32 + lane;
                if (ch < kBmChunks) {
                    const uint4 w = lds128(sa + (uint32_t)ch * 16u);
                    const uint32_t bt = pack16(zero_bytes_exact(w.x ^ 0x09090909u), zero_bytes_exact(w.y ^ 0x09090909u), zero_bytes_exact(w.z ^ 0x09090909u),
                                               zero_bytes_exact(w.w ^ 0x09090909u));
                    const uint32_t bn = pack16(zero_bytes_exact(w.x ^ kNL4), zero_bytes_exact(w.y ^ kNL4), zero_bytes_exact(w.z ^ kNL4), zero_bytes_exact(w.w ^ kNL4));
                    const uint32_t bs = pack16(zero_bytes_exact(w.x ^ 0x3B3B3B3Bu), zero_bytes_exact(w.y ^ 0x3B3B3B3Bu), zero_bytes_exact(w.z ^ 0x3B3B3B3Bu),
                                               zero_bytes_exact(w.w ^ 0x3B3B3B3Bu));
                    sts16(bm_sa + 2u * (uint32_t)ch, bt);
                    sts16(bm_sa + kBmWords * 4 + 2u * (uint32_t)ch, bn);
                    sts16(bm_sa + 2 * kBmWords * 4 + 2u * (uint32_t)ch, bs);
                } else if (ch < kBmWords * 2) {
                    sts16(bm_sa + 2u * (uint32_t)ch, 0u);
                    sts16(bm_sa + kBmWords * 4 + 2u * (uint32_t)ch, 0u);
                    sts16(bm_sa + 2 * kBmWords * 4 + 2u * (uint32_t)ch, 0u);
                }
            }
            __syncwarp();
        }

        unsigned long long row0 = 0, vb = 0;
        if (EMIT) {
            using TS = typename std::conditional<WIDE, TileSum, TileSum2>::type;
            row0 = __ldg(&static_cast<const TS *>(a.tile_prefix)[T].rows);
            vb = __ldg(&static_cast<const TS *>(a.tile_prefix)[T].bytes);
        }
        uint32_t t_rows = 0;           // rows of this tile already drained (warp-uniform)
        unsigned long long t_bytes = 0;  // EMIT: CHROM bytes already placed (warp-uniform); pass A: this lane's partial sum
        uint32_t lane_rows = 0;        // pass A without CHROM: this lane's line count
        int qn = 0;
        // columns 2..6: EMIT: entries / bytes already placed in this tile (warp-uniform); pass A: this lane's partial sums
        uint32_t w_idE = 0, w_idB = 0, w_refB = 0, w_fiE = 0, w_fiB = 0;
        TileSum pre = TileSum{0, 0, 0, 0, 0, 0, 0, 0};
        if (EMIT && WIDE) pre = static_cast<const TileSum *>(a.tile_prefix)[T];

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
                WideRow R;
                uint32_t refB = 0;
                if (WIDE) {
                    R.idE = R.idB = R.fiE = R.fiB = 0;
                    R.rf = 0;
                    R.q = 0.0f;
                    R.err = 0;
                    R.t1 = R.t2 = R.t3 = R.t4 = R.t5 = R.t6 = 0;
                    if (act && !e) {
                        const int t0 = ls + (int)clen;
                        bool done = false;
                        if (interior) done = wide_fields_bitmap(bm_sa, sa, ls, R) && R.t1 > t0;
                        if (!done && interior) {
                            RdFast rd{sa, sm_hi, false};
                            wide_fields(rd, t0, R);
                            done = !rd.over;
                        }
                        if (!done) {
                            RdView rd{TileView{sm, g, seg_lo, hi, sm_lo, sm_hi}};
                            wide_fields(rd, t0, R);
                        }
                        if (R.err) {
                            R.idE = R.idB = R.fiE = R.fiB = 0;
                            R.rf = 0;
                        } else {
                            refB = (uint32_t)(R.t3 - R.t2 - 1);
                        }
                    }
                    const uint32_t slow_q = __ballot_sync(0xFFFFFFFFu, (R.rf & 8u) != 0u);
                    if (!EMIT) {
                        w_idE += (uint32_t)R.idE, w_idB += (uint32_t)R.idB, w_refB += refB, w_fiE += (uint32_t)R.fiE, w_fiB += (uint32_t)R.fiB;
                        if (slow_q && lane == 0 && a.want_qual) atomicAdd(a.qual_list_n, (unsigned long long)__popc(slow_q));
                    }  // (a malformed line contributes nothing here and is reported, with its row number, by pass B)
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
                    if (WIDE) {
                        uint32_t s_idE = (uint32_t)R.idE, s_idB = (uint32_t)R.idB, s_refB = refB, s_fiE = (uint32_t)R.fiE, s_fiB = (uint32_t)R.fiB;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t v0 = __shfl_up_sync(0xFFFFFFFFu, s_idE, d), v1 = __shfl_up_sync(0xFFFFFFFFu, s_idB, d),
                                           v2 = __shfl_up_sync(0xFFFFFFFFu, s_refB, d), v3 = __shfl_up_sync(0xFFFFFFFFu, s_fiE, d),
                                           v4 = __shfl_up_sync(0xFFFFFFFFu, s_fiB, d);
                            if (lane >= d) s_idE += v0, s_idB += v1, s_refB += v2, s_fiE += v3, s_fiB += v4;
                        }
                        const unsigned long long row = row0 + t_rows + (uint32_t)i;
                        if (act) {
                            auto emit_cells = [&](auto &rd) {
                                if (a.want_id) {
                                    const unsigned long long ea = pre.idE + w_idE + (s_idE - (uint32_t)R.idE), va = pre.idB + w_idB + (s_idB - (uint32_t)R.idB);
                                    a.id_eabs[row] = (uint32_t)ea;
                                    if (R.rf & 1u) wide_items(rd, R.t1 + 1, R.t2, a.id_vabs, a.id_val, ea, va);
                                }
                                if (a.want_ref) {
                                    const unsigned long long va = pre.refB + w_refB + (s_refB - refB);
                                    a.ref_vabs[row] = (uint32_t)va;
                                    for (uint32_t j = 0; j < refB; ++j) a.ref_val[va + j] = (uint8_t)rd(R.t2 + 1 + (int)j);
                                }
                                if (a.want_filter) {
                                    const unsigned long long ea = pre.fiE + w_fiE + (s_fiE - (uint32_t)R.fiE), va = pre.fiB + w_fiB + (s_fiB - (uint32_t)R.fiB);
                                    a.fi_eabs[row] = (uint32_t)ea;
                                    if (R.fiE) wide_items(rd, R.t5 + 1, R.t6, a.fi_vabs, a.fi_val, ea, va);
                                }
                            };
                            if (interior && R.t6 < sm_hi) {  // every byte the cells hold is staged: no range checks
                                RdFast rd{sa, sm_hi, false};
                                emit_cells(rd);
                            } else {
                                RdView rd{TileView{sm, g, seg_lo, hi, sm_lo, sm_hi}};
                                emit_cells(rd);
                            }
                            if (a.want_qual) {
                                a.qual[row] = R.q;
                                if (R.rf & 8u) {
                                    const unsigned long long k = atomicAdd(a.qual_list_n, 1ull);
                                    if (k < a.qual_list_cap) a.qual_list[k] = QualSlow{g + R.t4 + 1, (uint32_t)(R.t5 - R.t4 - 1), 0u, row};
                                }
                            }
                            if (R.err) {
                                atomicOr(a.flags + 1, R.err);
                                atomicMin(a.first_bad_row, row);
                            }
                        }
                        // validity: the 32 rows of this step are consecutive, so their bits fall into two bitmap words
                        const unsigned long long rfirst = row0 + t_rows + (uint32_t)i0;
                        const uint32_t sh = (uint32_t)(rfirst & 31ull);
                        const unsigned long long wd = rfirst >> 5;
                        auto put_valid = [&](uint32_t *bm, uint32_t bit) {
                            const uint32_t m = __ballot_sync(0xFFFFFFFFu, (R.rf & bit) != 0u);
                            if (lane == 0) {
                                if (m << sh) atomicOr(bm + wd, m << sh);
                                if (sh && (m >> (32u - sh))) atomicOr(bm + wd + 1, m >> (32u - sh));
                            }
                        };
                        if (a.want_id) put_valid(a.id_valid_abs, 1u);
                        if (a.want_alt) put_valid(a.alt_valid_abs, 2u);
                        if (a.want_qual) put_valid(a.qual_valid_abs, 4u);
                        w_idE += __shfl_sync(0xFFFFFFFFu, s_idE, 31), w_idB += __shfl_sync(0xFFFFFFFFu, s_idB, 31), w_refB += __shfl_sync(0xFFFFFFFFu, s_refB, 31);
                        w_fiE += __shfl_sync(0xFFFFFFFFu, s_fiE, 31), w_fiB += __shfl_sync(0xFFFFFFFFu, s_fiB, 31);
                    }
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
            unsigned long long s0 = w_idE, s1 = w_idB, s2 = w_refB, s3 = w_fiE, s4 = w_fiB;
            if (WIDE) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, d), s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, d), s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, d);
                    s3 += __shfl_xor_sync(0xFFFFFFFFu, s3, d), s4 += __shfl_xor_sync(0xFFFFFFFFu, s4, d);
                }
            }
            if (lane == 0) {
                if (WIDE) static_cast<TileSum *>(a.tile_stats)[T] = TileSum{rows, bytes, s0, s1, s2, s3, s4, 0ull};
                else static_cast<TileSum2 *>(a.tile_stats)[T] = TileSum2{rows, bytes};
            }
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

template <bool EMIT, bool WIDE>
cudaError_t launch_cols(const ColArgs &args, int sm_count, cudaStream_t stream) {
    constexpr size_t smem = ((SmemLayout<kColU, kColS, kColW>::total + 15) & ~(size_t)15) + (WIDE ? (size_t)kColW * kBmBytes : 0);
    auto kern = vcf_cols_kernel<EMIT, WIDE, kColU, kColS, kColW>;
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
template <class TS>
__global__ void gather_prefix(const TS *prefix, const long long *idx, int n, TS *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = prefix[idx[i]];
}

// Full 64-bit value offset of the first row of every batch (and the grand total at index n_batches): the tile
// that holds the row is found by binary search over the tile prefix; the row's u32 offset supplies the low bits.
template <class TS>
__global__ void batch_value_offsets(const TS *prefix, int64_t n_tiles, const long long *batch_row0, int64_t n_batches,
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

// ---- pass C of the wide columns: absolute numbering -> the per-batch layout arrow-rs' builders emit ----
struct WideAbs {
    const uint32_t *id_eabs, *id_vabs, *ref_vabs, *fi_eabs, *fi_vabs;
    unsigned long long tot[5];  // kIdE, kIdB, kRefB, kFiE, kFiB
    int32_t want_id, want_ref, want_filter;
};
// 64-bit entry / byte number of the first row of every batch (and the totals at index n_batches): the tile that holds the
// row comes from a binary search over the tile prefix, the low 32 bits from the row's absolute offsets
__global__ void wide_batch_bases(const TileSum *prefix, int64_t n_tiles, const long long *brow, int64_t n_batches, int64_t n_rows, WideAbs w,
                                 long long *base /* [5][n_batches + 1] */) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_batches) return;
    const int64_t nb1 = n_batches + 1;
    const unsigned long long row = (unsigned long long)brow[b];
    if ((int64_t)row >= n_rows) {
        for (int k = 0; k < 5; ++k) base[k * nb1 + b] = (long long)w.tot[k];
        return;
    }
    int64_t lo = 0, hi = n_tiles;  // last tile t with prefix[t].rows <= row
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (prefix[mid].rows <= row) lo = mid;
        else hi = mid;
    }
    const TileSum p = prefix[lo];
    auto widen = [](unsigned long long tile_base, uint32_t low) { return tile_base + (uint32_t)(low - (uint32_t)tile_base); };
    unsigned long long idE = 0, idB = 0, refB = 0, fiE = 0, fiB = 0;
    if (w.want_id) {
        idE = widen(p.idE, w.id_eabs[row]);
        idB = widen(p.idB, w.id_vabs[idE]);
    }
    if (w.want_ref) refB = widen(p.refB, w.ref_vabs[row]);
    if (w.want_filter) {
        fiE = widen(p.fiE, w.fi_eabs[row]);
        fiB = widen(p.fiB, w.fi_vabs[fiE]);
    }
    base[0 * nb1 + b] = (long long)idE, base[1 * nb1 + b] = (long long)idB, base[2 * nb1 + b] = (long long)refB, base[3 * nb1 + b] = (long long)fiE,
                   base[4 * nb1 + b] = (long long)fiB;
}
// offsets[b * (batch_rows + 1) + i] = abs[row0(b) + i] - abs[row0(b)], i = 0 .. rows of the batch (abs has n_rows + 1 entries)
__global__ void __launch_bounds__(256) wide_rebase_rows(const long long *brow, int batch_rows, const uint32_t *abs, int32_t *out) {
    const int64_t b = blockIdx.x;
    const long long r0 = brow[b];
    const int n = (int)(brow[b + 1] - r0);
    const uint32_t base = abs[r0];
    int32_t *o = out + b * (int64_t)(batch_rows + 1);
    for (int i = threadIdx.x; i <= n; i += blockDim.x) o[i] = (int32_t)(abs[r0 + i] - base);
}
// child offsets of a batch: entries e0(b) .. e0(b + 1) (inclusive: the closing offset) at out[e0(b) + b ..]
__global__ void __launch_bounds__(256) wide_rebase_children(const long long *ebase, const uint32_t *vabs, int32_t *out) {
    const int64_t b = blockIdx.x;
    const long long e0 = ebase[b], n = ebase[b + 1] - e0;
    const uint32_t base = vabs[e0];
    int32_t *o = out + e0 + b;
    for (long long j = threadIdx.x; j <= n; j += blockDim.x) o[j] = (int32_t)(vabs[e0 + j] - base);
}
// validity words of a batch from the one-bit-per-absolute-row bitmap (which has a spare word at its end)
__global__ void __launch_bounds__(256) wide_repack_valid(const long long *brow, int wpb, const uint32_t *abs_bits, uint32_t *out) {
    const int64_t b = blockIdx.x;
    const unsigned long long r0 = (unsigned long long)brow[b];
    const int n = (int)(brow[b + 1] - (long long)r0);
    for (int w = threadIdx.x; w < (n + 31) / 32; w += blockDim.x) {
        const unsigned long long r = r0 + 32ull * (unsigned)w;
        uint32_t v = __funnelshift_r(abs_bits[r >> 5], abs_bits[(r >> 5) + 1], (uint32_t)(r & 31ull));
        const int left = n - 32 * w;
        if (left < 32) v &= (1u << left) - 1u;
        out[b * (int64_t)wpb + w] = v;
    }
}
__global__ void wide_set_terminals(uint32_t *id_eabs, uint32_t *id_vabs, uint32_t *ref_vabs, uint32_t *fi_eabs, uint32_t *fi_vabs, int64_t n_rows, WideAbs w) {
    if (w.want_id) id_eabs[n_rows] = (uint32_t)w.tot[0], id_vabs[w.tot[0]] = (uint32_t)w.tot[1];
    if (w.want_ref) ref_vabs[n_rows] = (uint32_t)w.tot[2];
    if (w.want_filter) fi_eabs[n_rows] = (uint32_t)w.tot[3], fi_vabs[w.tot[3]] = (uint32_t)w.tot[4];
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
    // Columns 2..6 ride in K2's two passes (WIDE instantiations of the tile kernel).  INFO / FORMAT (7, 8) are built row-parallel on
    // the line index (vcf_wide.cu); when one of them is projected that build takes columns 2..6 along.
    bool text78 = false;
    for (int p : s->projection) text78 = text78 || p >= 7;
    const bool tile_wide = wide_wanted(c->projection) && !text78;
    if (!c->want_chrom && !c->want_pos && wide_wanted(c->projection) && !tile_wide) {
        // only columns 2..8: the line index of the wide build also yields the batch table
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
    size_t cub_bytes = 0, cub_bytes2 = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, cub_bytes, (TileSum *)nullptr, (TileSum *)nullptr, TileSumAdd(), TileSum{0, 0, 0, 0, 0, 0, 0, 0},
                                            (int)(n_tiles + 1), st));
    CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, cub_bytes2, (TileSum2 *)nullptr, (TileSum2 *)nullptr, TileSum2Add(), TileSum2{0, 0}, (int)(n_tiles + 1), st));
    cub_bytes = std::max(cub_bytes, cub_bytes2);
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
    const size_t ts_bytes = tile_wide ? sizeof(TileSum) : sizeof(TileSum2);  // element of the stats / prefix / gather arrays
    CUDA_TRY(cudaMemsetAsync(reinterpret_cast<uint8_t *>(d_stats) + (size_t)n_tiles * ts_bytes, 0, ts_bytes, st));
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
    a.flags = reinterpret_cast<uint32_t *>(d_misc);  // low word: chrom / pos errors, high word: columns 2..6 (kWErr*)
    a.first_bad_row = d_misc + 1;
    a.qual_list_n = d_misc + 3;
    bool want_col[9] = {false, false, false, false, false, false, false, false, false};
    for (int p : s->projection) want_col[p] = true;
    if (tile_wide) a.want_id = want_col[2], a.want_ref = want_col[3], a.want_alt = want_col[4], a.want_qual = want_col[5], a.want_filter = want_col[6];

    // ---- pass A + scan ----
    CUDA_TRY((tile_wide ? launch_cols<false, true>(a, ctx->sm_count, st) : launch_cols<false, false>(a, ctx->sm_count, st)));
    ctx->launches.fetch_add(1);
    if (tile_wide) {
        CUDA_TRY(cub::DeviceScan::ExclusiveScan(scr + o_cub, cub_bytes, d_stats, d_prefix, TileSumAdd(), TileSum{0, 0, 0, 0, 0, 0, 0, 0}, (int)(n_tiles + 1), st));
        gather_prefix<TileSum><<<(unsigned)((file_tiles.size() + 127) / 128), 128, 0, st>>>(d_prefix, d_gidx, (int)file_tiles.size(), d_gout);
    } else {
        CUDA_TRY(cub::DeviceScan::ExclusiveScan(scr + o_cub, cub_bytes, reinterpret_cast<TileSum2 *>(d_stats), reinterpret_cast<TileSum2 *>(d_prefix), TileSum2Add(),
                                                TileSum2{0, 0}, (int)(n_tiles + 1), st));
        gather_prefix<TileSum2><<<(unsigned)((file_tiles.size() + 127) / 128), 128, 0, st>>>(reinterpret_cast<const TileSum2 *>(d_prefix), d_gidx, (int)file_tiles.size(),
                                                                                           reinterpret_cast<TileSum2 *>(d_gout));
    }
    ctx->launches.fetch_add(2);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(ctx->h_scratch, d_gout, file_tiles.size() * ts_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<TileSum> h_gv(file_tiles.size());
    for (size_t f = 0; f < file_tiles.size(); ++f) {
        if (tile_wide) {
            h_gv[f] = static_cast<const TileSum *>(ctx->h_scratch)[f];
        } else {
            const TileSum2 t2 = static_cast<const TileSum2 *>(ctx->h_scratch)[f];
            h_gv[f] = TileSum{t2.rows, t2.bytes, 0, 0, 0, 0, 0, 0};
        }
    }
    const TileSum *h_g = h_gv.data();
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
    if (!c->want_chrom && !c->want_pos && !tile_wide) return EXON_GPU_OK;  // empty projection: row counts only

    // ---- outputs + scratch B: absolute u32 offsets | batch_row0 | batch_v0 ----
    const size_t nb1 = (size_t)c->n_batches + 1;
    const size_t ob_off32 = 0;
    const size_t ob_brow = ob_off32 + align256(c->want_chrom ? sizeof(uint32_t) * (size_t)(n_rows + 1) : 0);
    const size_t ob_bv0 = ob_brow + align256(nb1 * sizeof(long long));
    // columns 2..6: absolute offsets per row / per entry (+ 1 closing element each), three one-bit-per-row bitmaps, batch bases, QUAL list
    const TileSum tot = h_g[n_files_max];
    unsigned long long h_misc_a[4] = {0, 0, 0, 0};
    if (tile_wide) {
        CUDA_TRY(cudaMemcpyAsync(h_misc_a, d_misc, sizeof(h_misc_a), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemsetAsync(d_misc + 3, 0, 8, st));
    }
    const unsigned long long n_qslow = h_misc_a[3];
    const size_t rows1 = (size_t)n_rows + 1, bm_words = ((size_t)n_rows + 31) / 32 + 2;
    const size_t ow_ide = ob_bv0 + align256(nb1 * sizeof(long long));
    const size_t ow_idv = ow_ide + align256(tile_wide && want_col[2] ? rows1 * 4 : 0);
    const size_t ow_ref = ow_idv + align256(tile_wide && want_col[2] ? ((size_t)tot.idE + 1) * 4 : 0);
    const size_t ow_fie = ow_ref + align256(tile_wide && want_col[3] ? rows1 * 4 : 0);
    const size_t ow_fiv = ow_fie + align256(tile_wide && want_col[6] ? rows1 * 4 : 0);
    const size_t ow_bm = ow_fiv + align256(tile_wide && want_col[6] ? ((size_t)tot.fiE + 1) * 4 : 0);
    const size_t ow_base = ow_bm + align256(tile_wide ? 3 * bm_words * 4 : 0);
    const size_t ow_ql = ow_base + align256(tile_wide ? 5 * nb1 * 8 : 0);
    const size_t ow_end = ow_ql + align256((size_t)n_qslow * sizeof(QualSlow));
    if (int rc = ctx->ensure_scratch_b(ow_end)) return rc;
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
    if (!c->want_chrom && !c->want_pos) CUDA_TRY(cudaMemcpyAsync(d_brow, c->batch_row0.data(), nb1 * sizeof(long long), cudaMemcpyHostToDevice, st));
    else if (tile_wide && !c->want_chrom) CUDA_TRY(cudaMemcpyAsync(d_brow, c->batch_row0.data(), nb1 * sizeof(long long), cudaMemcpyHostToDevice, st));
    WideStore *w = nullptr;
    uint32_t *bm_abs = reinterpret_cast<uint32_t *>(scb + ow_bm);
    long long *d_wbase = reinterpret_cast<long long *>(scb + ow_base);
    auto wide_alloc = [&](WideBuf &b, size_t bytes, bool zero) -> int {
        b.bytes = std::max<size_t>(bytes, 8);
        CUDA_TRY(cudaMallocAsync(&b.d, b.bytes, st));
        if (zero) CUDA_TRY(cudaMemsetAsync(b.d, 0, b.bytes, st));
        return EXON_GPU_OK;
    };
    if (tile_wide) {
        w = new (std::nothrow) WideStore();
        if (!w) return fail(EXON_GPU_ERR_OOM, "next_batch: out of host memory");
        c->wide = w;
        w->device = ctx->device;
        w->on_device = s->columns_on_device;
        w->batch_rows = s->batch_rows;
        w->wpb = ((s->batch_rows + 63) / 64) * 2;
        for (int p : s->projection) w->want[p] = true;
        w->n_rows = n_rows;
        w->n_batches = c->n_batches;
        w->batch_row0 = c->batch_row0;
        const size_t valid_bytes = (size_t)w->n_batches * (size_t)w->wpb * 4, loff_bytes = (size_t)w->n_batches * (size_t)(w->batch_rows + 1) * 4;
        if (want_col[2]) {
            if (int rc = wide_alloc(w->id_loff, loff_bytes, false)) return rc;
            if (int rc = wide_alloc(w->id_coff, ((size_t)tot.idE + nb1) * 4, false)) return rc;
            if (int rc = wide_alloc(w->id_val, (size_t)tot.idB, false)) return rc;
            if (int rc = wide_alloc(w->id_valid, valid_bytes, true)) return rc;
            a.id_eabs = reinterpret_cast<uint32_t *>(scb + ow_ide), a.id_vabs = reinterpret_cast<uint32_t *>(scb + ow_idv);
            a.id_val = (uint8_t *)w->id_val.d, a.id_valid_abs = bm_abs;
        }
        if (want_col[3]) {
            if (int rc = wide_alloc(w->ref_off, loff_bytes, false)) return rc;
            if (int rc = wide_alloc(w->ref_val, (size_t)tot.refB, false)) return rc;
            a.ref_vabs = reinterpret_cast<uint32_t *>(scb + ow_ref), a.ref_val = (uint8_t *)w->ref_val.d;
        }
        if (want_col[4]) {
            if (int rc = wide_alloc(w->alt_valid, valid_bytes, true)) return rc;
            if (int rc = wide_alloc(w->zeros, (size_t)(w->batch_rows + 1) * 4, true)) return rc;
            a.alt_valid_abs = bm_abs + bm_words;
        }
        if (want_col[5]) {
            if (int rc = wide_alloc(w->qual, (size_t)n_rows * 4, false)) return rc;
            if (int rc = wide_alloc(w->qual_valid, valid_bytes, true)) return rc;
            a.qual = (float *)w->qual.d, a.qual_valid_abs = bm_abs + 2 * bm_words;
            a.qual_list = reinterpret_cast<QualSlow *>(scb + ow_ql);
            a.qual_list_cap = n_qslow;
        }
        if (want_col[6]) {
            if (int rc = wide_alloc(w->fi_loff, loff_bytes, false)) return rc;
            if (int rc = wide_alloc(w->fi_coff, ((size_t)tot.fiE + nb1) * 4, false)) return rc;
            if (int rc = wide_alloc(w->fi_val, (size_t)tot.fiB, false)) return rc;
            a.fi_eabs = reinterpret_cast<uint32_t *>(scb + ow_fie), a.fi_vabs = reinterpret_cast<uint32_t *>(scb + ow_fiv);
            a.fi_val = (uint8_t *)w->fi_val.d;
        }
        CUDA_TRY(cudaMemsetAsync(bm_abs, 0, 3 * bm_words * 4, st));
    }

    // ---- pass B (+ C) ----
    CUDA_TRY((tile_wide ? launch_cols<true, true>(a, ctx->sm_count, st) : launch_cols<true, false>(a, ctx->sm_count, st)));
    ctx->launches.fetch_add(1);
    if (tile_wide) {
        WideAbs wa;
        wa.id_eabs = a.id_eabs, wa.id_vabs = a.id_vabs, wa.ref_vabs = a.ref_vabs, wa.fi_eabs = a.fi_eabs, wa.fi_vabs = a.fi_vabs;
        wa.tot[0] = tot.idE, wa.tot[1] = tot.idB, wa.tot[2] = tot.refB, wa.tot[3] = tot.fiE, wa.tot[4] = tot.fiB;
        wa.want_id = want_col[2], wa.want_ref = want_col[3], wa.want_filter = want_col[6];
        wide_set_terminals<<<1, 1, 0, st>>>(a.id_eabs, a.id_vabs, a.ref_vabs, a.fi_eabs, a.fi_vabs, n_rows, wa);
        if (want_col[5] && n_qslow)
            if (int rc = wide_qual_list(ctx, a.qual_list, n_qslow, a.qual, a.qual_valid_abs, a.flags + 1, a.first_bad_row)) return rc;
        wide_batch_bases<<<(unsigned)((nb1 + 127) / 128), 128, 0, st>>>(d_prefix, n_tiles, d_brow, c->n_batches, n_rows, wa, d_wbase);
        const unsigned nbat = (unsigned)c->n_batches;
        if (want_col[2]) {
            wide_rebase_rows<<<nbat, 256, 0, st>>>(d_brow, c->batch_rows, a.id_eabs, (int32_t *)w->id_loff.d);
            wide_rebase_children<<<nbat, 256, 0, st>>>(d_wbase + 0 * nb1, a.id_vabs, (int32_t *)w->id_coff.d);
            wide_repack_valid<<<nbat, 256, 0, st>>>(d_brow, w->wpb, a.id_valid_abs, (uint32_t *)w->id_valid.d);
            ctx->launches.fetch_add(3);
        }
        if (want_col[3]) {
            wide_rebase_rows<<<nbat, 256, 0, st>>>(d_brow, c->batch_rows, a.ref_vabs, (int32_t *)w->ref_off.d);
            ctx->launches.fetch_add(1);
        }
        if (want_col[4]) {
            wide_repack_valid<<<nbat, 256, 0, st>>>(d_brow, w->wpb, a.alt_valid_abs, (uint32_t *)w->alt_valid.d);
            ctx->launches.fetch_add(1);
        }
        if (want_col[5]) {
            wide_repack_valid<<<nbat, 256, 0, st>>>(d_brow, w->wpb, a.qual_valid_abs, (uint32_t *)w->qual_valid.d);
            ctx->launches.fetch_add(1);
        }
        if (want_col[6]) {
            wide_rebase_rows<<<nbat, 256, 0, st>>>(d_brow, c->batch_rows, a.fi_eabs, (int32_t *)w->fi_loff.d);
            wide_rebase_children<<<nbat, 256, 0, st>>>(d_wbase + 3 * nb1, a.fi_vabs, (int32_t *)w->fi_coff.d);
            ctx->launches.fetch_add(2);
        }
        ctx->launches.fetch_add(2);
        CUDA_TRY(cudaGetLastError());
        static const int kOf[5] = {kIdE, kIdB, kRefB, kFiE, kFiB};
        for (int k = 0; k < 5; ++k) {
            w->base[kOf[k]].resize(nb1);
            CUDA_TRY(cudaMemcpyAsync(w->base[kOf[k]].data(), d_wbase + (size_t)k * nb1, nb1 * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    if (c->want_chrom) {
        if (tile_wide)
            batch_value_offsets<TileSum><<<(unsigned)((nb1 + 127) / 128), 128, 0, st>>>(d_prefix, n_tiles, d_brow, c->n_batches, n_rows, total_values, d_off32, d_bv0);
        else
            batch_value_offsets<TileSum2><<<(unsigned)((nb1 + 127) / 128), 128, 0, st>>>(reinterpret_cast<const TileSum2 *>(d_prefix), n_tiles, d_brow, c->n_batches,
                                                                                        n_rows, total_values, d_off32, d_bv0);
        batch_offsets<<<(unsigned)c->n_batches, 256, 0, st>>>(d_brow, d_bv0, n_rows, total_values, c->batch_rows, d_off32, c->d_offsets);
        ctx->launches.fetch_add(2);
        CUDA_TRY(cudaGetLastError());
        c->batch_v0.resize(nb1);
        CUDA_TRY(cudaMemcpyAsync(c->batch_v0.data(), d_bv0, nb1 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_misc[4];
    CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (const uint32_t we = (uint32_t)(h_misc[0] >> 32)) return wide_fail(we, h_misc[1]);
    if (tile_wide && h_misc[3] != n_qslow) return fail(EXON_GPU_ERR_STATE, "vcf_next_batch: the two passes disagree on the rows with a non-trivial QUAL");
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
    if (tile_wide) {
        static const int kOf[5] = {kIdE, kIdB, kRefB, kFiE, kFiB};
        for (int k = 0; k < 5; ++k)
            for (int64_t b = 0; b < w->n_batches; ++b)
                if (w->base[kOf[k]][(size_t)b + 1] - w->base[kOf[k]][(size_t)b] > 0x7FFFFFFFll)
                    return fail(EXON_GPU_ERR_UNSUPPORTED, "vcf_next_batch: batch %lld overflows int32 offsets", (long long)b);
        if (!w->on_device) {
            WideBuf *b[WideStore::kBufs];
            w->all(b);
            for (int i = 0; i < WideStore::kBufs; ++i) {
                if (!b[i]->d) continue;
                CUDA_TRY(cudaHostAlloc(&b[i]->h, b[i]->bytes, cudaHostAllocDefault));
                CUDA_TRY(cudaMemcpyAsync(b[i]->h, b[i]->d, b[i]->bytes, cudaMemcpyDeviceToHost, st));
            }
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        return EXON_GPU_OK;
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
