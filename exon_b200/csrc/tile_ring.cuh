// tile_ring.cuh -- the warp-private TMA tile pipeline as a reusable device-side object.
//
// Every warp of a persistent grid owns an S-stage ring of TILE-byte tiles in shared memory.  Lane 0 posts 1-D bulk
// async copies (cp.async.bulk -> TMA engine, SASS UBLKCP) of [PRE bytes before | tile | HALO bytes after] and arms
// an mbarrier with the byte count; all lanes wait on the barrier and then read the staged bytes with LDS.  Tiles
// are numbered launch-wide across all segments of the table (ScanSeg, vcf_scan.cuh) and dealt round-robin to
// warps.  No block-wide barrier is involved.  K1/K2 (vcf_scan.cu, vcf_columns.cu) carry an inlined copy of this
// loop; the FASTQ and BAM kernels use this header.
#pragma once
#include "common.cuh"
#include "vcf_scan.cuh"

namespace exon {

template <int TILE_, int S_, int WARPS_, int PRE_, int HALO_, int EXTRA_PER_WARP_>
struct TileRing {
    static constexpr int TILE = TILE_, S = S_, WARPS = WARPS_, PRE = PRE_, HALO = HALO_;
    static constexpr int STAGE = ((PRE + TILE + HALO + 127) / 128) * 128;
    struct Meta {
        const uint8_t *g;  // global address of tile byte 0
        int lo;            // >= 0: first tile of its segment, first valid tile-relative index; else -PRE
        int hi;            // bytes from tile byte 0 to the end of the segment (clamped to 2^30)
        int seg;           // index of the tile's segment in the table
    };
    static constexpr size_t o_ring = 0;
    static constexpr size_t o_bars = (size_t)WARPS * S * STAGE;
    static constexpr size_t o_meta = o_bars + (size_t)WARPS * S * sizeof(uint64_t);
    static constexpr size_t o_extra = o_meta + (size_t)WARPS * S * sizeof(Meta);
    static constexpr size_t smem_bytes = o_extra + (size_t)WARPS * EXTRA_PER_WARP_;

    // what the consumer sees of the current tile
    struct View {
        const uint8_t *sm;  // shared-memory address of tile byte 0
        uint32_t sa;        // the same as a 32-bit shared-window address
        const uint8_t *g;   // global address of tile byte 0
        int seg_lo;         // smallest tile-relative index inside the segment (-2^30 when the tile is not the first)
        int hi;             // one past the largest tile-relative index inside the segment
        int sm_lo, sm_hi;   // tile-relative index range present in shared memory
        bool first;         // first tile of its segment
        int seg;            // index of the tile's segment
        bool interior;      // every staged byte (PRE excluded for first tiles) is segment data and tile byte 0 is a segment byte
    };

    uint8_t *ring;
    uint64_t *bars;
    Meta *meta;
    uint8_t *extra;  // EXTRA_PER_WARP_ bytes of per-warp scratch (queues)
    const ScanSeg *segs;
    int64_t n_tiles, nw, wg;
    int pc;
    int64_t p_tile0, p_next0;
    int s;
    uint32_t parity;
    int lane;

    __device__ __forceinline__ void init(uint8_t *smem_raw, const ScanSeg *segs_, int64_t n_tiles_) {
        const int warp = threadIdx.x >> 5;
        lane = threadIdx.x & 31;
        ring = smem_raw + o_ring + (size_t)warp * (S * STAGE);
        bars = reinterpret_cast<uint64_t *>(smem_raw + o_bars) + warp * S;
        meta = reinterpret_cast<Meta *>(smem_raw + o_meta) + warp * S;
        extra = smem_raw + o_extra + (size_t)warp * EXTRA_PER_WARP_;
        segs = segs_;
        n_tiles = n_tiles_;
        nw = (int64_t)gridDim.x * WARPS;
        wg = (int64_t)blockIdx.x * WARPS + warp;
        pc = 0;
        s = 0;
        parity = 0;
        p_tile0 = p_next0 = 0;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < S; ++i) mbar_init(&bars[i], 1);
            mbar_fence_init();
            p_tile0 = __ldg(&segs[0].tile0);
            p_next0 = __ldg(&segs[1].tile0);
#pragma unroll 1
            for (int i = 0; i < S; ++i) {
                const int64_t T = wg + i * nw;
                if (T < n_tiles) issue(T, i);
            }
        }
        __syncwarp();
    }

    __device__ __forceinline__ void issue(int64_t T, int st) {  // lane 0 only
        while (T >= p_next0) {
            ++pc;
            p_tile0 = p_next0;
            p_next0 = __ldg(&segs[pc + 1].tile0);
        }
        const uint8_t *base = segs[pc].base;
        const int skip = __ldg(&segs[pc].skip);
        const int64_t off = (T - p_tile0) * TILE;
        const int64_t rem = skip + __ldg(&segs[pc].len) - off;
        const int pre = off ? PRE : 0;
        const int64_t body = (rem + 15) & ~(int64_t)15;
        const uint32_t bytes = (uint32_t)(body < TILE + HALO ? body : TILE + HALO) + pre;
        meta[st].g = base + off;
        meta[st].lo = off ? -PRE : skip;
        meta[st].hi = rem > (1 << 30) ? (1 << 30) : (int)rem;
        meta[st].seg = pc;
        mbar_arrive_expect_tx(&bars[st], bytes);
        bulk_g2s(ring + st * STAGE + (PRE - pre), base + off - pre, bytes, &bars[st]);
    }

    // first tile of this warp; tiles advance by nw
    __device__ __forceinline__ int64_t first_tile() const { return wg; }

    __device__ __forceinline__ View acquire() {
        mbar_wait(&bars[s], parity);  // lane 0 wrote meta[s] before it armed the barrier
        View v;
        v.sm = ring + s * STAGE + PRE;
        v.sa = smem_u32(v.sm);
        v.g = meta[s].g;
        const int lo = meta[s].lo;
        v.hi = meta[s].hi;
        v.seg = meta[s].seg;
        v.first = lo >= 0;
        v.seg_lo = v.first ? lo : -(1 << 30);
        v.sm_lo = v.first ? 0 : -PRE;
        v.sm_hi = v.hi < TILE + HALO ? ((v.hi + 15) & ~15) : TILE + HALO;
        v.interior = v.hi >= TILE + HALO && lo <= 0;
        return v;
    }

    // the consumer is done with the current tile T: refill its stage with tile T + S * nw
    __device__ __forceinline__ void release(int64_t T) {
        __syncwarp();
        if (lane == 0) {
            const int64_t Tn = T + (int64_t)S * nw;
            if (Tn < n_tiles) issue(Tn, s);
        }
        if (++s == S) {
            s = 0;
            parity ^= 1;
        }
    }
};

// Byte at tile-relative index i of a view: '\n' outside the segment, shared memory inside the staged window,
// global memory otherwise.
template <class V>
__device__ __forceinline__ uint32_t view_byte(const V &v, int i) {
    if (i < v.seg_lo || i >= v.hi) return '\n';
    if (i >= v.sm_lo && i < v.sm_hi) return v.sm[i];
    return __ldg(v.g + i);
}

// 0x80 flags in up to 16 bytes -> 16-bit mask, bit i = byte i flagged
__device__ __forceinline__ uint32_t pack_flags16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    uint32_t a = __dp4a(f0, 0x08040201u, 0u);
    a = __dp4a(f1, 0x80402010u, a);
    uint32_t b = __dp4a(f2, 0x08040201u, 0u);
    b = __dp4a(f3, 0x80402010u, b);
    return (a >> 7) | (b << 1);
}

// 16-bit mask of the '\n' bytes in one staged 16-byte chunk
__device__ __forceinline__ uint32_t newline_mask16(const uint4 w) {
    return pack_flags16(zero_bytes_exact(w.x ^ kNL4), zero_bytes_exact(w.y ^ kNL4), zero_bytes_exact(w.z ^ kNL4),
                        zero_bytes_exact(w.w ^ kNL4));
}

// restricts a chunk's newline mask to the '\n' that start a line inside the segment: tile index p with
// p >= seg_lo and p + 1 < hi (chunk starts at tile index c0)
__device__ __forceinline__ uint32_t clip_mask16(uint32_t m, int c0, int seg_lo, int hi) {
    const int j_lo = seg_lo - c0 > 0 ? seg_lo - c0 : 0;
    const int j_hi = hi - 1 - c0 < 16 ? hi - 1 - c0 : 16;
    return (j_hi > j_lo) ? (m & ((1u << j_hi) - 1u) & ~((1u << j_lo) - 1u)) : 0u;
}

}  // namespace exon
