// vcf_scan.cu -- K1: fused VCF text -> (chrom, pos) predicate -> COUNT, one pass over the body bytes.
//
// Replaces, for the COUNT query, the whole chain
//   AsyncBatchStream::read_batch        exon/exon-vcf/src/async_batch_stream.rs:80-109
//   LazyVCFArrayBuilder::append col 0/1 exon/exon-vcf/src/array_builder/lazy_array_builder.rs:157-168
//   FilterExec (eq / gt_eq / lt_eq / and_kleene) + AggregateExec(Partial) count   (DataFusion 44, third party)
// without materialising columns: each body byte is read from HBM exactly once.
//
// Execution model (B200): persistent grid, every WARP runs its own S-stage pipeline.  Lane 0 posts 1-D bulk
// async copies (cp.async.bulk -> TMA engine, SASS UBLKCP) of [16 B pre-halo | tile | 48 B post-halo] into the
// warp's shared-memory ring and arms an mbarrier with the byte count; all lanes wait on the barrier, read
// their 16-byte chunks with conflict-free LDS.128, and test them with SWAR integer ops.  No block-wide
// barrier exists anywhere in the kernel, so a warp that falls into the (rare) slow path never stalls its
// neighbours.  Tiles are numbered launch-wide across all resident segments (shard bodies) and dealt
// round-robin to warps, so one launch covers a whole partition's file group.
//
// Fast path: a record can only satisfy `chrom = lit` if the bytes "\n" lit "\t" occur in the text, so the
// hot loop is a multi-byte pattern search, not a line parser:
//   chrom of 1 byte  -> 3-byte pattern: z = (w ^ NL) | (w>>8 ^ C) | (w>>16 ^ TAB) per 32-bit word (byte-shifted
//                       windows via funnel shifts); a zero byte in z marks a hit (8 integer ops per word);
//   chrom >= 2 bytes -> the last 4 pattern bytes are compared as one 32-bit window per byte position
//                       (SHF + ISETP), the preceding bytes are verified only on a hit.
// Slow path (hit): verify the full pattern, parse POS digits (Rust usize::from_str rules), range-test, count.
// Dense modes (no chrom literal, or strict validation) visit every line start instead.
#include "vcf_scan.cuh"

#include "common.cuh"

namespace exon {

namespace {

constexpr int kPre = 16;   // bytes staged before the tile (pattern bytes that precede the anchor)
constexpr int kHalo = 48;  // bytes staged after the tile (window overhang + POS digits)

struct TileView {
    const uint8_t *sm;  // shared-memory address of tile byte 0
    const uint8_t *g;   // global address of tile byte 0
    int lo;             // smallest tile-relative index inside the segment (<= 0)
    int hi;             // one past the largest (> 0)
    int sm_lo, sm_hi;   // tile-relative index range present in shared memory
};

// Byte at tile-relative index i.  Outside the segment reads as '\n' (a record can neither start before the
// segment nor continue past its end); outside the staged window falls back to a global load.
__device__ __forceinline__ uint32_t ld_byte(const TileView &t, int i) {
    if (i < t.lo || i >= t.hi) return '\n';
    if (i >= t.sm_lo && i < t.sm_hi) return t.sm[i];
    return __ldg(t.g + i);
}

// POS field starting at index d: Rust `usize::from_str` (optional '+', >= 1 digit, no other bytes) terminated
// by '\t'; 0 is rejected (noodles maps "0" to None and the column is non-nullable).  Returns the predicate
// value; malformed input raises an error bit and yields 0.
__device__ __forceinline__ uint32_t test_pos(const TileView &t, int d, const ScanArgs &a, uint32_t &err) {
    uint32_t c = ld_byte(t, d);
    if (c == '+') c = ld_byte(t, ++d);
    if (c - '0' > 9u) {
        err |= (c == '\n') ? kErrShortLine : kErrBadPos;
        return 0;
    }
    unsigned long long v = 0;
    int sig = 0;
    bool ovf = false;
    do {
        const uint32_t dg = c - '0';
        if (v | dg) ++sig;
        if (sig > 19) ovf = true;
        v = v * 10ull + dg;
        c = ld_byte(t, ++d);
    } while (c - '0' <= 9u);
    if (c != '\t') {
        err |= (c == '\n') ? kErrShortLine : kErrBadPos;
        return 0;
    }
    if (ovf || v == 0ull || v > 0x7FFFFFFFFFFFFFFFull) {
        err |= kErrBadPos;
        return 0;
    }
    const long long p = (long long)v;
    return (p >= a.lo) & (p <= a.hi);
}

// A candidate whose pattern ('\n' + chrom + '\t') starts at index q (the '\n').
__device__ __forceinline__ uint32_t candidate(const TileView &t, int q, const ScanArgs &a, uint32_t &err) {
    if (q < t.lo) return 0;  // the segment's first line has no real '\n' before it; handled by first_line()
    for (int j = 0; j < a.pat_len; ++j)
        if (ld_byte(t, q + j) != a.pat[j]) return 0;
    if (!a.has_interval) return 1;
    return test_pos(t, q + a.pat_len, a, err);
}

// Full treatment of the line starting at index ls: chrom compare (if any), POS validation, predicate.
__device__ __forceinline__ uint32_t dense_line(const TileView &t, int ls, const ScanArgs &a, uint32_t &err) {
    bool chrom_ok = true;
    int d = ls;
    if (a.has_chrom) {
        for (int j = 0; j <= a.chrom_len; ++j)
            if (ld_byte(t, ls + j) != a.pat[1 + j]) { chrom_ok = false; break; }
        if (chrom_ok) d = ls + a.chrom_len + 1;
    }
    if (!chrom_ok || !a.has_chrom) {
        uint32_t c;
        while ((c = ld_byte(t, d)) != '\t') {
            if (c == '\n') { err |= kErrShortLine; return 0; }
            ++d;
        }
        if (d == ls) { err |= kErrShortLine; return 0; }  // empty CHROM
        ++d;
    }
    const uint32_t r = test_pos(t, d, a, err);
    return chrom_ok ? (a.has_interval ? r : 1u) : 0u;
}

// First line of a segment in the key modes (validated only if its chrom matches, like every other line there).
__device__ __forceinline__ uint32_t first_line(const TileView &t, int ls, const ScanArgs &a, uint32_t &err) {
    for (int j = 0; j <= a.chrom_len; ++j)
        if (ld_byte(t, ls + j) != a.pat[1 + j]) return 0;
    if (!a.has_interval) return 1;
    return test_pos(t, ls + a.chrom_len + 1, a, err);
}

// ---- per-chunk tests -----------------------------------------------------------------------------------
// Fast test of one 16-byte chunk (words w.x..w.w plus the following word w4): non-zero iff the chunk MAY
// contain a pattern anchor.  Fully unrolled, registers only.
template <int MODE>
__device__ __forceinline__ uint32_t chunk_may_hit(const uint4 w, const uint32_t w4, const uint32_t key,
                                                  const uint32_t c4) {
    const uint32_t ws[5] = {w.x, w.y, w.z, w.w, w4};
    if (MODE == kScanKey3) {
        uint32_t h = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t z = (ws[k] ^ kNL4) | (__funnelshift_r(ws[k], ws[k + 1], 8) ^ c4) |
                               (__funnelshift_r(ws[k], ws[k + 1], 16) ^ kTAB4);
            h |= (z - 0x01010101u) & ~z;
        }
        return h & 0x80808080u;
    } else {
        bool hit = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hit |= (ws[k] == key);
            hit |= (__funnelshift_r(ws[k], ws[k + 1], 8) == key);
            hit |= (__funnelshift_r(ws[k], ws[k + 1], 16) == key);
            hit |= (__funnelshift_r(ws[k], ws[k + 1], 24) == key);
        }
        return hit;
    }
}

// Exact treatment of one chunk (cold for the key modes; the whole story for the dense modes).  Re-reads the
// chunk from shared memory so the hot loop keeps no indexable register array alive.
// Returns count | (error bits << 32).
template <int MODE>
__device__ __noinline__ unsigned long long chunk_exact(const uint8_t *sm, const uint8_t *g, int lo, int hi, int sm_lo,
                                                       int sm_hi, int c0, const ScanArgs *ap, uint32_t key,
                                                       uint32_t c4) {
    const ScanArgs &a = *ap;
    TileView t{sm, g, lo, hi, sm_lo, sm_hi};
    uint32_t cnt = 0, err = 0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(sm + c0 + 4 * k);
        const uint32_t w1 = *reinterpret_cast<const uint32_t *>(sm + c0 + 4 * k + 4);
        if (MODE == kScanKey3) {
            const uint32_t z = (w0 ^ kNL4) | (__funnelshift_r(w0, w1, 8) ^ c4) | (__funnelshift_r(w0, w1, 16) ^ kTAB4);
            uint32_t f = zero_bytes_exact(z);
            while (f) {
                const int j = (__ffs(f) - 1) >> 3;
                f &= f - 1;
                cnt += candidate(t, c0 + 4 * k + j, a, err);
            }
        } else if (MODE == kScanKey4) {
            const int back = a.pat_len - 4;  // pattern bytes that precede the 4-byte key
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t x = b ? __funnelshift_r(w0, w1, 8 * b) : w0;
                if (x == key) cnt += candidate(t, c0 + 4 * k + b - back, a, err);
            }
        } else if (MODE == kScanDense) {
            uint32_t f = zero_bytes_exact(w0 ^ kNL4);
            while (f) {
                const int j = (__ffs(f) - 1) >> 3;
                f &= f - 1;
                const int ls = c0 + 4 * k + j + 1;
                if (ls > lo && ls < hi) cnt += dense_line(t, ls, a, err);
            }
        } else {  // kScanLines, ragged end of a segment: a '\n' that is the last byte starts no line
            uint32_t f = zero_bytes_exact(w0 ^ kNL4);
            while (f) {
                const int j = (__ffs(f) - 1) >> 3;
                f &= f - 1;
                const int ls = c0 + 4 * k + j + 1;
                cnt += (ls > lo && ls < hi);
            }
        }
    }
    return (unsigned long long)cnt | ((unsigned long long)err << 32);
}

template <int MODE>
__device__ __noinline__ unsigned long long first_line_exact(const uint8_t *sm, const uint8_t *g, int lo, int hi,
                                                            int sm_hi, const ScanArgs *ap) {
    TileView t{sm, g, lo, hi, 0, sm_hi};
    uint32_t err = 0, cnt;
    if (MODE == kScanDense) cnt = dense_line(t, lo, *ap, err);
    else cnt = first_line(t, lo, *ap, err);
    return (unsigned long long)cnt | ((unsigned long long)err << 32);
}

template <int MODE, int U, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) vcf_scan_kernel(const __grid_constant__ ScanArgs a) {
    constexpr int TILE = 512 * U;
    constexpr int STAGE = ((kPre + TILE + kHalo + 127) / 128) * 128;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *ring = smem_raw + (size_t)warp * (S * STAGE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)WARPS * S * STAGE) + warp * S;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int64_t nw = (int64_t)gridDim.x * WARPS;
    const int64_t wg = (int64_t)blockIdx.x * WARPS + warp;

    // pattern constants
    uint32_t key = 0, c4 = 0;
    if (MODE == kScanKey3) c4 = 0x01010101u * a.pat[1];
    if (MODE == kScanKey4)
        key = (uint32_t)a.pat[a.pat_len - 4] | ((uint32_t)a.pat[a.pat_len - 3] << 8) |
              ((uint32_t)a.pat[a.pat_len - 2] << 16) | ((uint32_t)a.pat[a.pat_len - 1] << 24);

    int pc = 0;  // producer's segment cursor
    auto issue = [&](int64_t T, int s) {
        while (T >= __ldg(&a.segs[pc + 1].tile0)) ++pc;
        if (lane == 0) {
            const uint8_t *base = a.segs[pc].base;
            const int64_t off = (T - __ldg(&a.segs[pc].tile0)) * TILE;
            const int64_t rem = __ldg(&a.segs[pc].skip) + __ldg(&a.segs[pc].len) - off;
            const int pre = off ? kPre : 0;
            const int64_t body = (rem + 15) & ~(int64_t)15;
            const uint32_t bytes = (uint32_t)(body < TILE + kHalo ? body : TILE + kHalo) + pre;
            mbar_arrive_expect_tx(&bars[s], bytes);
            bulk_g2s(ring + s * STAGE + (kPre - pre), base + off - pre, bytes, &bars[s]);
        }
    };

#pragma unroll 1
    for (int s = 0; s < S; ++s) {
        const int64_t T = wg + s * nw;
        if (T < a.n_tiles) issue(T, s);
    }

    uint32_t cnt = 0, err = 0;
    int cc = 0;  // consumer's segment cursor
    int s = 0;
    uint32_t parity = 0;
#pragma unroll 1
    for (int64_t T = wg; T < a.n_tiles; T += nw) {
        while (T >= __ldg(&a.segs[cc + 1].tile0)) ++cc;
        const int64_t off = (T - __ldg(&a.segs[cc].tile0)) * TILE;
        const int skip = __ldg(&a.segs[cc].skip);
        const int64_t rem = skip + __ldg(&a.segs[cc].len) - off;
        const uint8_t *sm = ring + s * STAGE + kPre;
        const uint8_t *g = a.segs[cc].base + off;
        const int lo = off ? (off > (1 << 30) ? -(1 << 30) : skip - (int)off) : skip;
        const int hi = rem > (1 << 30) ? (1 << 30) : (int)rem;
        const int sm_lo = off ? -kPre : 0;
        const int64_t body = (rem + 15) & ~(int64_t)15;
        const int sm_hi = (int)(body < TILE + kHalo ? body : TILE + kHalo);

        mbar_wait(&bars[s], parity);

        if (off == 0 && lane == 0 && hi > lo) {
            if (MODE == kScanLines) {
                cnt += 1;
            } else {
                const unsigned long long r = first_line_exact<MODE>(sm, g, lo, hi, sm_hi, &a);
                cnt += (uint32_t)r;
                err |= (uint32_t)(r >> 32);
            }
        }
        const bool full = hi > TILE;  // every chunk (and its successor word) lies inside the segment
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c0 = (u * 32 + lane) * 16;
            if (full || c0 < hi) {
                if (MODE == kScanKey3 || MODE == kScanKey4) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(sm + c0);
                    const uint32_t w4 = *reinterpret_cast<const uint32_t *>(sm + c0 + 16);
                    if (chunk_may_hit<MODE>(w, w4, key, c4)) {
                        const unsigned long long r = chunk_exact<MODE>(sm, g, lo, hi, sm_lo, sm_hi, c0, &a, key, c4);
                        cnt += (uint32_t)r;
                        err |= (uint32_t)(r >> 32);
                    }
                } else if (MODE == kScanLines && full && c0 >= lo) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(sm + c0);
                    cnt += __popc(zero_bytes_exact(w.x ^ kNL4)) + __popc(zero_bytes_exact(w.y ^ kNL4)) +
                           __popc(zero_bytes_exact(w.z ^ kNL4)) + __popc(zero_bytes_exact(w.w ^ kNL4));
                } else {
                    const unsigned long long r = chunk_exact<MODE>(sm, g, lo, hi, sm_lo, sm_hi, c0, &a, key, c4);
                    cnt += (uint32_t)r;
                    err |= (uint32_t)(r >> 32);
                }
            }
        }
        __syncwarp();
        const int64_t Tn = T + (int64_t)S * nw;
        if (Tn < a.n_tiles) issue(Tn, s);
        if (++s == S) { s = 0; parity ^= 1; }
    }

    cnt = warp_sum(cnt);
    err = __reduce_or_sync(0xFFFFFFFFu, err);
    if (lane == 0) {
        if (cnt) atomicAdd(a.out_count, (unsigned long long)cnt);
        if (err) atomicOr(a.out_flags, err);
    }
}

struct Variant {
    const char *name;
    int U, S, W;
};
constexpr Variant kVariants[] = {
    {"u4s4w8", 4, 4, 8}, {"u8s3w8", 8, 3, 8}, {"u8s4w4", 8, 4, 4},
    {"u4s6w8", 4, 6, 8}, {"u2s6w8", 2, 6, 8}, {"u16s3w4", 16, 3, 4},
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

template <int MODE, int U, int S, int W>
cudaError_t launch_one(const ScanArgs &args, int ctas, int sm_count, cudaStream_t stream) {
    constexpr int TILE = 512 * U;
    constexpr int STAGE = ((kPre + TILE + kHalo + 127) / 128) * 128;
    constexpr size_t smem = (size_t)W * S * STAGE + (size_t)W * S * sizeof(uint64_t);
    auto kern = vcf_scan_kernel<MODE, U, S, W>;
    static int occ = 0;
    if (!occ) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, W * 32, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
    }
    int64_t grid = ctas > 0 ? ctas : (int64_t)occ * sm_count;
    const int64_t need = (args.n_tiles + W - 1) / W;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, W * 32, smem, stream>>>(args);
    return cudaGetLastError();
}

template <int U, int S, int W>
cudaError_t launch_mode(const ScanArgs &args, ScanMode mode, int ctas, int sm_count, cudaStream_t stream) {
    switch (mode) {
        case kScanKey3: return launch_one<kScanKey3, U, S, W>(args, ctas, sm_count, stream);
        case kScanKey4: return launch_one<kScanKey4, U, S, W>(args, ctas, sm_count, stream);
        case kScanDense: return launch_one<kScanDense, U, S, W>(args, ctas, sm_count, stream);
        default: return launch_one<kScanLines, U, S, W>(args, ctas, sm_count, stream);
    }
}

}  // namespace

int scan_variant_count() { return kNumVariants; }
const char *scan_variant_name(int v) { return (v >= 0 && v < kNumVariants) ? kVariants[v].name : "?"; }
int scan_tile_bytes(int v) { return 512 * kVariants[(v >= 0 && v < kNumVariants) ? v : 0].U; }

cudaError_t launch_vcf_scan(const ScanArgs &args, ScanMode mode, const ScanConfig &cfg, int sm_count,
                            cudaStream_t stream) {
    if (args.n_tiles <= 0) return cudaSuccess;
    switch (cfg.variant) {
        case 1: return launch_mode<8, 3, 8>(args, mode, cfg.ctas, sm_count, stream);
        case 2: return launch_mode<8, 4, 4>(args, mode, cfg.ctas, sm_count, stream);
        case 3: return launch_mode<4, 6, 8>(args, mode, cfg.ctas, sm_count, stream);
        case 4: return launch_mode<2, 6, 8>(args, mode, cfg.ctas, sm_count, stream);
        case 5: return launch_mode<16, 3, 4>(args, mode, cfg.ctas, sm_count, stream);
        default: return launch_mode<4, 4, 8>(args, mode, cfg.ctas, sm_count, stream);
    }
}

}  // namespace exon
