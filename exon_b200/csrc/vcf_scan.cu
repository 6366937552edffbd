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
// warp's shared-memory ring and arms an mbarrier with the byte count; all lanes wait on the barrier and read
// their 16-byte chunks with conflict-free LDS.128.  The tile loop has no block-wide barrier; the CTA meets once, at the
// very end, to add its partial to the query accumulator -- and the CTA that arrives last at that accumulator runs the
// query's tail (publish to mapped host memory, exchange with the peer GPUs): one launch per query, nothing after it.
// Tiles are numbered launch-wide across all resident segments (shard bodies) and dealt round-robin to warps,
// so one launch covers a whole partition's file group.
//
// The kernel is instruction-issue bound, not latency bound (ncu: profiles/), so everything on the per-chunk
// path is SWAR on 32-bit words with no per-byte loops:
//   * line starts: exact '\n' flags per word (4 ops), packed into a 16-bit mask with two IDP.4A per 8 bytes;
//   * a line is parsed by ONE lane from a 16-byte window fetched at the line start (5 LDS.32 + 4 funnel
//     shifts): separator flags -> positions of the first two tabs; CHROM compared as masked words; the POS
//     digits are re-fetched right-aligned (12 bytes ending at the second tab), validated and converted with
//     IDP.4A (4 digits per 3 ops); `lo <= pos <= hi` is one 64-bit unsigned compare;
//   * anything unusual (a field that does not fit the window, '+', > 12 digits, POS 0, a short line, a tile
//     that touches a segment boundary) falls back to the byte-exact scalar routines below, which also own
//     all error reporting.
// Lazy modes (a chrom literal, not strict): a record can only satisfy `chrom = lit` if the bytes
// '\n' lit '\t' occur in the text, so chunks are first screened with a 3-byte SWAR pattern test (1-byte names)
// or a 4-byte window compare (longer names); only screened chunks run the line parser, and POS is parsed
// only for rows whose CHROM matches -- the lazy-record behaviour of noodles, taken one step further.
// Dense mode (strict, or no chrom literal) parses and validates every line.
#include "vcf_scan.cuh"

#include <type_traits>

#include "common.cuh"
#include "vcf_tile.cuh"

namespace exon {

namespace {

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

// Full treatment of the line starting at index ls.  LAZY: the row is validated only as far as the predicate
// reads it (CHROM first; POS only if CHROM matches and an interval is asked for).  Otherwise CHROM must be
// non-empty and POS valid on every row.
template <bool LAZY>
__device__ __forceinline__ uint32_t line_scalar(const TileView &t, int ls, const ScanArgs &a, uint32_t &err) {
    bool chrom_ok = true;
    int d = ls;
    if (a.has_chrom) {
        for (int j = 0; j <= a.chrom_len; ++j)
            if (ld_byte(t, ls + j) != a.pat[1 + j]) { chrom_ok = false; break; }
        if (chrom_ok) d = ls + a.chrom_len + 1;
        if (LAZY && !chrom_ok) return 0;
    }
    if (LAZY && !a.has_interval) return 1;
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

template <bool LAZY>
__device__ __noinline__ unsigned long long line_exact(const uint8_t *sm, const uint8_t *g, int lo, int hi, int sm_lo,
                                                      int sm_hi, int ls, const ScanArgs *ap) {
    TileView t{sm, g, lo, hi, sm_lo, sm_hi};
    uint32_t err = 0;
    const uint32_t cnt = line_scalar<LAZY>(t, ls, *ap, err);
    return (unsigned long long)cnt | ((unsigned long long)err << 32);
}

// Careful treatment of one 16-byte chunk of a boundary tile: every '\n' inside the segment whose successor
// byte is also inside the segment starts a line.  Returns count | (error bits << 32).
template <int MODE>
__device__ __noinline__ unsigned long long chunk_careful(const uint8_t *sm, const uint8_t *g, int lo, int hi, int sm_lo,
                                                         int sm_hi, int c0, const ScanArgs *ap) {
    constexpr bool LAZY = (MODE == kScanKey3 || MODE == kScanKey4);
    TileView t{sm, g, lo, hi, sm_lo, sm_hi};
    uint32_t cnt = 0, err = 0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        uint32_t f = zero_bytes_exact(*reinterpret_cast<const uint32_t *>(sm + c0 + 4 * k) ^ kNL4);
        while (f) {
            const int j = (__ffs(f) - 1) >> 3;
            f &= f - 1;
            const int p = c0 + 4 * k + j;  // position of the '\n'
            if (p < lo || p + 1 >= hi) continue;
            if (MODE == kScanLines) cnt += 1;
            else cnt += line_scalar<LAZY>(t, p + 1, *ap, err);
        }
    }
    return (unsigned long long)cnt | ((unsigned long long)err << 32);
}

// ===================================================================================================
// Pipe-aware SWAR helpers.  Measured on B200 (tools/ubench_pipes.cu, profiles/r2_ubench_pipes.txt): LOP3 / SHF / PRMT issue
// on the ALU pipe and IMAD / IMAD.WIDE / IDP.4A on the FMA pipe, each at one warp-instruction per 2 cycles per SM
// sub-partition, and the two pipes overlap.  K1 was ALU-bound (79 % ALU pipe, FMA idle), so byte shifts and the
// "minus 0x01010101" of the zero-byte tests are written as multiplies, and every 3-input boolean is ONE LOP3: its
// constants live in (uniform) registers derived from a value the compiler cannot see through -- as immediates they would
// split each LOP3 in two (one immediate per instruction) and turn the power-of-two multiplies back into ALU shifts.
// ===================================================================================================
constexpr uint32_t kC80 = 0x80808080u;
struct Konst {
    uint32_t one, m24, m16;                            // 1, 2^24, 2^16
    uint32_t nl4, tab4, c80, m01, c30, c76, cfc, c08;  // byte patterns
};
__device__ __forceinline__ Konst make_konst(uint32_t one) {
    Konst C;
    C.one = one;
    C.m24 = one << 24;
    C.m16 = one << 16;
    C.nl4 = one * kNL4;
    C.tab4 = one * kTAB4;
    C.c80 = one * 0x80808080u;
    C.m01 = one * 0xFEFEFEFFu;  // -0x01010101
    C.c30 = one * 0x30303030u;
    C.c76 = one * 0x76767676u;
    C.cfc = one * 0xFCFCFCFCu;
    C.c08 = one * 0x08080808u;
    return C;
}
// one LOP3 with truth table LUT over (a, b, c) = (0xF0, 0xCC, 0xAA)
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
__device__ __forceinline__ uint32_t fma_add(uint32_t a, uint32_t b, uint32_t one) {  // a + b (IMAD)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}
// (lo >> k) | (hi << (32 - k)) for m = 2^(32 - k): IMAD.WIDE (high half = lo >> k) + IMAD (hi * m adds hi << (32 - k))
__device__ __forceinline__ uint32_t fma_funnel(uint32_t lo, uint32_t hi, uint32_t m) {
    uint32_t h32, r;
    asm("{.reg .u64 t; .reg .u32 l; mul.wide.u32 t, %1, %2; mov.b64 {l, %0}, t;}" : "=r"(h32) : "r"(lo), "r"(m));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(m), "r"(h32));
    return r;
}
// 0x80 in every byte lane of w that is '\n' -- exact and borrow-free: ((w ^ NL) | 0x80) - 1 clears a lane's top bit only
// when the lane was 0x00 or 0x80, and ~w rules out 0x80 (the top bit of w ^ NL is the top bit of w).  2 ALU + 1 FMA.
__device__ __forceinline__ uint32_t nl_flags(uint32_t w, const Konst &C) {
    const uint32_t t = fma_add(lop3<0xBE>(w, C.nl4, C.c80), C.m01, C.one);  // ((w ^ nl) | 0x80) - 0x01..
    return lop3<0x02>(t, w, C.c80);                                         // ~t & ~w & 0x80..
}
// conservative zero-byte accumulator: h |= flags of zero bytes of z (the lowest flag of a word is exact, higher ones may be
// borrow artefacts -- fine for a screen and for "first flag" searches); the caller masks with 0x80808080.  1 FMA + 1 ALU.
__device__ __forceinline__ uint32_t zero_acc(uint32_t h, uint32_t z, const Konst &C) { return lop3<0xF4>(h, fma_add(z, C.m01, C.one), z); }
// 16 byte flags (0x80 each) -> mask with bit 7 + i set for byte i (IDP.4A x4 + IMAD: FMA pipe only)
__device__ __forceinline__ uint32_t pack16_7(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    uint32_t a = __dp4a(f0, 0x08040201u, 0u);
    a = __dp4a(f1, 0x80402010u, a);
    uint32_t b = __dp4a(f2, 0x08040201u, 0u);
    b = __dp4a(f3, 0x80402010u, b);
    return b * 256u + a;
}

// ===================================================================================================
// SWAR line parser (every byte of the 28-byte window read from the line start belongs to the segment)
// ===================================================================================================
struct LineConsts {
    uint32_t P[3], M[3];  // ((window ^ P) & M) == 0  <=>  the line starts with chrom + '\t' (chrom of at most 11 bytes); M = 0 without a literal
    uint32_t lo_lo, lo_hi, span_lo, span_hi;  // pos in [lo, lo + span]
    int has_chrom, has_interval;
    int p0;                                   // chrom_len + 1: where POS starts on a line whose CHROM matched
};

// Parses the line whose first byte is tile byte `ls` from a 24-byte window; `sa` is the shared-window address of tile
// byte 0.  Returns the predicate (0/1); sets `slow` when the line needs the scalar routine instead (nothing has been decided
// or reported then): a CHROM of more than 11 bytes, a field that leaves the window, '+', POS with a leading zero or more
// than 12 digits, anything malformed.  Not LAZY (strict streams, or no chrom literal): CHROM must be non-empty and POS a
// positive decimal on EVERY line (what the reference's builder checks while it fills the columns,
// lazy_array_builder.rs:157-168); LAZY: only a line whose CHROM matches is looked at.
// NW = 6: the 24-byte window; NW = 4: the first 16 bytes only (CHROM, POS and both tabs of an ordinary record fit: "chr10",
// nine digits) -- two loads, two shifts and a third of the digit flags less; a line it cannot settle is given to NW = 6.
template <bool LAZY, int NW>
__device__ __forceinline__ uint32_t line_swar(uint32_t sa, int ls, const LineConsts &K, const Konst &C, bool &slow) {
    const uint32_t la = sa + (uint32_t)ls;  // address of the line's first byte
    const uint32_t a0 = la & ~3u;
    const uint32_t sh = (la & 3u) << 3;     // tile byte 0 is 16-byte aligned
    const uint32_t w0 = lds32(a0), w1 = lds32(a0 + 4), w2 = lds32(a0 + 8), w3 = lds32(a0 + 12), w4 = lds32(a0 + 16);
    uint32_t w5 = 0, w6 = 0;
    if (NW == 6) w5 = lds32(a0 + 20), w6 = lds32(a0 + 24);
    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh),
                   v3 = __funnelshift_r(w3, w4, sh);
    uint32_t v4 = 0, v5 = 0;
    if (NW == 6) v4 = __funnelshift_r(w4, w5, sh), v5 = __funnelshift_r(w5, w6, sh);
    const bool chrom_ok = lop3<0xFE>(lop3<0x28>(v0, K.P[0], K.M[0]), lop3<0x28>(v1, K.P[1], K.M[1]), lop3<0x28>(v2, K.P[2], K.M[2])) == 0;
    if (LAZY && !chrom_ok) return 0;
    if (LAZY && !K.has_interval) return 1;
    int p0 = K.p0;
    bool ok = true;
    if (!LAZY) {
        // first separator (bytes 0x08..0x0B) of the first 12 bytes; it must be a tab and must not be byte 0
        const uint32_t e0 = lop3<0x6A>(v0, C.cfc, C.c08), e1 = lop3<0x6A>(v1, C.cfc, C.c08), e2 = lop3<0x6A>(v2, C.cfc, C.c08);
        const uint32_t f0 = lop3<0x20>(fma_add(e0, C.m01, C.one), e0, C.c80), f1 = lop3<0x20>(fma_add(e1, C.m01, C.one), e1, C.c80),
                       f2 = lop3<0x20>(fma_add(e2, C.m01, C.one), e2, C.c80);
        uint32_t sm7 = __dp4a(f0, 0x08040201u, 0u);
        sm7 = __dp4a(f1, 0x80402010u, sm7);
        sm7 = __dp4a(f2, 0x08040201u, 0u) * 256u + sm7;
        const int s1 = __ffs(sm7) - 8;   // -8: no separator in reach
        ok = s1 >= 1 && lds8(la + (uint32_t)s1) == '\t';  // (s1 < 0 reads the staged bytes before the line: harmless)
        p0 = s1 + 1;
    }
    // non-digit flags of the whole window; the first one at or after p0 ends POS
    const uint32_t d0 = v0 ^ C.c30, d1 = v1 ^ C.c30, d2 = v2 ^ C.c30, d3 = v3 ^ C.c30;
    const uint32_t n0 = lop3<0xA8>(fma_add(d0, C.c76, C.one), d0, C.c80), n1 = lop3<0xA8>(fma_add(d1, C.c76, C.one), d1, C.c80),
                   n2 = lop3<0xA8>(fma_add(d2, C.c76, C.one), d2, C.c80), n3 = lop3<0xA8>(fma_add(d3, C.c76, C.one), d3, C.c80);
    uint32_t nm = pack16_7(n0, n1, n2, n3);                       // bit 7 + i: byte i is not a digit
    if (NW == 6) {
        const uint32_t d4 = v4 ^ C.c30, d5 = v5 ^ C.c30;
        const uint32_t n4 = lop3<0xA8>(fma_add(d4, C.c76, C.one), d4, C.c80), n5 = lop3<0xA8>(fma_add(d5, C.c76, C.one), d5, C.c80);
        uint32_t nh = __dp4a(n4, 0x08040201u, 0u);
        nh = __dp4a(n5, 0x80402010u, nh);                         // bits 7..14 for bytes 16..23
        nm = nh * 65536u + nm;
    }
    const uint32_t nd = nm >> (7 + p0);
    const int n = __ffs(nd) - 1;                                  // digits of POS (-1: none in reach)
    const uint32_t c_end = lds8(la + (uint32_t)(p0 + n)), c_first = lds8(la + (uint32_t)p0);
    ok = ok && n >= 1 && n <= 12 && c_end == '\t' && c_first != '0';
    if (!ok) {
        slow = true;  // POS leaves the window, is empty or longer than 12 digits, starts with '+' or '0', is not followed by a tab, ...
        return 0;
    }
    if (!chrom_ok || !K.has_interval) return chrom_ok;  // validated; the value is not needed
    // value: the 12 bytes that end right before the second tab; the first 12 - n of them are not POS
    const uint32_t b = la + (uint32_t)(p0 + n) - 12u;
    const uint32_t b0 = b & ~3u;
    const uint32_t sh2 = (b & 3u) << 3;
    const uint32_t x0 = lds32(b0), x1 = lds32(b0 + 4), x2 = lds32(b0 + 8), x3 = lds32(b0 + 12);
    const uint32_t s = (uint32_t)(12 - n) << 3;
    const uint32_t q0d = (__funnelshift_r(x0, x1, sh2) ^ 0x30303030u) & shl_clamp(0xFFFFFFFFu, s);
    const uint32_t q1d = (__funnelshift_r(x1, x2, sh2) ^ 0x30303030u) & shl_clamp(0xFFFFFFFFu, s > 32u ? s - 32u : 0u);
    const uint32_t q2d = (__funnelshift_r(x2, x3, sh2) ^ 0x30303030u) & shl_clamp(0xFFFFFFFFu, s > 64u ? s - 64u : 0u);
    // 4 digits per word, most significant in byte 0
    const uint32_t q0 = __dp4a(q0d, 0x00010A64u, 0u) * 10u + (q0d >> 24);
    const uint32_t q1 = __dp4a(q1d, 0x00010A64u, 0u) * 10u + (q1d >> 24);
    const uint32_t q2 = __dp4a(q2d, 0x00010A64u, 0u) * 10u + (q2d >> 24);
    const unsigned long long v = (unsigned long long)(q0 * 10000u + q1) * 10000ull + q2;
    const unsigned long long lo = ((unsigned long long)K.lo_hi << 32) | K.lo_lo;
    const unsigned long long span = ((unsigned long long)K.span_hi << 32) | K.span_lo;
    return (v - lo) <= span;
}

// Key4 screen (chrom of 2+ bytes): non-zero iff the 16-byte chunk (words w.x..w.w plus the following word w4) MAY contain
// the start of the last four pattern bytes.
__device__ __forceinline__ uint32_t chunk_may_hit_key4(const uint4 w, const uint32_t w4, const uint32_t key) {
    const uint32_t ws[5] = {w.x, w.y, w.z, w.w, w4};
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

// Key3 screen, stage 1: flags (top bits) where '\n' is followed by the name byte c.  z[] keeps the per-word test values
// for stage 2.  5 + 4 + 4 ALU, 4 + 4 + 4 FMA per 16 bytes.
__device__ __forceinline__ uint32_t key3_stage1(const uint32_t (&ws)[5], uint32_t c4, uint32_t (&z)[4], const Konst &C) {
    uint32_t v[5], h = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = ws[k] ^ c4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        z[k] = lop3<0xBE>(ws[k], C.nl4, fma_funnel(v[k], v[k + 1], C.m24));  // (w ^ '\n') | ((w ^ c) >> 1 byte)
        h = zero_acc(h, z[k], C);
    }
    return h;
}
// stage 2: ... and a tab right behind it (the exact 3-byte pattern, up to borrow artefacts)
__device__ __forceinline__ uint32_t key3_stage2(const uint32_t (&ws)[5], const uint32_t (&z)[4], const Konst &C) {
    uint32_t y[5], h = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) y[k] = ws[k] ^ C.tab4;
#pragma unroll
    for (int k = 0; k < 4; ++k) h = zero_acc(h, z[k] | fma_funnel(y[k], y[k + 1], C.m16), C);
    return h;
}

// ===================================================================================================
// The tail: what the last CTA does after the last partial has been added (fused into K1)
// ===================================================================================================
// Executed by ONE full warp after every partial of the query is visible in *acc.  Publishes the partition's count,
// exchanges it with the other ranks over peer memory and hands {local, flags, global} to the host through mapped pinned
// memory -- one launch per query: no memset before it, no second kernel, no device-to-host copy after it.
//   peer slots (per rank): [2 parities][n ranks] x {value, seq}; parity = seq & 1, so a rank that is one query ahead never
//   overwrites a value a slower peer is still reading (it cannot be two ahead: every query needs everybody's partial).
__device__ __forceinline__ void scan_finalize(ScanAcc *acc, const ScanTail &t) {
    const int lane = threadIdx.x & 31;
    unsigned long long local = 0, flags = 0;
    if (lane == 0) {
        if (t.finalize == 1) {  // leave the accumulator zeroed for the next query on this stream
            local = atomicExch(&acc->count, 0ull);
            flags = atomicExch(&acc->flags, 0ull);
        } else {                // pushdown accumulator: keeps growing with later feeds
            local = atomicAdd(&acc->count, 0ull);
            flags = atomicAdd(&acc->flags, 0ull);
        }
    }
    local = __shfl_sync(0xFFFFFFFFu, local, 0);
    flags = __shfl_sync(0xFFFFFFFFu, flags, 0);
    unsigned long long global = local, xerr = 0;
    if (t.n_ranks > 1) {
        const int n = t.n_ranks, p = lane;
        const size_t par = (size_t)(t.xseq & 1ull);
        long long part = 0;
        bool ok = true;
        if (p < n) {
            volatile unsigned long long *theirs = t.peers[p] + (par * n + t.rank) * 2;
            theirs[0] = local;
            __threadfence_system();
            theirs[1] = t.xseq;
            volatile unsigned long long *mine = t.peers[t.rank] + (par * n + p) * 2;
            const long long t0 = clock64();
            while (mine[1] != t.xseq) {
                if (clock64() - t0 > 240000000000ll) {  // ~2 minutes: a peer died or never launched (ranks may reach their
                    ok = false;                          // first exchange seconds apart); the host reports EXON_GPU_ERR_NCCL
                    break;
                }
            }
            __threadfence_system();
            part = ok ? (long long)mine[0] : 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
        global = (unsigned long long)part;
        xerr = __all_sync(0xFFFFFFFFu, ok) ? 0ull : 1ull;
    }
    if (lane == 0) {
        if (t.device_out) *t.device_out = (long long)local;
        if (t.host_out) {
            volatile unsigned long long *h = t.host_out;
            h[kHostLocal] = local;
            h[kHostFlags] = flags;
            h[kHostGlobal] = global;
            h[kHostXchgErr] = xerr;
            __threadfence_system();
            h[kHostSeq] = t.host_seq;
        }
    }
}

// The tail on its own: queries that launch no scan (nothing resident, an unmatchable literal) or whose scans ran ahead
// of the query (pushdown feeds) still publish / exchange through the same code.
__global__ void __launch_bounds__(32) scan_finish_kernel(ScanAcc *acc, const ScanTail t) { scan_finalize(acc, t); }

// Adds one CTA's partials to the query accumulator.  The CTA that arrives last re-arms the launch-scoped counters (CTA
// ticket, tile ticket) for the next launch on this accumulator and, in a finalising launch, runs the tail.
// Called by warp 0 after a block-wide barrier.
__device__ __forceinline__ void scan_block_epilogue(const ScanArgs &a, unsigned long long cnt, unsigned long long err) {
    const int lane = threadIdx.x & 31;
    bool last = false;
    if (lane == 0) {
        if (cnt) atomicAdd(&a.acc->count, cnt);
        if (err) atomicOr(&a.acc->flags, err);
        __threadfence();
        last = atomicAdd(&a.acc->ticket, 1ull) == (unsigned long long)gridDim.x - 1ull;
        if (last) {
            atomicExch(&a.acc->ticket, 0ull);
            atomicExch(&a.acc->next_tile, 0ull);
        }
    }
    last = __shfl_sync(0xFFFFFFFFu, last, 0);
    if (last && a.tail.finalize) {
        __threadfence();
        scan_finalize(a.acc, a.tail);
    }
}

// ===================================================================================================
// Tile descriptors: one 16-byte record per 4 KiB tile, built on the device from the segment table whenever the table
// changes, so that the scan kernel's producer lane spends a dozen instructions per tile instead of walking the table
// with 64-bit arithmetic (ncu, round 2: ~135 of the ~560 warp-instructions per tile were per-tile bookkeeping).
// ===================================================================================================
__global__ void build_tile_descs_kernel(const ScanSeg *segs, int n_segs, int64_t n_tiles, int tile, TileDesc *out) {
    const int64_t T = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (T >= n_tiles) return;
    int lo = 0, hi = n_segs - 1;  // last segment with tile0 <= T
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].tile0 <= T) lo = mid;
        else hi = mid - 1;
    }
    const ScanSeg sg = segs[lo];
    const int64_t t = T - sg.tile0;
    const int64_t off = t * tile;
    const int64_t rem = sg.skip + sg.len - off;
    TileDesc d;
    d.src = sg.base + off;
    d.hi = rem > (1 << 30) ? (1 << 30) : (int32_t)rem;
    d.lo = (int16_t)(t ? -kPre : sg.skip);
    d.flags = (uint16_t)((rem >= tile + kHalo && t) ? kTileInterior : 0);  // a segment's first tile is never interior: its first line has no '\n' before it
    out[T] = d;
}

// ===================================================================================================
// The kernel
// ===================================================================================================
template <int MODE, int U, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, ctas_per_sm<U, S, WARPS>()) vcf_scan_kernel(const __grid_constant__ ScanArgs a) {
    constexpr bool LAZY = (MODE == kScanKey3 || MODE == kScanKey4);
    using L = SmemLayout<U, S, WARPS>;
    constexpr int TILE = L::TILE, STAGE = L::STAGE;
    constexpr int kWindow = 28;  // bytes line_swar may touch from a line start (24-byte window + alignment slack)
    static_assert(32 * U <= kQueue, "the per-warp line queue holds one entry per 16-byte chunk of a tile");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ unsigned long long s_part[2];  // this CTA's count / error bits (the only block-wide state)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (threadIdx.x == 0) s_part[0] = s_part[1] = 0ull;
    __syncthreads();
    uint8_t *ring = smem_raw + L::ring + (size_t)warp * (S * STAGE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + L::bars) + warp * S;
    TileDesc *meta = reinterpret_cast<TileDesc *>(smem_raw + L::meta) + warp * S;
    static_assert(sizeof(TileDesc) == sizeof(StageMeta), "the ring's meta slots hold tile descriptors");
    const uint32_t ring_sa = smem_u32(ring);
    const uint32_t meta_sa = smem_u32(meta);
    const uint32_t queue_sa = smem_u32(smem_raw + L::queue) + (uint32_t)(warp * kQueue * sizeof(uint16_t));

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const uint32_t nw = gridDim.x * WARPS;
    const uint32_t wg = blockIdx.x * WARPS + warp;

    // ---- per-launch constants ----
    const Konst C = make_konst(a.n_segs >= 0 ? 1u : 0u);  // opaque to the compiler (see Konst)
    uint32_t key = 0, c4 = 0;
    if (MODE == kScanKey3) c4 = 0x01010101u * a.pat[1];
    if (MODE == kScanKey4) key = (uint32_t)a.pat[0] | ((uint32_t)a.pat[1] << 8) | ((uint32_t)a.pat[2] << 16) | ((uint32_t)a.pat[3] << 24);
    LineConsts K;
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint32_t p = 0, m = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = 4 * k + j;
                if (a.has_chrom && i <= a.chrom_len && a.chrom_len <= 11) {
                    p |= (uint32_t)a.pat[1 + i] << (8 * j);
                    m |= 0xFFu << (8 * j);
                }
            }
            K.P[k] = p;
            K.M[k] = m;
        }
        unsigned long long lo = (unsigned long long)(a.lo < 1 ? 1 : a.lo);  // every valid POS is >= 1
        unsigned long long span = (unsigned long long)a.hi - lo;
        if (a.hi < (long long)lo) {  // empty interval: (v - lo) <= span must never hold
            lo = ~0ull;
            span = 0;
        }
        K.lo_lo = (uint32_t)lo;
        K.lo_hi = (uint32_t)(lo >> 32);
        K.span_lo = (uint32_t)span;
        K.span_hi = (uint32_t)(span >> 32);
        K.has_chrom = a.has_chrom;
        K.has_interval = a.has_interval;
        K.p0 = a.chrom_len + 1;
    }
    const bool swar_ok = !a.has_chrom || a.chrom_len <= 11;  // longer names do not fit the window's pattern words

    // ---- producer (lane 0) ----------------------------------------------------------------------------------------------
    // Tiles [0, R * nw) are dealt round-robin (tile wg + r * nw in round r, no communication); the rest -- about an eighth of
    // the launch -- is handed out through an atomic ticket, so that warps on faster SMs (and warps that met no slow tile) take
    // more of the tail instead of waiting for the slowest one.  Everything is fetched ahead: while tile i is parsed, the
    // descriptor of the tile to stage next is already in registers and the ticket after that is in flight.
    const uint32_t n_tiles = (uint32_t)a.n_tiles, R = a.static_rounds, dyn_base = R * nw;
    unsigned int *ticket_ctr = reinterpret_cast<unsigned int *>(&a.acc->next_tile);
    uint32_t p_round = 0;         // static rounds handed out so far
    uint32_t p_next = 0xFFFFFFFFu;  // tile whose descriptor is fetched next (>= n_tiles: none)
    bool pd_valid = false;
    uint4 pd = make_uint4(0, 0, 0, 0);
    auto advance = [&]() {  // lane 0: the tile after p_next
        if (p_round < R) p_next = wg + (p_round++) * nw;
        else p_next = dyn_base + atomicAdd(ticket_ctr, 1u);
    };
    auto fetch = [&]() {  // lane 0: descriptor of tile p_next into pd, then ask for the tile after it
        pd_valid = p_next < n_tiles;
        if (pd_valid) {
            pd = __ldg(reinterpret_cast<const uint4 *>(a.tiles + p_next));
            advance();
        }
    };
    auto issue = [&](int s) {  // lane 0 only: stage the tile described by pd into slot s
        const uint8_t *src = reinterpret_cast<const uint8_t *>(((unsigned long long)pd.y << 32) | pd.x);
        const int hi = (int)pd.z;
        const int pre = (int)(int16_t)(pd.w & 0xFFFFu) < 0 ? kPre : 0;
        const uint32_t body = hi >= TILE + kHalo ? (uint32_t)(TILE + kHalo) : ((uint32_t)hi + 15u) & ~15u;
        const uint32_t bytes = body + pre;
        *reinterpret_cast<uint4 *>(&meta[s]) = pd;
        mbar_arrive_expect_tx(&bars[s], bytes);
        bulk_g2s(ring + s * STAGE + (kPre - pre), src - pre, bytes, &bars[s]);
    };

    int pending = 0;  // slots in flight (warp-uniform)
    if (lane == 0) {
        advance();
        fetch();
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
            if (pd_valid) {
                issue(s);
                ++pending;
                fetch();
            }
        }
    }
    pending = __shfl_sync(0xFFFFFFFFu, pending, 0);

    uint32_t cnt = 0, err = 0, nl128 = 0;
    uint32_t parity = 0;
    int s = 0;

    // One tile through the vector path.  EDGE: the segment begins and / or ends inside the staged window: '\n' flags outside
    // the segment are masked and a line that starts within kWindow bytes of the segment's end takes the scalar routine.
    auto scan_tile = [&](auto edge_tag, const uint8_t *sm, const uint8_t *g, int seg_lo, int hi, int sm_lo, int sm_hi, uint32_t sa) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        const int u_end = (!EDGE || hi >= TILE) ? U : (hi + 511) >> 9;  // 512-byte rows that hold segment bytes
        // 1. every lane scans its chunks and appends the line starts it finds to the warp's queue;
        // 2. the queue is drained one line per lane, so the parser runs with (nearly) all lanes busy.
        int qn = 0;
        auto drain = [&]() {
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < qn; i += 32) {
                const int ls = (int)lds16(queue_sa + 2u * (uint32_t)i);
                bool slow = EDGE && ls + kWindow > hi;  // the window would leave the segment
                if (!slow) {
                    if (LAZY) {
                        cnt += line_swar<LAZY, 6>(sa, ls, K, C, slow);
                    } else {
                        bool wider = false;
                        cnt += line_swar<LAZY, 4>(sa, ls, K, C, wider);
                        if (wider) cnt += line_swar<LAZY, 6>(sa, ls, K, C, slow);
                    }
                }
                if (slow) {
                    const unsigned long long r = line_exact<LAZY>(sm, g, seg_lo, hi, sm_lo, sm_hi, ls, &a);
                    cnt += (uint32_t)r;
                    err |= (uint32_t)(r >> 32);
                }
            }
            __syncwarp();
            qn = 0;
        };
        // line starts of one chunk (m: bit 7 + i = '\n' at byte c0 + i) -> queue; extra starts in the same 16 bytes -> scalar
        auto push = [&](uint32_t m, int c0) {
            if (EDGE) {
                // keep '\n' at positions p with seg_lo <= p and p + 1 < hi
                const int lo_i = seg_lo - c0, hi_i = hi - 1 - c0;  // valid i: lo_i <= i < hi_i
                uint32_t keep = hi_i >= 16 ? 0xFFFFu : (hi_i <= 0 ? 0u : (1u << hi_i) - 1u);
                if (lo_i > 0) keep &= lo_i >= 16 ? 0u : ~((1u << lo_i) - 1u);
                m &= keep << 7;
            }
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, m != 0);
            if (m) {
                sts16(queue_sa + 2u * (uint32_t)(qn + __popc(b & lt_mask)), (uint32_t)(c0 + __ffs(m) - 7));  // line = byte after '\n'
                m &= m - 1;
            }
            qn += __popc(b);
            while (m) {  // lines shorter than 16 bytes
                const unsigned long long r = line_exact<LAZY>(sm, g, seg_lo, hi, sm_lo, sm_hi, c0 + __ffs(m) - 7, &a);
                m &= m - 1;
                cnt += (uint32_t)r;
                err |= (uint32_t)(r >> 32);
            }
            // (no overflow check: a chunk queues at most one line, so a tile queues at most 32 * U <= kQueue lines)
        };
        if (MODE == kScanKey3) {
            // two chunks per step: both screens are computed before the warp votes once (twice the independent work in
            // flight, half the votes); on a hit each chunk goes through stage 2 and the line-start search on its own
#pragma unroll 1
            for (int u = 0; u < u_end; u += 2) {
                const int c0 = (u * 32 + lane) * 16;
                const bool two = u + 1 < u_end;
                const uint4 wa = lds128(sa + (uint32_t)c0);
                const uint4 wb = two ? lds128(sa + (uint32_t)c0 + 512u) : make_uint4(0, 0, 0, 0);
                // the word after each chunk: lane + 1 holds it in a register (the last lane reads it from the window)
                uint32_t wa4 = __shfl_down_sync(0xFFFFFFFFu, wa.x, 1), wb4 = __shfl_down_sync(0xFFFFFFFFu, wb.x, 1);
                if (lane == 31) {
                    wa4 = lds32(sa + (uint32_t)c0 + 16u);
                    wb4 = lds32(sa + (uint32_t)c0 + 528u);
                }
                const uint32_t wsa[5] = {wa.x, wa.y, wa.z, wa.w, wa4}, wsb[5] = {wb.x, wb.y, wb.z, wb.w, wb4};
                uint32_t za[4], zb[4];
                const uint32_t ha = key3_stage1(wsa, c4, za, C), hb = key3_stage1(wsb, c4, zb, C);
                if (__ballot_sync(0xFFFFFFFFu, ((ha | hb) & kC80) != 0) == 0) continue;
                bool look = (ha & kC80) != 0;
                if (__ballot_sync(0xFFFFFFFFu, look)) {
                    look = look && (key3_stage2(wsa, za, C) & kC80) != 0;
                    if (__ballot_sync(0xFFFFFFFFu, look))
                        push(look ? pack16_7(nl_flags(wa.x, C), nl_flags(wa.y, C), nl_flags(wa.z, C), nl_flags(wa.w, C)) : 0u, c0);
                }
                look = two && (hb & kC80) != 0;
                if (__ballot_sync(0xFFFFFFFFu, look)) {
                    look = look && (key3_stage2(wsb, zb, C) & kC80) != 0;
                    if (__ballot_sync(0xFFFFFFFFu, look))
                        push(look ? pack16_7(nl_flags(wb.x, C), nl_flags(wb.y, C), nl_flags(wb.z, C), nl_flags(wb.w, C)) : 0u, c0 + 512);
                }
            }
        } else {
#pragma unroll 2
            for (int u = 0; u < u_end; ++u) {
                const int c0 = (u * 32 + lane) * 16;
                const uint4 w = lds128(sa + (uint32_t)c0);
                bool look = true;
                if (MODE == kScanKey4) {
                    uint32_t w4 = __shfl_down_sync(0xFFFFFFFFu, w.x, 1);
                    if (lane == 31) w4 = lds32(sa + (uint32_t)c0 + 16u);
                    look = chunk_may_hit_key4(w, w4, key) != 0;
                    if (__ballot_sync(0xFFFFFFFFu, look) == 0) continue;
                }
                push(look ? pack16_7(nl_flags(w.x, C), nl_flags(w.y, C), nl_flags(w.z, C), nl_flags(w.w, C)) : 0u, c0);
            }
        }
        if (qn) drain();
    };

#pragma unroll 1
    while (pending > 0) {
        const uint8_t *sm = ring + s * STAGE + kPre;
        mbar_wait(&bars[s], parity);  // lane 0 wrote meta[s] before it armed the barrier
        const uint4 md = lds128(meta_sa + (uint32_t)s * (uint32_t)sizeof(TileDesc));
        const uint32_t sa = ring_sa + (uint32_t)(s * STAGE + kPre);  // shared-window address of tile byte 0
        const bool interior = (md.w >> 16) & kTileInterior;
        if (interior && MODE == kScanLines) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint4 w = lds128(sa + (uint32_t)((u * 32 + lane) * 16));
                nl128 = __dp4a(zero_bytes_exact(w.x ^ kNL4), 0x01010101u, nl128);
                nl128 = __dp4a(zero_bytes_exact(w.y ^ kNL4), 0x01010101u, nl128);
                nl128 = __dp4a(zero_bytes_exact(w.z ^ kNL4), 0x01010101u, nl128);
                nl128 = __dp4a(zero_bytes_exact(w.w ^ kNL4), 0x01010101u, nl128);
            }
        } else if (interior && swar_ok) {
            scan_tile(std::false_type{}, sm, nullptr, -(1 << 30), 1 << 30, -kPre, TILE + kHalo, sa);
        } else {
            const uint8_t *g = reinterpret_cast<const uint8_t *>(((unsigned long long)md.y << 32) | md.x);
            const int lo = (int)(int16_t)(md.w & 0xFFFFu), hi = (int)md.z;
            const bool first = lo >= 0;  // first tile of its segment: line 0 has no '\n' before it
            const int seg_lo = first ? lo : -(1 << 30);
            const int sm_lo = first ? 0 : -kPre;
            const int sm_hi = hi < TILE + kHalo ? ((hi + 15) & ~15) : TILE + kHalo;
            if (first && lane == 0 && hi > lo) {
                if (MODE == kScanLines) {
                    cnt += 1;
                } else {
                    const unsigned long long r = line_exact<LAZY>(sm, g, seg_lo, hi, sm_lo, sm_hi, lo, &a);
                    cnt += (uint32_t)r;
                    err |= (uint32_t)(r >> 32);
                }
            }
            if (swar_ok && MODE != kScanLines) {
                scan_tile(std::true_type{}, sm, g, seg_lo, hi, sm_lo, sm_hi, sa);
            } else {
#pragma unroll 1
                for (int u = 0; u < U; ++u) {
                    const int c0 = (u * 32 + lane) * 16;
                    if (c0 < hi && c0 + 16 > seg_lo) {
                        const unsigned long long r = chunk_careful<MODE>(sm, g, seg_lo, hi, sm_lo, sm_hi, c0, &a);
                        cnt += (uint32_t)r;
                        err |= (uint32_t)(r >> 32);
                    }
                }
            }
        }
        __syncwarp();
        int staged = 0;
        if (lane == 0 && pd_valid) {
            issue(s);
            fetch();
            staged = 1;
        }
        pending += __shfl_sync(0xFFFFFFFFu, staged, 0) - 1;
        if (++s == S) {
            s = 0;
            parity ^= 1;
        }
    }

    if (MODE == kScanLines) cnt += nl128 >> 7;
    cnt = warp_sum(cnt);
    err = __reduce_or_sync(0xFFFFFFFFu, err);
    if (lane == 0) {
        if (cnt) atomicAdd(&s_part[0], (unsigned long long)cnt);
        if (err) atomicOr(&s_part[1], (unsigned long long)err);
    }
    __syncthreads();  // the only barrier after the prologue: every warp has drained its tiles
    if (warp == 0) scan_block_epilogue(a, s_part[0], s_part[1]);
}

struct Variant {
    const char *name;
    int U, S, W;
};
// 0 is the default: 4 KiB tiles, TWO stages per warp, 8 warps -> 3 CTAs (24 warps) per SM.  Round-2 measurements on 100 M rows
// (region query, lazy / strict): u8s2w8 0.566 / 0.80 ms, u6s2w8 (4 CTAs at 64 registers) 0.652 / 0.879, u4s2w8 0.739 / 0.999;
// round 1 (older kernel): u8s3w8 0.611 / 1.17, u8s2w4 0.592 / 1.06, u12s2w8 0.594 / 1.16.  The other entries keep odd
// geometries under test (tile edges, ring depth, warps per CTA).
constexpr Variant kVariants[] = {
    {"u8s2w8", 8, 2, 8}, {"u4s4w8", 4, 4, 8}, {"u8s4w4", 8, 4, 4}, {"u4s3w8", 4, 3, 8}, {"u2s4w8", 2, 4, 8}, {"u8s2w6", 8, 2, 6},
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

template <int MODE, int U, int S, int W>
cudaError_t launch_one(const ScanArgs &args, int ctas, int sm_count, cudaStream_t stream) {
    constexpr size_t smem = SmemLayout<U, S, W>::total;
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
    if (args.n_tiles >= (int64_t)0xF0000000u) return cudaErrorInvalidValue;  // tile numbers are 32-bit (4 KiB tiles: 15 TB per launch)
    ScanArgs a2 = args;
    const int64_t nw = grid * W, rounds = args.n_tiles / nw;  // full rounds of the round-robin deal
    a2.static_rounds = (uint32_t)(rounds - (rounds + 7) / 8);  // the last eighth (at least one round) goes through the ticket
    kern<<<(unsigned)grid, W * 32, smem, stream>>>(a2);
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

cudaError_t launch_build_tile_descs(const ScanSeg *d_segs, int n_segs, int64_t n_tiles, int variant, TileDesc *d_out, cudaStream_t stream) {
    if (n_tiles <= 0) return cudaSuccess;
    const int tile = scan_tile_bytes(variant);
    build_tile_descs_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, stream>>>(d_segs, n_segs, n_tiles, tile, d_out);
    return cudaGetLastError();
}

int scan_variant_count() { return kNumVariants; }
const char *scan_variant_name(int v) { return (v >= 0 && v < kNumVariants) ? kVariants[v].name : "?"; }
int scan_tile_bytes(int v) { return 512 * kVariants[(v >= 0 && v < kNumVariants) ? v : 0].U; }

cudaError_t launch_vcf_scan(const ScanArgs &args, ScanMode mode, const ScanConfig &cfg, int sm_count,
                            cudaStream_t stream) {
    if (args.n_tiles <= 0) {
        if (!args.tail.finalize) return cudaSuccess;
        scan_finish_kernel<<<1, 32, 0, stream>>>(args.acc, args.tail);
        return cudaGetLastError();
    }
    switch (cfg.variant) {
        case 1: return launch_mode<4, 4, 8>(args, mode, cfg.ctas, sm_count, stream);
        case 2: return launch_mode<8, 4, 4>(args, mode, cfg.ctas, sm_count, stream);
        case 3: return launch_mode<4, 3, 8>(args, mode, cfg.ctas, sm_count, stream);
        case 4: return launch_mode<2, 4, 8>(args, mode, cfg.ctas, sm_count, stream);
        case 5: return launch_mode<8, 2, 6>(args, mode, cfg.ctas, sm_count, stream);
        default:
            // COUNT(*) only counts newlines: nothing but the copies' latency to hide, and four stages behind four warps hide
            // more of it than two behind eight (0.44 vs 0.51 ms per 2.75 GB); same 4 KiB tiles, so the tile table is shared
            if (mode == kScanLines) return launch_one<kScanLines, 8, 4, 4>(args, cfg.ctas, sm_count, stream);
            return launch_mode<8, 2, 8>(args, mode, cfg.ctas, sm_count, stream);
    }
}

}  // namespace exon
