// inflate.cu -- DEFLATE (RFC 1951) on the device as two kernels (the decompression in front of
// every parser: noodles-bgzf 0.34 `Reader` / async-compression `GzipDecoder`, see bgzf.cu for the call sites).
//
// Huffman decoding is serial per stream and LZ77 copying is wide but ordered; one kernel that does both leaves the
// lanes idle in turn (the round-1 first cut: 16 lanes per member, ~8 active threads per instruction).  Here the two
// halves are separate kernels with the layout each one wants:
//
//   K_dec  inflate_decode_kernel   ONE LANE PER MEMBER, 32 members per warp in lock step.  Each lane runs the bit reader
//          (two 32-bit words + one prefetched, peek = one funnel shift; the stream's next L1 line is requested ahead),
//          builds its own tables in shared memory (lit/len 2^LB x u16, distance 2^7 x u8, plus the canonical limits and
//          offsets of the lengths 8..15: a code longer than the primary table costs two shared loads, a branch-free
//          length search and one dependent local load of the symbol), and writes
//            * every LITERAL byte straight to its final place in the output, and
//            * every MATCH as a 3-byte token (len - 3 | (dist - 1) << 8) INTO THE FIRST THREE BYTES OF THE MATCH'S OWN
//              OUTPUT RANGE (a match is >= 3 bytes, so the token always fits and needs no memory of its own), plus
//              one bit per match start in a bitmap (1 bit per output byte).
//          LB = 8 puts 320 lanes on an SM, LB = 9 decodes faster per lane: chosen per launch (bgzf_inflate_launch).
//   K_copy inflate_copy_kernel     ONE WARP PER MEMBER.  Walks the output in 1 KiB segments kept in a 4 KiB shared-memory
//          ring: loads the segment (literals in place), turns the segment's bitmap words into a list of match
//          positions, requests the lines of every source that lies behind the ring, reads the tokens (one lane per
//          match), copies every match whose source lies wholly before the segment in parallel (one lane per match),
//          executes the remaining (dependent) matches in output order with all 32 lanes -- found 32 at a time by
//          ballot --, and writes the finished segment back with 16-byte stores.  A match that crosses the segment end
//          is continued in the next segment.
//
// Output is bit-exact DEFLATE; ISIZE of every member is checked, CRC32 is not.
#include <algorithm>
#include <cstdlib>

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

constexpr uint32_t kInfErrData = 1u;  // invalid DEFLATE data
constexpr uint32_t kInfErrSize = 2u;  // output does not match ISIZE

// ------------------------------------------------------------------------------------------------------------------
// K_dec
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDecWarps = 2;

__constant__ uint8_t c_clen_order2[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// LSB-first bit reader: `lo` holds the current word, `hi` the next, `nxt` one more (loaded ahead so that its latency
// is off the critical path).  0 <= bp < 32 always, so peek() returns 32 valid bits.
__device__ __forceinline__ void prefetch_line(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct LaneBits {
    const uint32_t *wp;    // next word to load
    const uint32_t *w0;    // aligned word the MEMBER started in (bits_consumed() counts from there, across stored blocks)
    const uint32_t *wend;  // last word that holds payload bytes: nothing beyond it is ever read (zeros are fed instead), so a
                           // corrupt or truncated stream cannot run off the staging buffer; the decoder notices the overrun
                           // through overrun() / bits_consumed()
    uint32_t lo, hi, nxt;
    int bp;
    int mis8;
    __device__ __forceinline__ uint32_t load(const uint32_t *p) const { return p <= wend ? __ldg(p) : 0u; }
    // position the reader at byte p of the same member (p >= the member's first byte)
    __device__ __forceinline__ void seek(const uint8_t *p) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const int mis = (int)(a & 3);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(a - mis);
        lo = load(w);
        hi = load(w + 1);
        nxt = load(w + 2);
        wp = w + 3;
        bp = 8 * mis;
        prefetch_line(w + 32);
        prefetch_line(w + 64);
    }
    __device__ __forceinline__ void init(const uint8_t *p, uint32_t payload_bytes) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const int mis = (int)(a & 3);
        w0 = reinterpret_cast<const uint32_t *>(a - mis);
        mis8 = 8 * mis;
        wend = payload_bytes ? reinterpret_cast<const uint32_t *>((a + payload_bytes - 1) & ~(uintptr_t)3) : w0 - 1;
        seek(p);
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, bp); }
    __device__ __forceinline__ void skip(int n) {  // n <= 32
        bp += n;
        if (bp >= 32) {
            lo = hi;
            hi = nxt;
            nxt = load(wp);
            ++wp;
            bp -= 32;
            // every lane streams its own member: without this, some lane of the warp misses L1 at almost every refill
            if ((reinterpret_cast<uintptr_t>(wp) & 127u) == 0) prefetch_line(wp + 32);
        }
    }
    __device__ __forceinline__ uint32_t take(int n) {  // n <= 16
        const uint32_t v = peek() & ((1u << n) - 1u);
        skip(n);
        return v;
    }
    __device__ __forceinline__ int64_t bits_consumed() const { return ((int64_t)(wp - 3 - w0) << 5) + bp - mis8; }
    // the word being consumed lies wholly beyond the payload: every further bit is a fed zero
    __device__ __forceinline__ bool overrun() const { return wp - 3 > wend; }
    __device__ __forceinline__ const uint8_t *byte_ptr() const {  // only meaningful when bp % 8 == 0
        return reinterpret_cast<const uint8_t *>(wp - 3) + (bp >> 3);
    }
};

// per-lane canonical code description (local memory): used to build the primary table; `sym` also serves codes longer than it
template <int N>
struct Canon {
    uint16_t sym[N];
    uint16_t cnt[16];
};

// Builds cnt / sym from lens[0..n) and fills the primary table tab[0 .. 1 << PB): entry = sym << LS | len, 0 = longer code
// (LS = 4 for the 16-bit tables, 3 for the 8-bit distance table whose lengths are <= 7).  For lengths 8..15 it also leaves,
// in shared memory, lim[L - 8] = (first code of length L + cnt[L]) << (15 - L) -- a 15-bit MSB-first prefix below it has
// length <= L -- and off[L - 8] = index of the first symbol of length L in sym[] - first code of length L.
// false on an over-subscribed code.
template <class E, int LS, class CanonT>
__device__ bool canon_table(const uint8_t *lens, int n, int PB, E *tab, CanonT &C, uint16_t *lim, int16_t *off) {
    uint16_t offs[16], next[16];
#pragma unroll
    for (int l = 0; l < 16; ++l) C.cnt[l] = 0;
    for (int s = 0; s < n; ++s) C.cnt[lens[s]]++;
    C.cnt[0] = 0;
    int left = 1;
    uint32_t code = 0;
    offs[0] = 0;
    offs[1] = 0;
    next[0] = 0;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= C.cnt[l];
        if (left < 0) return false;
        next[l] = (uint16_t)code;
        if (lim && l >= 8) {
            lim[l - 8] = (uint16_t)((code + C.cnt[l]) << (15 - l));
            off[l - 8] = (int16_t)((int)offs[l] - (int)code);
        }
        code = (code + C.cnt[l]) << 1;
        if (l < 15) offs[l + 1] = (uint16_t)(offs[l] + C.cnt[l]);
    }
    uint32_t *t32 = reinterpret_cast<uint32_t *>(tab);
    for (int i = 0; i < (int)((sizeof(E) << PB) / 4); ++i) t32[i] = 0u;
    for (int s = 0; s < n; ++s) {
        const int L = lens[s];
        if (!L) continue;
        C.sym[offs[L]++] = (uint16_t)s;
        const uint32_t c = next[L]++;
        if (L <= PB) {
            const uint32_t rev = __brev(c) >> (32 - L);
            const E e = (E)((s << LS) | L);
            for (uint32_t i = rev; i < (1u << PB); i += (1u << L)) tab[i] = e;
        }
    }
    return true;
}

// The code at bit 0 of `bits` (LSB first) is longer than PB (>= 7) bits.  lim[] is non-decreasing in the length, so the
// length is PB + 1 + the number of lengths whose limit the 15-bit MSB-first prefix has reached: no branches, the limits
// come from shared memory in one go, and a single dependent local load fetches the symbol.
// sym | len << 16, or 0xFFFFFFFF when no code matches.
template <int PB, class CanonT>
__device__ __forceinline__ uint32_t canon_long(uint32_t bits, const uint16_t *lim, const int16_t *off, const CanonT &C) {
    const uint32_t c15 = __brev(bits) >> 17;
    const uint4 L = *reinterpret_cast<const uint4 *>(lim);
    const uint32_t lw[4] = {L.x, L.y, L.z, L.w};
    int len = PB + 1;
#pragma unroll
    for (int l = PB + 1; l <= 15; ++l) {
        const uint32_t v = (lw[(l - 8) >> 1] >> (((l - 8) & 1) * 16)) & 0xFFFFu;
        len += c15 >= v ? 1 : 0;
    }
    if (len > 15) return 0xFFFFFFFFu;
    return (uint32_t)C.sym[(int)off[len - 8] + (int)(c15 >> (15 - len))] | ((uint32_t)len << 16);
}

// per-lane tables in shared memory: 2^LB x u16 + 2^DB x u8 + 64 B of limits / offsets
template <int LB, int DB>
struct LaneTabs {
    uint16_t lit[1 << LB];  // sym << 4 | len
    uint8_t dist[1 << DB];  // sym << 3 | len (DB <= 7)
    __align__(16) uint16_t llim[8];
    __align__(16) uint16_t dlim[8];
    int16_t loff[8];
    int16_t doff[8];
};

template <int LB, int DB>
__global__ void __launch_bounds__(kDecWarps * 32) inflate_decode_kernel(const uint8_t *comp, const BgzfMember *members, int n_members,
                                                                        uint32_t *bitmap, uint32_t *flags, int *first_bad) {
    extern __shared__ __align__(16) uint8_t dec_smem_raw[];
    LaneTabs<LB, DB> &T = reinterpret_cast<LaneTabs<LB, DB> *>(dec_smem_raw)[threadIdx.x];
    const int gt = blockIdx.x * (kDecWarps * 32) + threadIdx.x, nt = gridDim.x * (kDecWarps * 32);
    static_assert(LB >= 7 && LB <= 9 && DB == 7, "limits / offsets in shared memory start at length 8; the 8-bit distance table holds lengths <= 7");
    Canon<288> CL;  // lit/len code of the current block (and, while the header is read, the code-length code)
    Canon<32> CD;   // distance code
    uint8_t lens[320];
#pragma unroll 1
    for (int mi = gt; mi < n_members; mi += nt) {
        const BgzfMember M = members[mi];
        if (M.isize == 0) continue;
        uint8_t *out = reinterpret_cast<uint8_t *>((uintptr_t)M.out_addr);
        const uint32_t isize = M.isize;
        const uint32_t q0 = (uint32_t)(M.out_addr & 15u);
        uint32_t *bm = bitmap + M.bm_off;
        uint32_t bm_wi = 0, bm_w = 0;
        LaneBits br;
        br.init(comp + M.in_off, M.in_len);
        const int64_t in_bits = (int64_t)M.in_len * 8;
        uint32_t pos = 0, err = 0;
        bool last = false;
#pragma unroll 1
        while (!last && !err) {
            const uint32_t hdr = br.peek();
            last = (hdr & 1u) != 0u;
            const int btype = (int)((hdr >> 1) & 3u);
            br.skip(3);
            if (br.bits_consumed() > in_bits || btype == 3) {  // a stream that runs past its payload is corrupt
                err = kInfErrData;
                break;
            }
            if (btype == 0) {
                br.skip((8 - (br.bp & 7)) & 7);
                const uint32_t w = br.peek();
                br.skip(32);
                const uint32_t len = w & 0xFFFFu;
                if ((len ^ (w >> 16)) != 0xFFFFu || pos + len > isize || br.bits_consumed() + (int64_t)len * 8 > in_bits) {
                    err = kInfErrData;
                    break;
                }
                const uint8_t *src = br.byte_ptr();
                for (uint32_t j = 0; j < len; ++j) out[pos + j] = __ldg(src + j);
                pos += len;
                br.seek(src + len);  // bits_consumed() keeps counting from the member's start
                continue;
            }
            if (btype == 1) {
                for (int s = 0; s < 144; ++s) lens[s] = 8;
                for (int s = 144; s < 256; ++s) lens[s] = 9;
                for (int s = 256; s < 280; ++s) lens[s] = 7;
                for (int s = 280; s < 288; ++s) lens[s] = 8;
                for (int s = 0; s < 30; ++s) lens[288 + s] = 5;
                lens[318] = lens[319] = 0;
            } else {
                const uint32_t h = br.peek();
                br.skip(14);
                const int nlit = (int)(h & 31u) + 257, ndist = (int)((h >> 5) & 31u) + 1, ncl = (int)((h >> 10) & 15u) + 4;
                uint8_t cl[19];
#pragma unroll
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < ncl; ++i) cl[c_clen_order2[i]] = (uint8_t)br.take(3);
                // the code-length code (7-bit codes at most) borrows the literal table's place and CL's arrays
                if (nlit > 286 || ndist > 30 || !canon_table<uint16_t, 4>(cl, 19, 7, T.lit, CL, nullptr, nullptr)) {
                    err = kInfErrData;
                    break;
                }
                int i = 0;
                const int total = nlit + ndist;
                while (i < total) {
                    const uint32_t e = T.lit[br.peek() & 127u];
                    if (!e) {
                        err = kInfErrData;
                        break;
                    }
                    br.skip((int)(e & 15u));
                    const int sym = (int)(e >> 4);
                    if (sym < 16) {
                        lens[i++] = (uint8_t)sym;
                        continue;
                    }
                    int rep, val = 0;
                    if (sym == 16) {
                        if (i == 0) {
                            err = kInfErrData;
                            break;
                        }
                        val = lens[i - 1];
                        rep = 3 + (int)br.take(2);
                    } else if (sym == 17) {
                        rep = 3 + (int)br.take(3);
                    } else {
                        rep = 11 + (int)br.take(7);
                    }
                    if (i + rep > total) {
                        err = kInfErrData;
                        break;
                    }
                    while (rep--) lens[i++] = (uint8_t)val;
                }
                if (err) break;
                if (lens[256] == 0) {  // no end-of-block code
                    err = kInfErrData;
                    break;
                }
                // distance lengths follow the literal/length lengths: move them to a fixed place
                uint8_t tmp[32];
                for (int s = 0; s < 32; ++s) tmp[s] = s < ndist ? lens[nlit + s] : (uint8_t)0;
                for (int s = nlit; s < 288; ++s) lens[s] = 0;
                for (int s = 0; s < 32; ++s) lens[288 + s] = tmp[s];
            }
            // an incomplete distance code with a single symbol is legal (RFC 1951 3.2.7); over-subscription is not
            if (!canon_table<uint16_t, 4>(lens, 288, LB, T.lit, CL, T.llim, T.loff) ||
                !canon_table<uint8_t, 3>(lens + 288, 30, DB, T.dist, CD, T.dlim, T.doff)) {
                err = kInfErrData;
                break;
            }
            // ---- symbols ----
#pragma unroll 1
            while (true) {
                if (br.overrun()) {  // the stream ran off its payload (truncated or corrupt member): stop now, not after ISIZE symbols
                    err = kInfErrData;
                    break;
                }
                uint32_t w = br.peek();
                uint32_t e = T.lit[w & ((1u << LB) - 1u)];
                if (!e) {
                    const uint32_t r = canon_long<LB>(w, T.llim, T.loff, CL);
                    if (r == 0xFFFFFFFFu) {
                        err = kInfErrData;
                        break;
                    }
                    e = ((r & 0xFFFFu) << 4) | (r >> 16);
                }
                const int clen = (int)(e & 15u);
                const int sym = (int)(e >> 4);
                if (sym < 256) {
                    br.skip(clen);
                    if (pos >= isize) {
                        err = kInfErrData;
                        break;
                    }
                    out[pos++] = (uint8_t)sym;
                    continue;
                }
                if (sym == 256) {
                    br.skip(clen);
                    break;
                }
                const int ls = sym - 257;
                if (ls >= 29) {
                    err = kInfErrData;
                    break;
                }
                // length base / extra bits in closed form (RFC 1951 3.2.5)
                w >>= clen;
                int lx = ls < 8 ? 0 : (ls - 4) >> 2;
                uint32_t len = ls < 8 ? (uint32_t)(3 + ls) : (uint32_t)(3 + ((4 + (ls & 3)) << lx));
                if (ls == 28) {
                    lx = 0;
                    len = 258;
                }
                len += w & ((1u << lx) - 1u);
                br.skip(clen + lx);  // <= 15 + 5
                w = br.peek();
                e = T.dist[w & ((1u << DB) - 1u)];
                int dlen = (int)(e & 7u), ds = (int)(e >> 3);
                if (!e) {
                    const uint32_t r = canon_long<DB>(w, T.dlim, T.doff, CD);
                    if (r == 0xFFFFFFFFu) {
                        err = kInfErrData;
                        break;
                    }
                    dlen = (int)(r >> 16);
                    ds = (int)(r & 0xFFFFu);
                }
                if (ds >= 30) {
                    err = kInfErrData;
                    break;
                }
                w >>= dlen;
                const int dx = ds < 4 ? 0 : (ds - 2) >> 1;
                const uint32_t dist = (ds < 4 ? (uint32_t)(1 + ds) : (uint32_t)(1 + ((2 + (ds & 1)) << dx))) + (w & ((1u << dx) - 1u));
                br.skip(dlen + dx);  // <= 15 + 13
                if (dist > pos || pos + len > isize) {
                    err = kInfErrData;
                    break;
                }
                // the token goes where the match will be, its start is marked in the bitmap
                const uint32_t tok = (len - 3u) | ((dist - 1u) << 8);
                out[pos] = (uint8_t)tok;
                out[pos + 1] = (uint8_t)(tok >> 8);
                out[pos + 2] = (uint8_t)(tok >> 16);
                const uint32_t q = q0 + pos, wi = q >> 5;
                if (wi != bm_wi) {
                    if (bm_w) bm[bm_wi] = bm_w;
                    bm_wi = wi;
                    bm_w = 0;
                }
                bm_w |= 1u << (q & 31u);
                pos += len;
            }
            if (!err && br.bits_consumed() > in_bits) err = kInfErrData;
        }
        if (bm_w) bm[bm_wi] = bm_w;
        if (!err && pos != isize) err = kInfErrSize;
        if (err) {
            atomicOr(flags, err);
            atomicMin(first_bad, mi);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K_copy
// ------------------------------------------------------------------------------------------------------------------
constexpr int kCopyWarps = 4;
constexpr uint32_t kSeg = 1024;
#ifndef EXON_INF_RING
#define EXON_INF_RING 4096
#endif
constexpr uint32_t kRingBytes = EXON_INF_RING;
constexpr uint32_t kRingMask = kRingBytes - 1;
constexpr int kMaxMatches = 352;  // matches that can start inside one segment (1024 / 3, rounded up)

struct CopySmem {
    __align__(16) uint8_t ring[kRingBytes];  // ring[q & kRingMask] = output byte q (q counted from the member's 16-byte aligned origin)
    uint32_t mtok[kMaxMatches];              // ordered phase: bytes inside the segment | dist << 9; 0 = already copied
    uint16_t mpos[kMaxMatches];              // match start inside the segment
};

__global__ void __launch_bounds__(kCopyWarps * 32) inflate_copy_kernel(const BgzfMember *members, int n_members, const uint32_t *bitmap) {
    __shared__ CopySmem smem_all[kCopyWarps];
    CopySmem &S = smem_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kCopyWarps + (threadIdx.x >> 5), nw = gridDim.x * kCopyWarps;
#pragma unroll 1
    for (int mi = gw; mi < n_members; mi += nw) {
        const BgzfMember M = members[mi];
        if (M.isize == 0) continue;
        const uint32_t q0 = (uint32_t)(M.out_addr & 15u);
        uint8_t *O = reinterpret_cast<uint8_t *>((uintptr_t)(M.out_addr - q0));
        const uint32_t qend = q0 + M.isize;
        const uint32_t *bm = bitmap + M.bm_off;
        const uint32_t nseg = (qend + kSeg - 1) / kSeg;
        uint32_t carry_len = 0, carry_dist = 0;
#pragma unroll 1
        for (uint32_t s = 0; s < nseg; ++s) {
            const uint32_t segq = s * kSeg;
            const uint32_t lq = segq + 32u * (uint32_t)lane;
            // 1. the segment as K_dec left it (literals final, match ranges hold tokens / garbage)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t uq = lq + 16u * h;
                if (uq < qend) *reinterpret_cast<uint4 *>(&S.ring[uq & kRingMask]) = __ldcg(reinterpret_cast<const uint4 *>(O + uq));
            }
            // 2. match starts of the segment
            const uint32_t w = lq < qend ? __ldg(bm + (lq >> 5)) : 0u;
            const int c = __popc(w);
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            const int n_match = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (n_match == 0 && carry_len == 0) {
                __syncwarp();
                continue;  // literals only: already final in global memory, the ring keeps them as history
            }
            {
                uint32_t ww = w;
                int i = incl - c;
                while (ww) {
                    S.mpos[i++] = (uint16_t)(32 * lane + __ffs((int)ww) - 1);
                    ww &= ww - 1u;
                }
            }
            __syncwarp();
            // 3. tokens, one lane per match; matches whose source ends before the segment are copied right away.
            // Sources behind the ring come from global memory: their lines are requested for every match of the lane before
            // the first copy starts, so that the L2 round trips overlap instead of being paid one match after the other.
            // (Those bytes were written back at least three segments ago, by this warp, with st.cg; the loads below may use
            // the L1: no line they touch is written again.)
            const uint32_t ring_lo = segq > (kRingBytes - kSeg) ? segq - (kRingBytes - kSeg) : 0u;
            if (ring_lo > 0u) {
                for (int k = lane; k < n_match; k += 32) {
                    const uint32_t p = S.mpos[k], q = segq + p;
                    if (p + 2 >= kSeg) continue;
                    const uint32_t len = (uint32_t)S.ring[q & kRingMask] + 3u;
                    const uint32_t dist = ((uint32_t)S.ring[(q + 1) & kRingMask] | ((uint32_t)S.ring[(q + 2) & kRingMask] << 8)) + 1u;
                    const uint32_t src = q - dist;
                    if (src < ring_lo && dist <= q) {
                        prefetch_line(O + src);
                        prefetch_line(O + src + min(len, kSeg - p) - 1u);
                    }
                }
            }
            uint32_t spill_len = 0, spill_dist = 0;
            for (int k = lane; k < n_match; k += 32) {
                const uint32_t p = S.mpos[k], q = segq + p;
                uint32_t b0, b1, b2;
                if (p + 2 < kSeg) {
                    b0 = S.ring[q & kRingMask];
                    b1 = S.ring[(q + 1) & kRingMask];
                    b2 = S.ring[(q + 2) & kRingMask];
                } else {
                    b0 = __ldcg(O + q);
                    b1 = __ldcg(O + q + 1);
                    b2 = __ldcg(O + q + 2);
                }
                const uint32_t len = b0 + 3u, dist = (b1 | (b2 << 8)) + 1u;
                const uint32_t lseg = min(len, kSeg - p);
                if (len > lseg) {
                    spill_len = len - lseg;
                    spill_dist = dist;
                }
                const uint32_t src = q - dist;
                if (src + lseg <= segq) {
                    for (uint32_t j = 0; j < lseg; ++j) {
                        const uint32_t sq = src + j;
                        const uint8_t b = sq >= ring_lo ? S.ring[sq & kRingMask] : O[sq];
                        S.ring[(q + j) & kRingMask] = b;
                    }
                    S.mtok[k] = 0u;
                } else {
                    S.mtok[k] = lseg | (dist << 9);
                }
            }
            __syncwarp();
            // 4. the rest in output order, all lanes on one match
            if (carry_len) {
                const uint32_t src = segq - carry_dist;
                for (uint32_t j = lane; j < carry_len; j += 32) {  // the only ordered copy whose source may lie behind the ring
                    const uint32_t sq = src + (carry_dist >= carry_len ? j : j % carry_dist);
                    S.ring[(segq + j) & kRingMask] = sq >= ring_lo ? S.ring[sq & kRingMask] : __ldcg(O + sq);
                }
                __syncwarp();
            }
            // 32 matches at a time: every lane looks at one, the warp then walks only those that still have to be copied
#pragma unroll 1
            for (int k0 = 0; k0 < n_match; k0 += 32) {
                const int k = k0 + lane;
                const uint32_t t_l = k < n_match ? S.mtok[k] : 0u, p_l = k < n_match ? (uint32_t)S.mpos[k] : 0u;
                uint32_t pend = __ballot_sync(0xFFFFFFFFu, t_l != 0u);
                while (pend) {
                    const int b = __ffs((int)pend) - 1;
                    pend &= pend - 1u;
                    const uint32_t t = __shfl_sync(0xFFFFFFFFu, t_l, b), q = segq + __shfl_sync(0xFFFFFFFFu, p_l, b);
                    const uint32_t len = t & 511u, dist = t >> 9, src = q - dist;
                    for (uint32_t j = lane; j < len; j += 32) {
                        const uint32_t sj = dist >= len ? j : j % dist;
                        S.ring[(q + j) & kRingMask] = S.ring[(src + sj) & kRingMask];
                    }
                    __syncwarp();
                }
            }
            // the spill of this segment's last match, if any (at most one lane has it)
            const uint32_t sp = __ballot_sync(0xFFFFFFFFu, spill_len != 0u);
            carry_len = 0;
            if (sp) {
                const int owner = __ffs((int)sp) - 1;
                carry_len = __shfl_sync(0xFFFFFFFFu, spill_len, owner);
                carry_dist = __shfl_sync(0xFFFFFFFFu, spill_dist, owner);
            }
            // 5. the finished segment back to global memory
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t uq = lq + 16u * h;
                if (uq >= qend) continue;
                if (uq >= q0 && uq + 16u <= qend) {
                    __stcg(reinterpret_cast<uint4 *>(O + uq), *reinterpret_cast<const uint4 *>(&S.ring[uq & kRingMask]));
                } else {
                    const uint32_t a = max(uq, q0), b = min(uq + 16u, qend);
                    for (uint32_t x = a; x < b; ++x) __stcg(O + x, S.ring[x & kRingMask]);
                }
            }
            __syncwarp();
        }
    }
}

template <int LB, int DB>
int launch_decode(Ctx *c, const uint8_t *d_comp, const BgzfMember *d_table, int n_members, uint32_t *d_bitmap, uint32_t *d_flags) {
    static int occ = 0;
    constexpr size_t smem = sizeof(LaneTabs<LB, DB>) * kDecWarps * 32;
    if (!occ) {
        CUDA_TRY(cudaFuncSetAttribute(inflate_decode_kernel<LB, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, inflate_decode_kernel<LB, DB>, kDecWarps * 32, smem));
        if (occ < 1) occ = 1;
    }
    const int per_cta = kDecWarps * 32;
    const int grid = std::min((n_members + per_cta - 1) / per_cta, occ * c->sm_count);
    inflate_decode_kernel<LB, DB><<<grid, per_cta, smem, c->stream>>>(d_comp, d_table, n_members, d_bitmap, d_flags, (int *)(d_flags + 1));
    c->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return EXON_GPU_OK;
}

}  // namespace

// Gives every member its place in the match bitmap (1 bit per output byte, counted from the member's 16-byte aligned
// origin, whole 32-bit words per member); returns the number of words.
size_t bgzf_assign_bitmap(BgzfMember *m, size_t n) {
    size_t words = 0;
    for (size_t i = 0; i < n; ++i) {
        m[i].bm_off = (uint32_t)words;
        m[i].pad_ = 0;
        if (m[i].isize) words += ((size_t)(m[i].out_addr & 15u) + m[i].isize + 31) / 32;
    }
    return words;
}

// Enqueues the inflate of `n_members` members (table in device memory; in_off relative to d_comp, out_addr absolute,
// bm_off from bgzf_assign_bitmap) on the context's stream.  d_flags: two words of device scratch, {0, INT_MAX} before
// the launch.
int bgzf_inflate_launch(Ctx *c, const uint8_t *d_comp, const BgzfMember *d_table, int n_members, uint32_t *d_flags, size_t bitmap_words,
                        size_t comp_bytes) {
    if (n_members <= 0) return EXON_GPU_OK;
    const size_t bm_bytes = (bitmap_words + 64) * sizeof(uint32_t);
    if (bm_bytes > c->inf_bitmap_cap) {
        if (c->inf_bitmap) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            CUDA_TRY(cudaFree(c->inf_bitmap));
            c->inf_bitmap = nullptr;
            c->inf_bitmap_cap = 0;
        }
        const size_t cap = bm_bytes + bm_bytes / 4;
        CUDA_TRY(cudaMalloc(&c->inf_bitmap, cap));
        c->inf_bitmap_cap = cap;
    }
    uint32_t *bm = (uint32_t *)c->inf_bitmap;
    CUDA_TRY(cudaMemsetAsync(bm, 0, bm_bytes, c->stream));
    // Table size: 9-bit literal tables decode faster (fewer long-code fallbacks, and 2 CTAs per SM leave the L1 to the
    // input streams) but hold only 128 lanes per SM; 8-bit tables put 320 lanes on an SM.  A launch that fits one wave of
    // the 9-bit kernel, or that is literal-heavy (compressed > 40 % of the output: BAM), takes the 9-bit tables.
    // EXON_GPU_INFLATE_TABLES = 87 | 97 forces one.
    static const int forced = [] {
        const char *e = getenv("EXON_GPU_INFLATE_TABLES");
        return e ? atoi(e) : 0;
    }();
    const bool wide9 = forced ? forced == 97 : (n_members <= 2 * c->sm_count * kDecWarps * 32 || (double)comp_bytes > 0.4 * 32.0 * (double)bitmap_words);
    const int rc = wide9 ? launch_decode<9, 7>(c, d_comp, d_table, n_members, bm, d_flags) : launch_decode<8, 7>(c, d_comp, d_table, n_members, bm, d_flags);
    if (rc) return rc;
    static int occ = 0;
    if (!occ) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, inflate_copy_kernel, kCopyWarps * 32, 0));
        if (occ < 1) occ = 1;
    }
    const int grid = std::min((n_members + kCopyWarps - 1) / kCopyWarps, occ * c->sm_count);
    inflate_copy_kernel<<<grid, kCopyWarps * 32, 0, c->stream>>>(d_table, n_members, bm);
    c->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return EXON_GPU_OK;
}

}  // namespace exon
