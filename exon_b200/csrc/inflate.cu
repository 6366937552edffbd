// inflate.cu -- DEFLATE (RFC 1951) on the device as two kernels (the decompression in front of
// every parser: noodles-bgzf 0.34 `Reader` / async-compression `GzipDecoder`, see bgzf.cu for the call sites).
//
// Huffman decoding is serial per stream and wants tens of thousands of streams in flight; LZ77 copying is wide but ordered
// and wants a stream's recent output on chip.  One kernel that does both starves one half or the other, so the two halves
// are separate kernels joined by a COMPACT TOKEN STREAM in device memory (round 2; round 1 wrote literals and 3-byte tokens
// in place into the output and marked match starts in a bitmap: 16.9 GB of DRAM traffic for 3.3 GB of algorithmic bytes,
// profiles/r1_ncu_full_bgzf_inflate.txt -- every sector of the output was written partially, read back, and written again):
//
//   member token area (bgzf_token_units(isize) 16-byte units, assigned by bgzf_assign_tokens):
//       unit 0              header {n_tokens, n_literals, 0, 0}
//       units 1 ..          literal bytes, in output order, growing upwards
//       .. last unit        32-bit tokens in groups of four, group k in unit (last - k):
//                           literals before the match (0..255) | match length (0 = none, 3..258) << 8 | (distance - 1) << 17
//       (literals + 4 x tokens never meet: a match is worth >= 3 output bytes, so the area holds 4/3 of ISIZE + slack)
//
//   K_dec  inflate_decode_kernel   ONE LANE PER MEMBER, 32 members per warp in lock step.  Each lane pulls its stream through
//          registers in 16-byte loads issued one chunk ahead (the L1 left beside the tables is far smaller than 320 lanes x
//          one line: per-word loads missed it constantly in round 1), keeps a 64-bit bit buffer, builds its own tables in
//          shared memory -- element i of lane l at [i][l], so that the 32 lanes of a lookup fall into different banks
//          -- and collects literals and tokens in registers, sixteen bytes / four tokens per 16-byte store.
//   K_copy inflate_copy_kernel     ONE WARP PER MEMBER with the member's last RB bytes of output in a shared-memory ring.
//          32 tokens at a time (fewer when they span more than kSpan bytes): one packed warp scan gives every lane its output
//          and literal-stream positions; the group's literal bytes are read coalesced and scattered into the ring; matches
//          whose source lies behind the ring come from global memory (written there by this warp at least RB - kSpan bytes
//          ago); the rest resolve in rounds -- a lane copies as soon as its source lies below the first unfinished match of
//          the group, so independent matches go 32 at a time and a chain costs one round per link.  Finished output leaves
//          the ring in 16-byte stores, 512 bytes per warp instruction.
//
// DRAM traffic per launch: compressed in + tokens out (K_dec), tokens in + output out (K_copy).  Output is bit-exact DEFLATE;
// ISIZE of every member is checked, CRC32 is not.
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
constexpr int kDecWarps = 4;

__constant__ uint8_t c_clen_order2[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ void prefetch_line(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// The stream is read 16 bytes at a time, eight times per 128-byte line and microseconds apart: it has to STAY in the L2 in
// between (an evict-first hint -- L1::no_allocate turns into one -- made every one of the eight a DRAM access), and it is of
// no use in the L1, which is far smaller than the lines 288 lanes are reading.
__device__ __forceinline__ uint4 ld_stream16(const uint4 *p) { return __ldcg(p); }
// one whole 32-byte sector per store: a sector written in two halves makes the L2 fetch it from DRAM first (round 1 and the
// first cut of this file read 5x the compressed bytes that way)
__device__ __forceinline__ void st_sector(void *p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]),
                 "r"(w[6]), "r"(w[7])
                 : "memory");
}

constexpr int kInWords = 16;  // a lane's input FIFO in shared memory: four 16-byte chunks

// LSB-first bit reader.  `bb` holds `cnt` valid bits; refill() brings cnt to 33..64 one 32-bit word at a time.  The words
// come out of a small per-lane FIFO in shared memory (word w of the stream at fifo[w % kInWords][lane]); `nxtw` is the next
// one, fetched ahead so that the shared-memory latency is off the refill's critical path.
//
// The FIFO is fed 16 bytes at a time, and the feeding is what the structure is about: a warp's scoreboards do not know
// lanes apart, so a load that some lanes issue in one round makes the next round's read of the same registers wait for it,
// whichever lanes read.  The symbol loop therefore tops the FIFO up at a FIXED place every second round: topup() first
// moves the chunk requested two rounds ago from registers to shared memory, then requests the next one for every lane
// whose level has fallen below eight words.  A round consumes at most 48 bits, so between two top-ups a lane takes at most
// three words and the level never falls below two.  Outside the symbol loop (block headers, stored blocks) fill() loads
// synchronously.
//
// Chunks that lie wholly beyond the payload are never read (zeros are fed instead), so a corrupt or truncated stream cannot
// run off the staging buffer; `spent` says that the reader is certainly past the payload.
struct LaneBits {
    uint32_t *fifo;     // &W.in[0][lane]
    const uint4 *np;    // next chunk to request
    const uint4 *endp;  // the chunk that holds the payload's last byte
    uint4 fl;           // chunk in flight
    uint32_t lo, hi;    // the word being consumed and the one after it
    uint32_t nxtw;      // the one after that, fetched ahead
    uint32_t rd, wr;    // words fetched from / written to the FIFO
    int bp;             // bits of `lo` already consumed; refill() brings it below 32, so that peek() returns 32 valid bits
    int skew;           // bits of the first word that precede the payload
    bool infl, spent;
    __device__ __forceinline__ uint4 load(const uint4 *p) const { return p <= endp ? ld_stream16(p) : make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ void land() {
        if (infl) {
            fifo[((wr + 0u) & (kInWords - 1)) * 32] = fl.x;
            fifo[((wr + 1u) & (kInWords - 1)) * 32] = fl.y;
            fifo[((wr + 2u) & (kInWords - 1)) * 32] = fl.z;
            fifo[((wr + 3u) & (kInWords - 1)) * 32] = fl.w;
            wr += 4u;
            infl = false;
        }
    }
    __device__ __forceinline__ void request() {
        if (wr - rd < 8u) {
            fl = load(np);
            spent = spent || np > endp + 5;  // more than the FIFO and the three words in registers hold together lies beyond the payload
            ++np;
            infl = true;
        }
    }
    __device__ __forceinline__ void topup() {
        land();
        request();
    }
    __device__ __forceinline__ void fill() {  // synchronous: block headers and stored blocks
        land();
        while (wr - rd < 8u) {
            request();
            land();
        }
    }
    __device__ __forceinline__ void next_word() {
        lo = hi;
        hi = nxtw;
        nxtw = fifo[(rd & (kInWords - 1)) * 32];
        ++rd;
    }
    __device__ __forceinline__ void init(uint32_t *lane_fifo, const uint8_t *p, uint32_t payload_bytes) {
        fifo = lane_fifo;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const int mis = (int)(a & 15);
        const uint4 *c0 = reinterpret_cast<const uint4 *>(a - mis);
        endp = payload_bytes ? reinterpret_cast<const uint4 *>((a + payload_bytes - 1) & ~(uintptr_t)15) : c0 - 1;
        np = c0;
        rd = wr = 0u;
        infl = spent = false;
        fill();
        rd = (uint32_t)(mis >> 2);  // the words before the payload are never fetched
        next_word();
        next_word();
        next_word();                // lo = the payload's first word
        skew = 32 * (mis >> 2) + 8 * (mis & 3);
        bp = 8 * (mis & 3);
    }
    // inside the symbol loop (topup() keeps the FIFO ahead): no branch -- a few lanes need a word in any given round, and a
    // divergent branch would cost the warp more than the four predicated instructions do
    __device__ __forceinline__ void refill() {
        const bool need = bp >= 32;
        lo = need ? hi : lo;
        hi = need ? nxtw : hi;
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(fifo) + (rd & (kInWords - 1)) * 128u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u32 %0, [%1];\n\t}" : "+r"(nxtw) : "r"(a), "r"((uint32_t)need));
        rd += need ? 1u : 0u;
        bp -= need ? 32 : 0;
    }
    __device__ __forceinline__ void refill_sync() {  // anywhere else
        if (bp >= 32) {
            if (wr - rd < 2u) fill();
            next_word();
            bp -= 32;
        }
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_r(lo, hi, bp); }  // bp < 32
    __device__ __forceinline__ void skip(int n) { bp += n; }                                   // n <= 32 - (bits used since the last refill)
    __device__ __forceinline__ uint32_t take(int n) {                                          // n <= 16, after a refill
        const uint32_t v = peek() & ((1u << n) - 1u);
        skip(n);
        return v;
    }
    // three words have been fetched beyond the consumed ones (lo, hi, nxtw)
    __device__ __forceinline__ int64_t bits_consumed() const { return (int64_t)(rd - 3u) * 32 + bp - skew; }
};

// per-lane canonical code description (local memory): used to build the primary table; `sym` also serves codes longer than it
template <int N>
struct Canon {
    uint16_t sym[N];
    uint16_t cnt[16];
};

// Builds cnt / sym from lens[0..n) and fills the lane's primary table (element i at tab[i * 32]), 1 << PB entries:
// entry = sym << LS | len, 0 = longer code (LS = 4 for the 16-bit tables, 3 for the 8-bit distance table whose lengths are
// <= 7).  For lengths 7..15 it also leaves lim[L - 7] = (first code of length L + cnt[L]) << (15 - L) -- a 15-bit MSB-first
// prefix below it has length <= L -- and off[L - 7] = index of the first symbol of length L in sym[] - first code of length
// L.  false on an over-subscribed code.
template <class E, int LS, class Enc, class CanonT>
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
        if (lim && l >= 7) {
            lim[l - 7] = (uint16_t)((code + C.cnt[l]) << (15 - l));
            off[l - 7] = (int16_t)((int)offs[l] - (int)code);
        }
        code = (code + C.cnt[l]) << 1;
        if (l < 15) offs[l + 1] = (uint16_t)(offs[l] + C.cnt[l]);
    }
    for (int i = 0; i < (1 << PB); ++i) tab[i * 32] = (E)0;
    for (int s = 0; s < n; ++s) {
        const int L = lens[s];
        if (!L) continue;
        C.sym[offs[L]++] = (uint16_t)s;
        const uint32_t c = next[L]++;
        if (L <= PB) {
            const uint32_t rev = __brev(c) >> (32 - L);
            const E e = (E)((Enc()((uint32_t)s) << LS) | L);
            for (uint32_t i = rev; i < (1u << PB); i += (1u << L)) tab[i * 32] = e;
        }
    }
    return true;
}

// The code at bit 0 of `bits` (LSB first) is longer than PB (>= 6) bits.  lim[] is non-decreasing in the length, so the
// length is PB + 1 + the number of lengths whose limit the 15-bit MSB-first prefix has reached: no branches, and a single
// dependent local load fetches the symbol.  sym | len << 16, or 0xFFFFFFFF when no code matches.
template <int PB, class CanonT>
__device__ __forceinline__ uint32_t canon_long(uint32_t bits, const uint16_t *lim, const int16_t *off, const CanonT &C) {
    const uint32_t c15 = __brev(bits) >> 17;
    int len = PB + 1;
#pragma unroll
    for (int l = PB + 1; l <= 15; ++l) len += c15 >= (uint32_t)lim[l - 7] ? 1 : 0;
    if (len > 15) return 0xFFFFFFFFu;
    return (uint32_t)C.sym[(int)off[len - 7] + (int)(c15 >> (15 - len))] | ((uint32_t)len << 16);
}

// one warp's shared memory: the primary tables, the input FIFOs and the output staging, all interleaved by lane (element i
// of lane l at [i][l]: the 32 lanes of an access fall into different banks -- two lanes per bank word for the 16-bit table,
// four for the 8-bit one).  The long-code limits / offsets are per-lane local memory (Limits): a code longer than the
// primary table is rare, and the 2 KiB they took is what lets a ninth warp fit on the SM.
template <int LB, int DB>
struct WarpTabs {
    uint16_t lit[1 << LB][32];  // sym << 4 | len
    uint8_t dist[1 << DB][32];  // sym << 3 | len (DB <= 7)
    uint32_t in[kInWords][32];
    uint32_t tok[16][32];
    uint32_t lw[16][32];
};
struct Limits {  // lengths 7..15
    uint16_t llim[10];
    uint16_t dlim[10];
    int16_t loff[10];
    int16_t doff[10];
};

// What a lane has produced and not yet stored: literal bytes (the word being filled in a register, finished words in a ring of
// 16 in shared memory) and tokens (a ring of 16).  32 bytes / eight tokens leave as one sector store -- not when a lane
// happens to have them (some lane of the warp always does: the whole warp would step through the store sequence every
// round) but at a fixed place every eight rounds, where each lane stores the sector it has complete, if any.  A round adds at
// most one literal and one token, so the rings never hold more than 7 + 8 entries.
struct LaneOut {
    uint32_t *tokq, *litq;  // &W.tok[0][lane], &W.lw[0][lane]
    uint8_t *lit_dst;       // next 32 bytes of the literal stream
    uint8_t *tok_dst;       // next (lower) group of eight tokens
    uint32_t acc;
    uint32_t n_lit, n_tok;  // produced
    uint32_t f_lw, f_tok;   // literal words / tokens stored (multiples of 8)
    uint32_t run;           // literals since the last token
    __device__ __forceinline__ void push_byte(uint32_t b) {
        acc = __byte_perm(acc, b, 0x4321);  // bytes enter at the top: four of them later the word reads in stream order
        ++n_lit;
        if ((n_lit & 3u) == 0u) litq[(((n_lit >> 2) - 1u) & 15u) * 32] = acc;
    }
    __device__ __forceinline__ void push_token_raw(uint32_t t) {
        tokq[(n_tok & 15u) * 32] = t;
        ++n_tok;
    }
    __device__ __forceinline__ void literal(uint32_t b) {
        push_byte(b);
        if (++run == 255u) {
            push_token_raw(255u);
            run = 0;
        }
    }
    __device__ __forceinline__ void match(uint32_t len, uint32_t dist) {
        push_token_raw(run | (len << 8) | ((dist - 1u) << 17));
        run = 0;
    }
    __device__ __forceinline__ void store_if_complete() {
        if (n_tok - f_tok >= 8u) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = tokq[((f_tok + i) & 15u) * 32];
            st_sector(tok_dst, w);
            tok_dst -= 32;
            f_tok += 8u;
        }
        if ((n_lit >> 2) - f_lw >= 8u) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = litq[((f_lw + i) & 15u) * 32];
            st_sector(lit_dst, w);
            lit_dst += 32;
            f_lw += 8u;
        }
    }
    __device__ void drain() {  // outside the symbol loop's rhythm (stored blocks, block ends)
        while (n_tok - f_tok >= 8u || (n_lit >> 2) - f_lw >= 8u) store_if_complete();
    }
    // the pending literals become a token, partial sectors are padded with zeros (a zero token copies nothing)
    __device__ void finish(uint4 *hdr, bool ok) {
        if (run) push_token_raw(run);
        run = 0;
        const uint32_t nl = n_lit, nt = n_tok;
        drain();
        while (n_lit & 31u) push_byte(0u);
        while (n_tok & 7u) push_token_raw(0u);
        drain();
        *hdr = make_uint4(ok ? nt : 0u, ok ? nl : 0u, 0u, 0u);
    }
};

// Literal/length table payload (12 bits above the 4-bit code length): what the symbol loop needs, ready to use.
//   0x000..0x0FF literal byte | 0x100 end of block | 0x200 invalid symbol (286, 287)
//   0x800 | big << 6 | lx << 3 | m : a length, = 3 + (m << lx) + lx extra bits (+ 255 when `big`: symbol 285 = 258)
__device__ __forceinline__ uint32_t litlen_payload(uint32_t s) {
    if (s <= 256u) return s;
    const uint32_t ls = s - 257u;
    if (ls >= 29u) return 0x200u;
    if (ls == 28u) return 0x800u | 0x40u;
    if (ls < 8u) return 0x800u | ls;
    return 0x800u | (((ls - 4u) >> 2) << 3) | (4u + (ls & 3u));
}
struct PlainSym {
    __device__ __forceinline__ uint32_t operator()(uint32_t s) const { return s; }
};
struct LitLenSym {
    __device__ __forceinline__ uint32_t operator()(uint32_t s) const { return litlen_payload(s); }
};

template <int LB, int DB>
__global__ void __launch_bounds__(kDecWarps * 32, LB == 7 ? 4 : 2) inflate_decode_kernel(const uint8_t *comp, const BgzfMember *members, int n_members, int lanes_per_warp,
                                                                                       uint4 *tokens, uint32_t *flags, int *first_bad) {
    extern __shared__ __align__(16) uint8_t dec_smem_raw[];
    WarpTabs<LB, DB> &W = reinterpret_cast<WarpTabs<LB, DB> *>(dec_smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    uint16_t *const lit = &W.lit[0][lane];
    uint8_t *const dtab = &W.dist[0][lane];
    // With fewer members than the machine holds lanes, the members are spread over MORE warps with fewer working lanes each:
    // a warp's round costs the sum of the paths its lanes take, and the SM hides one warp's latency behind the others.
    if (lane >= lanes_per_warp) return;
    const int gt = (blockIdx.x * kDecWarps + (threadIdx.x >> 5)) * lanes_per_warp + lane, nt = gridDim.x * kDecWarps * lanes_per_warp;
    static_assert(LB >= 7 && LB <= 9 && DB >= 6 && DB <= 7, "limits / offsets start at length 7; the 8-bit distance table holds lengths <= 7");
    Canon<288> CL;  // lit/len code of the current block (and, while the header is read, the code-length code)
    Canon<32> CD;   // distance code
    Limits LM;
    uint8_t lens[320];
#pragma unroll 1
    for (int mi = gt; mi < n_members; mi += nt) {
        const BgzfMember M = members[mi];
        if (M.isize == 0) continue;
        const uint32_t isize = M.isize;
        uint4 *area = tokens + M.tok_off;
        LaneOut out;
        out.tokq = &W.tok[0][lane];
        out.litq = &W.lw[0][lane];
        out.lit_dst = reinterpret_cast<uint8_t *>(area + 2);
        out.tok_dst = reinterpret_cast<uint8_t *>(area + bgzf_token_units(isize) - 2);
        out.acc = 0u;
        out.n_lit = out.n_tok = out.run = out.f_lw = out.f_tok = 0u;
        LaneBits br;
        br.init(&W.in[0][lane], comp + M.in_off, M.in_len);
        const int64_t in_bits = (int64_t)M.in_len * 8;
        uint32_t pos = 0, err = 0;
        bool last = false;
#pragma unroll 1
        while (!last && !err) {
            br.refill_sync();
            const uint32_t hdr = br.peek();
            last = (hdr & 1u) != 0u;
            const int btype = (int)((hdr >> 1) & 3u);
            br.skip(3);
            if (br.spent || br.bits_consumed() > in_bits || btype == 3) {  // a stream that runs past its payload is corrupt
                err = kInfErrData;
                break;
            }
            if (btype == 0) {
                br.skip((8 - (br.bp & 7)) & 7);
                br.refill_sync();
                const uint32_t len = br.take(16);
                br.refill_sync();
                const uint32_t nlen = br.take(16);
                if ((len ^ nlen) != 0xFFFFu || pos + len > isize || br.bits_consumed() + (int64_t)len * 8 > in_bits) {
                    err = kInfErrData;
                    break;
                }
                for (uint32_t j = 0; j < len; ++j) {
                    br.refill_sync();
                    out.literal(br.take(8));
                    out.drain();
                }
                pos += len;
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
                br.refill_sync();
                const uint32_t h = br.peek();
                br.skip(14);
                const int nlit = (int)(h & 31u) + 257, ndist = (int)((h >> 5) & 31u) + 1, ncl = (int)((h >> 10) & 15u) + 4;
                uint8_t cl[19];
#pragma unroll
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < ncl; ++i) {
                    br.refill_sync();
                    cl[c_clen_order2[i]] = (uint8_t)br.take(3);
                }
                // the code-length code (7-bit codes at most) borrows the literal table's place and CL's arrays
                if (nlit > 286 || ndist > 30 || !canon_table<uint16_t, 4, PlainSym>(cl, 19, 7, lit, CL, nullptr, nullptr)) {
                    err = kInfErrData;
                    break;
                }
                int i = 0;
                const int total = nlit + ndist;
                while (i < total) {
                    br.refill_sync();
                    const uint32_t e = lit[(br.peek() & 127u) * 32];
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
                    br.refill_sync();  // peek() wants bp < 32
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
            if (!canon_table<uint16_t, 4, LitLenSym>(lens, 288, LB, lit, CL, LM.llim, LM.loff) || !canon_table<uint8_t, 3, PlainSym>(lens + 288, 30, DB, dtab, CD, LM.dlim, LM.doff)) {
                err = kInfErrData;
                break;
            }
            // ---- symbols: one per round; the FIFO is topped up every second round, finished sectors leave every eighth.
            // One exit, at the bottom: the lanes of the warp meet again after the literal / match fork of every round. ----
            br.fill();
            uint32_t round = 0;
#pragma unroll 1
            while (true) {
                if ((round & 1u) == 0u) br.topup();
                if ((round & 7u) == 7u) out.store_if_complete();
                ++round;
                br.refill();  // bp < 32: a literal/length code and its extra bits (<= 20) are there
                uint32_t w = br.peek();
                uint32_t e = lit[(w & ((1u << LB) - 1u)) * 32];
                if (!e) {
                    const uint32_t r = canon_long<LB>(w, LM.llim, LM.loff, CL);
                    e = r == 0xFFFFFFFFu ? (0x200u << 4) : (litlen_payload(r & 0xFFFFu) << 4) | (r >> 16);
                }
                const int clen = (int)(e & 15u);
                const uint32_t pay = e >> 4;
                bool stop = br.spent;
                if (pay < 0x100u) {
                    br.skip(clen);
                    if (pos < isize) {
                        ++pos;
                        out.literal(pay);
                    } else {
                        err = kInfErrData;
                    }
                } else if (pay & 0x800u) {
                    // length = 3 + (m << lx) + extra (+ 255 for symbol 285), RFC 1951 3.2.5
                    const uint32_t lx = (pay >> 3) & 7u;
                    const uint32_t len = 3u + ((pay & 7u) << lx) + ((w >> clen) & ((1u << lx) - 1u)) + ((pay >> 6) & 1u) * 255u;
                    br.skip(clen + (int)lx);  // <= 15 + 5
                    br.refill();              // bp < 32 again: a distance code and its extra bits (<= 28)
                    w = br.peek();
                    e = dtab[(w & ((1u << DB) - 1u)) * 32];
                    int dlen = (int)(e & 7u), ds = (int)(e >> 3);
                    if (!e) {
                        const uint32_t r = canon_long<DB>(w, LM.dlim, LM.doff, CD);
                        dlen = r == 0xFFFFFFFFu ? 0 : (int)(r >> 16);
                        ds = r == 0xFFFFFFFFu ? 30 : (int)(r & 0xFFFFu);
                    }
                    const int dx = (max(ds, 2) - 2) >> 1;
                    const uint32_t dist = 1u + ((uint32_t)(ds < 2 ? ds : 2 + (ds & 1)) << dx) + ((w >> dlen) & ((1u << dx) - 1u));
                    br.skip(dlen + dx);  // <= 15 + 13
                    if (ds < 30 && dist <= pos && pos + len <= isize) {
                        out.match(len, dist);
                        pos += len;
                    } else {
                        err = kInfErrData;
                    }
                } else {  // end of block, or a symbol that does not exist
                    br.skip(clen);
                    if (pay != 0x100u) err = kInfErrData;
                    stop = true;
                }
                if (stop || err) break;
            }
            out.drain();
            br.land();
            if (!err && (br.spent || br.bits_consumed() > in_bits)) err = kInfErrData;
        }
        if (!err && pos != isize) err = kInfErrSize;
        out.finish(area, err == 0);
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
// kRing: bytes of output a warp keeps in shared memory (8 KiB: 24 warps per SM; 16 KiB: 12, but fewer sources behind the ring)
// kSpan: most output bytes one group of tokens may produce (one token: <= 513)
// kNear: a source at most this far behind the group's first byte is read from the ring
constexpr uint32_t kLaneCopyMax = 32;       // an independent match up to this long is copied by its own lane
constexpr uint32_t kFlush = 512;            // finished bytes leave the ring as soon as there are this many

__device__ __forceinline__ uint32_t member_token(const uint32_t *tokw, uint32_t last_sector, uint32_t i) {
    // read once, 128 bytes per warp instruction: evict-first, so that the L2 is left to the output the far sources come from
    return __ldcs(tokw + 8u * (last_sector - (i >> 3)) + (i & 7u));
}

template <uint32_t kRing>
__global__ void __launch_bounds__(kCopyWarps * 32) inflate_copy_kernel(const BgzfMember *members, int n_members, const uint4 *tokens) {
    constexpr uint32_t kRingMask = kRing - 1, kSpan = kRing / 4, kNear = kRing - kSpan - 64;
    static_assert(kSpan >= 1024 && kNear >= kFlush + 16 + 128, "ring too small");
    extern __shared__ __align__(16) uint8_t copy_smem_raw[];
    uint8_t *const ring = copy_smem_raw + (size_t)(threadIdx.x >> 5) * kRing;  // ring[q & kRingMask] = output byte q (q counted from the member's 16-byte aligned origin)
    const uint32_t *const ringw = reinterpret_cast<const uint32_t *>(ring);
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * kCopyWarps + (threadIdx.x >> 5), nw = gridDim.x * kCopyWarps;
    constexpr uint32_t kFull = 0xFFFFFFFFu;
#pragma unroll 1
    for (int mi = gw; mi < n_members; mi += nw) {
        const BgzfMember M = members[mi];
        if (M.isize == 0) continue;
        const uint4 *area = tokens + M.tok_off;
        const uint32_t n_tok = __ldg(reinterpret_cast<const uint32_t *>(area));
        if (n_tok == 0) continue;  // the decoder gave up on this member
        const uint8_t *lits = reinterpret_cast<const uint8_t *>(area + 2);
        const uint32_t *tokw = reinterpret_cast<const uint32_t *>(area);
        const uint32_t last_sector = bgzf_token_units(M.isize) / 2u - 1u;
        const uint32_t q0 = (uint32_t)(M.out_addr & 15u);
        uint8_t *O = reinterpret_cast<uint8_t *>((uintptr_t)(M.out_addr - q0));
        const uint32_t *Ow = reinterpret_cast<const uint32_t *>(O);
        uint32_t outq = q0;    // next output position
        uint32_t flushed = 0;  // everything below has left the ring (a multiple of 16)
        uint32_t lit_off = 0, ti = 0;
        uint32_t cur = lane < (int)n_tok ? member_token(tokw, last_sector, (uint32_t)lane) : 0u;
        uint32_t nxt = 32u + lane < n_tok ? member_token(tokw, last_sector, 32u + lane) : 0u;
        uint32_t nx2 = 64u + lane < n_tok ? member_token(tokw, last_sector, 64u + lane) : 0u;  // two groups ahead: a loaded DRAM round trip outlasts one group
        __syncwarp();
#pragma unroll 1
        while (ti < n_tok) {
            // ---- positions: one scan of (literals + match bytes) | literals << 16 ----
            const uint32_t L = cur & 255u, ml = (cur >> 8) & 511u, dist = (cur >> 17) + 1u;
            const uint32_t v = (L + ml) | (L << 16);
            uint32_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(kFull, inc, d);
                if (lane >= d) inc += o;
            }
            const int n = __popc(__ballot_sync(kFull, (inc & 0xFFFFu) <= kSpan));  // >= 1; lanes n.. wait for the next group
            const uint32_t tot = __shfl_sync(kFull, inc, n - 1);
            const uint32_t gs = tot & 0xFFFFu, gl = tot >> 16;
            const bool mine = lane < n;
            const uint32_t exc = inc - v;
            const uint32_t o = outq + (exc & 0xFFFFu), lo = exc >> 16, ms = o + L;
            const uint32_t src = ms - dist;
            const uint32_t ring_lo = outq > kNear ? outq - kNear : 0u;  // <= flushed - 128: everything below is in global memory
            const bool has = mine && ml != 0u;
            // a match whose source ends before the group's first byte depends on nothing the group produces; the long ones
            // go the cooperative way all the same (one lane would spend 64 steps on 258 bytes while 31 wait)
            const bool indep = has && src + ml <= outq && ml <= kLaneCopyMax;
            if (has && src < ring_lo) {
                prefetch_line(O + src);
                prefetch_line(O + src + ml - 1u);
            }
            if (gl) prefetch_line(lits + lit_off + 256u);
            // ---- literals: read in stream order, each byte finds its token by binary search over the literal offsets ----
            for (uint32_t j0 = 0; j0 < gl; j0 += 32u) {
                const uint32_t j = j0 + lane;
                const uint32_t b = j < gl ? __ldg(lits + lit_off + j) : 0u;
                int t = 0;
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    const int c = t + s;
                    const uint32_t lc = __shfl_sync(kFull, lo, c & 31);
                    if (c < n && lc <= j) t = c;
                }
                const uint32_t dst = __shfl_sync(kFull, o, t) + j - __shfl_sync(kFull, lo, t);
                if (j < gl) ring[dst & kRingMask] = (uint8_t)b;
            }
            // ---- independent matches, one lane each; a source behind the ring left for global memory long ago ----
            // head bytes up to the first aligned destination word, whole words (each one funnel shift of two source words),
            // tail bytes: half the shared-memory stores of a bytewise copy.  The first 16 body bytes without a loop.
            if (indep) {
                const bool far = src < ring_lo;  // (then the whole source lies below flushed: ring_lo <= flushed - 128)
                const uint32_t head = min((0u - ms) & 3u, ml);
                uint32_t wi = src >> 2;
                uint32_t sw[6];  // the source words the head and the first 16 body bytes can touch
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const bool live = (uint32_t)(4 * i) < (src & 3u) + ml;
                    sw[i] = !live ? 0u : far ? Ow[wi + i] : ringw[(wi + i) & (kRingMask >> 2)];
                }
                const uint32_t sh = (src & 3u) * 8u;
                const uint32_t x = __funnelshift_r(sw[0], sw[1], sh);  // source bytes 0..3
                if (head > 0u) ring[ms & kRingMask] = (uint8_t)x;
                if (head > 1u) ring[(ms + 1u) & kRingMask] = (uint8_t)(x >> 8);
                if (head > 2u) ring[(ms + 2u) & kRingMask] = (uint8_t)(x >> 16);
                // from `head` on the destination is word aligned; the source then starts (src + head) & 3 bytes into word
                // (src + head) >> 2, which is sw[0] or sw[1]
                const uint32_t a0 = src + head, sh2 = (a0 & 3u) * 8u;
                const bool up = (a0 >> 2) != wi;
                const uint32_t body = ml - head;
                uint32_t *dwp = reinterpret_cast<uint32_t *>(ring);
                const uint32_t dw = (ms + head) >> 2;
                uint32_t y[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) y[i] = __funnelshift_r(up ? sw[i + 1] : sw[i], up ? sw[i + 2] : sw[i + 1], sh2);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t rem = body - min(body, (uint32_t)(4 * i));
                    const uint32_t d = (ms + head + 4u * i);
                    if (rem >= 4u) {
                        dwp[(dw + i) & (kRingMask >> 2)] = y[i];
                    } else {
                        if (rem > 0u) ring[d & kRingMask] = (uint8_t)y[i];
                        if (rem > 1u) ring[(d + 1u) & kRingMask] = (uint8_t)(y[i] >> 8);
                        if (rem > 2u) ring[(d + 2u) & kRingMask] = (uint8_t)(y[i] >> 16);
                    }
                }
                for (uint32_t k = 16u; k < body; k += 4u) {  // the few matches of 17..32 bytes
                    const uint32_t w_i = (a0 + k) >> 2;
                    const uint32_t w0 = far ? Ow[w_i] : ringw[w_i & (kRingMask >> 2)];
                    const uint32_t w1 = far ? Ow[w_i + 1u] : ringw[(w_i + 1u) & (kRingMask >> 2)];
                    const uint32_t yy = __funnelshift_r(w0, w1, sh2);
                    const uint32_t rem = body - k, d = ms + head + k;
                    if (rem >= 4u) {
                        dwp[(d >> 2) & (kRingMask >> 2)] = yy;
                    } else {
                        ring[d & kRingMask] = (uint8_t)yy;
                        if (rem > 1u) ring[(d + 1u) & kRingMask] = (uint8_t)(yy >> 8);
                        if (rem > 2u) ring[(d + 2u) & kRingMask] = (uint8_t)(yy >> 16);
                    }
                }
            }
            __syncwarp();
            // ---- long independent matches, the whole warp on one (rare; the source may lie behind the ring) ----
            uint32_t lm = __ballot_sync(kFull, has && !indep && src + ml <= outq);
#pragma unroll 1
            while (lm) {
                const int b = __ffs((int)lm) - 1;
                lm &= lm - 1u;
                const uint32_t tk = __shfl_sync(kFull, cur, b), m0 = __shfl_sync(kFull, ms, b);
                const uint32_t len = (tk >> 8) & 511u, s0 = m0 - ((tk >> 17) + 1u);
                for (uint32_t j = lane; j < len; j += 32u) ring[(m0 + j) & kRingMask] = s0 + j < ring_lo ? O[s0 + j] : ring[(s0 + j) & kRingMask];
            }
            __syncwarp();
            // ---- the others in token order, the whole warp on one match: a chain of matches costs one short step per link.
            // Their sources reach into the group itself, so they lie in the ring. ----
            const bool dep = has && src + ml > outq;
            uint32_t dm = __ballot_sync(kFull, dep);
            const uint32_t pk = (ms & kRingMask) | (ml << 16);  // what the warp needs of a match, packed by its owner
            const uint32_t sk = src & kRingMask;
#pragma unroll 1
            while (dm) {
                const int b = __ffs((int)dm) - 1;
                dm &= dm - 1u;
                const uint32_t p = __shfl_sync(kFull, pk, b), s0 = __shfl_sync(kFull, sk, b);
                const uint32_t len = p >> 16, m0 = p & 0xFFFFu, dd = (m0 - s0) & kRingMask;
                if (len <= 32u && dd >= len) {  // the common case: one step, no overlap
                    if ((uint32_t)lane < len) ring[(m0 + lane) & kRingMask] = ring[(s0 + lane) & kRingMask];
                } else if (dd >= 32u) {  // every step's 32 source bytes lie below its 32 destination bytes
                    for (uint32_t j0 = 0; j0 < len; j0 += 32u) {
                        const uint32_t j = j0 + lane;
                        if (j < len) ring[(m0 + j) & kRingMask] = ring[(s0 + j) & kRingMask];
                        __syncwarp();
                    }
                } else {  // the match overlaps itself: its first dd bytes repeat
                    for (uint32_t j = lane; j < len; j += 32u) ring[(m0 + j) & kRingMask] = ring[(s0 + j % dd) & kRingMask];
                }
                __syncwarp();
            }
            outq += gs;
            lit_off += gl;
            ti += (uint32_t)n;
            // ---- finished bytes leave the ring ----
            const bool at_end = ti >= n_tok;
            if (outq - flushed >= kFlush || at_end) {
                const uint32_t lim = at_end ? outq : (outq & ~15u);
                for (uint32_t u = flushed + 16u * (uint32_t)lane; u < lim; u += 512u) {
                    if (u >= q0 && u + 16u <= lim) {
                        *reinterpret_cast<uint4 *>(O + u) = *reinterpret_cast<const uint4 *>(&ring[u & kRingMask]);
                    } else {  // the member's first and last unit are shared with its neighbours
                        const uint32_t a = max(u, q0), b = min(u + 16u, lim);
                        for (uint32_t x = a; x < b; ++x) O[x] = ring[x & kRingMask];
                    }
                }
                flushed = lim & ~15u;
                __syncwarp();
            }
            // ---- next tokens ----
            {
                const uint32_t fresh = ti + 64u + lane < n_tok ? member_token(tokw, last_sector, ti + 64u + lane) : 0u;
                if (n == 32) {
                    cur = nxt;
                    nxt = nx2;
                    nx2 = fresh;
                } else {
                    const int from = lane + n;
                    const uint32_t x = __shfl_sync(kFull, cur, from & 31), y = __shfl_sync(kFull, nxt, from & 31), z = __shfl_sync(kFull, nx2, from & 31);
                    cur = from < 32 ? x : y;
                    nxt = from < 32 ? y : z;
                    nx2 = from < 32 ? z : fresh;
                }
            }
        }
    }
}

template <int LB, int DB>
int launch_decode(Ctx *c, const uint8_t *d_comp, const BgzfMember *d_table, int n_members, uint4 *d_tokens, uint32_t *d_flags) {
    static int occ = 0;
    constexpr size_t smem = sizeof(WarpTabs<LB, DB>) * kDecWarps;
    if (!occ) {
        CUDA_TRY(cudaFuncSetAttribute(inflate_decode_kernel<LB, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, inflate_decode_kernel<LB, DB>, kDecWarps * 32, smem));
        if (occ < 1) occ = 1;
    }
    // one wave when the members fit: spread them over every warp the machine holds (at least 8 lanes per warp)
    const int max_ctas = occ * c->sm_count, max_warps = max_ctas * kDecWarps;
    int lpw = std::min(32, std::max(8, (n_members + max_warps - 1) / max_warps));
    static const int forced_lpw = [] {
        const char *e = getenv("EXON_GPU_INFLATE_LANES");
        return e ? atoi(e) : 0;
    }();
    if (forced_lpw >= 1 && forced_lpw <= 32) lpw = forced_lpw;
    const int per_cta = kDecWarps * lpw;
    const int grid = std::min((n_members + per_cta - 1) / per_cta, max_ctas);
    inflate_decode_kernel<LB, DB><<<grid, kDecWarps * 32, smem, c->stream>>>(d_comp, d_table, n_members, lpw, d_tokens, d_flags, (int *)(d_flags + 1));
    c->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return EXON_GPU_OK;
}

template <uint32_t kRing>
int launch_copy(Ctx *c, const BgzfMember *d_table, int n_members, const uint4 *tok) {
    static int occ = 0;
    constexpr size_t copy_smem = (size_t)kCopyWarps * kRing;
    if (!occ) {
        CUDA_TRY(cudaFuncSetAttribute(inflate_copy_kernel<kRing>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)copy_smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, inflate_copy_kernel<kRing>, kCopyWarps * 32, copy_smem));
        if (occ < 1) occ = 1;
    }
    static const int copy_ctas = [] {  // experiment knob: CTAs per SM (fewer members in flight = more of their history still in the L2)
        const char *e = getenv("EXON_GPU_INFLATE_COPY_CTAS");
        return e ? atoi(e) : 0;
    }();
    const int ctas_per_sm = copy_ctas > 0 ? std::min(copy_ctas, occ) : occ;
    const int grid = std::min((n_members + kCopyWarps - 1) / kCopyWarps, ctas_per_sm * c->sm_count);
    inflate_copy_kernel<kRing><<<grid, kCopyWarps * 32, copy_smem, c->stream>>>(d_table, n_members, tok);
    c->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return EXON_GPU_OK;
}

}  // namespace

// Gives every member its place in the token scratch; returns the number of 16-byte units.
size_t bgzf_assign_tokens(BgzfMember *m, size_t n) {
    size_t units = 0;
    for (size_t i = 0; i < n; ++i) {
        m[i].tok_off = (uint32_t)units;
        m[i].pad_ = 0;
        if (m[i].isize) units += bgzf_token_units(m[i].isize);
    }
    return units;
}

// Enqueues the inflate of `n_members` members (table in device memory; in_off relative to d_comp, out_addr absolute,
// tok_off from bgzf_assign_tokens or its device-side equivalent) on the context's stream.  d_flags: two words of device
// scratch, {0, INT_MAX} before the launch.  d_comp must be 16-byte aligned.
int bgzf_inflate_launch(Ctx *c, const uint8_t *d_comp, const BgzfMember *d_table, int n_members, uint32_t *d_flags, size_t token_units,
                        size_t comp_bytes) {
    if (n_members <= 0) return EXON_GPU_OK;
    if (token_units > 0xFFFFFFFFull) return fail(EXON_GPU_ERR_UNSUPPORTED, "inflate: more than 48 GiB of output in one launch");
    const size_t tok_bytes = (token_units + 4) * 16;
    if (tok_bytes > c->inf_tokens_cap) {
        if (c->inf_tokens) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            CUDA_TRY(cudaFree(c->inf_tokens));
            c->inf_tokens = nullptr;
            c->inf_tokens_cap = 0;
        }
        const size_t cap = tok_bytes + tok_bytes / 8;
        CUDA_TRY(cudaMalloc(&c->inf_tokens, cap));
        c->inf_tokens_cap = cap;
    }
    uint4 *tok = (uint4 *)c->inf_tokens;
    // Table size.  The decoder is bound by the latency of each lane's serial chain, so what counts is warps per SM, and the
    // tables are what limits them: 7-bit literal / 6-bit distance tables put 16 warps on an SM, 8 / 7 bits 8, 9 / 7 bits 4.
    // Codes longer than the table take the slow canonical search -- 1.7 % of the symbols of VCF text at 7 bits, 4.4 % of
    // BAM records (tools/deflate_stats.c).  EXON_GPU_INFLATE_TABLES = 76 | 87 | 97 forces one.
    static const int forced = [] {
        const char *e = getenv("EXON_GPU_INFLATE_TABLES");
        return e ? atoi(e) : 0;
    }();
    const int pick = forced ? forced : 76;
    (void)comp_bytes;
    const int rc = pick == 97   ? launch_decode<9, 7>(c, d_comp, d_table, n_members, tok, d_flags)
                   : pick == 87 ? launch_decode<8, 7>(c, d_comp, d_table, n_members, tok, d_flags)
                                : launch_decode<7, 6>(c, d_comp, d_table, n_members, tok, d_flags);
    if (rc) return rc;
    static const int ring_kib = [] {
        const char *e = getenv("EXON_GPU_INFLATE_RING_KIB");
        return e ? atoi(e) : 8;
    }();
    return ring_kib == 16 ? launch_copy<16384>(c, d_table, n_members, tok) : launch_copy<8192>(c, d_table, n_members, tok);
}

}  // namespace exon
