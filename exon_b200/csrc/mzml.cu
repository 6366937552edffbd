// mzml.cu -- mzML text -> m/z range predicate -> SUM(intensity), fused (BASELINE.json configs[4]).
//
// Replaces, for `SELECT SUM(i) FROM (SELECT unnest(mz.mz) m, unnest(intensity.intensity) i FROM mzml) WHERE m BETWEEN a AND b`:
//   MzMLReader::read_spectrum        exon/exon-mzml/src/mzml_reader/parser.rs:43-109 (quick-xml events of every <spectrum>)
//   decode_binary_array              exon/exon-mzml/src/mzml_reader/binary_conversion.rs:26-95 (base64 -> LE f32 / f64 -> f64)
//   MzMLArrayBuilder::append         exon/exon-mzml/src/array_builder.rs:236-416 (List<Float64> per array kind, classified by
//                                    the cvParam accessions of exon/exon-mzml/src/mzml_reader/types.rs:119-121,207-208,274-275)
//   unnest + FilterExec + AggregateExec(Partial) sum                                 (DataFusion 44, third party)
// Three kernels, none of which materialises a List<Float64>:
//   M1 events   the warp-private TMA tile pipeline (tile_ring.cuh) reads the text once.  '<' and ':' never occur inside a
//               base64 payload, so 16-byte chunks are screened for those two bytes; the few hits are classified
//               (<spectrum, <binaryDataArray, <binary>, </binary>, accession="MS:1000xxx") and appended to an event list
//               (segment, offset, kind packed into 64 bits), which cub then sorts into file order
//   M2 spectra  one thread per <spectrum> event walks the events up to the next spectrum and leaves a descriptor of its
//               m/z and intensity payloads (address, trimmed length, 32/64 bit, compression)
//   M3 sum      one warp per spectrum: lane i decodes value i of both payloads straight from the base64 text (16 characters
//               cover any 8-byte value: 5 aligned word loads, a 256-entry table in shared memory, byte permutes), applies
//               lo <= mz <= hi and accumulates intensity in f64; warp shuffle reduction, one atomicAdd per warp
// zlib-compressed arrays (MS:1000574; binary_conversion.rs:44-60 wraps the base64 bytes in a flate2 ZlibDecoder) take a
// detour: their base64 text is decoded into a staging buffer (Z2), a member table is built on the device (Z3: 2-byte zlib
// header checked and skipped, 4-byte Adler-32 trailer dropped, output size = the spectrum's defaultArrayLength x the value
// width) and the DEFLATE streams go through the two inflate kernels of inflate.cu; M3 then reads those values from the
// inflated bytes instead of the base64 text.  A stream whose output differs from the declared length is an error here
// (the reference decodes to the end of the stream whatever the attribute says).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstring>

#include "common.cuh"
#include "internal.h"
#include "mzml.cuh"
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

constexpr int kMzBuf = 30;  // events a warp collects in shared memory before it reserves space in the global list
using MzRing = TileRing<4096, 2, 8, 16, 368, 16 + 8 * kMzBuf>;  // 2 stages + a small buffer: 3 CTAs per SM
constexpr int kMzU = MzRing::TILE / 512;

enum : uint32_t {
    kEvSpectrum = 1, kEvBda = 2, kEvMz = 3, kEvIntensity = 4, kEvWave = 5, kEvF32 = 6, kEvF64 = 7, kEvZlib = 8, kEvNoComp = 9,
    kEvBinStart = 10, kEvBinEnd = 11, kEvSpecEnd = 12  // </spectrum>: arrays after it (chromatograms) belong to no spectrum
};
constexpr uint32_t kMzErrFormat = 1u;   // malformed structure (a <binary> without </binary>, missing data type, bad base64)
constexpr uint32_t kMzErrZlib = 2u;     // a zlib array without a usable defaultArrayLength, or with a bad zlib header
constexpr uint32_t kMzHasZlib = 4u;     // (not an error) some array is zlib-compressed: the host runs the inflate detour

struct MzArgs {
    const ScanSeg *segs;
    int32_t n_segs;
    int64_t n_tiles;
    unsigned long long *events;   // key = seg << 44 | offset << 4 | kind
    unsigned long long cap;
    unsigned long long *n_events;  // [0] events appended (may exceed cap: the host retries) ... [5] <spectrum> events
};

__device__ __forceinline__ bool is_delim(uint32_t c) { return c == ' ' || c == '>' || c == '/' || c == '\n' || c == '\t' || c == '\r'; }

// four characters as the little-endian word an unaligned load yields
__host__ __device__ constexpr uint32_t cc4(char a, char b, char c, char d) {
    return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24);
}
// 32-bit load from shared memory at any alignment
__device__ __forceinline__ uint32_t lds_unaligned32(uint32_t sa) {
    const uint32_t a0 = sa & ~3u;
    return __funnelshift_r(lds32(a0), lds32(a0 + 4), (sa & 3u) * 8u);
}

template <class V>
__device__ __forceinline__ bool match_at(const V &v, int p, const char *lit, int n) {
    for (int i = 0; i < n; ++i)
        if (view_byte(v, p + i) != (uint32_t)(uint8_t)lit[i]) return false;
    return true;
}

__global__ void __launch_bounds__(MzRing::WARPS * 32, 3) mzml_events_kernel(const __grid_constant__ MzArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    MzRing ring;
    ring.init(smem_raw, a.segs, a.n_tiles);
    const int lane = ring.lane;
    constexpr uint32_t kLT4 = 0x3C3C3C3Cu, kCOL4 = 0x3A3A3A3Au;
    // per-warp event buffer: [0] count (shared-memory atomic), [2..] keys; one global reservation per flush instead of one
    // global atomic per event (the list head is a single address)
    unsigned int *buf_n = reinterpret_cast<unsigned int *>(ring.extra);
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(ring.extra + 16);
    if (lane == 0) *buf_n = 0;
    __syncwarp();
    auto flush = [&]() {
        __syncwarp();
        const unsigned int n = *buf_n < (unsigned)kMzBuf ? *buf_n : (unsigned)kMzBuf;
        unsigned long long base = 0;
        if (lane == 0 && n) base = atomicAdd(a.n_events, (unsigned long long)n);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        for (unsigned int i = lane; i < n; i += 32)
            if (base + i < a.cap) a.events[base + i] = buf[i];
        __syncwarp();
        if (lane == 0) *buf_n = 0;
        __syncwarp();
    };
    uint32_t n_spectrum = 0;
#pragma unroll 1
    for (int64_t T = ring.first_tile(); T < a.n_tiles; T += ring.nw) {
        const MzRing::View v = ring.acquire();
        // offset of tile byte 0 inside its segment
        const long long tile_off = (long long)(v.g - (a.segs[v.seg].base + a.segs[v.seg].skip));
        const unsigned long long seg_key = (unsigned long long)v.seg << 44;
#pragma unroll 1
        for (int u = 0; u < kMzU; ++u) {
            const int c0 = (u * 32 + lane) * 16;
            uint32_t m = 0;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (v.interior || c0 < v.sm_hi) {
            w = lds128(v.sa + (uint32_t)c0);
            uint32_t hit = zero_bytes_fast(w.x ^ kLT4) | zero_bytes_fast(w.y ^ kLT4) | zero_bytes_fast(w.z ^ kLT4) | zero_bytes_fast(w.w ^ kLT4) |
                           zero_bytes_fast(w.x ^ kCOL4) | zero_bytes_fast(w.y ^ kCOL4) | zero_bytes_fast(w.z ^ kCOL4) | zero_bytes_fast(w.w ^ kCOL4);
            if (hit) m = pack_flags16(zero_bytes_exact(w.x ^ kLT4) | zero_bytes_exact(w.x ^ kCOL4), zero_bytes_exact(w.y ^ kLT4) | zero_bytes_exact(w.y ^ kCOL4),
                                      zero_bytes_exact(w.z ^ kLT4) | zero_bytes_exact(w.z ^ kCOL4), zero_bytes_exact(w.w ^ kLT4) | zero_bytes_exact(w.w ^ kCOL4));
            }
            while (m) {
                const int p = c0 + __ffs(m) - 1;
                m &= m - 1;
                if (p < v.seg_lo || p >= v.hi) continue;
                uint32_t kind = 0;
                long long at = tile_off + p;
                const int lo_ok = v.seg_lo > v.sm_lo ? v.seg_lo : v.sm_lo, hi_ok = v.hi < v.sm_hi ? v.hi : v.sm_hi;
                const uint32_t pa = v.sa + (uint32_t)p;  // shared-window address of the hit (two's complement for p < 0)
                const bool is_lt = view_byte(v, p) == '<';
                if (is_lt && p + 20 <= hi_ok) {
                    // fast path: the tag name as unaligned words from the staged bytes; most '<' open a tag we do not track
                    const uint32_t w0 = lds_unaligned32(pa);
                    if (w0 == cc4('<', 's', 'p', 'e')) {
                        if (lds_unaligned32(pa + 4) == cc4('c', 't', 'r', 'u') && lds8(pa + 8) == 'm' && is_delim(lds8(pa + 9))) kind = kEvSpectrum;
                    } else if (w0 == cc4('<', 'b', 'i', 'n')) {
                        const uint32_t w1 = lds_unaligned32(pa + 4);
                        if (w1 == cc4('a', 'r', 'y', '>')) {
                            kind = kEvBinStart;
                            at += 8;  // the payload starts after the tag
                        } else if (w1 == cc4('a', 'r', 'y', 'D') && lds_unaligned32(pa + 8) == cc4('a', 't', 'a', 'A') && lds_unaligned32(pa + 12) == cc4('r', 'r', 'a', 'y') &&
                                   is_delim(lds8(pa + 16))) {
                            kind = kEvBda;
                        }
                    } else if (w0 == cc4('<', '/', 'b', 'i')) {
                        if (lds_unaligned32(pa + 4) == cc4('n', 'a', 'r', 'y') && lds8(pa + 8) == '>') kind = kEvBinEnd;
                    } else if (w0 == cc4('<', '/', 's', 'p')) {
                        if (lds_unaligned32(pa + 4) == cc4('e', 'c', 't', 'r') && (lds_unaligned32(pa + 8) & 0x00FFFFFFu) == cc4('u', 'm', '>', 0)) kind = kEvSpecEnd;
                    }
                } else if (is_lt) {
                    const uint32_t c1 = view_byte(v, p + 1);
                    if (c1 == 's') {
                        if (match_at(v, p + 1, "spectrum", 8) && is_delim(view_byte(v, p + 9))) kind = kEvSpectrum;
                    } else if (c1 == 'b') {
                        if (match_at(v, p + 1, "binary", 6)) {
                            if (view_byte(v, p + 7) == '>') {
                                kind = kEvBinStart;
                                at += 8;  // the payload starts after the tag
                            } else if (match_at(v, p + 7, "DataArray", 9) && is_delim(view_byte(v, p + 16))) {
                                kind = kEvBda;
                            }
                        }
                    } else if (c1 == '/') {
                        if (match_at(v, p + 2, "binary>", 7)) kind = kEvBinEnd;
                        else if (match_at(v, p + 2, "spectrum>", 9)) kind = kEvSpecEnd;
                    }
                } else if (p - 14 >= lo_ok && p + 9 <= hi_ok) {
                    // ... accession="MS:1000xxx": p is the ':'; 16 bytes before and 8 after it as unaligned words
                    if (lds_unaligned32(pa - 2) == cc4('M', 'S', ':', '1') && lds_unaligned32(pa - 6) == cc4('o', 'n', '=', '"') &&
                        lds_unaligned32(pa - 10) == cc4('e', 's', 's', 'i') && lds_unaligned32(pa - 14) == cc4(' ', 'a', 'c', 'c') &&
                        (lds_unaligned32(pa + 2) & 0x00FFFFFFu) == cc4('0', '0', '0', 0) && lds8(pa + 8) == '"') {
                        const uint32_t d0 = lds8(pa + 5) - '0', d1 = lds8(pa + 6) - '0', d2 = lds8(pa + 7) - '0';
                        if (d0 <= 9u && d1 <= 9u && d2 <= 9u) {
                            const uint32_t code = d0 * 100u + d1 * 10u + d2;
                            kind = code == 514u ? kEvMz : code == 515u ? kEvIntensity : code == 617u ? kEvWave : code == 521u ? kEvF32
                                 : code == 523u ? kEvF64 : code == 574u ? kEvZlib : code == 576u ? kEvNoComp : 0u;
                        }
                    }
                } else {
                    // ... accession="MS:1000xxx": p is the ':'
                    if (match_at(v, p - 14, " accession=\"MS:1000", 19) && view_byte(v, p + 8) == '"') {
                        const uint32_t d0 = view_byte(v, p + 5) - '0', d1 = view_byte(v, p + 6) - '0', d2 = view_byte(v, p + 7) - '0';
                        if (d0 <= 9u && d1 <= 9u && d2 <= 9u) {
                            const uint32_t code = d0 * 100u + d1 * 10u + d2;
                            kind = code == 514u ? kEvMz : code == 515u ? kEvIntensity : code == 617u ? kEvWave : code == 521u ? kEvF32
                                 : code == 523u ? kEvF64 : code == 574u ? kEvZlib : code == 576u ? kEvNoComp : 0u;
                        }
                    }
                }
                if (kind) {
                    if (kind == kEvSpectrum) ++n_spectrum;
                    const unsigned long long key = seg_key | ((unsigned long long)at << 4) | kind;
                    const unsigned int slot = atomicAdd(buf_n, 1u);
                    if (slot < (unsigned)kMzBuf) {
                        buf[slot] = key;
                    } else {  // buffer full (an XML-dense stretch): straight to the global list
                        const unsigned long long i = atomicAdd(a.n_events, 1ull);
                        if (i < a.cap) a.events[i] = key;
                    }
                }
            }
            __syncwarp();
            // events beyond the buffer go straight to the global list; flush once the buffer is half full
            if (*buf_n > (unsigned)(kMzBuf / 2)) flush();
        }
        flush();
        ring.release(T);
    }
    n_spectrum = warp_sum(n_spectrum);
    if (lane == 0 && n_spectrum) atomicAdd(a.n_events + 5, (unsigned long long)n_spectrum);
}

// (SpecDesc lives in mzml.cuh: the column build of mzml_columns.cu reads the same descriptors)

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

// 1 for every <spectrum> event (its exclusive scan is the spectrum's rank in file order)
__global__ void mzml_spec_flags_kernel(const unsigned long long *ev, unsigned long long n, int32_t *flag) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) flag[i] = i < n && (ev[i] & 15u) == kEvSpectrum ? 1 : 0;
}

__global__ void mzml_spectra_kernel(const unsigned long long *ev, unsigned long long n, const ScanSeg *segs, const long long *rank, SpecDesc *out,
                                    uint32_t *flags) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (ev[i] & 15u) != kEvSpectrum) return;
    const unsigned long long seg = ev[i] >> 44;
    SpecDesc d;
    memset(&d, 0, sizeof(d));
    const uint8_t *base = segs[seg].base + segs[seg].skip;
    d.seg = (uint32_t)seg;
    d.tag = base + ((ev[i] >> 4) & ((1ull << 40) - 1ull));
    d.end = base + segs[seg].len;  // until a </spectrum> event says otherwise
    // defaultArrayLength="N" inside the <spectrum ...> tag (the only place a zlib array's decoded size is declared); parsed
    // only when the spectrum turns out to hold a zlib array
    auto default_array_length = [&]() -> uint32_t {
        const uint8_t *t = d.tag;
        const uint8_t *tend = base + segs[seg].len;
        const char lit[] = "defaultArrayLength=\"";
        for (int k = 0; k < 4096 && t + k + 20 < tend && t[k] != '>'; ++k) {
            if (t[k] != 'd') continue;
            bool m = true;
            for (int q = 1; q < 20 && m; ++q) m = t[k + q] == (uint8_t)lit[q];
            if (!m) continue;
            unsigned long long v = 0;
            for (const uint8_t *c = t + k + 20; c < tend && *c >= '0' && *c <= '9' && v < (1ull << 31); ++c) v = v * 10ull + (*c - '0');
            return v < (1ull << 31) ? (uint32_t)v : 0u;
        }
        return 0u;
    };
    uint32_t kind = 0, f32 = 0, f64 = 0, zl = 0, nc = 0, err = 0;
    unsigned long long start = 0;
    bool open = false;
    for (unsigned long long j = i + 1; j < n; ++j) {
        const unsigned long long e = ev[j];
        const uint32_t k = (uint32_t)(e & 15u);
        if ((e >> 44) != seg || k == kEvSpectrum) break;
        const unsigned long long off = (e >> 4) & ((1ull << 40) - 1ull);
        if (k == kEvSpecEnd) {
            d.end = base + off;
            break;
        }
        if (k == kEvBda) { kind = f32 = f64 = zl = nc = 0; open = false; }
        else if (k == kEvMz || k == kEvIntensity || k == kEvWave) { if (!kind) kind = k; }
        else if (k == kEvF32) f32 = 1;
        else if (k == kEvF64) f64 = 1;
        else if (k == kEvZlib) zl = 1;
        else if (k == kEvNoComp) nc = 1;
        else if (k == kEvBinStart) { start = off; open = true; }
        else if (k == kEvBinEnd) {
            if (!open) { err |= kMzErrFormat; continue; }
            open = false;
            const uint8_t *b0 = base + start, *b1 = base + off;
            while (b0 < b1 && is_ws(*b0)) ++b0;   // quick-xml trim_text(true)
            while (b1 > b0 && is_ws(b1[-1])) --b1;
            if (!kind) continue;  // no array type
            if (b1 == b0) {       // <binary/> or empty content: the builder appends an EMPTY list for the array (array_builder.rs:276-296)
                const int a = kind == kEvMz ? 0 : kind == kEvIntensity ? 1 : 2;
                d.arr[a] = b0;
                d.len[a] = 0;
                d.f32[a] = d.zl[a] = 0;
                continue;
            }
            if ((!f32 && !f64) || (!zl && !nc)) { err |= kMzErrFormat; continue; }
            if (((b1 - b0) & 3) != 0 || (b1 - b0) > 0x7FFFFFFFll) { err |= kMzErrFormat; continue; }
            if (zl) {
                if (!d.n_default) d.n_default = default_array_length();
                err |= d.n_default ? kMzHasZlib : kMzErrZlib;
            }
            const int a = kind == kEvMz ? 0 : kind == kEvIntensity ? 1 : 2;
            d.arr[a] = b0;
            d.len[a] = (uint32_t)(b1 - b0);
            d.f32[a] = f32 && !f64;
            d.zl[a] = zl;
        }
    }
    if (open) err |= kMzErrFormat;
    if (err) atomicOr(flags, err);
    out[rank[i]] = d;
}

__global__ void __launch_bounds__(256) mzml_sum_kernel(const SpecDesc *specs, unsigned long long n_spec, int has_pred, double lo, double hi,
                                                      double *out_sum, unsigned long long *out_cnt, uint32_t *flags) {
    __shared__ uint8_t lut[256];
    b64_lut_init(lut);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned long long wid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    unsigned long long cnt = 0;
    uint32_t bad = 0;
    for (unsigned long long sidx = wid; sidx < n_spec; sidx += nw) {
        const SpecDesc d = specs[sidx];
        if (!d.arr[0] || !d.arr[1]) continue;
        const int wm = d.f32[0] ? 4 : 8, wi = d.f32[1] ? 4 : 8;
        const uint32_t nm = d.zl[0] ? d.n_default : b64_bytes(d.arr[0], d.len[0]) / (uint32_t)wm;
        const uint32_t ni = d.zl[1] ? d.n_default : b64_bytes(d.arr[1], d.len[1]) / (uint32_t)wi;
        const uint32_t n = nm < ni ? nm : ni;   // the two unnested lists are zipped; the longer one's tail meets NULLs
        for (uint32_t i = lane; i < n; i += 32) {
            const unsigned long long mb = d.zl[0] ? raw_value(d.raw[0], i, wm) : b64_value(d.arr[0], i, wm, lut, bad);
            const double m = d.f32[0] ? (double)__uint_as_float((uint32_t)mb) : __longlong_as_double((long long)mb);
            if (has_pred && !(m >= lo && m <= hi)) continue;
            const unsigned long long ib = d.zl[1] ? raw_value(d.raw[1], i, wi) : b64_value(d.arr[1], i, wi, lut, bad);
            acc += d.f32[1] ? (double)__uint_as_float((uint32_t)ib) : __longlong_as_double((long long)ib);
            ++cnt;
        }
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, dlt);
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, dlt);
        bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, dlt);
    }
    if (lane == 0) {
        if (cnt) {
            atomicAdd(out_sum, acc);
            atomicAdd(out_cnt, cnt);
        }
        if (bad & 0x80u) atomicOr(flags, kMzErrFormat);
    }
}


// ---- zlib detour ----------------------------------------------------------------------------------------------------
// Inflated size of a zlib array: defaultArrayLength x value width, computed in 64 bits and clamped to what DEFLATE can
// expand `comp` bytes to (at most 1032 : 1) and to 1 GiB.  defaultArrayLength is an untrusted XML attribute: an absurd
// value must not wrap the int32 size scans (and with them the output and token scratch allocations); a clamped size differs
// from what the stream inflates to, so the array fails the inflate kernels' size check and the query reports it.
__device__ __forceinline__ uint32_t zarr_isize(uint32_t n_default, bool f32, uint32_t comp) {
    const unsigned long long want = (unsigned long long)n_default * (f32 ? 4ull : 8ull);
    const unsigned long long cap = ((1032ull * comp + 64ull) < (1ull << 30) ? (1032ull * comp + 64ull) : (1ull << 30));
    return (uint32_t)(want < cap ? want : cap);
}
struct ZArr {
    uint32_t spec, which;  // which: 0 = m/z, 1 = intensity, 2 = wavelength
};
// Z0: the list of zlib arrays + their sizes: [0] compressed bytes (padded), [1] inflated bytes (padded to 16), [2] token scratch units
__global__ void mzml_zlist_kernel(const SpecDesc *specs, unsigned long long n_spec, ZArr *list, unsigned long long *n_list, int32_t *sizes,
                                  unsigned long long cap) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_spec) return;
    const SpecDesc d = specs[i];
    for (uint32_t which = 0; which < 3; ++which) {
        const bool zl = d.zl[which] != 0;
        const uint8_t *p = d.arr[which];
        if (!zl || !p) continue;
        const unsigned long long k = atomicAdd(n_list, 1ull);
        if (k >= cap) continue;
        list[k] = ZArr{(uint32_t)i, which};
        const uint32_t comp = b64_bytes(p, d.len[which]);
        const uint32_t isize = zarr_isize(d.n_default, d.f32[which] != 0, comp);
        sizes[k] = (int32_t)((comp + 64u + 15u) & ~15u);
        sizes[cap + 1 + k] = (int32_t)((isize + 15u) & ~15u);
        sizes[2 * (cap + 1) + k] = (int32_t)bgzf_token_units(isize);
    }
}

// Z2: one warp per zlib array: base64 text -> bytes in the staging buffer
__global__ void __launch_bounds__(256) mzml_b64_kernel(const SpecDesc *specs, const ZArr *list, unsigned long long n_list, const long long *comp_off,
                                                      uint8_t *comp, uint32_t *flags) {
    __shared__ uint8_t lut[256];
    {
        const int c = threadIdx.x;
        uint8_t v = 0x80;
        if (c >= 'A' && c <= 'Z') v = (uint8_t)(c - 'A');
        else if (c >= 'a' && c <= 'z') v = (uint8_t)(c - 'a' + 26);
        else if (c >= '0' && c <= '9') v = (uint8_t)(c - '0' + 52);
        else if (c == '+') v = 62;
        else if (c == '/') v = 63;
        else if (c == '=') v = 0;
        lut[c] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned long long wid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    uint32_t bad = 0;
    for (unsigned long long k = wid; k < n_list; k += nw) {
        const SpecDesc d = specs[list[k].spec];
        const uint8_t *p = d.arr[list[k].which];
        const uint32_t len = d.len[list[k].which], nbytes = b64_bytes(p, len);
        uint8_t *dst = comp + comp_off[k];
        for (uint32_t g = lane; g < len / 4u; g += 32) {
            const uint32_t v0 = lut[p[4 * g]], v1 = lut[p[4 * g + 1]], v2 = lut[p[4 * g + 2]], v3 = lut[p[4 * g + 3]];
            bad |= v0 | v1 | v2 | v3;
            const uint32_t t = ((v0 & 63u) << 18) | ((v1 & 63u) << 12) | ((v2 & 63u) << 6) | (v3 & 63u);
            const uint32_t o = 3u * g;
            if (o < nbytes) dst[o] = (uint8_t)(t >> 16);
            if (o + 1 < nbytes) dst[o + 1] = (uint8_t)(t >> 8);
            if (o + 2 < nbytes) dst[o + 2] = (uint8_t)t;
        }
    }
    bad = __reduce_or_sync(0xFFFFFFFFu, bad);
    if (lane == 0 && (bad & 0x80u)) atomicOr(flags, kMzErrFormat);
}

// Z3: member table of the inflate kernels + the descriptors' pointers to the inflated values.  RFC 1950: CMF (deflate, window
// <= 32 KiB), FLG (check bits, no preset dictionary), DEFLATE data, Adler-32.
__global__ void mzml_ztable_kernel(SpecDesc *specs, const ZArr *list, unsigned long long n_list, const long long *comp_off, const long long *out_off,
                                   const long long *tok_off, const uint8_t *comp, uint8_t *out, BgzfMember *table, uint32_t *flags) {
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_list) return;
    SpecDesc &d = specs[list[k].spec];
    const uint32_t which = list[k].which;
    const uint32_t nbytes = b64_bytes(d.arr[which], d.len[which]);
    const uint32_t isize = zarr_isize(d.n_default, d.f32[which] != 0, nbytes);
    const uint8_t *z = comp + comp_off[k];
    BgzfMember m;
    m.in_off = (uint64_t)comp_off[k] + 2u;
    m.in_len = nbytes >= 6u ? nbytes - 6u : 0u;
    m.isize = isize;
    m.out_addr = (uint64_t)reinterpret_cast<uintptr_t>(out + out_off[k]);
    m.tok_off = (uint32_t)tok_off[k];
    m.pad_ = 0;
    if (nbytes < 6u || (z[0] & 0x0Fu) != 8u || (z[0] >> 4) > 7u || (((uint32_t)z[0] << 8) | z[1]) % 31u != 0u || (z[1] & 0x20u)) {
        atomicOr(flags, kMzErrZlib);
        m.isize = 0;  // skipped by the inflate kernels
    }
    table[k] = m;
    d.raw[which] = out + out_off[k];
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

MzScan::~MzScan() {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (void *p : z_pool)
        if (p) cudaFreeAsync(p, ctx->stream);
}

// rank of the first spectrum whose resident range is >= seg0[f] (descriptors are in file order: seg is non-decreasing)
__global__ void mzml_file_bounds_kernel(const SpecDesc *specs, unsigned long long n_spec, const uint32_t *seg0, int n_files, long long *out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_files) return;
    unsigned long long lo = 0, hi = n_spec;
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (specs[mid].seg < seg0[f]) lo = mid + 1;
        else hi = mid;
    }
    out[f] = (long long)lo;
}

int mzml_scan_spectra(VcfStream *s, MzScan *out) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    out->ctx = ctx;
    out->n_spec = 0;
    out->file_spec0.assign(1, 0);
    std::vector<Piece> pieces;
    s->cut_pieces(pieces);
    if (pieces.empty()) return EXON_GPU_OK;
    std::vector<ScanSeg> h_segs;
    std::vector<uint32_t> file_seg0;  // first resident range of every file
    int64_t n_tiles = 0, n_bytes = 0;
    for (const Piece &p : pieces) {
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)p.base & 15);
        sg.base = p.base - sg.skip;
        sg.len = p.len;
        sg.tile0 = n_tiles;
        sg.pad_ = 0;
        n_tiles += (sg.skip + p.len + MzRing::TILE - 1) / MzRing::TILE;
        n_bytes += p.len;
        if (p.starts_file || h_segs.empty()) file_seg0.push_back((uint32_t)h_segs.size());
        h_segs.push_back(sg);
    }
    if (h_segs.size() >= (1u << 20)) return fail(EXON_GPU_ERR_UNSUPPORTED, "mzml: too many resident ranges");
    ScanSeg sentinel;
    memset(&sentinel, 0, sizeof(sentinel));
    sentinel.tile0 = n_tiles;
    h_segs.push_back(sentinel);

    static int occ = 0;
    if (!occ) {
        CUDA_TRY(cudaFuncSetAttribute(mzml_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MzRing::smem_bytes));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mzml_events_kernel, MzRing::WARPS * 32, MzRing::smem_bytes));
        if (occ < 1) occ = 1;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t cap = (size_t)(attempt == 0 ? n_bytes / 96 + 4096 : n_bytes / 8 + 4096);
        size_t sort_bytes = 0, scan_bytes = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)cap, 0, 64, st));
        CUDA_TRY(exclusive_sum_i32_i64(nullptr, scan_bytes, (const int32_t *)nullptr, (long long *)nullptr, (int)(cap + 1), st));
        const size_t o_segs = 0, o_ev = o_segs + al256(h_segs.size() * sizeof(ScanSeg)), o_ev2 = o_ev + al256(cap * 8), o_sort = o_ev2 + al256(cap * 8),
                     o_flag = o_sort + al256(std::max(sort_bytes, scan_bytes)), o_rank = o_flag + al256((cap + 1) * 4), o_files = o_rank + al256((cap + 1) * 8),
                     o_out = o_files + al256(file_seg0.size() * 4) + al256(file_seg0.size() * 8 + 16);  // d_seg0 (u32 per file) | d_f0 (i64 per file), each on its own 256-byte boundary
        if (int rc = ctx->ensure_scratch(o_out + 256, 256 + file_seg0.size() * 8)) return rc;
        uint8_t *scr = (uint8_t *)ctx->scratch;
        unsigned long long *d_out = (unsigned long long *)(scr + o_out);
        out->d_out = d_out;
        CUDA_TRY(cudaMemcpyAsync(scr + o_segs, h_segs.data(), h_segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(d_out, 0, 64, st));
        MzArgs a;
        a.segs = (const ScanSeg *)(scr + o_segs);
        a.n_segs = (int32_t)h_segs.size() - 1;
        a.n_tiles = n_tiles;
        a.events = (unsigned long long *)(scr + o_ev);
        a.cap = cap;
        a.n_events = d_out;
        int64_t grid = std::min<int64_t>((int64_t)occ * ctx->sm_count, (n_tiles + MzRing::WARPS - 1) / MzRing::WARPS);
        if (grid < 1) grid = 1;
        if (attempt == 0) CUDA_TRY(ctx->timed_begin(st));
        mzml_events_kernel<<<(unsigned)grid, MzRing::WARPS * 32, MzRing::smem_bytes, st>>>(a);
        CUDA_TRY(cudaGetLastError());
        unsigned long long *h = (unsigned long long *)ctx->h_scratch;
        CUDA_TRY(cudaMemcpyAsync(h, d_out, 64, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const unsigned long long n_ev = h[0], n_spec = h[5];
        ctx->launches.fetch_add(1);
        if (n_ev > cap) {
            if (attempt == 0) continue;  // an XML-dense file: go again with room for every possible event
            return fail(EXON_GPU_ERR_UNSUPPORTED, "mzml: event buffer overflow");
        }
        out->file_spec0.assign(file_seg0.size() + 1, 0);
        if (n_ev == 0 || n_spec == 0) return EXON_GPU_OK;
        if (int rc = ctx->ensure_scratch_b((size_t)(n_spec + 1) * sizeof(SpecDesc))) return rc;
        SpecDesc *d_spec = (SpecDesc *)ctx->scratch_b;
        size_t sb = sort_bytes;
        int seg_bits = 1;
        while ((1ull << seg_bits) < h_segs.size()) ++seg_bits;
        // nothing lives above the segment index (the kind bits stay in: <binary></binary> puts two events at one offset)
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(scr + o_sort, sb, (const unsigned long long *)(scr + o_ev), (unsigned long long *)(scr + o_ev2), (int)n_ev, 0,
                                                44 + seg_bits, st));
        // file order: a spectrum's slot is the number of <spectrum> events before it
        int32_t *d_flag = (int32_t *)(scr + o_flag);
        long long *d_rank = (long long *)(scr + o_rank);
        mzml_spec_flags_kernel<<<(unsigned)((n_ev + 256) / 256), 256, 0, st>>>((const unsigned long long *)(scr + o_ev2), n_ev, d_flag);
        size_t tb = scan_bytes;
        CUDA_TRY(exclusive_sum_i32_i64(scr + o_sort, tb, (const int32_t *)d_flag, d_rank, (int)(n_ev + 1), st));
        mzml_spectra_kernel<<<(unsigned)((n_ev + 255) / 256), 256, 0, st>>>((const unsigned long long *)(scr + o_ev2), n_ev, a.segs, d_rank, d_spec,
                                                                              (uint32_t *)(d_out + 4));
        CUDA_TRY(cudaGetLastError());
        // where every file's spectra start
        uint32_t *d_seg0 = (uint32_t *)(scr + o_files);
        long long *d_f0 = (long long *)(scr + o_files + al256(file_seg0.size() * 4));
        CUDA_TRY(cudaMemcpyAsync(d_seg0, file_seg0.data(), file_seg0.size() * 4, cudaMemcpyHostToDevice, st));
        mzml_file_bounds_kernel<<<(unsigned)((file_seg0.size() + 127) / 128), 128, 0, st>>>(d_spec, n_spec, d_seg0, (int)file_seg0.size(), d_f0);
        CUDA_TRY(cudaMemcpyAsync(h, d_out, 64, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h + 32, d_f0, file_seg0.size() * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        ctx->launches.fetch_add(5);
        for (size_t f = 0; f < file_seg0.size(); ++f) out->file_spec0[f] = (long long)h[32 + f];
        out->file_spec0[file_seg0.size()] = (long long)n_spec;
        if ((uint32_t)h[4] & kMzErrFormat) return fail(EXON_GPU_ERR_PARSE, "malformed mzML: a <binary> element without data type / compression / end tag, or invalid base64");
        if ((uint32_t)h[4] & kMzErrZlib)
            return fail(EXON_GPU_ERR_PARSE, "mzml: a zlib-compressed binary array (MS:1000574) without defaultArrayLength on its <spectrum>");
        if ((uint32_t)h[4] & kMzHasZlib) {
            // ---- zlib detour: list -> sizes -> scans -> base64 decode -> member table -> inflate ----
            const unsigned long long zcap = 3 * n_spec;
            size_t cub2 = 0;
            CUDA_TRY(exclusive_sum_i32_i64(nullptr, cub2, (const int32_t *)nullptr, (long long *)nullptr, (int)(zcap + 1), st));
            const size_t o_list = 0, o_sizes = o_list + al256(zcap * sizeof(ZArr)), o_offs = o_sizes + al256(3 * (zcap + 1) * 4), o_cub = o_offs + al256(3 * (zcap + 1) * 8),
                         o_tab = o_cub + al256(cub2), o_misc = o_tab + al256(zcap * sizeof(BgzfMember)), z_bytes = o_misc + 256;
            CUDA_TRY(cudaMallocAsync(&out->z_pool[0], z_bytes, st));
            uint8_t *zb = (uint8_t *)out->z_pool[0];
            ZArr *d_list = (ZArr *)(zb + o_list);
            int32_t *d_sizes = (int32_t *)(zb + o_sizes);
            long long *d_offs = (long long *)(zb + o_offs);
            unsigned long long *d_zmisc = (unsigned long long *)(zb + o_misc);  // [0] n_list | [2] inflate flags (2 x u32)
            CUDA_TRY(cudaMemsetAsync(d_sizes, 0, 3 * (zcap + 1) * 4, st));
            CUDA_TRY(cudaMemsetAsync(d_zmisc, 0, 64, st));
            mzml_zlist_kernel<<<(unsigned)((n_spec + 255) / 256), 256, 0, st>>>(d_spec, n_spec, d_list, d_zmisc, d_sizes, zcap);
            for (int q = 0; q < 3; ++q) {
                size_t tb2 = cub2;
                CUDA_TRY(exclusive_sum_i32_i64(zb + o_cub, tb2, (const int32_t *)(d_sizes + q * (zcap + 1)), d_offs + q * (zcap + 1), (int)(zcap + 1), st));
            }
            long long h_tot[3];
            unsigned long long h_nz = 0;
            for (int q = 0; q < 3; ++q) CUDA_TRY(cudaMemcpyAsync(&h_tot[q], d_offs + q * (zcap + 1) + zcap, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(&h_nz, d_zmisc, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            ctx->launches.fetch_add(5);
            if (h_nz > zcap) return fail(EXON_GPU_ERR_STATE, "mzml: zlib array list overflow");
            if (h_nz) {
                const size_t comp_total = (size_t)h_tot[0] + 512, out_total = (size_t)h_tot[1] + 64;
                CUDA_TRY(cudaMallocAsync(&out->z_pool[1], al256(comp_total) + out_total, st));
                uint8_t *d_comp = (uint8_t *)out->z_pool[1], *d_raw = d_comp + al256(comp_total);
                const int b64_grid = (int)std::min<unsigned long long>((h_nz + 7) / 8, (unsigned long long)ctx->sm_count * 8);
                mzml_b64_kernel<<<b64_grid, 256, 0, st>>>(d_spec, d_list, h_nz, d_offs, d_comp, (uint32_t *)(d_out + 4));
                mzml_ztable_kernel<<<(unsigned)((h_nz + 127) / 128), 128, 0, st>>>(d_spec, d_list, h_nz, d_offs, d_offs + (zcap + 1), d_offs + 2 * (zcap + 1), d_comp,
                                                                                    d_raw, (BgzfMember *)(zb + o_tab), (uint32_t *)(d_out + 4));
                CUDA_TRY(cudaGetLastError());
                const int init_flags[2] = {0, 0x7FFFFFFF};
                CUDA_TRY(cudaMemcpyAsync(d_zmisc + 2, init_flags, sizeof(init_flags), cudaMemcpyHostToDevice, st));
                if (int rc = bgzf_inflate_launch(ctx, d_comp, (const BgzfMember *)(zb + o_tab), (int)h_nz, (uint32_t *)(d_zmisc + 2), (size_t)h_tot[2], (size_t)h_tot[0]))
                    return rc;
                ctx->launches.fetch_add(2);
                uint32_t h_inf[2] = {0, 0};
                CUDA_TRY(cudaMemcpyAsync(h_inf, d_zmisc + 2, 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaMemcpyAsync(h, d_out, 64, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                if (h_inf[0])
                    return fail(EXON_GPU_ERR_PARSE, "mzml: a zlib-compressed binary array does not inflate:%s%s", (h_inf[0] & 1u) ? " invalid DEFLATE data;" : "",
                                (h_inf[0] & 2u) ? " its size differs from defaultArrayLength x value width;" : "");
                if ((uint32_t)h[4] & kMzErrZlib) return fail(EXON_GPU_ERR_PARSE, "mzml: a zlib-compressed binary array with an invalid zlib header");
                if ((uint32_t)h[4] & kMzErrFormat) return fail(EXON_GPU_ERR_PARSE, "malformed mzML: invalid base64 in a zlib-compressed binary array");
            }
        }
        out->n_spec = n_spec;
        out->d_spec = d_spec;
        return EXON_GPU_OK;
    }
    return fail(EXON_GPU_ERR_STATE, "mzml: unreachable");
}

int mzml_filter_sum(VcfStream *s, const exon_gpu_mzml_pred *pred, double *out_sum, int64_t *out_selected, int64_t *out_spectra) {
    if (int rc = s->flush_gz()) return rc;
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    std::lock_guard<std::recursive_mutex> work(ctx->work_mu);
    if (out_sum) *out_sum = 0.0;
    if (out_selected) *out_selected = 0;
    if (out_spectra) *out_spectra = 0;
    MzScan sc;
    if (int rc = mzml_scan_spectra(s, &sc)) return rc;
    if (!sc.n_spec) return EXON_GPU_OK;
    unsigned long long *d_out = sc.d_out;
    unsigned long long *h = (unsigned long long *)ctx->h_scratch;
    const int sum_grid = (int)std::min<unsigned long long>((sc.n_spec + 7) / 8, (unsigned long long)ctx->sm_count * 8);
    mzml_sum_kernel<<<sum_grid, 256, 0, st>>>(sc.d_spec, sc.n_spec, pred != nullptr, pred ? pred->mz_lo : 0.0, pred ? pred->mz_hi : 0.0, (double *)(d_out + 2),
                                               d_out + 3, (uint32_t *)(d_out + 4));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(ctx->timed_end(st));
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaMemcpyAsync(h, d_out, 64, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint32_t flags = (uint32_t)h[4];
    if (flags & kMzErrFormat) return fail(EXON_GPU_ERR_PARSE, "malformed mzML: a <binary> element without data type / compression / end tag, or invalid base64");
    if (out_sum) memcpy(out_sum, &h[2], 8);
    if (out_selected) *out_selected = (int64_t)h[3];
    if (out_spectra) *out_spectra = (int64_t)sc.n_spec;
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_mzml_open(exon_gpu_ctx *c, exon_gpu_stream **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "mzml_open: NULL argument");
    exon_gpu_vcf_opts vo;
    memset(&vo, 0, sizeof(vo));
    if (int rc = exon_gpu_vcf_open(c, &vo, out)) return rc;
    (*out)->fmt = kFmtMzml;
    (*out)->hdr = VcfStream::kBody;
    return EXON_GPU_OK;
}

int exon_gpu_mzml_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last) {
    if (!s || s->fmt != kFmtMzml) return fail(EXON_GPU_ERR_ARG, "mzml_feed: not an mzML stream");
    return exon_gpu_vcf_feed(s, text, len, is_device_ptr, is_last);
}

int exon_gpu_mzml_filter_sum(exon_gpu_stream *s, const exon_gpu_mzml_pred *pred, double *out_sum, int64_t *out_selected, int64_t *out_spectra) {
    if (!s || s->fmt != kFmtMzml) return fail(EXON_GPU_ERR_ARG, "mzml_filter_sum: not an mzML stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return mzml_filter_sum(s, pred, out_sum, out_selected, out_spectra);
}

}  // extern "C"

// ---- record batches (exon_gpu_mzml_next_batch) are built in mzml_columns.cu ----

