// vcf_wide.cu -- VCF text -> Arrow columns 2..6 {id: list<utf8>, ref: utf8, alt: list<utf8>, qual: f32, filter: list<utf8>}.
//
// Replaces LazyVCFArrayBuilder::append / finish for those columns (exon/exon-vcf/src/array_builder/
// lazy_array_builder.rs:169-216, 451-484), with the lazy builder's own quirks kept:
//   id      "." -> NULL, else one list item per ';'-separated id                                      (:169-179)
//   ref     the bases, byte for byte                                                                   (:180-189)
//   alt     "." -> NULL, else a VALID BUT EMPTY list: the builder never appends the alleles            (:190-204)
//   qual    "." -> NULL, else Rust `f32::from_str`, correctly rounded (f32_parse.cuh)                  (:205-208)
//   filter  always valid: "." -> [], else one item per ';'-separated filter                            (:209-216)
//   info    (string mode) the field re-serialised: equal to its text plus "=true" after every flag, see info_walk      (:217-298)
// Columns 0 / 1 stay with K2 (vcf_columns.cu), whose batch table (batches restart at every file) this build shares.
//
// Row-parallel, on top of the partition's line index (build_line_index, fastq_scan.cu):
//   1. measure  one thread per record: walk to the 7th tab, item / byte counts of id, ref, filter, QUAL -> f32, validity flags
//   2. 5 exclusive scans (cub): global item / byte offsets of every row
//   3. emit     one thread per record: batch-relative int32 list offsets, child offsets and bytes at their final place,
//               validity bits; the batch of a row is found by binary search over the batch table
// Batches are zero-copy views of the store.  The child arrays of batch b restart at 0: its child offsets live at
// coff[e0(b) + b .. e0(b + 1) + b] (one extra entry per batch), its bytes at val[v0(b) ..].
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstring>
#include <new>

#include "common.cuh"
#include "f32_parse.cuh"
#include "internal.h"
#include "scan_i64.cuh"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

constexpr uint32_t kWErrFields = 1u;       // fewer than 8 tab-separated fields
constexpr uint32_t kWErrQual = 2u;         // QUAL is not a float literal
constexpr uint32_t kWErrQualDigits = 4u;   // QUAL has more than 36 significant digits
constexpr uint32_t kWErrFieldLen = 8u;     // a line of 2 GiB or more
constexpr uint32_t kWErrInfoValue = 16u;   // INFO: a non-flag key without a value (the reference unwraps a None there)
constexpr uint32_t kWErrInfoKey = 32u;     // INFO: a key the header does not define
constexpr uint32_t kWErrInfoForm = 64u;    // INFO: a value whose re-serialisation by the reference would differ from its text

enum { kIdE = 0, kIdB = 1, kRefB = 2, kFiE = 3, kFiB = 4, kInfoB = 5, kNScan = 6 };

struct WideArgs {
    int64_t n_rows;
    const uint8_t *const *line_start;
    const uint8_t *const *line_end;
    const long long *brow;  // n_batches + 1
    int64_t n_batches;
    int32_t batch_rows, wpb;
    int32_t want_id, want_ref, want_alt, want_qual, want_filter, want_info;
    // INFO definitions of the header (exon_gpu_vcf_set_header): FNV-1a hash, offset / length into the blob, type
    const uint32_t *info_hash;
    const int32_t *info_off, *info_len;
    const uint8_t *info_type, *info_blob;
    int32_t n_info;
    int32_t *info_offs;  // batch-relative offsets, n_batches * (batch_rows + 1)
    uint8_t *info_val;
    int32_t *cnt[kNScan];          // measure out, n_rows + 1 entries (the last one 0); NULL when not wanted
    const long long *pre[kNScan];  // emit in: exclusive scans of cnt
    uint8_t *rowflags;             // measure out: bit0 id valid, bit1 alt valid, bit2 qual valid
    float *qual;
    int32_t *id_loff, *id_coff, *ref_off, *fi_loff, *fi_coff;
    uint8_t *id_val, *ref_val, *fi_val;
    uint32_t *id_valid, *alt_valid, *qual_valid;
    uint32_t *flags;
    unsigned long long *first_bad_row;
};

// Offsets (from ls) of tabs 0..6 of the line [ls, le); false when the line has fewer than 8 fields.  Aligned 8-byte loads,
// exact SWAR tab flags, bytes outside the line masked off; the seven offsets stay in registers (static indices only).
__device__ __forceinline__ bool find_tabs(const uint8_t *ls, const uint8_t *le, int32_t t[7]) {
    const uint8_t *wp = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(ls) & ~(uintptr_t)7);
    int32_t off = (int32_t)(wp - ls);  // <= 0
    int nt = 0;
    while (wp < le) {
        const unsigned long long w = __ldg(reinterpret_cast<const unsigned long long *>(wp));
        const unsigned long long x = w ^ 0x0909090909090909ull;
        const unsigned long long y = (x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full;
        unsigned long long m = ~(y | x | 0x7F7F7F7F7F7F7F7Full);  // 0x80 in every byte that is a tab
        if (off < 0) m &= ~0ull << (8 * -off);
        const long long rem = le - wp;
        if (rem < 8) m &= (1ull << (8 * rem)) - 1ull;
        while (m) {
            const int32_t p = off + ((__ffsll((long long)m) - 1) >> 3);
#pragma unroll
            for (int k = 0; k < 7; ++k)
                if (nt == k) t[k] = p;
            if (++nt == 7) return true;
            m &= m - 1ull;
        }
        wp += 8;
        off += 8;
    }
    return false;
}

__device__ __forceinline__ bool is_missing(const uint8_t *f, int64_t n) { return n == 0 || (n == 1 && __ldg(f) == '.'); }

// items = 1 + number of ';', bytes = n - (items - 1)
__device__ __forceinline__ void count_items(const uint8_t *f, int32_t n, int32_t &items, int32_t &bytes) {
    int32_t semi = 0;
    for (int32_t i = 0; i < n; ++i) semi += __ldg(f + i) == ';';
    items = semi + 1;
    bytes = n - semi;
}


// ---- INFO (column 7, string mode) ---------------------------------------------------------------------------------------
// LazyVCFArrayBuilder::append :217-298 does not copy the field: it walks noodles' typed view of it (types from the header's
// ##INFO lines) and prints every entry again -- `key=value` joined by ';', flags as `key=true`, numbers through Rust's
// Display.  The text that comes out equals the text that went in, plus "=true" after every flag, PROVIDED every number is
// already in the form Display would give it.  That is checked here entry by entry; a value for which it does not hold (an
// exponent, trailing zeros, more than 6 significant digits in a fraction, '%' escapes in a string ...), a key the header
// does not define and a flag with a value are refused (EXON_GPU_ERR_UNSUPPORTED) instead of being approximated.
enum : uint8_t { kInfoInteger = 0, kInfoFloat = 1, kInfoFlag = 2, kInfoCharacter = 3, kInfoString = 4 };

__device__ __forceinline__ uint32_t fnv1a(const uint8_t *p, int n) {
    uint32_t h = 2166136261u;
    for (int i = 0; i < n; ++i) h = (h ^ __ldg(p + i)) * 16777619u;
    return h;
}

// -?(0|[1-9][0-9]*) within i32; "." is a missing element
__device__ __forceinline__ bool canon_int(const uint8_t *p, int n) {
    if (n == 1 && __ldg(p) == '.') return true;
    int i = 0;
    const bool neg = n > 0 && __ldg(p) == '-';
    if (neg) i = 1;
    const int nd = n - i;
    if (nd < 1 || nd > 10) return false;
    if (__ldg(p + i) == '0') return nd == 1 && !neg;
    unsigned long long v = 0;
    for (; i < n; ++i) {
        const uint32_t d = (uint32_t)__ldg(p + i) - '0';
        if (d > 9u) return false;
        v = v * 10ull + d;
    }
    return v <= (neg ? 2147483648ull : 2147483647ull);
}

// the text Rust's Display prints for the f32 nearest to it: an integer of at most 7 digits (exact below 2^24), or a plain
// decimal without trailing zeros whose significant digits number at most 6 (FLT_DIG: such decimals survive the round trip,
// so the shortest representation of their f32 is the decimal itself); no sign on zero, no exponent, no leading '+'
__device__ __forceinline__ bool canon_float(const uint8_t *p, int n) {
    if (n == 1 && __ldg(p) == '.') return true;
    int i = 0;
    const bool neg = n > 0 && __ldg(p) == '-';
    if (neg) i = 1;
    if (i >= n) return false;
    int int_digits = 0, sig = 0;
    bool nonzero = false, lead_zero = false;
    const int i0 = i;
    for (; i < n; ++i) {
        const uint32_t d = (uint32_t)__ldg(p + i) - '0';
        if (d > 9u) break;
        if (int_digits == 0 && d == 0) lead_zero = true;
        ++int_digits;
        if (nonzero || d) {
            nonzero = true;
            ++sig;
        }
    }
    if (int_digits == 0 || (lead_zero && int_digits > 1)) return false;
    (void)i0;
    if (i == n) return nonzero ? sig <= 7 : !neg;  // integer: "0" but not "-0"
    if (__ldg(p + i) != '.' || i + 1 >= n) return false;
    ++i;
    uint32_t last = 0;
    for (; i < n; ++i) {
        last = (uint32_t)__ldg(p + i) - '0';
        if (last > 9u) return false;
        if (nonzero || last) {
            nonzero = true;
            ++sig;
        }
    }
    return last != 0u && sig <= 6;
}

// Walks one INFO field.  Returns the length of the re-serialised string; *err collects kWErrInfo*.  When `dst` is set the
// string is written there.
__device__ __forceinline__ int32_t info_walk(const WideArgs &a, const uint8_t *f, int32_t n, uint8_t *dst, uint32_t *err) {
    if (n == 0 || (n == 1 && __ldg(f) == '.')) return 0;
    int32_t out = 0, i = 0;
    while (i <= n) {
        // entry [i, e)
        int32_t e = i, eq = -1;
        while (e < n && __ldg(f + e) != ';') {
            if (eq < 0 && __ldg(f + e) == '=') eq = e;
            ++e;
        }
        const int32_t klen = (eq < 0 ? e : eq) - i;
        const uint32_t h = fnv1a(f + i, klen);
        int type = -1;
        for (int k = 0; k < a.n_info; ++k) {
            if (a.info_hash[k] != h || a.info_len[k] != klen) continue;
            bool same = true;
            for (int q = 0; q < klen && same; ++q) same = a.info_blob[a.info_off[k] + q] == __ldg(f + i + q);
            if (same) {
                type = a.info_type[k];
                break;
            }
        }
        if (type < 0) *err |= kWErrInfoKey;
        if (type == kInfoFlag) {
            if (eq >= 0) *err |= kWErrInfoForm;
        } else if (type >= 0) {
            if (eq < 0) {
                *err |= kWErrInfoValue;
            } else {
                // elements of the value
                int32_t v = eq + 1;
                if (v == e) *err |= kWErrInfoForm;
                while (v <= e && v < e + 1) {
                    int32_t ve = v;
                    while (ve < e && __ldg(f + ve) != ',') ++ve;
                    const uint8_t *p = f + v;
                    const int32_t m = ve - v;
                    bool ok = true;
                    if (type == kInfoInteger) ok = canon_int(p, m);
                    else if (type == kInfoFloat) ok = canon_float(p, m);
                    else if (type == kInfoCharacter) ok = m == 1;
                    else
                        for (int32_t q = 0; q < m && ok; ++q) ok = __ldg(p + q) != '%';
                    if (!ok || m == 0) *err |= kWErrInfoForm;
                    if (ve >= e) break;
                    v = ve + 1;
                }
            }
        }
        // key[=value], "=true" for a flag, ';' between entries
        if (dst) {
            for (int32_t q = i; q < e; ++q) dst[out + (q - i)] = __ldg(f + q);
            if (type == kInfoFlag && eq < 0) {
                dst[out + (e - i)] = '=';
                dst[out + (e - i) + 1] = 't';
                dst[out + (e - i) + 2] = 'r';
                dst[out + (e - i) + 3] = 'u';
                dst[out + (e - i) + 4] = 'e';
            }
        }
        out += e - i;
        if (type == kInfoFlag && eq < 0) out += 5;
        if (e >= n) break;
        if (dst) dst[out] = ';';
        ++out;
        i = e + 1;
    }
    return out;
}

// end of the INFO field: the 8th tab, or the end of the line when the record has no FORMAT / sample columns
__device__ __forceinline__ const uint8_t *info_end(const uint8_t *info, const uint8_t *le) {
    const uint8_t *p = info;
    while (p < le && __ldg(p) != '\t') ++p;
    return p;
}

__global__ void __launch_bounds__(256) vw_measure_kernel(const __grid_constant__ WideArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= a.n_rows) return;
    const uint8_t *ls = a.line_start[r], *le = a.line_end[r];
    int32_t to[7];
    uint32_t err = 0;
    int32_t c[kNScan] = {0, 0, 0, 0, 0, 0};
    uint8_t rf = 0;
    float q = 0.0f;
    if (le - ls > 0x7FFFFFF0ll) {
        err = kWErrFieldLen;
    } else if (!find_tabs(ls, le, to)) {
        err = kWErrFields;
    } else {
        const uint8_t *tab[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) tab[k] = ls + to[k];
        if (a.want_id) {
            const uint8_t *f = tab[1] + 1;
            const int32_t n = (int32_t)(tab[2] - f);
            if (!is_missing(f, n)) {
                rf |= 1u;
                count_items(f, n, c[kIdE], c[kIdB]);
            }
        }
        if (a.want_ref) c[kRefB] = (int32_t)(tab[3] - tab[2] - 1);
        if (a.want_alt && !is_missing(tab[3] + 1, tab[4] - tab[3] - 1)) rf |= 2u;
        if (a.want_qual) {
            const uint8_t *f = tab[4] + 1;
            const int64_t n = tab[5] - f;
            if (!(n == 1 && __ldg(f) == '.')) {
                // the usual QUAL is a short unsigned integer: exact in f32 below 2^24; everything else goes to the parser
                uint32_t v = 0;
                bool plain = n >= 1 && n <= 7;
                for (int i = 0; plain && i < (int)n; ++i) {
                    const uint32_t d = (uint32_t)__ldg(f + i) - '0';
                    plain = d <= 9u;
                    v = v * 10u + d;
                }
                if (plain) {
                    q = (float)v;
                    rf |= 4u;
                } else {
                    rf |= 8u;  // left to vw_qual_kernel
                }
            }
        }
        if (a.want_filter) {
            const uint8_t *f = tab[5] + 1;
            const int32_t n = (int32_t)(tab[6] - f);
            if (!is_missing(f, n)) count_items(f, n, c[kFiE], c[kFiB]);
        }
    }
#pragma unroll
    for (int k = 0; k < kNScan; ++k)
        if (k != kInfoB && a.cnt[k]) a.cnt[k][r] = c[k];  // kInfoB belongs to vw_info_measure_kernel
    a.rowflags[r] = rf;
    if (a.want_qual) {
        a.qual[r] = q;
        // rows left to vw_qual_kernel: one atomic per warp (the host launches that kernel only when the count is not zero)
        const uint32_t pend = __ballot_sync(__activemask(), (rf & 8u) != 0u);
        if (pend && (threadIdx.x & 31) == __ffs((int)pend) - 1) atomicAdd(a.first_bad_row + 1, (unsigned long long)__popc(pend));
    }
    if (err) {
        atomicOr(a.flags, err);
        atomicMin(a.first_bad_row, (unsigned long long)r);
    }
}

// QUAL of the rows the measure pass did not settle (anything but a short unsigned integer): the exact parser
// (128-bit digits, 512-bit comparisons) lives in its own kernel so that it does not shape the register budget of the row pass.
__global__ void __launch_bounds__(256) vw_qual_kernel(const __grid_constant__ WideArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= a.n_rows) return;
    const uint8_t rf = a.rowflags[r];
    if (!(rf & 8u)) return;
    const uint8_t *ls = a.line_start[r], *le = a.line_end[r];
    int32_t to[7];
    if (!find_tabs(ls, le, to)) return;
    const uint8_t *f = ls + to[4] + 1;
    const int64_t n = to[5] - to[4] - 1;
    float q = 0.0f;
    const int rc = n > 4096 ? kF32Malformed : parse_f32_rust(f, (int)n, &q);
    if (rc == kF32Ok) {
        a.qual[r] = q;
        a.rowflags[r] = (uint8_t)((rf & ~8u) | 4u);
    } else {
        atomicOr(a.flags, rc == kF32Unsupported ? kWErrQualDigits : kWErrQual);
        atomicMin(a.first_bad_row, (unsigned long long)r);
    }
}

// writes the items of one list cell: child offsets (relative to the batch's first byte) and bytes
__device__ __forceinline__ void emit_items(const uint8_t *f, int32_t n, int32_t *coff, uint8_t *val, long long v_abs, long long v_rel) {
    int32_t k = 0;
    coff[0] = (int32_t)v_rel;
    long long w = 0;
    for (int32_t i = 0; i < n; ++i) {
        const uint8_t ch = __ldg(f + i);
        if (ch == ';') coff[++k] = (int32_t)(v_rel + w);
        else val[v_abs + w++] = ch;
    }
}

__global__ void __launch_bounds__(256) vw_emit_kernel(const __grid_constant__ WideArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    // batch of the row: last b with brow[b] <= r.  Rows of a block are consecutive: one binary search per block, then a
    // short walk (a block spans more than two batches only when batches are tiny)
    __shared__ long long s_b0;
    if (threadIdx.x == 0) {
        const int64_t rb = (int64_t)blockIdx.x * 256;
        int64_t lo = 0, hi = a.n_batches;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(&a.brow[mid]) <= rb) lo = mid;
            else hi = mid;
        }
        s_b0 = lo;
    }
    __syncthreads();
    if (r >= a.n_rows) return;
    int64_t b = s_b0;
    while (__ldg(&a.brow[b + 1]) <= r) ++b;
    const int64_t r0 = __ldg(&a.brow[b]);
    const int in_batch = (int)(r - r0);
    const bool last = r + 1 == __ldg(&a.brow[b + 1]);
    const uint8_t rf = a.rowflags[r];
    const uint32_t bit = 1u << (in_batch & 31);
    const int64_t word = b * a.wpb + (in_batch >> 5);
    {
        // validity bits: the lanes of a warp hold consecutive rows, so they fall into one or two bitmap words; one atomic per
        // word and column instead of one per row
        const uint32_t peers = __match_any_sync(__activemask(), word);
        const bool leader = (threadIdx.x & 31) == __ffs((int)peers) - 1;
        if (a.want_alt) {
            const uint32_t v = __reduce_or_sync(peers, (rf & 2u) ? bit : 0u);
            if (leader && v) atomicOr(a.alt_valid + word, v);
        }
        if (a.want_qual) {
            const uint32_t v = __reduce_or_sync(peers, (rf & 4u) ? bit : 0u);
            if (leader && v) atomicOr(a.qual_valid + word, v);
        }
        if (a.want_id) {
            const uint32_t v = __reduce_or_sync(peers, (rf & 1u) ? bit : 0u);
            if (leader && v) atomicOr(a.id_valid + word, v);
        }
    }
    if (!(a.want_id || a.want_ref || a.want_filter)) return;
    const uint8_t *ls = a.line_start[r], *le = a.line_end[r];
    int32_t to[7];
    if (le - ls > 0x7FFFFFF0ll || !find_tabs(ls, le, to)) return;  // reported by the measure pass
    const uint8_t *tab[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) tab[k] = ls + to[k];
    const int64_t lrow = b * (int64_t)(a.batch_rows + 1) + in_batch;
    if (a.want_id) {
        const long long e = a.pre[kIdE][r], e0 = a.pre[kIdE][r0], v = a.pre[kIdB][r], v0 = a.pre[kIdB][r0];
        a.id_loff[lrow] = (int32_t)(e - e0);
        int32_t *coff = a.id_coff + e0 + b;
        if (rf & 1u) emit_items(tab[1] + 1, (int32_t)(tab[2] - tab[1] - 1), coff + (e - e0), a.id_val, v, v - v0);
        if (last) {
            a.id_loff[lrow + 1] = (int32_t)(a.pre[kIdE][r + 1] - e0);
            coff[a.pre[kIdE][r + 1] - e0] = (int32_t)(a.pre[kIdB][r + 1] - v0);
        }
    }
    if (a.want_ref) {
        const long long v = a.pre[kRefB][r], v0 = a.pre[kRefB][r0];
        a.ref_off[lrow] = (int32_t)(v - v0);
        const uint8_t *f = tab[2] + 1;
        const int32_t n = (int32_t)(tab[3] - f);
        for (int32_t i = 0; i < n; ++i) a.ref_val[v + i] = __ldg(f + i);
        if (last) a.ref_off[lrow + 1] = (int32_t)(a.pre[kRefB][r + 1] - v0);
    }
    if (a.want_filter) {
        const long long e = a.pre[kFiE][r], e0 = a.pre[kFiE][r0], v = a.pre[kFiB][r], v0 = a.pre[kFiB][r0];
        a.fi_loff[lrow] = (int32_t)(e - e0);
        int32_t *coff = a.fi_coff + e0 + b;
        const uint8_t *f = tab[5] + 1;
        const int32_t n = (int32_t)(tab[6] - f);
        if (!is_missing(f, n)) emit_items(f, n, coff + (e - e0), a.fi_val, v, v - v0);
        if (last) {
            a.fi_loff[lrow + 1] = (int32_t)(a.pre[kFiE][r + 1] - e0);
            coff[a.pre[kFiE][r + 1] - e0] = (int32_t)(a.pre[kFiB][r + 1] - v0);
        }
    }
}

// INFO lives in its own two kernels (the entry walk with its key lookups must not shape the register budget of the row
// passes): length + checks, then -- after the scan -- offsets and bytes.
__global__ void __launch_bounds__(256) vw_info_measure_kernel(const __grid_constant__ WideArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= a.n_rows) return;
    const uint8_t *ls = a.line_start[r], *le = a.line_end[r];
    int32_t to[7];
    int32_t n = 0;
    uint32_t err = 0;
    if (le - ls <= 0x7FFFFFF0ll && find_tabs(ls, le, to)) {  // a short line is reported by the row pass
        const uint8_t *f = ls + to[6] + 1;
        n = info_walk(a, f, (int32_t)(info_end(f, le) - f), nullptr, &err);
    }
    a.cnt[kInfoB][r] = n;
    if (err) {
        atomicOr(a.flags, err);
        atomicMin(a.first_bad_row, (unsigned long long)r);
    }
}

__global__ void __launch_bounds__(256) vw_info_emit_kernel(const __grid_constant__ WideArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    __shared__ long long s_b0;
    if (threadIdx.x == 0) {
        const int64_t rb = (int64_t)blockIdx.x * 256;
        int64_t lo = 0, hi = a.n_batches;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(&a.brow[mid]) <= rb) lo = mid;
            else hi = mid;
        }
        s_b0 = lo;
    }
    __syncthreads();
    if (r >= a.n_rows) return;
    int64_t b = s_b0;
    while (__ldg(&a.brow[b + 1]) <= r) ++b;
    const int64_t r0 = __ldg(&a.brow[b]);
    const int in_batch = (int)(r - r0);
    const int64_t lrow = b * (int64_t)(a.batch_rows + 1) + in_batch;
    const long long v = a.pre[kInfoB][r], v0 = a.pre[kInfoB][r0];
    a.info_offs[lrow] = (int32_t)(v - v0);
    if (r + 1 == __ldg(&a.brow[b + 1])) a.info_offs[lrow + 1] = (int32_t)(a.pre[kInfoB][r + 1] - v0);
    const uint8_t *ls = a.line_start[r], *le = a.line_end[r];
    int32_t to[7];
    if (le - ls > 0x7FFFFFF0ll || !find_tabs(ls, le, to)) return;
    const uint8_t *f = ls + to[6] + 1;
    uint32_t ignored = 0;
    info_walk(a, f, (int32_t)(info_end(f, le) - f), a.info_val + v, &ignored);
}

__global__ void vw_gather_i64(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

size_t al256w(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

struct WideBuf {
    void *d = nullptr, *h = nullptr;
    size_t bytes = 0;
};

struct WideStore {
    int device = 0;
    bool on_device = false;
    int batch_rows = 8192, wpb = 256;
    int64_t n_batches = 0, n_rows = 0;
    bool want[9] = {false, false, false, false, false, false, false, false, false};
    WideBuf id_loff, id_coff, id_val, id_valid, ref_off, ref_val, alt_valid, zeros, qual, qual_valid, fi_loff, fi_coff, fi_val, info_off, info_val, info_tab;
    std::vector<long long> batch_row0, base[kNScan];  // per batch (+ total): global item / byte offset of the batch's first row
    static constexpr int kBufs = 16;
    void all(WideBuf *out[kBufs]) {
        WideBuf *v[kBufs] = {&id_loff, &id_coff, &id_val, &id_valid, &ref_off, &ref_val, &alt_valid, &zeros, &qual, &qual_valid, &fi_loff, &fi_coff, &fi_val, &info_off, &info_val, &info_tab};
        for (int i = 0; i < kBufs; ++i) out[i] = v[i];
    }
    template <class T>
    const T *p(const WideBuf &b) const { return static_cast<const T *>(on_device ? b.d : b.h); }
};

void wide_free(WideStore *w) {
    if (!w) return;
    cudaSetDevice(w->device);
    WideBuf *b[WideStore::kBufs];
    w->all(b);
    for (int i = 0; i < WideStore::kBufs; ++i) {
        cudaFree(b[i]->d);
        cudaFreeHost(b[i]->h);
    }
    delete w;
}

bool wide_wanted(const std::vector<int> &projection) {
    for (int p : projection)
        if (p >= 2) return true;
    return false;
}

// The caller (build_columns) holds ctx->work_mu.  *n_rows < 0: no K2 column is projected and the batch table is not known
// yet; it is derived here from the line index (batches restart at every file) and handed back.
int wide_build(VcfStream *s, std::vector<long long> *batch_row0, int64_t *n_rows_io, WideStore **out) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto *w = new (std::nothrow) WideStore();
    if (!w) return fail(EXON_GPU_ERR_OOM, "next_batch: out of host memory");
    *out = w;
    w->device = ctx->device;
    w->on_device = s->columns_on_device;
    w->batch_rows = s->batch_rows;
    w->wpb = ((s->batch_rows + 63) / 64) * 2;
    for (int p : s->projection) w->want[p] = true;
    if (w->want[8]) return fail(EXON_GPU_ERR_UNSUPPORTED, "vcf_next_batch: column 8 (formats) is re-serialised by the reference builder and is not built on the device yet");
    if (w->want[7] && !s->info_defs.set)
        return fail(EXON_GPU_ERR_STATE, "vcf_next_batch: the info column needs the header's ##INFO definitions: call exon_gpu_vcf_set_header first");
    if (*n_rows_io == 0) return EXON_GPU_OK;

    // per-row temporaries in scratch_b behind the line tables: 5 counts (i32) | 5 prefixes (i64) | flags
    const size_t per_line = kNScan * 4 + kNScan * 8 + 1;
    LineIndex li;
    if (int rc = build_line_index(s, per_line, (2 * kNScan + 4) * 256 + (1 << 20), &li)) return rc;
    if (*n_rows_io < 0) {
        batch_row0->clear();
        for (size_t f = 0; f + 1 < li.file_line0.size(); ++f)
            for (long long r = li.file_line0[f]; r < li.file_line0[f + 1]; r += w->batch_rows) batch_row0->push_back(r);
        batch_row0->push_back(li.n_lines);
        *n_rows_io = li.n_lines;
    }
    const int64_t n_rows = *n_rows_io;
    if (li.n_lines != n_rows) return fail(EXON_GPU_ERR_STATE, "vcf_next_batch: line index has %lld lines for %lld rows", (long long)li.n_lines, (long long)n_rows);
    w->n_rows = n_rows;
    w->n_batches = (int64_t)batch_row0->size() - 1;
    w->batch_row0 = *batch_row0;
    if (n_rows == 0) return EXON_GPU_OK;
    const size_t nb1 = (size_t)w->n_batches + 1, nr1 = (size_t)n_rows + 1;
    size_t cub_bytes = 0;
    CUDA_TRY(exclusive_sum_i32_i64(nullptr, cub_bytes, (const int32_t *)nullptr, (long long *)nullptr, (int)nr1, st));
    uint8_t *x = li.extra;
    auto take = [&](size_t bytes) {
        uint8_t *p = x;
        x += al256w(bytes);
        return p;
    };
    const bool need[kNScan] = {w->want[2], w->want[2], w->want[3], w->want[6], w->want[6], w->want[7]};
    WideArgs a;
    memset(&a, 0, sizeof(a));
    a.n_rows = n_rows;
    a.line_start = li.line_start;
    a.line_end = li.line_end;
    a.n_batches = w->n_batches;
    a.batch_rows = w->batch_rows;
    a.wpb = w->wpb;
    a.want_id = w->want[2], a.want_ref = w->want[3], a.want_alt = w->want[4], a.want_qual = w->want[5], a.want_filter = w->want[6];
    a.want_info = w->want[7];
    long long *pre[kNScan] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < kNScan; ++k) {
        if (!need[k]) continue;
        a.cnt[k] = (int32_t *)take(nr1 * 4);
        pre[k] = (long long *)take(nr1 * 8);
        a.pre[k] = pre[k];
        CUDA_TRY(cudaMemsetAsync(a.cnt[k] + n_rows, 0, 4, st));
    }
    a.rowflags = take(nr1);
    // scan scratch grows with the row count (64-bit tile states): from the pool, not from the fixed part of scratch_b
    uint8_t *cub_tmp = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&cub_tmp, cub_bytes + 256, st));
    struct CubFree {
        void *p;
        cudaStream_t st;
        ~CubFree() { cudaFreeAsync(p, st); }
    } cub_guard{cub_tmp, st};
    unsigned long long *d_misc = (unsigned long long *)take(64);
    // batch table | per-scan batch bases (small, sized by the batch count: from the pool)
    long long *d_brow = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&d_brow, (kNScan + 1) * al256w(nb1 * 8), st));
    struct PoolFree {
        void *p;
        cudaStream_t st;
        ~PoolFree() { cudaFreeAsync(p, st); }
    } d_brow_guard{d_brow, st};
    auto d_base = [&](int k) { return reinterpret_cast<long long *>(reinterpret_cast<uint8_t *>(d_brow) + (size_t)(k + 1) * al256w(nb1 * 8)); };
    a.brow = d_brow;
    a.flags = reinterpret_cast<uint32_t *>(d_misc);
    a.first_bad_row = d_misc + 1;
    const unsigned long long init_misc[3] = {0ull, ~0ull, 0ull};  // error flags | first bad row | rows with a QUAL left to the exact parser
    CUDA_TRY(cudaMemcpyAsync(d_misc, init_misc, sizeof(init_misc), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_brow, batch_row0->data(), nb1 * 8, cudaMemcpyHostToDevice, st));

    auto dev_alloc = [&](WideBuf &b, size_t bytes, bool zero) -> int {
        b.bytes = std::max<size_t>(bytes, 8);
        CUDA_TRY(cudaMallocAsync(&b.d, b.bytes, st));
        if (zero) CUDA_TRY(cudaMemsetAsync(b.d, 0, b.bytes, st));
        return EXON_GPU_OK;
    };
    const size_t valid_bytes = (size_t)w->n_batches * (size_t)w->wpb * 4;
    const size_t loff_bytes = (size_t)w->n_batches * (size_t)(w->batch_rows + 1) * 4;
    if (w->want[7]) {
        // the header's INFO definitions: hash | offset | length (u32 each) | type (u8) | key bytes
        const InfoDefs &defs = s->info_defs;
        const size_t nk = defs.ids.size();
        std::vector<uint32_t> hs(nk), off(nk), len(nk);
        std::string blob;
        for (size_t k = 0; k < nk; ++k) {
            uint32_t h = 2166136261u;
            for (unsigned char ch : defs.ids[k]) h = (h ^ ch) * 16777619u;
            hs[k] = h;
            off[k] = (uint32_t)blob.size();
            len[k] = (uint32_t)defs.ids[k].size();
            blob += defs.ids[k];
        }
        const size_t o_off = al256w(nk * 4), o_len = o_off + al256w(nk * 4), o_type = o_len + al256w(nk * 4), o_blob = o_type + al256w(nk);
        if (int rc = dev_alloc(w->info_tab, o_blob + blob.size() + 16, false)) return rc;
        uint8_t *t = (uint8_t *)w->info_tab.d;
        if (nk) {
            CUDA_TRY(cudaMemcpyAsync(t, hs.data(), nk * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_off, off.data(), nk * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_len, len.data(), nk * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_type, defs.types.data(), nk, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));  // the sources are locals
        }
        a.info_hash = (const uint32_t *)t;
        a.info_off = (const int32_t *)(t + o_off);
        a.info_len = (const int32_t *)(t + o_len);
        a.info_type = t + o_type;
        a.info_blob = t + o_blob;
        a.n_info = (int32_t)nk;
    }
    if (w->want[5]) {
        if (int rc = dev_alloc(w->qual, (size_t)n_rows * 4, false)) return rc;
        if (int rc = dev_alloc(w->qual_valid, valid_bytes, true)) return rc;
        a.qual = (float *)w->qual.d;
        a.qual_valid = (uint32_t *)w->qual_valid.d;
    }

    // ---- 1. measure ----
    const unsigned grid = (unsigned)((n_rows + 255) / 256);
    vw_measure_kernel<<<grid, 256, 0, st>>>(a);
    ctx->launches.fetch_add(1);
    if (w->want[7]) {
        vw_info_measure_kernel<<<grid, 256, 0, st>>>(a);
        ctx->launches.fetch_add(1);
    }

    CUDA_TRY(cudaGetLastError());
    // ---- 2. scans + per-batch bases ----
    for (int k = 0; k < kNScan; ++k) {
        if (!need[k]) continue;
        size_t tb = cub_bytes;
        CUDA_TRY(exclusive_sum_i32_i64(cub_tmp, tb, (const int32_t *)a.cnt[k], pre[k], (int)nr1, st));
        vw_gather_i64<<<(unsigned)((nb1 + 255) / 256), 256, 0, st>>>(pre[k], d_brow, (int64_t)nb1, d_base(k));
        ctx->launches.fetch_add(2);
        w->base[k].resize(nb1);
        CUDA_TRY(cudaMemcpyAsync(w->base[k].data(), d_base(k), nb1 * 8, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_misc[3];
    CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (w->want[5] && h_misc[2] && !h_misc[0]) {
        // some QUAL is not a short unsigned integer: the exact parser, then the error state again
        vw_qual_kernel<<<grid, 256, 0, st>>>(a);
        ctx->launches.fetch_add(1);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (const uint32_t e = (uint32_t)h_misc[0])
        return fail((e & ~(kWErrQualDigits | kWErrInfoKey | kWErrInfoForm)) ? EXON_GPU_ERR_PARSE : EXON_GPU_ERR_UNSUPPORTED, "VCF record at row %llu:%s%s%s%s%s%s%s", h_misc[1],
                    (e & kWErrFields) ? " fewer than 8 tab-separated fields;" : "", (e & kWErrQual) ? " QUAL is not a float literal;" : "",
                    (e & kWErrQualDigits) ? " QUAL has more than 36 significant digits;" : "", (e & kWErrFieldLen) ? " a line of 2 GiB or more;" : "",
                    (e & kWErrInfoValue) ? " an INFO key that is not a flag has no value;" : "", (e & kWErrInfoKey) ? " an INFO key the header does not define;" : "",
                    (e & kWErrInfoForm) ? " an INFO value the reference would print differently (number not in Rust's Display form, '%' escape, flag with a value);" : "");
    for (int k = 0; k < kNScan; ++k) {
        if (!need[k]) continue;
        for (int64_t b = 0; b < w->n_batches; ++b)
            if (w->base[k][(size_t)b + 1] - w->base[k][(size_t)b] > 0x7FFFFFFFll)
                return fail(EXON_GPU_ERR_UNSUPPORTED, "vcf_next_batch: batch %lld overflows int32 offsets", (long long)b);
    }
    // ---- outputs ----
    if (w->want[2]) {
        if (int rc = dev_alloc(w->id_loff, loff_bytes, false)) return rc;
        if (int rc = dev_alloc(w->id_coff, ((size_t)w->base[kIdE][nb1 - 1] + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(w->id_val, (size_t)w->base[kIdB][nb1 - 1], false)) return rc;
        if (int rc = dev_alloc(w->id_valid, valid_bytes, true)) return rc;
        a.id_loff = (int32_t *)w->id_loff.d, a.id_coff = (int32_t *)w->id_coff.d, a.id_val = (uint8_t *)w->id_val.d, a.id_valid = (uint32_t *)w->id_valid.d;
    }
    if (w->want[3]) {
        if (int rc = dev_alloc(w->ref_off, loff_bytes, false)) return rc;
        if (int rc = dev_alloc(w->ref_val, (size_t)w->base[kRefB][nb1 - 1], false)) return rc;
        a.ref_off = (int32_t *)w->ref_off.d, a.ref_val = (uint8_t *)w->ref_val.d;
    }
    if (w->want[4]) {
        if (int rc = dev_alloc(w->alt_valid, valid_bytes, true)) return rc;
        if (int rc = dev_alloc(w->zeros, (size_t)(w->batch_rows + 1) * 4, true)) return rc;
        a.alt_valid = (uint32_t *)w->alt_valid.d;
    }
    if (w->want[7]) {
        if (int rc = dev_alloc(w->info_off, loff_bytes, false)) return rc;
        if (int rc = dev_alloc(w->info_val, (size_t)w->base[kInfoB][nb1 - 1], false)) return rc;
        a.info_offs = (int32_t *)w->info_off.d, a.info_val = (uint8_t *)w->info_val.d;
    }
    if (w->want[6]) {
        if (int rc = dev_alloc(w->fi_loff, loff_bytes, false)) return rc;
        if (int rc = dev_alloc(w->fi_coff, ((size_t)w->base[kFiE][nb1 - 1] + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(w->fi_val, (size_t)w->base[kFiB][nb1 - 1], false)) return rc;
        a.fi_loff = (int32_t *)w->fi_loff.d, a.fi_coff = (int32_t *)w->fi_coff.d, a.fi_val = (uint8_t *)w->fi_val.d;
    }
    // ---- 3. emit ----
    vw_emit_kernel<<<grid, 256, 0, st>>>(a);
    ctx->launches.fetch_add(1);
    if (w->want[7]) {
        vw_info_emit_kernel<<<grid, 256, 0, st>>>(a);
        ctx->launches.fetch_add(1);
    }
    CUDA_TRY(cudaGetLastError());
    if (!w->on_device) {
        WideBuf *b[WideStore::kBufs];
        w->all(b);
        for (int i = 0; i < WideStore::kBufs; ++i) {
            if (!b[i]->d) continue;
            CUDA_TRY(cudaHostAlloc(&b[i]->h, b[i]->bytes, cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(b[i]->h, b[i]->d, b[i]->bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return EXON_GPU_OK;
}

// Arrow C Data Interface view of column `col` (2..6) of batch b.  The caller owns `a` and `slot` and sets the release callbacks.
void wide_export(const WideStore *w, int col, int64_t b, int64_t rows, ArrowArray *a, WideChildSlot *slot) {
    memset(a, 0, sizeof(*a));
    memset(slot, 0, sizeof(*slot));
    a->length = rows;
    a->buffers = slot->bufs;
    const int64_t row0 = w->batch_row0[(size_t)b];
    const size_t loff = (size_t)b * (size_t)(w->batch_rows + 1), vw = (size_t)b * (size_t)w->wpb;
    auto list_child = [&](int64_t n_items, const int32_t *coff, const uint8_t *val) {
        ArrowArray &it = slot->item;
        it.length = n_items;
        it.null_count = 0;
        it.n_buffers = 3;
        slot->item_bufs[0] = nullptr;
        slot->item_bufs[1] = coff;
        slot->item_bufs[2] = val;
        it.buffers = slot->item_bufs;
        slot->item_ptr = &it;
        a->n_children = 1;
        a->children = &slot->item_ptr;
        a->n_buffers = 2;
    };
    switch (col) {
        case 2:
            a->null_count = -1;
            slot->bufs[0] = w->p<uint32_t>(w->id_valid) + vw;
            slot->bufs[1] = w->p<int32_t>(w->id_loff) + loff;
            list_child(w->base[kIdE][(size_t)b + 1] - w->base[kIdE][(size_t)b], w->p<int32_t>(w->id_coff) + w->base[kIdE][(size_t)b] + b,
                       w->p<uint8_t>(w->id_val) + w->base[kIdB][(size_t)b]);
            break;
        case 3:
            a->null_count = 0;
            a->n_buffers = 3;
            slot->bufs[0] = nullptr;
            slot->bufs[1] = w->p<int32_t>(w->ref_off) + loff;
            slot->bufs[2] = w->p<uint8_t>(w->ref_val) + w->base[kRefB][(size_t)b];
            break;
        case 4:
            a->null_count = -1;
            slot->bufs[0] = w->p<uint32_t>(w->alt_valid) + vw;
            slot->bufs[1] = w->p<int32_t>(w->zeros);
            list_child(0, w->p<int32_t>(w->zeros), w->p<uint8_t>(w->zeros));
            break;
        case 5:
            a->null_count = -1;
            a->n_buffers = 2;
            slot->bufs[0] = w->p<uint32_t>(w->qual_valid) + vw;
            slot->bufs[1] = w->p<float>(w->qual) + row0;
            break;
        case 7:
            a->null_count = 0;
            a->n_buffers = 3;
            slot->bufs[0] = nullptr;
            slot->bufs[1] = w->p<int32_t>(w->info_off) + loff;
            slot->bufs[2] = w->p<uint8_t>(w->info_val) + w->base[kInfoB][(size_t)b];
            break;
        default:  // 6
            a->null_count = 0;
            slot->bufs[0] = nullptr;
            slot->bufs[1] = w->p<int32_t>(w->fi_loff) + loff;
            list_child(w->base[kFiE][(size_t)b + 1] - w->base[kFiE][(size_t)b], w->p<int32_t>(w->fi_coff) + w->base[kFiE][(size_t)b] + b,
                       w->p<uint8_t>(w->fi_val) + w->base[kFiB][(size_t)b]);
            break;
    }
}

}  // namespace exon
