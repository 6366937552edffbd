// vcf_wide.cu -- VCF text -> Arrow columns 2..6 {id: list<utf8>, ref: utf8, alt: list<utf8>, qual: f32, filter: list<utf8>}.
//
// Replaces LazyVCFArrayBuilder::append / finish for those columns (exon/exon-vcf/src/array_builder/
// lazy_array_builder.rs:169-216, 451-484), with the lazy builder's own quirks kept:
//   id      "." -> NULL, else one list item per ';'-separated id                                      (:169-179)
//   ref     the bases, byte for byte                                                                   (:180-189)
//   alt     "." -> NULL, else a VALID BUT EMPTY list: the builder never appends the alleles            (:190-204)
//   qual    "." -> NULL, else Rust `f32::from_str`, correctly rounded (f32_parse.cuh)                  (:205-208)
//   filter  always valid: "." -> [], else one item per ';'-separated filter                            (:209-216)
//   info    (string mode) the field re-serialised from noodles' typed view: numbers through Rust's Display, see info_walk  (:217-298)
//   formats (string mode) FORMAT keys + every sample's values re-serialised the same way, see formats_walk               (:310-432)
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
#include "f32_display.cuh"
#include "f32_parse.cuh"
#include "internal.h"
#include "scan_i64.cuh"
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

struct KeyDefs;  // defined with the INFO / FORMAT walks below
struct KeyDefsPod {  // KeyDefs by value inside the kernel argument
    const uint32_t *hash;
    const int32_t *off, *len;
    const uint8_t *type, *single, *blob;
    int32_t n;
};
struct WideArgs {
    int64_t n_rows;
    const uint8_t *const *line_start;
    const uint8_t *const *line_end;
    const long long *brow;  // n_batches + 1
    int64_t n_batches;
    int32_t batch_rows, wpb;
    int32_t want_id, want_ref, want_alt, want_qual, want_filter, want_info, want_formats;
    // INFO / FORMAT definitions of the header (exon_gpu_vcf_set_header): FNV-1a hash, offset / length into the blob, type, Number == 1
    KeyDefsPod info_defs, fmt_defs;
    int32_t *info_offs, *fmt_offs;  // batch-relative offsets, n_batches * (batch_rows + 1)
    uint8_t *info_val, *fmt_val;
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


// ---- INFO (column 7) and FORMAT / samples (column 8), string mode ----------------------------------------------------------
// LazyVCFArrayBuilder::append :217-298 / :310-432 does not copy these fields: it walks noodles' typed view of them (types
// from the header's ##INFO / ##FORMAT lines) and prints every value again -- integers through i32 Display, floats through
// f32 Display (f32_display.cuh), strings percent-decoded, an INFO flag as `key=true`, a genotype allele by allele.  Most
// values are already in the form Display gives them (checked with canon_int / canon_float) and are copied; the others are
// parsed (Rust grammar: `+7`, `007`, `1e3`, `0.50`, `inf` ...) and printed.  What the reference cannot print is an error
// here too: a value noodles cannot parse, and a MISSING value (`key=.`, a non-flag key without `=`, a sample value `.`), on
// which the builder's `value_option.unwrap()` panics.
// Key definitions: the header's; for a key the header does not define, the VCF specification's reserved key of that name
// (a stable subset of noodles' fallback table, below); otherwise a single String.
enum : uint8_t { kInfoInteger = 0, kInfoFloat = 1, kInfoFlag = 2, kInfoCharacter = 3, kInfoString = 4, kFmtGenotype = 5 };

__device__ __forceinline__ uint32_t fnv1a(const uint8_t *p, int n) {
    uint32_t h = 2166136261u;
    for (int i = 0; i < n; ++i) h = (h ^ __ldg(p + i)) * 16777619u;
    return h;
}

// -?(0|[1-9][0-9]*) within i32: the text i32 Display prints
__device__ __forceinline__ bool canon_int(const uint8_t *p, int n) {
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
    int i = 0;
    const bool neg = n > 0 && __ldg(p) == '-';
    if (neg) i = 1;
    if (i >= n) return false;
    int int_digits = 0, sig = 0;
    bool nonzero = false, lead_zero = false;
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

// Rust `i32::from_str`: [+-]?[0-9]+ within range
__device__ __forceinline__ bool parse_i32_rust(const uint8_t *p, int n, int32_t *out) {
    int i = 0;
    bool neg = false;
    if (n > 0 && (__ldg(p) == '-' || __ldg(p) == '+')) {
        neg = __ldg(p) == '-';
        i = 1;
    }
    if (i >= n) return false;
    unsigned long long v = 0;
    for (; i < n; ++i) {
        const uint32_t d = (uint32_t)__ldg(p + i) - '0';
        if (d > 9u) return false;
        v = v * 10ull + d;
        if (v > 2147483648ull) return false;
    }
    if (!neg && v > 2147483647ull) return false;
    *out = neg ? (int32_t)(0u - (uint32_t)v) : (int32_t)v;
    return true;
}

struct Emit {  // byte sink: counts always, writes when dst is set
    uint8_t *dst;
    int32_t n;
    __device__ __forceinline__ void put(uint8_t c) {
        if (dst) dst[n] = c;
        ++n;
    }
    __device__ __forceinline__ void copy(const uint8_t *p, int32_t m) {
        if (dst)
            for (int32_t q = 0; q < m; ++q) dst[n + q] = __ldg(p + q);
        n += m;
    }
};

__device__ __forceinline__ int hexval(uint32_t c) {
    if (c - '0' <= 9u) return (int)(c - '0');
    c |= 0x20u;
    return c - 'a' <= 5u ? (int)(c - 'a' + 10) : -1;
}

// the slow half of a number: parse with Rust's grammar, print with Rust's Display (kept out of line: the exact float parser
// and the shortest-digits printer are large, and almost no value comes here)
__device__ __noinline__ int32_t number_reprint(int type, const uint8_t *p, int32_t m, uint8_t *out /* >= kF32DisplayMax */) {
    if (m > 512) return -1;
    if (type == kInfoInteger) {
        int32_t v;
        if (!parse_i32_rust(p, m, &v)) return -1;
        return i32_display(v, out);
    }
    uint8_t tmp[64];
    if (m > 64) return -1;  // longer literals exist (a long string of zeros) but not in practice
    for (int32_t q = 0; q < m; ++q) tmp[q] = __ldg(p + q);
    float f;
    if (parse_f32_rust(tmp, m, &f) != kF32Ok) return -1;
    return f32_display(f, out);
}

// one element [p, p + m) of a value of type `type`; false: the reference cannot print it
__device__ __forceinline__ bool elem_emit(int type, const uint8_t *p, int32_t m, Emit &o) {
    if (m == 0) return false;
    if (type == kInfoInteger || type == kInfoFloat) {
        if (type == kInfoInteger ? canon_int(p, m) : canon_float(p, m)) {
            o.copy(p, m);
            return true;
        }
        uint8_t buf[kF32DisplayMax + 8];
        const int32_t k = number_reprint(type, p, m, buf);
        if (k < 0) return false;
        if (o.dst)
            for (int32_t q = 0; q < k; ++q) o.dst[o.n + q] = buf[q];
        o.n += k;
        return true;
    }
    // Character / String: percent-decoded (noodles: percent_encoding::percent_decode_str)
    int32_t out0 = o.n;
    for (int32_t q = 0; q < m; ++q) {
        uint32_t c = __ldg(p + q);
        if (c == '%' && q + 2 < m) {
            const int h = hexval(__ldg(p + q + 1)), l = hexval(__ldg(p + q + 2));
            if (h >= 0 && l >= 0) {
                c = (uint32_t)(h * 16 + l);
                q += 2;
            }
        }
        o.put((uint8_t)c);
    }
    return type != kInfoCharacter || o.n - out0 == 1;
}

// a whole value: one element (Number=1) or a ','-separated array whose missing elements print as '.'
// (Character arrays skip them in FORMAT, :371-381)
__device__ __forceinline__ bool value_emit(int type, bool single, bool skip_missing_chars, const uint8_t *p, int32_t m, Emit &o) {
    if (m == 0 || (m == 1 && __ldg(p) == '.')) return false;  // a missing value: the builder unwraps a None
    if (single || type == kInfoString) return elem_emit(type, p, m, o);  // (a String array joins back to its own text)
    int32_t v = 0;
    bool first = true;
    while (true) {
        int32_t ve = v;
        while (ve < m && __ldg(p + ve) != ',') ++ve;
        const bool missing = ve - v == 1 && __ldg(p + v) == '.';
        if (!(missing && skip_missing_chars)) {
            if (!first) o.put(',');
            first = false;
            if (missing) o.put('.');
            else if (!elem_emit(type, p + v, ve - v, o)) return false;
        }
        if (ve >= m) break;
        v = ve + 1;
    }
    return true;
}

// header definitions of one kind of key (INFO or FORMAT) + the specification's reserved keys as the fallback
struct KeyDefs : KeyDefsPod {};
struct Reserved {
    const char *key;
    uint8_t type, single;
};
// VCF 4.3 / 4.4 tables 1 and 2, the entries that do not differ between the versions noodles knows
__constant__ Reserved c_reserved_info[] = {
    {"AA", kInfoString, 1}, {"AC", kInfoInteger, 0}, {"AD", kInfoInteger, 0}, {"ADF", kInfoInteger, 0}, {"ADR", kInfoInteger, 0},
    {"AF", kInfoFloat, 0}, {"AN", kInfoInteger, 1}, {"BQ", kInfoFloat, 1}, {"CIGAR", kInfoString, 0}, {"DB", kInfoFlag, 0},
    {"DP", kInfoInteger, 1}, {"END", kInfoInteger, 1}, {"H2", kInfoFlag, 0}, {"H3", kInfoFlag, 0}, {"MQ", kInfoFloat, 1},
    {"MQ0", kInfoInteger, 1}, {"NS", kInfoInteger, 1}, {"SB", kInfoInteger, 0}, {"SOMATIC", kInfoFlag, 0}, {"VALIDATED", kInfoFlag, 0},
    {"1000G", kInfoFlag, 0}, {"IMPRECISE", kInfoFlag, 0}, {"NOVEL", kInfoFlag, 0}, {"SVTYPE", kInfoString, 1},
};
__constant__ Reserved c_reserved_format[] = {
    {"AD", kInfoInteger, 0}, {"ADF", kInfoInteger, 0}, {"ADR", kInfoInteger, 0}, {"DP", kInfoInteger, 1}, {"EC", kInfoInteger, 0},
    {"FT", kInfoString, 1}, {"GL", kInfoFloat, 0}, {"GP", kInfoFloat, 0}, {"GQ", kInfoInteger, 1}, {"GT", kFmtGenotype, 1},
    {"HQ", kInfoInteger, 0}, {"MQ", kInfoInteger, 1}, {"PL", kInfoInteger, 0}, {"PP", kInfoInteger, 0}, {"PQ", kInfoInteger, 1},
    {"PS", kInfoInteger, 1},
};

template <bool FORMAT>
__device__ __forceinline__ void lookup_key(const KeyDefs &d, const uint8_t *k, int32_t klen, int *type, bool *single) {
    if (FORMAT && klen == 2 && __ldg(k) == 'G' && __ldg(k + 1) == 'T') {  // GT is a genotype whatever the header calls it
        *type = kFmtGenotype;
        *single = true;
        return;
    }
    const uint32_t h = fnv1a(k, klen);
    for (int i = 0; i < d.n; ++i) {
        if (d.hash[i] != h || d.len[i] != klen) continue;
        bool same = true;
        for (int q = 0; q < klen && same; ++q) same = d.blob[d.off[i] + q] == __ldg(k + q);
        if (same) {
            *type = d.type[i];
            *single = d.single[i] != 0;
            return;
        }
    }
    const Reserved *tab = FORMAT ? c_reserved_format : c_reserved_info;
    const int nt = FORMAT ? (int)(sizeof(c_reserved_format) / sizeof(Reserved)) : (int)(sizeof(c_reserved_info) / sizeof(Reserved));
    for (int i = 0; i < nt; ++i) {
        const char *r = tab[i].key;
        int q = 0;
        while (q < klen && r[q] && (uint8_t)r[q] == __ldg(k + q)) ++q;
        if (q == klen && !r[q]) {
            *type = tab[i].type;
            *single = tab[i].single != 0;
            return;
        }
    }
    *type = kInfoString;
    *single = true;
}

// Walks one INFO field.  Returns the length of the re-serialised string; *err collects kWErrInfo*.  When `dst` is set the
// string is written there.
__device__ __forceinline__ int32_t info_walk(const KeyDefs &defs, const uint8_t *f, int32_t n, uint8_t *dst, uint32_t *err) {
    if (n == 0 || (n == 1 && __ldg(f) == '.')) return 0;
    Emit o{dst, 0};
    int32_t i = 0;
    while (i <= n) {
        // entry [i, e)
        int32_t e = i, eq = -1;
        while (e < n && __ldg(f + e) != ';') {
            if (eq < 0 && __ldg(f + e) == '=') eq = e;
            ++e;
        }
        const int32_t klen = (eq < 0 ? e : eq) - i;
        int type;
        bool single;
        lookup_key<false>(defs, f + i, klen, &type, &single);
        if (i) o.put(';');
        o.copy(f + i, klen);
        o.put('=');
        if (type == kInfoFlag) {
            if (eq >= 0 && eq + 1 < e) *err |= kWErrInfoForm;  // a flag with a value
            o.put('t'), o.put('r'), o.put('u'), o.put('e');
        } else if (eq < 0) {
            *err |= kWErrInfoValue;
        } else if (!value_emit(type, single, false, f + eq + 1, e - eq - 1, o)) {
            *err |= (e - eq - 1 == 0 || (e - eq - 1 == 1 && __ldg(f + eq + 1) == '.')) ? kWErrInfoValue : kWErrInfoForm;
        }
        if (e >= n) break;
        i = e + 1;
    }
    return o.n;
}

// end of the INFO field: the 8th tab, or the end of the line when the record has no FORMAT / sample columns
__device__ __forceinline__ const uint8_t *info_end(const uint8_t *info, const uint8_t *le) {
    const uint8_t *p = info;
    while (p < le && __ldg(p) != '\t') ++p;
    return p;
}

// A genotype [p, p + m): alleles are usize (printed again: "01" -> "1") or '.', separators '/' and '|'.  The builder prints
// allele k (k >= 1) behind the phasing of allele k - 1 (:338-362); allele 0's phasing is noodles' inferred one -- phased iff
// every separator is '|' -- so a diploid call prints as written and a mixed-phase polyploid call shifts its separators.
__device__ __forceinline__ bool genotype_emit(const uint8_t *p, int32_t m, Emit &o) {
    if (m == 0) return false;
    bool all_phased = true;
    for (int32_t q = 0; q < m; ++q)
        if (__ldg(p + q) == '/') all_phased = false;
    int32_t v = 0, k = 0;
    uint8_t prev_sep = all_phased ? '|' : '/';
    while (true) {
        int32_t ve = v;
        while (ve < m && __ldg(p + ve) != '/' && __ldg(p + ve) != '|') ++ve;
        if (k) o.put(prev_sep);
        if (k) prev_sep = __ldg(p + v - 1);
        if (ve - v == 1 && __ldg(p + v) == '.') {
            o.put('.');
        } else {
            if (ve == v || ve - v > 18) return false;
            unsigned long long a = 0;
            for (int32_t q = v; q < ve; ++q) {
                const uint32_t d = (uint32_t)__ldg(p + q) - '0';
                if (d > 9u) return false;
                a = a * 10ull + d;
            }
            uint8_t dg[20];
            int nd = 0;
            do {
                dg[nd++] = (uint8_t)('0' + a % 10ull);
                a /= 10ull;
            } while (a);
            while (nd) o.put(dg[--nd]);
        }
        ++k;
        if (ve >= m) break;
        v = ve + 1;
    }
    return true;
}

// Column 8: `keys ':'-joined` '\t' `every sample, its values re-serialised and ':'-joined, '\t'-joined` (:310-432).  [f, f + n)
// is the text behind the 8th tab (n == 0 or no 8th tab: a sites-only record prints "\t").  A sample with fewer values than
// keys prints the values it has (the builder zips keys with values).
__device__ __forceinline__ int32_t formats_walk(const KeyDefs &defs, const uint8_t *f, int32_t n, uint8_t *dst, uint32_t *err) {
    Emit o{dst, 0};
    int32_t ke = 0;
    while (ke < n && __ldg(f + ke) != '\t') ++ke;  // FORMAT = [0, ke)
    o.copy(f, ke);
    o.put('\t');
    int32_t v = ke + 1;
    bool first_sample = true;
    while (v <= n && ke < n) {
        int32_t se = v;
        while (se < n && __ldg(f + se) != '\t') ++se;  // sample = [v, se)
        if (!first_sample) o.put('\t');
        first_sample = false;
        int32_t kv = 0, sv = v;
        bool first_val = true;
        while (kv < ke && sv <= se) {
            int32_t kq = kv, sq = sv;
            while (kq < ke && __ldg(f + kq) != ':') ++kq;
            while (sq < se && __ldg(f + sq) != ':') ++sq;
            int type;
            bool single;
            lookup_key<true>(defs, f + kv, kq - kv, &type, &single);
            if (!first_val) o.put(':');
            first_val = false;
            const bool ok = type == kFmtGenotype ? (sq - sv == 1 && __ldg(f + sv) == '.' ? false : genotype_emit(f + sv, sq - sv, o))
                                                 : value_emit(type, single, type == kInfoCharacter, f + sv, sq - sv, o);
            if (!ok) *err |= (sq - sv == 0 || (sq - sv == 1 && __ldg(f + sv) == '.')) ? kWErrFmtValue : kWErrFmtForm;
            if (sq >= se) break;
            kv = kq + 1;
            sv = sq + 1;
        }
        if (se >= n) break;
        v = se + 1;
    }
    return o.n;
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
        if (k != kInfoB && k != kFmtB && a.cnt[k]) a.cnt[k][r] = c[k];  // kInfoB / kFmtB belong to vw_text_measure_kernel
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

// The same for the rows the tile-pipeline build (vcf_columns.cu) could not settle: it hands over (row, field) pairs.
__global__ void __launch_bounds__(256) vw_qual_list_kernel(const QualSlow *list, unsigned long long n_list, float *qual, uint32_t *valid_abs,
                                                          uint32_t *flags, unsigned long long *first_bad_row) {
    const unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_list) return;
    const QualSlow e = list[i];
    float q = 0.0f;
    const int rc = e.n > 4096u ? kF32Malformed : parse_f32_rust(e.p, (int)e.n, &q);
    if (rc == kF32Ok) {
        qual[e.row] = q;
        atomicOr(valid_abs + (e.row >> 5), 1u << (e.row & 31ull));
    } else {
        atomicOr(flags, rc == kF32Unsupported ? kWErrQualDigits : kWErrQual);
        atomicMin(first_bad_row, e.row);
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

// INFO and FORMAT live in their own kernels (the entry walks with their key lookups must not shape the register budget of
// the row passes): length + checks, then -- after the scan -- offsets and bytes.  WHICH: kInfoB or kFmtB.
template <int WHICH>
__device__ __forceinline__ int32_t walk_row(const WideArgs &a, const uint8_t *ls, const uint8_t *le, const int32_t *to, uint8_t *dst, uint32_t *err) {
    const uint8_t *f = ls + to[6] + 1;
    const uint8_t *ie = info_end(f, le);
    if (WHICH == kInfoB) return info_walk(static_cast<const KeyDefs &>(a.info_defs), f, (int32_t)(ie - f), dst, err);
    const uint8_t *g = ie < le ? ie + 1 : le;  // behind the 8th tab
    return formats_walk(static_cast<const KeyDefs &>(a.fmt_defs), g, (int32_t)(le - g), dst, err);
}

template <int WHICH>
__global__ void __launch_bounds__(256) vw_text_measure_kernel(const __grid_constant__ WideArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= a.n_rows) return;
    const uint8_t *ls = a.line_start[r], *le = a.line_end[r];
    int32_t to[7];
    int32_t n = 0;
    uint32_t err = 0;
    if (le - ls <= 0x7FFFFFF0ll && find_tabs(ls, le, to)) n = walk_row<WHICH>(a, ls, le, to, nullptr, &err);  // a short line is reported by the row pass
    a.cnt[WHICH][r] = n;
    if (err) {
        atomicOr(a.flags, err);
        atomicMin(a.first_bad_row, (unsigned long long)r);
    }
}

template <int WHICH>
__global__ void __launch_bounds__(256) vw_text_emit_kernel(const __grid_constant__ WideArgs a) {
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
    int32_t *offs = WHICH == kInfoB ? a.info_offs : a.fmt_offs;
    uint8_t *val = WHICH == kInfoB ? a.info_val : a.fmt_val;
    const long long v = a.pre[WHICH][r], v0 = a.pre[WHICH][r0];
    offs[lrow] = (int32_t)(v - v0);
    if (r + 1 == __ldg(&a.brow[b + 1])) offs[lrow + 1] = (int32_t)(a.pre[WHICH][r + 1] - v0);
    const uint8_t *ls = a.line_start[r], *le = a.line_end[r];
    int32_t to[7];
    if (le - ls > 0x7FFFFFF0ll || !find_tabs(ls, le, to)) return;
    uint32_t ignored = 0;
    walk_row<WHICH>(a, ls, le, to, val + v, &ignored);
}

__global__ void vw_gather_i64(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

size_t al256w(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int wide_qual_list(Ctx *ctx, const QualSlow *list, unsigned long long n, float *qual, uint32_t *valid_abs, uint32_t *flags,
                   unsigned long long *first_bad_row) {
    if (n == 0) return EXON_GPU_OK;
    vw_qual_list_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(list, n, qual, valid_abs, flags, first_bad_row);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    return EXON_GPU_OK;
}

int wide_fail(uint32_t e, unsigned long long row) {
    return fail((e & ~kWErrQualDigits) ? EXON_GPU_ERR_PARSE : EXON_GPU_ERR_UNSUPPORTED, "VCF record at row %llu:%s%s%s%s%s%s%s%s", row,
                (e & kWErrFields) ? " fewer than 8 tab-separated fields;" : "", (e & kWErrQual) ? " QUAL is not a float literal;" : "",
                (e & kWErrQualDigits) ? " QUAL has more than 36 significant digits;" : "", (e & kWErrFieldLen) ? " a line of 2 GiB or more;" : "",
                (e & kWErrInfoValue) ? " an INFO key that is not a flag has no value (or '.'): the reference's builder unwraps a None there;" : "",
                (e & kWErrInfoForm) ? " an INFO value that does not parse as its declared type (or a flag with a value);" : "",
                (e & kWErrFmtValue) ? " a sample value is missing ('.'): the reference's builder unwraps a None there;" : "",
                (e & kWErrFmtForm) ? " a sample value that does not parse as its declared type;" : "");
}

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
    if ((w->want[7] || w->want[8]) && !s->info_defs.set)
        return fail(EXON_GPU_ERR_STATE, "vcf_next_batch: the info / formats columns need the header's ##INFO / ##FORMAT definitions: call exon_gpu_vcf_set_header first");
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
    const bool need[kNScan] = {w->want[2], w->want[2], w->want[3], w->want[6], w->want[6], w->want[7], w->want[8]};
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
    a.want_formats = w->want[8];
    long long *pre[kNScan] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
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
    // the header's INFO / FORMAT definitions: hash | offset | length (u32 each) | type | Number == 1 (u8 each) | key bytes
    auto upload_defs = [&](const InfoDefs &defs, WideBuf &tab, KeyDefsPod &out) -> int {
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
        const size_t o_off = al256w(nk * 4), o_len = o_off + al256w(nk * 4), o_type = o_len + al256w(nk * 4), o_single = o_type + al256w(nk),
                     o_blob = o_single + al256w(nk);
        if (int rc = dev_alloc(tab, o_blob + blob.size() + 16, false)) return rc;
        uint8_t *t = (uint8_t *)tab.d;
        if (nk) {
            CUDA_TRY(cudaMemcpyAsync(t, hs.data(), nk * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_off, off.data(), nk * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_len, len.data(), nk * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_type, defs.types.data(), nk, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_single, defs.single.data(), nk, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(t + o_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));  // the sources are locals
        }
        out.hash = (const uint32_t *)t;
        out.off = (const int32_t *)(t + o_off);
        out.len = (const int32_t *)(t + o_len);
        out.type = t + o_type;
        out.single = t + o_single;
        out.blob = t + o_blob;
        out.n = (int32_t)nk;
        return EXON_GPU_OK;
    };
    if (w->want[7])
        if (int rc = upload_defs(s->info_defs, w->info_tab, a.info_defs)) return rc;
    if (w->want[8])
        if (int rc = upload_defs(s->format_defs, w->fmt_tab, a.fmt_defs)) return rc;
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
        vw_text_measure_kernel<kInfoB><<<grid, 256, 0, st>>>(a);
        ctx->launches.fetch_add(1);
    }
    if (w->want[8]) {
        vw_text_measure_kernel<kFmtB><<<grid, 256, 0, st>>>(a);
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
    if (const uint32_t e = (uint32_t)h_misc[0]) return wide_fail(e, h_misc[1]);
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
    if (w->want[8]) {
        if (int rc = dev_alloc(w->fmt_off, loff_bytes, false)) return rc;
        if (int rc = dev_alloc(w->fmt_val, (size_t)w->base[kFmtB][nb1 - 1], false)) return rc;
        a.fmt_offs = (int32_t *)w->fmt_off.d, a.fmt_val = (uint8_t *)w->fmt_val.d;
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
        vw_text_emit_kernel<kInfoB><<<grid, 256, 0, st>>>(a);
        ctx->launches.fetch_add(1);
    }
    if (w->want[8]) {
        vw_text_emit_kernel<kFmtB><<<grid, 256, 0, st>>>(a);
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
        case 8:
            a->null_count = 0;
            a->n_buffers = 3;
            slot->bufs[0] = nullptr;
            slot->bufs[1] = w->p<int32_t>(w->fmt_off) + loff;
            slot->bufs[2] = w->p<uint8_t>(w->fmt_val) + w->base[kFmtB][(size_t)b];
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
