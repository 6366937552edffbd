// mzml_columns.cu -- mzML text -> Arrow record batches (SURVEY.md 8f rank 4 / seam B3 for mzML).
//
// Replaces MzMLArrayBuilder::append / finish (exon/exon-mzml/src/array_builder.rs:236-439) over the spectra MzMLReader::
// read_spectrum yields (exon/exon-mzml/src/mzml_reader/parser.rs:43-109).  File schema (exon/exon-mzml/src/config.rs:92-147):
//   0 id               Utf8 !null                      the <spectrum id="..."> attribute
//   1 mz               Struct{mz: List<Float64>}       the MS:1000514 array, decoded (base64 [-> zlib] -> LE f32 / f64 -> f64,
//   2 intensity        Struct{intensity: List<..>}       binary_conversion.rs:26-95); NULL (struct and list) when the spectrum
//   3 wavelength       Struct{wavelength: List<..>}      has no such array; an empty <binary/> gives an empty list (:276-296)
//   4 cv_params        List<Struct{accession, name, value: Utf8}>   the <cvParam> children of the <spectrum> element itself;
//                                                        value NULL when the attribute is absent or empty (:330-356)
//   5 precursor_mz     Float64                         MS:1000744 of the first selected ion of the first precursor (:360-378)
//   6 precusor_charge  Int64                           MS:1000041 of the same ion (:380-396; the column name is the reference's)
//
// Built on the descriptors of mzml_scan_spectra (mzml.cu: event scan -> sort -> one descriptor per spectrum in file order,
// zlib arrays already inflated):
//   1. measure  one thread per spectrum walks the element's TAGS (the base64 payloads are jumped over): attribute lengths
//               after entity unescaping, cvParam count, precursor values, the number of values of every array
//   2. scans    byte / item / value offsets (cub, 64-bit); batches restart at every file and hold batch_rows spectra
//   3. emit     the same walk writes strings and offsets; one warp per spectrum decodes the three arrays to f64
// Unpinned (no reference test covers them): XML comments / CDATA / processing instructions inside a spectrum
// are skipped as tags; a precursor m/z with more than 15 significant digits or an exponent beyond +-22 is refused (the
// exact decimal -> f64 conversion implemented here is Clinger's fast path).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "mzml.cuh"
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

constexpr uint32_t kMcErrXml = 1u;        // a tag that does not close, an attribute without quotes, an id-less <spectrum>
constexpr uint32_t kMcErrPrecursor = 2u;  // MS:1000744 / MS:1000041 value that is not a number (the reference unwraps the parse)
constexpr uint32_t kMcErrPrecDigits = 4u; // precursor m/z outside the exact fast path (see the header)
constexpr uint32_t kMcErrBase64 = 8u;
constexpr uint32_t kMcErrNoType = 16u;    // a <binaryDataArray> with content but no array-type cvParam ("No binary array type found")

// per-spectrum scan slots
enum { kIdB = 0, kCvN = 1, kAccB = 2, kNameB = 3, kValB = 4, kMzN = 5, kInN = 6, kWlN = 7, kSlots = 8 };

struct McArgs {
    const SpecDesc *specs;
    int64_t n_rows;
    const long long *brow;  // n_batches + 1: first spectrum of every batch
    int64_t n_batches;
    int32_t batch_rows, wpb;
    int32_t *cnt[kSlots];          // measure out (n_rows + 1 entries, the last one 0)
    const long long *pre[kSlots];  // exclusive scans
    uint8_t *rowflags;             // bit k: array k present; bit 3: precursor m/z valid; bit 4: charge valid
    double *prec_mz;
    long long *prec_charge;
    // outputs (emit)
    int32_t *id_off;               // batch layout: n_batches * (batch_rows + 1)
    uint8_t *id_val;
    int32_t *arr_off[3];           // list offsets, batch layout
    double *arr_val[3];
    uint32_t *arr_valid[3];        // per batch wpb words
    int32_t *cv_loff;              // list offsets (items), batch layout
    int32_t *acc_off, *name_off, *val_off;  // item-level string offsets, index item + batch (one extra entry per batch)
    uint8_t *acc_val, *name_val, *val_val;
    uint32_t *val_valid;           // item-level bitmap: batch b starts at word (e0(b) >> 5) + b
    uint32_t *pm_valid, *pc_valid; // per batch wpb words
    uint32_t *flags;
    unsigned long long *first_bad;
};

__device__ __forceinline__ bool is_space(uint32_t c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }
__device__ __forceinline__ bool name_is(const uint8_t *p, int n, const char *lit) {
    int i = 0;
    for (; i < n && lit[i]; ++i)
        if (p[i] != (uint8_t)lit[i]) return false;
    return i == n && !lit[i];
}

// byte sink with XML entity unescaping (quick-xml `unescape`: the five named entities and numeric references)
struct Sink {
    uint8_t *dst;
    int32_t n;
    __device__ __forceinline__ void put(uint8_t c) {
        if (dst) dst[n] = c;
        ++n;
    }
    __device__ void put_cp(uint32_t cp) {  // UTF-8
        if (cp < 0x80u) put((uint8_t)cp);
        else if (cp < 0x800u) put((uint8_t)(0xC0u | (cp >> 6))), put((uint8_t)(0x80u | (cp & 63u)));
        else if (cp < 0x10000u) put((uint8_t)(0xE0u | (cp >> 12))), put((uint8_t)(0x80u | ((cp >> 6) & 63u))), put((uint8_t)(0x80u | (cp & 63u)));
        else put((uint8_t)(0xF0u | (cp >> 18))), put((uint8_t)(0x80u | ((cp >> 12) & 63u))), put((uint8_t)(0x80u | ((cp >> 6) & 63u))), put((uint8_t)(0x80u | (cp & 63u)));
    }
    __device__ void unescape(const uint8_t *p, int m) {
        for (int i = 0; i < m; ++i) {
            if (p[i] != '&') {
                put(p[i]);
                continue;
            }
            int e = i + 1;
            while (e < m && e - i <= 10 && p[e] != ';') ++e;
            if (e >= m || p[e] != ';') {
                put('&');
                continue;
            }
            const uint8_t *q = p + i + 1;
            const int k = e - i - 1;
            if (name_is(q, k, "amp")) put('&');
            else if (name_is(q, k, "lt")) put('<');
            else if (name_is(q, k, "gt")) put('>');
            else if (name_is(q, k, "quot")) put('"');
            else if (name_is(q, k, "apos")) put('\'');
            else if (k >= 2 && q[0] == '#') {
                uint32_t cp = 0;
                bool ok = true;
                if (q[1] == 'x' || q[1] == 'X') {
                    for (int j = 2; j < k && ok; ++j) {
                        const uint32_t c = q[j] | 0x20u;
                        if (q[j] - '0' <= 9u) cp = cp * 16u + (q[j] - '0');
                        else if (c - 'a' <= 5u) cp = cp * 16u + (c - 'a' + 10u);
                        else ok = false;
                    }
                    ok = ok && k > 2;
                } else {
                    for (int j = 1; j < k && ok; ++j) {
                        if (q[j] - '0' <= 9u) cp = cp * 10u + (q[j] - '0');
                        else ok = false;
                    }
                }
                if (ok && cp < 0x110000u) put_cp(cp);
                else {
                    for (int j = i; j <= e; ++j) put(p[j]);
                }
            } else {
                for (int j = i; j <= e; ++j) put(p[j]);
            }
            i = e;
        }
    }
};

// attribute `name` of the start tag whose attributes lie in [p, end): value range or false
__device__ bool find_attr(const uint8_t *p, const uint8_t *end, const char *name, const uint8_t **v, int *vn) {
    while (p < end) {
        while (p < end && is_space(*p)) ++p;
        const uint8_t *k = p;
        while (p < end && *p != '=' && !is_space(*p) && *p != '>' && *p != '/') ++p;
        const int kn = (int)(p - k);
        while (p < end && is_space(*p)) ++p;
        if (p >= end || *p != '=') {
            if (kn == 0) ++p;
            continue;
        }
        ++p;
        while (p < end && is_space(*p)) ++p;
        if (p >= end || (*p != '"' && *p != '\'')) return false;
        const uint8_t q = *p++;
        const uint8_t *v0 = p;
        while (p < end && *p != q) ++p;
        if (p >= end) return false;
        if (name_is(k, kn, name)) {
            *v = v0;
            *vn = (int)(p - v0);
            return true;
        }
        ++p;
    }
    return false;
}

// Rust `f64::from_str` on the exact fast path: <= 15 significant digits and |exponent| <= 22 (one correctly rounded multiply
// or divide of two exactly representable doubles).  0: ok, 1: not a float literal, 2: outside the fast path.
__device__ int parse_f64_fast(const uint8_t *p, int n, double *out) {
    int i = 0;
    bool neg = false;
    if (i < n && (p[i] == '+' || p[i] == '-')) neg = p[i++] == '-';
    if (i >= n) return 1;
    if (n - i == 3 && (p[i] | 0x20) == 'i' && (p[i + 1] | 0x20) == 'n' && (p[i + 2] | 0x20) == 'f') {
        *out = neg ? -__longlong_as_double(0x7FF0000000000000ll) : __longlong_as_double(0x7FF0000000000000ll);
        return 0;
    }
    if (n - i == 3 && (p[i] | 0x20) == 'n' && (p[i + 1] | 0x20) == 'a' && (p[i + 2] | 0x20) == 'n') {
        *out = __longlong_as_double(0x7FF8000000000000ll);
        return 0;
    }
    unsigned long long m = 0;
    int sig = 0, e10 = 0;
    bool any = false, dot = false;
    for (; i < n; ++i) {
        const uint32_t d = (uint32_t)p[i] - '0';
        if (d <= 9u) {
            any = true;
            if (m || d) {
                if (sig < 19) {
                    m = m * 10ull + d;
                    ++sig;
                    if (dot) --e10;
                } else {
                    if (d) return 2;  // more digits than the accumulator holds
                    if (!dot) ++e10;
                }
            } else if (dot) {
                --e10;
            }
        } else if (p[i] == '.' && !dot) {
            dot = true;
        } else {
            break;
        }
    }
    if (!any) return 1;
    if (i < n && (p[i] == 'e' || p[i] == 'E')) {
        ++i;
        bool eneg = false;
        if (i < n && (p[i] == '+' || p[i] == '-')) eneg = p[i++] == '-';
        if (i >= n) return 1;
        int ex = 0;
        for (; i < n; ++i) {
            const uint32_t d = (uint32_t)p[i] - '0';
            if (d > 9u) return 1;
            if (ex < 100000) ex = ex * 10 + (int)d;
        }
        e10 += eneg ? -ex : ex;
    }
    if (i != n) return 1;
    double v;
    if (m == 0) v = 0.0;
    else {
        while (m % 10ull == 0ull && e10 < 0) {  // "810.7890" style trailing zeros
            m /= 10ull;
            ++e10;
        }
        if (m > (1ull << 53) || e10 < -22 || e10 > 22) return 2;
        static const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        v = e10 < 0 ? __ddiv_rn((double)m, p10[-e10]) : __dmul_rn((double)m, p10[e10]);
    }
    *out = neg ? -v : v;
    return 0;
}

__device__ bool parse_i64_rust(const uint8_t *p, int n, long long *out) {
    int i = 0;
    bool neg = false;
    if (n > 0 && (p[0] == '-' || p[0] == '+')) neg = p[i++] == '-';
    if (i >= n) return false;
    unsigned long long v = 0;
    for (; i < n; ++i) {
        const uint32_t d = (uint32_t)p[i] - '0';
        if (d > 9u || v > 922337203685477580ull) return false;
        v = v * 10ull + d;
        if (v > 9223372036854775808ull) return false;
    }
    if (!neg && v > 9223372036854775807ull) return false;
    *out = neg ? (long long)(0ull - v) : (long long)v;
    return true;
}

struct RowOut {  // where the emit pass writes one spectrum's strings (NULL members: measure pass)
    uint8_t *id;
    int32_t *acc_off, *name_off, *val_off;  // entry of the spectrum's first cvParam
    uint8_t *acc, *name, *val;              // batch bases
    int32_t acc0, name0, val0;              // byte offsets (batch-relative) of the spectrum's first cvParam
    uint32_t *val_valid;                    // the batch's first bitmap word
    int32_t item0;                          // index of the spectrum's first cvParam inside its batch
};

// Walks the tags of one <spectrum> element.  c[] receives the measured sizes.
__device__ uint32_t walk_spectrum(const SpecDesc &d, int32_t *c, double *prec_mz, long long *prec_charge, uint32_t *pflags, const RowOut *o) {
    uint32_t err = 0;
    const uint8_t *p = d.tag, *end = d.end;
    int depth = 0;        // 1 inside <spectrum>
    int n_prec = 0, n_ion = 0;
    bool in_ion1 = false, have_mz = false, have_z = false;
    int ion_depth = 0;
    int32_t n_cv = 0, acc_b = 0, name_b = 0, val_b = 0;
    c[kIdB] = 0;
    while (p < end) {
        if (*p != '<') {
            ++p;
            continue;
        }
        const uint8_t *t = p + 1;
        if (t >= end) break;
        if (*t == '/') {  // end tag
            while (p < end && *p != '>') ++p;
            if (in_ion1 && depth == ion_depth) in_ion1 = false;
            --depth;
            if (depth <= 0) break;
            continue;
        }
        if (*t == '?' || *t == '!') {  // declaration / comment / CDATA: skipped
            while (p < end && *p != '>') ++p;
            continue;
        }
        const uint8_t *nm = t;
        while (t < end && !is_space(*t) && *t != '>' && *t != '/') ++t;
        const int nn = (int)(t - nm);
        // end of the tag (quotes may hold '>')
        const uint8_t *a0 = t, *te = t;
        uint8_t q = 0;
        while (te < end && (q || *te != '>')) {
            if (q) {
                if (*te == q) q = 0;
            } else if (*te == '"' || *te == '\'') {
                q = *te;
            }
            ++te;
        }
        if (te >= end) {
            err |= kMcErrXml;
            break;
        }
        const bool empty = te > a0 && te[-1] == '/';
        const uint8_t *ae = empty ? te - 1 : te;
        if (depth == 0) {  // the <spectrum ...> tag itself
            const uint8_t *v;
            int vn;
            if (!find_attr(a0, ae, "id", &v, &vn)) err |= kMcErrXml;
            else {
                Sink s{o ? o->id : nullptr, 0};
                s.unescape(v, vn);
                c[kIdB] = s.n;
            }
        } else if (nn == 7 && name_is(nm, nn, "cvParam")) {
            if (depth == 1) {
                const uint8_t *v;
                int vn;
                const bool wr = o && o->acc_off;
                Sink sa{wr ? o->acc + o->acc0 + acc_b : nullptr, 0}, sn{wr ? o->name + o->name0 + name_b : nullptr, 0}, sv{wr ? o->val + o->val0 + val_b : nullptr, 0};
                if (find_attr(a0, ae, "accession", &v, &vn)) sa.unescape(v, vn);
                else err |= kMcErrXml;
                if (find_attr(a0, ae, "name", &v, &vn)) sn.unescape(v, vn);
                else err |= kMcErrXml;
                bool has_val = false;
                if (find_attr(a0, ae, "value", &v, &vn) && vn > 0) {
                    sv.unescape(v, vn);
                    has_val = sv.n > 0;
                }
                if (wr) {
                    o->acc_off[n_cv] = o->acc0 + acc_b;
                    o->name_off[n_cv] = o->name0 + name_b;
                    o->val_off[n_cv] = o->val0 + val_b;
                    if (has_val) atomicOr(o->val_valid + ((o->item0 + n_cv) >> 5), 1u << ((o->item0 + n_cv) & 31));
                }
                acc_b += sa.n;
                name_b += sn.n;
                val_b += sv.n;
                ++n_cv;
            } else if (in_ion1 && depth == ion_depth && !o) {
                const uint8_t *v, *w;
                int vn, wn;
                if (find_attr(a0, ae, "accession", &v, &vn) && vn == 10 && find_attr(a0, ae, "value", &w, &wn)) {
                    if (!have_mz && name_is(v, vn, "MS:1000744")) {
                        double x;
                        const int rc = parse_f64_fast(w, wn, &x);
                        if (rc) err |= rc == 1 ? kMcErrPrecursor : kMcErrPrecDigits;
                        *prec_mz = x;
                        have_mz = true;
                    } else if (!have_z && name_is(v, vn, "MS:1000041")) {
                        long long z;
                        if (!parse_i64_rust(w, wn, &z)) err |= kMcErrPrecursor;
                        *prec_charge = z;
                        have_z = true;
                    }
                }
            }
        } else if (nn == 9 && name_is(nm, nn, "precursor")) {
            ++n_prec;
        } else if (nn == 11 && name_is(nm, nn, "selectedIon") && n_prec == 1) {
            if (++n_ion == 1 && !empty) {
                in_ion1 = true;
                ion_depth = depth + 1;
            }
        }
        if (!empty) ++depth;
        p = te + 1;
        if (nn == 6 && name_is(nm, nn, "binary") && !empty) {
            // jump over the payload (no '<' inside base64, but thousands of bytes)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (d.arr[k] && d.arr[k] >= p && d.arr[k] - p < 256) p = d.arr[k] + d.len[k];
        }
    }
    c[kCvN] = n_cv;
    c[kAccB] = acc_b;
    c[kNameB] = name_b;
    c[kValB] = val_b;
    if (pflags) *pflags = (have_mz ? 8u : 0u) | (have_z ? 16u : 0u);
    return err;
}

__global__ void __launch_bounds__(128) mc_measure_kernel(const __grid_constant__ McArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_rows) return;
    const SpecDesc d = a.specs[r];
    int32_t c[kSlots];
    double pm = 0.0;
    long long pz = 0;
    uint32_t pf = 0;
    uint32_t err = walk_spectrum(d, c, &pm, &pz, &pf, nullptr);
    uint32_t rf = pf;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int32_t n = 0;
        if (d.arr[k]) {
            const uint32_t w = d.f32[k] ? 4u : 8u;
            n = (int32_t)(d.zl[k] ? d.n_default : b64_bytes(d.arr[k], d.len[k]) / w);
            rf |= 1u << k;
        }
        c[kMzN + k] = n;
    }
#pragma unroll
    for (int k = 0; k < kSlots; ++k)
        if (a.cnt[k]) a.cnt[k][r] = c[k];
    a.rowflags[r] = (uint8_t)rf;
    a.prec_mz[r] = pm;
    a.prec_charge[r] = pz;
    if (err) {
        atomicOr(a.flags, err);
        atomicMin(a.first_bad, (unsigned long long)r);
    }
}

__device__ __forceinline__ int64_t batch_of(const McArgs &a, int64_t r) {
    int64_t lo = 0, hi = a.n_batches;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&a.brow[mid]) <= r) lo = mid;
        else hi = mid;
    }
    return lo;
}

// strings, offsets, validity bits of one spectrum (thread per spectrum)
__global__ void __launch_bounds__(128) mc_emit_rows_kernel(const __grid_constant__ McArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_rows) return;
    const int64_t b = batch_of(a, r);
    const int64_t r0 = __ldg(&a.brow[b]), r1 = __ldg(&a.brow[b + 1]);
    const int in_batch = (int)(r - r0);
    const int64_t lrow = b * (int64_t)(a.batch_rows + 1) + in_batch;
    const bool last = r + 1 == r1;
    const uint8_t rf = a.rowflags[r];
    const uint32_t bit = 1u << (in_batch & 31);
    const int64_t vw = b * (int64_t)a.wpb + (in_batch >> 5);
    RowOut o;
    memset(&o, 0, sizeof(o));
    if (a.id_off) {
        a.id_off[lrow] = (int32_t)(a.pre[kIdB][r] - a.pre[kIdB][r0]);
        if (last) a.id_off[lrow + 1] = (int32_t)(a.pre[kIdB][r1] - a.pre[kIdB][r0]);
        o.id = a.id_val + a.pre[kIdB][r];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!a.arr_off[k]) continue;
        a.arr_off[k][lrow] = (int32_t)(a.pre[kMzN + k][r] - a.pre[kMzN + k][r0]);
        if (last) a.arr_off[k][lrow + 1] = (int32_t)(a.pre[kMzN + k][r1] - a.pre[kMzN + k][r0]);
        if (rf & (1u << k)) atomicOr(a.arr_valid[k] + vw, bit);
    }
    if (a.pm_valid && (rf & 8u)) atomicOr(a.pm_valid + vw, bit);
    if (a.pc_valid && (rf & 16u)) atomicOr(a.pc_valid + vw, bit);
    if (a.cv_loff) {
        const long long e = a.pre[kCvN][r], e0 = a.pre[kCvN][r0];
        a.cv_loff[lrow] = (int32_t)(e - e0);
        if (last) a.cv_loff[lrow + 1] = (int32_t)(a.pre[kCvN][r1] - e0);
        o.item0 = (int32_t)(e - e0);
        o.acc_off = a.acc_off + e + b;
        o.name_off = a.name_off + e + b;
        o.val_off = a.val_off + e + b;
        o.acc = a.acc_val + a.pre[kAccB][r0];
        o.name = a.name_val + a.pre[kNameB][r0];
        o.val = a.val_val + a.pre[kValB][r0];
        o.acc0 = (int32_t)(a.pre[kAccB][r] - a.pre[kAccB][r0]);
        o.name0 = (int32_t)(a.pre[kNameB][r] - a.pre[kNameB][r0]);
        o.val0 = (int32_t)(a.pre[kValB][r] - a.pre[kValB][r0]);
        o.val_valid = a.val_valid + (e0 >> 5) + b;
        if (last) {  // the closing entries of the batch's three string children
            const long long e1 = a.pre[kCvN][r1];
            a.acc_off[e1 + b] = (int32_t)(a.pre[kAccB][r1] - a.pre[kAccB][r0]);
            a.name_off[e1 + b] = (int32_t)(a.pre[kNameB][r1] - a.pre[kNameB][r0]);
            a.val_off[e1 + b] = (int32_t)(a.pre[kValB][r1] - a.pre[kValB][r0]);
        }
    }
    if (!a.id_off && !a.cv_loff) return;
    int32_t c[kSlots];
    const SpecDesc d = a.specs[r];
    double pm;
    long long pz;
    walk_spectrum(d, c, &pm, &pz, nullptr, &o);  // members of `o` that are NULL are measured, not written
}

// values of the three arrays, decoded to f64 (warp per spectrum)
__global__ void __launch_bounds__(256) mc_emit_values_kernel(const __grid_constant__ McArgs a) {
    __shared__ uint8_t lut[256];
    b64_lut_init(lut);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t bad = 0;
    for (int64_t r = wid; r < a.n_rows; r += nw) {
        const SpecDesc d = a.specs[r];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!a.arr_val[k] || !d.arr[k]) continue;
            const int w = d.f32[k] ? 4 : 8;
            const uint32_t n = (uint32_t)(a.pre[kMzN + k][r + 1] - a.pre[kMzN + k][r]);
            double *dst = a.arr_val[k] + a.pre[kMzN + k][r];
            for (uint32_t i = lane; i < n; i += 32) {
                const unsigned long long bits = d.zl[k] ? raw_value(d.raw[k], i, w) : b64_value(d.arr[k], i, w, lut, bad);
                dst[i] = d.f32[k] ? (double)__uint_as_float((uint32_t)bits) : __longlong_as_double((long long)bits);
            }
        }
    }
    bad = __reduce_or_sync(0xFFFFFFFFu, bad);
    if (lane == 0 && (bad & 0x80u)) atomicOr(a.flags, kMcErrBase64);
}

__global__ void mc_gather_i64(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

size_t al256m(size_t x) { return (x + 255) & ~(size_t)255; }

struct Buf {
    void *d = nullptr, *h = nullptr;
    size_t bytes = 0;
};

}  // namespace

// ---- column store ------------------------------------------------------------------------------------------------------
struct MzColumns {
    std::atomic<int> refs{1};
    int device = 0;
    bool on_device = false;
    int batch_rows = 8192, wpb = 256;
    int64_t n_rows = 0, n_batches = 0, next = 0;
    std::vector<int> projection;
    std::vector<long long> batch_row0, base[kSlots];
    static constexpr int kBufs = 23;
    Buf id_off, id_val, arr_off[3], arr_val[3], arr_valid[3], cv_loff, acc_off, name_off, val_off, acc_val, name_val, val_val, val_valid, pm, pm_valid, pc, pc_valid;
    void all(Buf *out[kBufs]) {
        Buf *v[kBufs] = {&id_off, &id_val, &arr_off[0], &arr_off[1], &arr_off[2], &arr_val[0], &arr_val[1], &arr_val[2], &arr_valid[0], &arr_valid[1], &arr_valid[2],
                         &cv_loff, &acc_off, &name_off, &val_off, &acc_val, &name_val, &val_val, &val_valid, &pm, &pm_valid, &pc, &pc_valid};
        for (int i = 0; i < kBufs; ++i) out[i] = v[i];
    }
    template <class T>
    const T *p(const Buf &b) const { return static_cast<const T *>(on_device ? b.d : b.h); }
    void unref() {
        if (refs.fetch_sub(1) != 1) return;
        cudaSetDevice(device);
        Buf *b[kBufs];
        all(b);
        for (int i = 0; i < kBufs; ++i) {
            cudaFree(b[i]->d);
            cudaFreeHost(b[i]->h);
        }
        delete this;
    }
};

void mzml_columns_free(VcfStream *s) {
    if (s->mz_cols) {
        s->mz_cols->unref();
        s->mz_cols = nullptr;
    }
}

static int mz_build_columns(VcfStream *s) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto *c = new (std::nothrow) MzColumns();
    if (!c) return fail(EXON_GPU_ERR_OOM, "mzml_next_batch: out of host memory");
    s->mz_cols = c;
    c->device = ctx->device;
    c->on_device = s->columns_on_device;
    c->batch_rows = s->batch_rows;
    c->wpb = ((s->batch_rows + 63) / 64) * 2;
    c->projection = s->projection;
    bool want[7] = {false, false, false, false, false, false, false};
    for (int p : s->projection) want[p] = true;
    MzScan sc;
    if (int rc = mzml_scan_spectra(s, &sc)) return rc;
    CUDA_TRY(ctx->timed_end(st));
    const int64_t n_rows = (int64_t)sc.n_spec;
    c->n_rows = n_rows;
    c->batch_row0.clear();
    for (size_t f = 0; f + 1 < sc.file_spec0.size(); ++f)
        for (long long r = sc.file_spec0[f]; r < sc.file_spec0[f + 1]; r += c->batch_rows) c->batch_row0.push_back(r);
    c->batch_row0.push_back(n_rows);
    c->n_batches = (int64_t)c->batch_row0.size() - 1;
    if (n_rows == 0) {
        c->n_batches = 0;
        return EXON_GPU_OK;
    }
    const size_t nb1 = (size_t)c->n_batches + 1, nr1 = (size_t)n_rows + 1;
    // temporaries from the stream-ordered pool: counts, prefixes, flags, precursor values, batch tables
    size_t cub_bytes = 0;
    CUDA_TRY(exclusive_sum_i32_i64(nullptr, cub_bytes, (const int32_t *)nullptr, (long long *)nullptr, (int)nr1, st));
    const size_t o_cnt = 0, o_pre = o_cnt + kSlots * al256m(nr1 * 4), o_rf = o_pre + kSlots * al256m(nr1 * 8), o_pm = o_rf + al256m(nr1), o_pz = o_pm + al256m(nr1 * 8),
                 o_brow = o_pz + al256m(nr1 * 8), o_base = o_brow + al256m(nb1 * 8), o_cub = o_base + kSlots * al256m(nb1 * 8), o_misc = o_cub + al256m(cub_bytes),
                 tmp_bytes = o_misc + 256;
    uint8_t *tmp = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&tmp, tmp_bytes, st));
    struct PoolFree {
        void *p;
        cudaStream_t st;
        ~PoolFree() { cudaFreeAsync(p, st); }
    } tmp_guard{tmp, st};
    McArgs a;
    memset(&a, 0, sizeof(a));
    a.specs = sc.d_spec;
    a.n_rows = n_rows;
    a.n_batches = c->n_batches;
    a.batch_rows = c->batch_rows;
    a.wpb = c->wpb;
    long long *pre[kSlots];
    for (int k = 0; k < kSlots; ++k) {
        a.cnt[k] = (int32_t *)(tmp + o_cnt + (size_t)k * al256m(nr1 * 4));
        pre[k] = (long long *)(tmp + o_pre + (size_t)k * al256m(nr1 * 8));
        a.pre[k] = pre[k];
        CUDA_TRY(cudaMemsetAsync(a.cnt[k] + n_rows, 0, 4, st));
    }
    a.rowflags = tmp + o_rf;
    a.prec_mz = (double *)(tmp + o_pm);
    a.prec_charge = (long long *)(tmp + o_pz);
    long long *d_brow = (long long *)(tmp + o_brow);
    a.brow = d_brow;
    unsigned long long *d_misc = (unsigned long long *)(tmp + o_misc);
    a.flags = (uint32_t *)d_misc;
    a.first_bad = d_misc + 1;
    const unsigned long long init_misc[2] = {0ull, ~0ull};
    CUDA_TRY(cudaMemcpyAsync(d_misc, init_misc, sizeof(init_misc), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_brow, c->batch_row0.data(), nb1 * 8, cudaMemcpyHostToDevice, st));
    // ---- 1. measure ----
    mc_measure_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    // ---- 2. scans + per-batch bases ----
    for (int k = 0; k < kSlots; ++k) {
        size_t tb = cub_bytes;
        CUDA_TRY(exclusive_sum_i32_i64(tmp + o_cub, tb, (const int32_t *)a.cnt[k], pre[k], (int)nr1, st));
        long long *d_base = (long long *)(tmp + o_base + (size_t)k * al256m(nb1 * 8));
        mc_gather_i64<<<(unsigned)((nb1 + 255) / 256), 256, 0, st>>>(pre[k], d_brow, (int64_t)nb1, d_base);
        c->base[k].resize(nb1);
        CUDA_TRY(cudaMemcpyAsync(c->base[k].data(), d_base, nb1 * 8, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_misc[2];
    CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ctx->launches.fetch_add(1 + 2 * kSlots);
    if (const uint32_t e = (uint32_t)h_misc[0])
        return fail((e & kMcErrPrecDigits) && !(e & ~kMcErrPrecDigits) ? EXON_GPU_ERR_UNSUPPORTED : EXON_GPU_ERR_PARSE, "mzML spectrum %llu:%s%s%s%s", h_misc[1],
                    (e & kMcErrXml) ? " malformed tag, or a <spectrum> / <cvParam> without its required attributes;" : "",
                    (e & kMcErrPrecursor) ? " a precursor m/z (MS:1000744) or charge (MS:1000041) that is not a number;" : "",
                    (e & kMcErrPrecDigits) ? " a precursor m/z with more than 15 significant digits or a decimal exponent beyond 22;" : "",
                    (e & kMcErrNoType) ? " a binary array without an array type;" : "");
    for (int k = 0; k < kSlots; ++k)
        for (int64_t b = 0; b < c->n_batches; ++b)
            if (c->base[k][(size_t)b + 1] - c->base[k][(size_t)b] > 0x7FFFFFFFll) return fail(EXON_GPU_ERR_UNSUPPORTED, "mzml_next_batch: batch %lld overflows int32 offsets", (long long)b);
    // ---- outputs ----
    auto dev_alloc = [&](Buf &b, size_t bytes, bool zero) -> int {
        b.bytes = std::max<size_t>(bytes, 8);
        CUDA_TRY(cudaMalloc(&b.d, b.bytes));
        if (zero) CUDA_TRY(cudaMemsetAsync(b.d, 0, b.bytes, st));
        return EXON_GPU_OK;
    };
    const size_t valid_bytes = (size_t)c->n_batches * (size_t)c->wpb * 4;
    const size_t loff_bytes = (size_t)c->n_batches * (size_t)(c->batch_rows + 1) * 4;
    if (want[0]) {
        if (int rc = dev_alloc(c->id_off, loff_bytes, false)) return rc;
        if (int rc = dev_alloc(c->id_val, (size_t)c->base[kIdB][nb1 - 1], false)) return rc;
        a.id_off = (int32_t *)c->id_off.d, a.id_val = (uint8_t *)c->id_val.d;
    }
    for (int k = 0; k < 3; ++k) {
        if (!want[1 + k]) continue;
        if (int rc = dev_alloc(c->arr_off[k], loff_bytes, false)) return rc;
        if (int rc = dev_alloc(c->arr_val[k], (size_t)c->base[kMzN + k][nb1 - 1] * 8, false)) return rc;
        if (int rc = dev_alloc(c->arr_valid[k], valid_bytes, true)) return rc;
        a.arr_off[k] = (int32_t *)c->arr_off[k].d, a.arr_val[k] = (double *)c->arr_val[k].d, a.arr_valid[k] = (uint32_t *)c->arr_valid[k].d;
    }
    if (want[4]) {
        const size_t items = (size_t)c->base[kCvN][nb1 - 1];
        if (int rc = dev_alloc(c->cv_loff, loff_bytes, false)) return rc;
        if (int rc = dev_alloc(c->acc_off, (items + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->name_off, (items + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->val_off, (items + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->acc_val, (size_t)c->base[kAccB][nb1 - 1], false)) return rc;
        if (int rc = dev_alloc(c->name_val, (size_t)c->base[kNameB][nb1 - 1], false)) return rc;
        if (int rc = dev_alloc(c->val_val, (size_t)c->base[kValB][nb1 - 1], false)) return rc;
        if (int rc = dev_alloc(c->val_valid, ((items >> 5) + nb1 + 1) * 4, true)) return rc;
        a.cv_loff = (int32_t *)c->cv_loff.d;
        a.acc_off = (int32_t *)c->acc_off.d, a.name_off = (int32_t *)c->name_off.d, a.val_off = (int32_t *)c->val_off.d;
        a.acc_val = (uint8_t *)c->acc_val.d, a.name_val = (uint8_t *)c->name_val.d, a.val_val = (uint8_t *)c->val_val.d;
        a.val_valid = (uint32_t *)c->val_valid.d;
    }
    if (want[5]) {
        if (int rc = dev_alloc(c->pm, (size_t)n_rows * 8, false)) return rc;
        if (int rc = dev_alloc(c->pm_valid, valid_bytes, true)) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->pm.d, a.prec_mz, (size_t)n_rows * 8, cudaMemcpyDeviceToDevice, st));
        a.pm_valid = (uint32_t *)c->pm_valid.d;
    }
    if (want[6]) {
        if (int rc = dev_alloc(c->pc, (size_t)n_rows * 8, false)) return rc;
        if (int rc = dev_alloc(c->pc_valid, valid_bytes, true)) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->pc.d, a.prec_charge, (size_t)n_rows * 8, cudaMemcpyDeviceToDevice, st));
        a.pc_valid = (uint32_t *)c->pc_valid.d;
    }
    // ---- 3. emit ----
    mc_emit_rows_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(a);
    if (want[1] || want[2] || want[3]) {
        const int grid = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)ctx->sm_count * 8);
        mc_emit_values_kernel<<<grid, 256, 0, st>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->launches.fetch_add(2);
    CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    if (!c->on_device) {
        Buf *b[MzColumns::kBufs];
        c->all(b);
        for (int i = 0; i < MzColumns::kBufs; ++i) {
            if (!b[i]->d) continue;
            CUDA_TRY(cudaHostAlloc(&b[i]->h, b[i]->bytes, cudaHostAllocDefault));
            CUDA_TRY(cudaMemcpyAsync(b[i]->h, b[i]->d, b[i]->bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((uint32_t)h_misc[0] & kMcErrBase64) return fail(EXON_GPU_ERR_PARSE, "malformed mzML: invalid base64 in a binary array");
    return EXON_GPU_OK;
}

// ---- Arrow export ------------------------------------------------------------------------------------------------------
namespace {

struct Node {
    ArrowArray arr;
    const void *bufs[3];
    ArrowArray *kids[3];
};
struct MzBatchPriv {
    MzColumns *cols;
    Node top;
    Node nodes[7 * 4];  // per projected column at most: column, list, values / struct, 3 strings
    ArrowArray *col_ptrs[7];
    const void *top_bufs[1];
};
void mz_release_child(ArrowArray *a) { a->release = nullptr; }
void mz_release_batch(ArrowArray *a) {
    auto *p = static_cast<MzBatchPriv *>(a->private_data);
    p->cols->unref();
    delete p;
    a->release = nullptr;
}
void node_init(Node &n, int64_t length, int64_t null_count, int n_buffers, int n_children) {
    memset(&n, 0, sizeof(n));
    n.arr.length = length;
    n.arr.null_count = null_count;
    n.arr.n_buffers = n_buffers;
    n.arr.buffers = n.bufs;
    n.arr.n_children = n_children;
    n.arr.children = n_children ? n.kids : nullptr;
    n.arr.release = mz_release_child;
}

struct MzSchemaPriv {
    ArrowSchema nodes[7 * 5];
    ArrowSchema *ptrs[7 * 5];
    ArrowSchema *col_ptrs[7];
    int n = 0;
};
void mz_release_schema_child(ArrowSchema *s) { s->release = nullptr; }
void mz_release_schema(ArrowSchema *s) {
    delete static_cast<MzSchemaPriv *>(s->private_data);
    s->release = nullptr;
}
ArrowSchema *schema_node(MzSchemaPriv *p, const char *fmt, const char *name, bool nullable) {
    ArrowSchema &s = p->nodes[p->n];
    memset(&s, 0, sizeof(s));
    s.format = fmt;
    s.name = name;
    s.flags = nullable ? ARROW_FLAG_NULLABLE : 0;
    s.release = mz_release_schema_child;
    p->ptrs[p->n] = &s;
    return p->ptrs[p->n++];
}
void set_children(ArrowSchema *s, ArrowSchema **first, int n) {
    s->n_children = n;
    s->children = first;
}

}  // namespace

// exon/exon-mzml/src/config.rs:92-147
static void mz_fill_schema(const std::vector<int> &projection, ArrowSchema *out) {
    static const char *arr_names[3] = {"mz", "intensity", "wavelength"};
    auto *p = new MzSchemaPriv();
    int ncol = 0;
    for (int col : projection) {
        ArrowSchema *c = nullptr;
        if (col == 0) c = schema_node(p, "u", "id", false);
        else if (col >= 1 && col <= 3) {
            c = schema_node(p, "+s", arr_names[col - 1], true);
            const int first = p->n;
            ArrowSchema *l = schema_node(p, "+l", arr_names[col - 1], true);
            const int item_at = p->n;
            schema_node(p, "g", "item", true);
            set_children(l, &p->ptrs[item_at], 1);
            set_children(c, &p->ptrs[first], 1);
        } else if (col == 4) {
            c = schema_node(p, "+l", "cv_params", true);
            const int item_at = p->n;
            ArrowSchema *it = schema_node(p, "+s", "item", true);
            const int f0 = p->n;
            schema_node(p, "u", "accession", true);
            schema_node(p, "u", "name", true);
            schema_node(p, "u", "value", true);
            set_children(it, &p->ptrs[f0], 3);
            set_children(c, &p->ptrs[item_at], 1);
        } else if (col == 5) c = schema_node(p, "g", "precursor_mz", true);
        else c = schema_node(p, "l", "precusor_charge", true);
        p->col_ptrs[ncol++] = c;
    }
    memset(out, 0, sizeof(*out));
    out->format = "+s";
    out->name = "";
    out->n_children = ncol;
    out->children = p->col_ptrs;
    out->release = mz_release_schema;
    out->private_data = p;
}

int mzml_stream_schema(VcfStream *s, ArrowSchema *out) {
    mz_fill_schema(s->projection, out);
    return EXON_GPU_OK;
}

static int mzml_next_batch(VcfStream *s, ArrowArray *out, ArrowSchema *out_schema) {
    {
        if (int rc = s->flush_gz()) return rc;
        std::lock_guard<std::recursive_mutex> work(s->ctx->work_mu);
        if (!s->mz_cols) {
            const int rc = mz_build_columns(s);
            if (rc != EXON_GPU_OK) {
                mzml_columns_free(s);
                return rc;
            }
            s->drained = true;
        }
    }
    MzColumns *c = s->mz_cols;
    if (out_schema) mz_fill_schema(s->projection, out_schema);
    memset(out, 0, sizeof(*out));
    if (c->next >= c->n_batches) return EXON_GPU_OK;
    const int64_t b = c->next++;
    const int64_t row0 = c->batch_row0[(size_t)b], rows = c->batch_row0[(size_t)b + 1] - row0;
    const size_t loff = (size_t)b * (size_t)(c->batch_rows + 1), vw = (size_t)b * (size_t)c->wpb;
    auto *p = new MzBatchPriv();
    p->cols = c;
    c->refs.fetch_add(1);
    int nn = 0, ncol = 0;
    for (int col : s->projection) {
        Node &n = p->nodes[nn++];
        if (col == 0) {
            node_init(n, rows, 0, 3, 0);
            n.bufs[0] = nullptr;
            n.bufs[1] = c->p<int32_t>(c->id_off) + loff;
            n.bufs[2] = c->p<uint8_t>(c->id_val) + c->base[kIdB][(size_t)b];
        } else if (col >= 1 && col <= 3) {
            const int k = col - 1;
            const uint32_t *valid = c->p<uint32_t>(c->arr_valid[k]) + vw;
            node_init(n, rows, -1, 1, 1);  // struct
            n.bufs[0] = valid;
            Node &l = p->nodes[nn++];
            node_init(l, rows, -1, 2, 1);  // list
            l.bufs[0] = valid;
            l.bufs[1] = c->p<int32_t>(c->arr_off[k]) + loff;
            Node &v = p->nodes[nn++];
            node_init(v, c->base[kMzN + k][(size_t)b + 1] - c->base[kMzN + k][(size_t)b], 0, 2, 0);
            v.bufs[0] = nullptr;
            v.bufs[1] = c->p<double>(c->arr_val[k]) + c->base[kMzN + k][(size_t)b];
            l.kids[0] = &v.arr;
            n.kids[0] = &l.arr;
        } else if (col == 4) {
            const long long e0 = c->base[kCvN][(size_t)b], items = c->base[kCvN][(size_t)b + 1] - e0;
            node_init(n, rows, 0, 2, 1);  // list
            n.bufs[0] = nullptr;
            n.bufs[1] = c->p<int32_t>(c->cv_loff) + loff;
            Node &st = p->nodes[nn++];
            node_init(st, items, 0, 1, 3);  // struct
            st.bufs[0] = nullptr;
            const Buf *offs[3] = {&c->acc_off, &c->name_off, &c->val_off};
            const Buf *vals[3] = {&c->acc_val, &c->name_val, &c->val_val};
            const int slots[3] = {kAccB, kNameB, kValB};
            for (int f = 0; f < 3; ++f) {
                Node &sn = p->nodes[nn++];
                node_init(sn, items, f == 2 ? -1 : 0, 3, 0);
                sn.bufs[0] = f == 2 ? c->p<uint32_t>(c->val_valid) + (e0 >> 5) + b : nullptr;
                sn.bufs[1] = c->p<int32_t>(*offs[f]) + e0 + b;
                sn.bufs[2] = c->p<uint8_t>(*vals[f]) + c->base[slots[f]][(size_t)b];
                st.kids[f] = &sn.arr;
            }
            n.kids[0] = &st.arr;
        } else if (col == 5) {
            node_init(n, rows, -1, 2, 0);
            n.bufs[0] = c->p<uint32_t>(c->pm_valid) + vw;
            n.bufs[1] = c->p<double>(c->pm) + row0;
        } else {
            node_init(n, rows, -1, 2, 0);
            n.bufs[0] = c->p<uint32_t>(c->pc_valid) + vw;
            n.bufs[1] = c->p<long long>(c->pc) + row0;
        }
        p->col_ptrs[ncol++] = &n.arr;
    }
    p->top_bufs[0] = nullptr;
    out->length = rows;
    out->null_count = 0;
    out->offset = 0;
    out->n_buffers = 1;
    out->buffers = p->top_bufs;
    out->n_children = ncol;
    out->children = p->col_ptrs;
    out->release = mz_release_batch;
    out->private_data = p;
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_mzml_open_columns(exon_gpu_ctx *c, const exon_gpu_fastq_opts *o, exon_gpu_stream **out) {
    if (!c || !o || !out) return fail(EXON_GPU_ERR_ARG, "mzml_open_columns: NULL argument");
    if (o->batch_rows < 0 || o->n_projection <= 0 || !o->projection) return fail(EXON_GPU_ERR_ARG, "mzml_open_columns: bad batch_rows / projection");
    for (int i = 0; i < o->n_projection; ++i) {
        if (o->projection[i] < 0 || o->projection[i] > 6)
            return fail(EXON_GPU_ERR_ARG, "mzml_open_columns: projection index %d is not an mzML file-schema column (0 id, 1 mz, 2 intensity, 3 wavelength, "
                                          "4 cv_params, 5 precursor_mz, 6 precusor_charge)", o->projection[i]);
        for (int j = 0; j < i; ++j)
            if (o->projection[j] == o->projection[i]) return fail(EXON_GPU_ERR_ARG, "mzml_open_columns: column %d is projected twice", o->projection[i]);
    }
    if (int rc = exon_gpu_mzml_open(c, out)) return rc;
    if (o->batch_rows > 0) (*out)->batch_rows = o->batch_rows;
    (*out)->projection.assign(o->projection, o->projection + o->n_projection);
    (*out)->columns_on_device = o->columns_on_device != 0;
    return EXON_GPU_OK;
}

int exon_gpu_mzml_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema) {
    if (!s || !out || s->fmt != kFmtMzml) return fail(EXON_GPU_ERR_ARG, "mzml_next_batch: not an mzML stream");
    if (s->projection.empty()) return fail(EXON_GPU_ERR_STATE, "mzml_next_batch: the stream was opened without a projection (exon_gpu_mzml_open_columns)");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return mzml_next_batch(s, out, out_schema);
}

}  // extern "C"
