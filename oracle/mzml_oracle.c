/*
 * oracle/mzml_oracle.c -- CPU restatement of the reference's mzML scan -> m/z range filter -> SUM(intensity) path
 * (BASELINE.json configs[4]).
 *
 * TEST INFRASTRUCTURE ONLY (same rule as vcf_oracle.c).
 *
 * Restates (paths relative to /root/reference):
 *   - exon/exon-mzml/src/mzml_reader/parser.rs:43-109           read_spectrum: every <spectrum ...> ... </spectrum> element
 *   - exon/exon-mzml/src/mzml_reader/types.rs:119-121,207-208,274-275  cvParam accessions that classify a binaryDataArray:
 *       MS:1000514 m/z, MS:1000515 intensity, MS:1000617 wavelength; MS:1000521 32-bit, MS:1000523 64-bit float;
 *       MS:1000574 zlib, MS:1000576 no compression
 *   - exon/exon-mzml/src/mzml_reader/binary_conversion.rs:26-95  base64 STANDARD -> optional zlib -> little-endian
 *       f32 / f64 -> f64 (trailing bytes that do not fill a value are dropped)
 *   - exon/exon-mzml/src/array_builder.rs:236-323                one List<Float64> per array kind and spectrum, NULL when absent
 *   - the config-5 query  SELECT SUM(i) FROM (SELECT unnest(mz.mz) m, unnest(intensity.intensity) i FROM mzml)
 *       WHERE m BETWEEN lo AND hi : the two lists are zipped (the shorter one padded with NULL, which neither passes the
 *       predicate nor adds to SUM); f64 running sum in file order (parity tolerance 1e-6 relative, north_star)
 * quick-xml 0.37 / base64 0.22 / flate2 (Cargo.lock, not vendored) are restated by a tag scanner, RFC 4648 and zlib.
 * Pinned by tests/test_mzml_golden.py: the two decode vectors of binary_conversion.rs:126-135, 2 spectra in test.mzML and
 * pyoteomics.mzML (slt/mzml-functions.slt:41-49), contains_peak(mz, 200, 1) = true on pyoteomics spectrum 0 (:9-12).
 * UNPINNED: XML comments / CDATA / entity references inside spectra (not handled; mzML writers do not emit them there).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static const uint8_t *find(const uint8_t *p, const uint8_t *end, const char *needle) {
    const size_t n = strlen(needle);
    if ((size_t)(end - p) < n) return NULL;
    return (const uint8_t *)memmem(p, (size_t)(end - p), needle, n);
}

/* next "<name" followed by ' ', '>' or '/' (so <spectrum does not match <spectrumList) */
static const uint8_t *find_tag(const uint8_t *p, const uint8_t *end, const char *name) {
    const size_t n = strlen(name);
    while ((p = find(p, end, name)) != NULL) {
        const uint8_t *q = p + n;
        if (q < end && (*q == ' ' || *q == '>' || *q == '/' || *q == '\n' || *q == '\t' || *q == '\r')) return p;
        p = q;
    }
    return NULL;
}

static int b64val(uint8_t c) {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
}

/* RFC 4648 standard alphabet with canonical padding; returns decoded length or -1 */
static int64_t b64decode(const uint8_t *s, int64_t n, uint8_t *out) {
    if (n % 4) return -1;
    int64_t w = 0;
    for (int64_t i = 0; i < n; i += 4) {
        int v[4], pad = 0;
        for (int k = 0; k < 4; k++) {
            if (s[i + k] == '=' && i + 4 == n && k >= 2) { v[k] = 0; pad++; }
            else { if (pad) return -1; v[k] = b64val(s[i + k]); if (v[k] < 0) return -1; }
        }
        const uint32_t t = ((uint32_t)v[0] << 18) | ((uint32_t)v[1] << 12) | ((uint32_t)v[2] << 6) | (uint32_t)v[3];
        out[w++] = (uint8_t)(t >> 16);
        if (pad < 2) out[w++] = (uint8_t)(t >> 8);
        if (pad < 1) out[w++] = (uint8_t)t;
    }
    return w;
}

/* decode_binary_array: base64 text -> f64 values; returns count or -1.  *out is malloc'ed. */
int64_t exo_mzml_decode_binary(const uint8_t *b64, int64_t n, int32_t is_zlib, int32_t is_f32, double **out) {
    uint8_t *raw = (uint8_t *)malloc((size_t)(n / 4 * 3 + 8));
    int64_t rn = b64decode(b64, n, raw);
    if (rn < 0) { free(raw); return -1; }
    if (is_zlib) {
        uLongf cap = (uLongf)(rn * 20 + 4096);
        uint8_t *dec = NULL;
        for (;;) {
            dec = (uint8_t *)malloc(cap);
            uLongf got = cap;
            int rc = uncompress(dec, &got, raw, (uLong)rn);
            if (rc == Z_OK) { free(raw); raw = dec; rn = (int64_t)got; break; }
            free(dec);
            if (rc != Z_BUF_ERROR) { free(raw); return -1; }
            cap *= 4;
        }
    }
    const int w = is_f32 ? 4 : 8;
    const int64_t cnt = rn / w;
    double *v = (double *)malloc(sizeof(double) * (size_t)(cnt > 0 ? cnt : 1));
    for (int64_t i = 0; i < cnt; i++) {
        if (is_f32) { float f; memcpy(&f, raw + 4 * i, 4); v[i] = (double)f; }
        else memcpy(&v[i], raw + 8 * i, 8);
    }
    free(raw);
    *out = v;
    return cnt;
}

typedef struct {
    int64_t n_spectra, n_selected;
    double sum;
    /* per-array-kind totals over the whole file, for golden checks: [0] mz [1] intensity [2] wavelength */
    double kind_sum[3];
    int64_t kind_count[3];
} exo_mzml_result;

static int has_acc(const uint8_t *p, const uint8_t *end, const char *acc) {
    char pat[40];
    strcpy(pat, " accession=\"");
    strcat(pat, acc);
    strcat(pat, "\"");
    return find(p, end, pat) != NULL;
}

/* 0 ok, -1 malformed.  spectrum_filter >= 0 restricts kind_sum / kind_count to that spectrum index. */
int exo_mzml_scan(const uint8_t *text, int64_t len, int32_t has_pred, double lo, double hi, int64_t spectrum_filter, exo_mzml_result *res) {
    const uint8_t *p = text, *end = text + len;
    memset(res, 0, sizeof(*res));
    while ((p = find_tag(p, end, "<spectrum")) != NULL) {
        const uint8_t *se = find(p, end, "</spectrum>");
        if (!se) return -1; /* "Unexpected Eof Event" */
        double *arr[3] = {NULL, NULL, NULL};
        int64_t cnt[3] = {-1, -1, -1};
        const uint8_t *q = p;
        while ((q = find_tag(q, se, "<binaryDataArray")) != NULL) {
            const uint8_t *be = find(q, se, "</binaryDataArray>");
            if (!be) { q += 16; continue; }
            int kind = has_acc(q, be, "MS:1000514") ? 0 : has_acc(q, be, "MS:1000515") ? 1 : has_acc(q, be, "MS:1000617") ? 2 : -1;
            const int f32 = has_acc(q, be, "MS:1000521"), f64 = has_acc(q, be, "MS:1000523");
            const int zl = has_acc(q, be, "MS:1000574"), nz = has_acc(q, be, "MS:1000576");
            const uint8_t *b0 = find(q, be, "<binary>");
            if (kind >= 0 && b0) {
                b0 += 8;
                const uint8_t *b1 = find(b0, be, "</binary>");
                if (!b1) return -1;
                while (b0 < b1 && (*b0 == ' ' || *b0 == '\n' || *b0 == '\t' || *b0 == '\r')) b0++; /* trim_text(true) */
                while (b1 > b0 && (b1[-1] == ' ' || b1[-1] == '\n' || b1[-1] == '\t' || b1[-1] == '\r')) b1--;
                if (b1 > b0) {
                    if ((!f32 && !f64) || (!zl && !nz)) return -1;
                    free(arr[kind]);
                    cnt[kind] = exo_mzml_decode_binary(b0, b1 - b0, zl, f32 && !f64, &arr[kind]);
                    if (cnt[kind] < 0) return -1;
                }
            }
            q = be;
        }
        for (int k = 0; k < 3; k++)
            if (cnt[k] > 0 && (spectrum_filter < 0 || spectrum_filter == res->n_spectra))
                for (int64_t i = 0; i < cnt[k]; i++) { res->kind_sum[k] += arr[k][i]; res->kind_count[k]++; }
        if (cnt[0] > 0 && cnt[1] > 0) {
            const int64_t n = cnt[0] < cnt[1] ? cnt[0] : cnt[1];
            for (int64_t i = 0; i < n; i++)
                if (!has_pred || (arr[0][i] >= lo && arr[0][i] <= hi)) { res->sum += arr[1][i]; res->n_selected++; }
        }
        for (int k = 0; k < 3; k++) free(arr[k]);
        res->n_spectra++;
        p = se + 11;
    }
    return 0;
}
