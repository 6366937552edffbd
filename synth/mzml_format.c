/* synth/mzml_format.c -- synthetic mzML for BASELINE configs[4]: spectra of `peaks` peaks, m/z
 * sorted uniform in [100, 2000) and intensity lognormal(mu = 8, sigma = 2), both 64-bit float, uncompressed, base64.
 * Counter-based splitmix64 per spectrum.  Reports, per spectrum, the f64 sum of the intensities whose m/z lies in
 * [lo, hi] and their number (the truth of the bench query).  Bench/test infrastructure; not part of the product path. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double unif(uint64_t *s) { return ((double)(splitmix(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static int cmp_d(const void *a, const void *b) { const double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

static uint8_t *b64(uint8_t *p, const uint8_t *src, size_t n) {
    static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    size_t i = 0;
    for (; i + 3 <= n; i += 3) {
        const uint32_t t = ((uint32_t)src[i] << 16) | ((uint32_t)src[i + 1] << 8) | src[i + 2];
        *p++ = (uint8_t)T[t >> 18]; *p++ = (uint8_t)T[(t >> 12) & 63]; *p++ = (uint8_t)T[(t >> 6) & 63]; *p++ = (uint8_t)T[t & 63];
    }
    if (n - i == 1) { const uint32_t t = (uint32_t)src[i] << 16; *p++ = (uint8_t)T[t >> 18]; *p++ = (uint8_t)T[(t >> 12) & 63]; *p++ = '='; *p++ = '='; }
    if (n - i == 2) { const uint32_t t = ((uint32_t)src[i] << 16) | ((uint32_t)src[i + 1] << 8); *p++ = (uint8_t)T[t >> 18]; *p++ = (uint8_t)T[(t >> 12) & 63]; *p++ = (uint8_t)T[(t >> 6) & 63]; *p++ = '='; }
    return p;
}

int64_t synth_mzml_spectrum_max_bytes(int32_t peaks) { return 1400 + 2 * (((int64_t)peaks * 8 + 2) / 3 * 4); }

/* spectra [first, first + n) as <spectrum> elements (no file header / footer); returns bytes written */
int64_t synth_mzml_format(uint64_t seed, int64_t first, int64_t n, int32_t peaks, double lo, double hi, uint8_t *out,
                          double *sel_sum, int64_t *sel_cnt) {
    uint8_t *p = out;
    double *mz = (double *)malloc(sizeof(double) * (size_t)peaks), *in = (double *)malloc(sizeof(double) * (size_t)peaks);
    const int64_t enc = ((int64_t)peaks * 8 + 2) / 3 * 4;
    for (int64_t i = 0; i < n; i++) {
        const int64_t id = first + i;
        uint64_t s = seed ^ ((uint64_t)id * 0xD1342543DE82EF95ull);
        for (int32_t k = 0; k < peaks; k++) mz[k] = 100.0 + 1900.0 * unif(&s);
        qsort(mz, (size_t)peaks, sizeof(double), cmp_d);
        for (int32_t k = 0; k < peaks; k += 2) {
            const double u1 = unif(&s), u2 = unif(&s), r = 2.0 * sqrt(-2.0 * log(u1));
            in[k] = exp(8.0 + r * cos(6.283185307179586 * u2));
            if (k + 1 < peaks) in[k + 1] = exp(8.0 + r * sin(6.283185307179586 * u2));
        }
        double acc = 0.0;
        int64_t c = 0;
        for (int32_t k = 0; k < peaks; k++)
            if (mz[k] >= lo && mz[k] <= hi) { acc += in[k]; c++; }
        sel_sum[i] = acc;
        sel_cnt[i] = c;
        p += sprintf((char *)p,
                     "      <spectrum index=\"%lld\" id=\"scan=%lld\" defaultArrayLength=\"%d\">\n"
                     "        <cvParam cvRef=\"MS\" accession=\"MS:1000511\" name=\"ms level\" value=\"1\"/>\n"
                     "        <cvParam cvRef=\"MS\" accession=\"MS:1000127\" name=\"centroid spectrum\" value=\"\"/>\n"
                     "        <binaryDataArrayList count=\"2\">\n"
                     "          <binaryDataArray encodedLength=\"%lld\">\n"
                     "            <cvParam cvRef=\"MS\" accession=\"MS:1000523\" name=\"64-bit float\" value=\"\"/>\n"
                     "            <cvParam cvRef=\"MS\" accession=\"MS:1000576\" name=\"no compression\" value=\"\"/>\n"
                     "            <cvParam cvRef=\"MS\" accession=\"MS:1000514\" name=\"m/z array\" value=\"\" unitCvRef=\"MS\" unitAccession=\"MS:1000040\" unitName=\"m/z\"/>\n"
                     "            <binary>",
                     (long long)id, (long long)id, peaks, (long long)enc);
        p = b64(p, (const uint8_t *)mz, (size_t)peaks * 8);
        p += sprintf((char *)p,
                     "</binary>\n"
                     "          </binaryDataArray>\n"
                     "          <binaryDataArray encodedLength=\"%lld\">\n"
                     "            <cvParam cvRef=\"MS\" accession=\"MS:1000523\" name=\"64-bit float\" value=\"\"/>\n"
                     "            <cvParam cvRef=\"MS\" accession=\"MS:1000576\" name=\"no compression\" value=\"\"/>\n"
                     "            <cvParam cvRef=\"MS\" accession=\"MS:1000515\" name=\"intensity array\" value=\"\" unitCvRef=\"MS\" unitAccession=\"MS:1000131\" unitName=\"number of counts\"/>\n"
                     "            <binary>",
                     (long long)enc);
        p = b64(p, (const uint8_t *)in, (size_t)peaks * 8);
        p += sprintf((char *)p, "</binary>\n          </binaryDataArray>\n        </binaryDataArrayList>\n      </spectrum>\n");
    }
    free(mz);
    free(in);
    return p - out;
}
