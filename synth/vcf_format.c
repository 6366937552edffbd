/* synth/vcf_format.c -- serialises the synthetic VCF integer columns (synth/vcf.py) into VCF 4.2 text.
 * Bench/test infrastructure; not part of the product path. */
#include <stdint.h>
#include <string.h>

static inline uint8_t *put_u64(uint8_t *p, uint64_t v) {
    char tmp[20];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = (uint8_t)tmp[--n];
    return p;
}

/* One line per record: CHROM POS . REF ALT QUAL PASS . ; qual < 0 prints '.'.
 * names: n_contigs fixed-width (8 byte, NUL padded) contig names.  Returns bytes written. */
int64_t synth_vcf_format(const uint8_t *contig, const int64_t *pos, const uint8_t *ref, const uint8_t *alt,
                         const int8_t *qual, int64_t n, const char *names, uint8_t *out) {
    static const char B[4] = {'A', 'C', 'G', 'T'};
    uint8_t *p = out;
    for (int64_t i = 0; i < n; i++) {
        const char *nm = names + 8 * contig[i];
        for (int k = 0; k < 8 && nm[k]; k++) *p++ = (uint8_t)nm[k];
        *p++ = '\t';
        p = put_u64(p, (uint64_t)pos[i]);
        memcpy(p, "\t.\t", 3); p += 3;
        *p++ = (uint8_t)B[ref[i] & 3];
        *p++ = '\t';
        *p++ = (uint8_t)B[alt[i] & 3];
        *p++ = '\t';
        if (qual[i] < 0) *p++ = '.'; else p = put_u64(p, (uint64_t)qual[i]);
        memcpy(p, "\tPASS\t.\n", 8); p += 8;
    }
    return p - out;
}
