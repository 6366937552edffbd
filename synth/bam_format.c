/* synth/bam_format.c -- synthetic BAM records for BASELINE configs[3]: n_ref references,
 * l_seq bases, one <l_seq>M CIGAR operation, name r%09d, flags from a fixed table with fixed probabilities, MAPQ uniform
 * 0..60 with 2 % 255, no tags.  Counter-based splitmix64 per record (any row range can be produced independently).
 * Also reports where an htslib-style writer would cut BGZF blocks (a record is never split unless larger than a block).
 * Bench/test infrastructure; not part of the product path. */
#include <stdint.h>
#include <string.h>

static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

static const uint16_t FLAGS[12] = {0, 16, 83, 99, 147, 163, 4, 1024 + 99, 256 + 16, 2048, 77, 141};
static const uint8_t FLAG_CDF[12] = {30, 60, 90, 120, 150, 180, 200, 215, 230, 240, 248, 255}; /* out of 256 */

int64_t synth_bam_record_bytes(int32_t l_seq) { return 4 + 32 + 11 + 4 + (l_seq + 1) / 2 + l_seq; }

/* records [first, first + n): out gets the bytes, ref_id / flag / mapq get the per-record truth.
 * ref_cdf: n_ref cumulative weights scaled to 2^32.  Returns bytes written. */
int64_t synth_bam_format(uint64_t seed, int64_t first, int64_t n, int32_t l_seq, int32_t n_ref, const uint32_t *ref_cdf,
                         const int32_t *ref_len, uint8_t *out, int32_t *ref_id_out, uint16_t *flag_out, uint8_t *mapq_out) {
    uint8_t *p = out;
    const int32_t block_size = (int32_t)synth_bam_record_bytes(l_seq) - 4;
    for (int64_t i = 0; i < n; i++) {
        const int64_t id = first + i;
        uint64_t s = seed ^ ((uint64_t)id * 0xD1342543DE82EF95ull);
        uint64_t r = splitmix(&s);
        int fi = 0;
        while (FLAG_CDF[fi] < (r & 255)) fi++;
        const uint16_t flag = FLAGS[fi];
        r >>= 8;
        uint8_t mapq = (uint8_t)((r & 0xFFFF) % 61);
        if (((r >> 16) & 0xFF) < 5) mapq = 255; /* ~2 % */
        r = splitmix(&s);
        const uint32_t u = (uint32_t)r;
        int32_t ref = 0;
        while (ref < n_ref - 1 && ref_cdf[ref] < u) ref++;
        if ((flag & 4) && ((r >> 32) & 1)) ref = -1; /* half of the unmapped reads are unplaced */
        const int32_t pos = ref >= 0 ? (int32_t)((r >> 33) % (uint64_t)(ref_len[ref] > l_seq ? ref_len[ref] - l_seq : 1)) : -1;
        put32(p, (uint32_t)block_size);
        put32(p + 4, (uint32_t)ref);
        put32(p + 8, (uint32_t)pos);
        p[12] = 11; /* l_read_name incl. NUL */
        p[13] = mapq;
        p[14] = 0x48; p[15] = 0x12; /* bin (not used by the path) */
        p[16] = 1; p[17] = 0;       /* n_cigar_op */
        p[18] = (uint8_t)flag; p[19] = (uint8_t)(flag >> 8);
        put32(p + 20, (uint32_t)l_seq);
        put32(p + 24, 0xFFFFFFFFu); /* next_refID -1 */
        put32(p + 28, 0xFFFFFFFFu);
        put32(p + 32, 0);
        uint8_t *q = p + 36;
        *q++ = 'r';
        { int64_t v = id % 1000000000ll; for (int k = 8; k >= 0; k--) { q[k] = (uint8_t)('0' + v % 10); v /= 10; } q += 9; }
        *q++ = 0;
        put32(q, ((uint32_t)l_seq << 4) | 0u); /* <l_seq>M */
        q += 4;
        for (int32_t j = 0; j < (l_seq + 1) / 2; j++) { if ((j & 7) == 0) r = splitmix(&s); *q++ = (uint8_t)((1u << (r & 3)) | ((1u << ((r >> 2) & 3)) << 4)); r >>= 4; }
        for (int32_t j = 0; j < l_seq; j++) { if ((j & 7) == 0) r = splitmix(&s); *q++ = (uint8_t)(2 + (r & 31)); r >>= 5; }
        p = q;
        ref_id_out[i] = ref;
        flag_out[i] = flag;
        mapq_out[i] = mapq;
    }
    return p - out;
}
