/* synth/fastq_format.c -- synthetic FASTQ reads for BASELINE configs[1]: fixed length L,
 * name r%09d, no description, bases uniform ACGT, per-read mean mu ~ N(30, 5) clipped to [5, 40], per-base
 * q = clip(round(N(mu, 3)), 2, 41), byte = q + 33.  The generator is a counter-based splitmix64 stream (one
 * stream per read, so any row range can be produced independently and in parallel); SURVEY.md names numpy's
 * default_rng, which is too slow for 1.5e9 normal deviates -- the distribution is the same, the truth is computed
 * from the generated integers either way.  Bench/test infrastructure; not part of the product path. */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double unif(uint64_t *s) { return ((double)(splitmix(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

/* Writes reads [first, first + n) into out (n * (read_len * 2 + 16) bytes are enough); sum_q[i] receives the
 * integer sum of the read's Phred scores.  Returns bytes written. */
int64_t synth_fastq_format(uint64_t seed, int64_t first, int64_t n, int32_t read_len, uint8_t *out, int32_t *sum_q) {
    static const char B[4] = {'A', 'C', 'G', 'T'};
    uint8_t *p = out;
    for (int64_t i = 0; i < n; i++) {
        const int64_t id = first + i;
        uint64_t s = seed ^ ((uint64_t)id * 0xD1342543DE82EF95ull);
        /* name */
        *p++ = '@';
        *p++ = 'r';
        {
            char tmp[9];
            int64_t v = id % 1000000000ll;
            for (int k = 8; k >= 0; k--) { tmp[k] = (char)('0' + v % 10); v /= 10; }
            memcpy(p, tmp, 9);
            p += 9;
        }
        *p++ = '\n';
        /* bases: 32 per 64-bit draw */
        for (int32_t j = 0; j < read_len; j += 32) {
            uint64_t r = splitmix(&s);
            for (int32_t k = 0; k < 32 && j + k < read_len; k++, r >>= 2) *p++ = (uint8_t)B[r & 3];
        }
        *p++ = '\n';
        *p++ = '+';
        *p++ = '\n';
        /* qualities: Box-Muller, two deviates per pair of uniforms */
        double mu;
        {
            const double u1 = unif(&s), u2 = unif(&s);
            mu = 30.0 + 5.0 * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            if (mu < 5.0) mu = 5.0;
            if (mu > 40.0) mu = 40.0;
        }
        int32_t sum = 0;
        for (int32_t j = 0; j < read_len; j += 2) {
            const double u1 = unif(&s), u2 = unif(&s);
            const double rad = 3.0 * sqrt(-2.0 * log(u1));
            const double z[2] = {rad * cos(6.283185307179586 * u2), rad * sin(6.283185307179586 * u2)};
            for (int k = 0; k < 2 && j + k < read_len; k++) {
                long q = lround(mu + z[k]);
                if (q < 2) q = 2;
                if (q > 41) q = 41;
                sum += (int32_t)q;
                *p++ = (uint8_t)(q + 33);
            }
        }
        *p++ = '\n';
        sum_q[i] = sum;
    }
    return p - out;
}
