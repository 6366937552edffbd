// f32_parse.cuh -- decimal text -> f32, correctly rounded (round to nearest, ties to even), with the grammar of Rust's
// `f32::from_str` (core::num::dec2flt), which is what noodles-vcf 0.70 applies to the QUAL field
// (`Record::quality_score`, called at exon/exon-vcf/src/array_builder/lazy_array_builder.rs:205-208):
//     [+-]? ( "inf" | "infinity" | "nan" (any case) | digits* ( "." digits* )? ( [eE] [+-]? digits+ )? )   with >= 1 mantissa digit
// Two paths, both exact:
//   fast   <= 19 significant digits, mantissa <= 2^24 and |exp10| <= 10: one IEEE f32 multiply or divide of two exactly
//          representable numbers (Clinger's fast path)
//   exact  a double-precision estimate picks a candidate f32; the decimal value D * 10^E (D < 10^36 as 128 bits) is then
//          compared, in integer arithmetic of 16 x 32-bit limbs, with the midpoints between the candidate and its
//          neighbours, and the candidate moves until the value lies between them (ties to the even bit pattern)
// More than 36 significant digits with a non-zero digit beyond the 36th are reported as unsupported, not approximated.
// Host + device: the host instance backs exon_gpu_parse_f32 (tests pin it against exact rational arithmetic).
#pragma once
#include <stdint.h>
#include <string.h>

namespace exon {

#ifdef __CUDACC__
#define EXON_HD __host__ __device__
#define EXON_HD_NOINLINE __host__ __device__ __noinline__
#else
#define EXON_HD
#define EXON_HD_NOINLINE
#endif

constexpr int kF32Ok = 0, kF32Malformed = 1, kF32Unsupported = 2;

struct Big512 {
    uint32_t w[16];
};

EXON_HD inline void big_mul_small(Big512 &b, uint32_t x) {
    uint64_t carry = 0;
    for (int i = 0; i < 16; ++i) {
        const uint64_t t = (uint64_t)b.w[i] * x + carry;
        b.w[i] = (uint32_t)t;
        carry = t >> 32;
    }
}
EXON_HD inline void big_shl(Big512 &b, int k) {  // k < 512
    const int ws = k >> 5, bs = k & 31;
    for (int i = 15; i >= 0; --i) {
        const uint32_t lo = i - ws >= 0 ? b.w[i - ws] : 0u;
        const uint32_t lo2 = i - ws - 1 >= 0 ? b.w[i - ws - 1] : 0u;
        b.w[i] = bs ? ((lo << bs) | (lo2 >> (32 - bs))) : lo;
    }
}
EXON_HD inline void big_mul_pow10(Big512 &b, int e) {
    while (e >= 9) {
        big_mul_small(b, 1000000000u);
        e -= 9;
    }
    uint32_t p = 1;
    while (e-- > 0) p *= 10u;
    if (p > 1) big_mul_small(b, p);
}
EXON_HD inline int big_cmp(const Big512 &a, const Big512 &b) {
    for (int i = 15; i >= 0; --i)
        if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
}

// sign of  D * 10^E  -  midpoint(bits, bits + 1)  for a non-negative finite f32 bit pattern
EXON_HD inline int f32_cmp_mid(const uint32_t D[4], int E, uint32_t bits) {
    uint32_t M;
    int e;
    if (bits < 0x00800000u) {
        M = bits;
        e = -149;
    } else {
        M = (bits & 0x007FFFFFu) | 0x00800000u;
        e = (int)(bits >> 23) - 150;
    }
    const int e2 = e - 1;  // midpoint = (2M + 1) * 2^(e - 1)
    Big512 lhs, rhs;
    for (int i = 0; i < 16; ++i) lhs.w[i] = i < 4 ? D[i] : 0u, rhs.w[i] = 0u;
    rhs.w[0] = 2u * M + 1u;
    if (E > 0) big_mul_pow10(lhs, E);
    else if (E < 0) big_mul_pow10(rhs, -E);
    if (e2 < 0) big_shl(lhs, -e2);
    else if (e2 > 0) big_shl(rhs, e2);
    return big_cmp(lhs, rhs);
}

EXON_HD inline bool f32_ieq(const uint8_t *s, int n, const char *lit, int m) {
    if (n != m) return false;
    for (int i = 0; i < n; ++i)
        if ((s[i] | 0x20u) != (uint8_t)lit[i]) return false;
    return true;
}

EXON_HD inline float f32_from_bits(uint32_t b) {
    float f;
    memcpy(&f, &b, 4);
    return f;
}

EXON_HD inline int parse_f32_rust(const uint8_t *s, int n, float *out) {
    int i = 0;
    uint32_t sign = 0;
    if (n > 0 && (s[0] == '+' || s[0] == '-')) {
        sign = s[0] == '-' ? 0x80000000u : 0u;
        i = 1;
    }
    if (i >= n) return kF32Malformed;
    if (f32_ieq(s + i, n - i, "inf", 3) || f32_ieq(s + i, n - i, "infinity", 8)) {
        *out = f32_from_bits(sign | 0x7F800000u);
        return kF32Ok;
    }
    if (f32_ieq(s + i, n - i, "nan", 3)) {
        *out = f32_from_bits(sign | 0x7FC00000u);
        return kF32Ok;
    }
    const int m0 = i;
    uint64_t m = 0;
    int nsig = 0, adj = 0;  // value = m * 10^(adj + exponent field) while nsig <= 19
    bool any = false, dot = false;
    for (; i < n; ++i) {
        const uint32_t c = s[i];
        if (c == '.') {
            if (dot) return kF32Malformed;
            dot = true;
            continue;
        }
        const uint32_t d = c - '0';
        if (d > 9u) break;
        any = true;
        if (nsig == 0 && d == 0) {
            if (dot) --adj;
            continue;
        }
        if (nsig < 19) {
            m = m * 10u + d;
            if (dot) --adj;
        } else if (!dot) {
            ++adj;
        }
        ++nsig;
    }
    const int m1 = i;
    if (!any) return kF32Malformed;
    int ex = 0;
    if (i < n && (s[i] | 0x20u) == 'e') {
        ++i;
        bool eneg = false;
        if (i < n && (s[i] == '+' || s[i] == '-')) {
            eneg = s[i] == '-';
            ++i;
        }
        if (i >= n) return kF32Malformed;
        for (; i < n; ++i) {
            const uint32_t d = (uint32_t)s[i] - '0';
            if (d > 9u) return kF32Malformed;
            if (ex < 100000) ex = ex * 10 + (int)d;
        }
        if (eneg) ex = -ex;
    }
    if (i != n) return kF32Malformed;
    if (nsig == 0) {
        *out = f32_from_bits(sign);
        return kF32Ok;
    }
    const int e10 = adj + ex;
    if (nsig <= 19 && m <= (1u << 24) && e10 >= -10 && e10 <= 10) {
        const float p10[11] = {1e0f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
        const float f = e10 >= 0 ? (float)(uint32_t)m * p10[e10] : (float)(uint32_t)m / p10[-e10];
        uint32_t b;
        memcpy(&b, &f, 4);
        *out = f32_from_bits(b | sign);
        return kF32Ok;
    }
    // ---- exact path: D = the first min(nsig, 36) significant digits, value = D * 10^E ----
    uint32_t D[4] = {(uint32_t)m, (uint32_t)(m >> 32), 0u, 0u};
    int E = e10, nd = nsig;
    if (nsig > 19) {
        D[0] = D[1] = 0u;
        int k = 0, adj36 = 0;
        bool dt = false;
        for (int j = m0; j < m1; ++j) {
            const uint32_t c = s[j];
            if (c == '.') {
                dt = true;
                continue;
            }
            const uint32_t d = c - '0';
            if (k == 0 && d == 0) {
                if (dt) --adj36;
                continue;
            }
            if (k < 36) {
                uint64_t carry = d;
                for (int q = 0; q < 4; ++q) {
                    const uint64_t t = (uint64_t)D[q] * 10u + carry;
                    D[q] = (uint32_t)t;
                    carry = t >> 32;
                }
                if (dt) --adj36;
            } else {
                if (d != 0) return kF32Unsupported;
                if (!dt) ++adj36;
            }
            ++k;
        }
        E = adj36 + ex;
        nd = nsig < 36 ? nsig : 36;
    }
    if (E + nd > 40) {
        *out = f32_from_bits(sign | 0x7F800000u);
        return kF32Ok;
    }
    if (E + nd < -46) {
        *out = f32_from_bits(sign);
        return kF32Ok;
    }
    // estimate
    double a = ((double)D[3] * 4294967296.0 + (double)D[2]) * 18446744073709551616.0 + ((double)D[1] * 4294967296.0 + (double)D[0]);
    {
        const double p22[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        int e = E;
        while (e > 22) a *= 1e22, e -= 22;
        while (e < -22) a /= 1e22, e += 22;
        a = e >= 0 ? a * p22[e] : a / p22[-e];
    }
    uint32_t b;
    if (a >= 3.5e38) {
        b = 0x7F800000u;
    } else {
        const float f = (float)a;
        memcpy(&b, &f, 4);
        b &= 0x7FFFFFFFu;
    }
    for (int it = 0; it < 8; ++it) {
        if (b > 0u) {
            const int c = f32_cmp_mid(D, E, b - 1u);
            if (c < 0) {
                --b;
                continue;
            }
            if (c == 0) {
                if (b & 1u) --b;
                break;
            }
        }
        if (b < 0x7F800000u) {
            const int c = f32_cmp_mid(D, E, b);
            if (c > 0) {
                ++b;
                continue;
            }
            if (c == 0) {
                if (b & 1u) ++b;
                break;
            }
        }
        *out = f32_from_bits(b | sign);
        return kF32Ok;
    }
    // tie exits land here; a candidate that never settles cannot happen with an estimate within one f32 ulp
    *out = f32_from_bits(b | sign);
    return kF32Ok;
}

}  // namespace exon
