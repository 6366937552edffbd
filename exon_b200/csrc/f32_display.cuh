// f32_display.cuh -- Rust `impl Display for f32` / `impl Display for i32` on the device (and on the host, for pinning).
//
// The reference's VCF builder does not copy INFO / FORMAT values: it prints noodles' typed view of them again -- integers
// through `i32::to_string`, floats through `f32::to_string` (exon/exon-vcf/src/array_builder/lazy_array_builder.rs:236-242,
// 253-262, 325-327, 379-389).  Rust's Display for a float prints the SHORTEST decimal digits that parse back to the same
// f32 (ties resolved towards the closest decimal), never with an exponent: 1.0 -> "1", 0.1 -> "0.1", 1e-7 -> "0.0000001",
// 3.4028235e38 -> "340282350000000000000000000000000000000", -0.0 -> "-0", NaN -> "NaN", +-inf -> "inf" / "-inf".
// The digit generation is the published Ryu algorithm for binary32 (Adams, PLDI 2018): exact 64-bit integer arithmetic
// against two tables of powers of five; the tables below were generated from their definitions
//   inv[i] = floor(2^(pow5bits(i) - 1 + 59) / 5^i) + 1,   spl[i] = 5^i scaled to 61 significant bits
// and the routine is pinned against numpy's shortest positional printing on random bit patterns
// (tests/test_vcf_wide_golden.py::test_format_f32_matches_shortest_repr) through exon_gpu_format_f32.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define EXON_FD_HD __host__ __device__ __forceinline__
#else
#define EXON_FD_HD inline
#endif

namespace exon {

constexpr int kF32DisplayMax = 56;  // "-0." + 44 zeros + 9 digits

#ifdef __CUDA_ARCH__
#define EXON_F32D_TABLE __constant__
#else
#define EXON_F32D_TABLE static const
#endif

namespace f32d {
EXON_F32D_TABLE uint64_t kPow5Inv[31] = {
    0x0800000000000001ull, 0x0666666666666667ull, 0x051eb851eb851eb9ull, 0x04189374bc6a7efaull,
    0x068db8bac710cb2aull, 0x053e2d6238da3c22ull, 0x0431bde82d7b634eull, 0x06b5fca6af2bd216ull,
    0x055e63b88c230e78ull, 0x044b82fa09b5a52dull, 0x06df37f675ef6eaeull, 0x057f5ff85e592558ull,
    0x0465e6604b7a8447ull, 0x0709709a125da071ull, 0x05a126e1a84ae6c1ull, 0x0480ebe7b9d58567ull,
    0x0734aca5f6226f0bull, 0x05c3bd5191b525a3ull, 0x049c97747490eae9ull, 0x0760f253edb4ab0eull,
    0x05e72843249088d8ull, 0x04b8ed0283a6d3e0ull, 0x078e480405d7b966ull, 0x060b6cd004ac9452ull,
    0x04d5f0a66a23a9dbull, 0x07bcb43d769f762bull, 0x063090312bb2c4efull, 0x04f3a68dbc8f03f3ull,
    0x07ec3daf94180651ull, 0x065697bfa9acd1daull, 0x051212ffbaf0a7e2ull};
EXON_F32D_TABLE uint64_t kPow5[47] = {
    0x1000000000000000ull, 0x1400000000000000ull, 0x1900000000000000ull, 0x1f40000000000000ull,
    0x1388000000000000ull, 0x186a000000000000ull, 0x1e84800000000000ull, 0x1312d00000000000ull,
    0x17d7840000000000ull, 0x1dcd650000000000ull, 0x12a05f2000000000ull, 0x174876e800000000ull,
    0x1d1a94a200000000ull, 0x12309ce540000000ull, 0x16bcc41e90000000ull, 0x1c6bf52634000000ull,
    0x11c37937e0800000ull, 0x16345785d8a00000ull, 0x1bc16d674ec80000ull, 0x1158e460913d0000ull,
    0x15af1d78b58c4000ull, 0x1b1ae4d6e2ef5000ull, 0x10f0cf064dd59200ull, 0x152d02c7e14af680ull,
    0x1a784379d99db420ull, 0x108b2a2c28029094ull, 0x14adf4b7320334b9ull, 0x19d971e4fe8401e7ull,
    0x1027e72f1f128130ull, 0x1431e0fae6d7217cull, 0x193e5939a08ce9dbull, 0x1f8def8808b02452ull,
    0x13b8b5b5056e16b3ull, 0x18a6e32246c99c60ull, 0x1ed09bead87c0378ull, 0x13426172c74d822bull,
    0x1812f9cf7920e2b6ull, 0x1e17b84357691b64ull, 0x12ced32a16a1b11eull, 0x178287f49c4a1d66ull,
    0x1d6329f1c35ca4bfull, 0x125dfa371a19e6f7ull, 0x16f578c4e0a060b5ull, 0x1cb2d6f618c878e3ull,
    0x11efc659cf7d4b8dull, 0x166bb7f0435c9e71ull, 0x1c06a5ec5433c60dull};

EXON_FD_HD uint32_t pow5bits(int32_t e) { return (uint32_t)(((uint32_t)e * 1217359u) >> 19) + 1u; }
EXON_FD_HD uint32_t log10pow2(int32_t e) { return ((uint32_t)e * 78913u) >> 18; }
EXON_FD_HD uint32_t log10pow5(int32_t e) { return ((uint32_t)e * 732923u) >> 20; }
EXON_FD_HD uint32_t pow5factor(uint32_t v) {
    uint32_t c = 0;
    while (v && v % 5u == 0u) {
        v /= 5u;
        ++c;
    }
    return c;
}
EXON_FD_HD uint32_t mulshift(uint32_t m, uint64_t factor, int32_t shift) {  // (m * factor) >> shift, shift > 32
    const uint64_t b0 = (uint64_t)m * (uint32_t)factor, b1 = (uint64_t)m * (uint32_t)(factor >> 32);
    return (uint32_t)(((b0 >> 32) + b1) >> (shift - 32));
}
}  // namespace f32d

// shortest decimal (digits, exponent) with value = digits * 10^exponent for a finite non-zero f32 given as mantissa / exponent fields
EXON_FD_HD void f32_shortest(uint32_t ieee_m, uint32_t ieee_e, uint32_t *digits, int32_t *exp10) {
    using namespace f32d;
    int32_t e2;
    uint32_t m2;
    if (ieee_e == 0) {
        e2 = 1 - 127 - 23 - 2;
        m2 = ieee_m;
    } else {
        e2 = (int32_t)ieee_e - 127 - 23 - 2;
        m2 = (1u << 23) | ieee_m;
    }
    const bool accept = (m2 & 1u) == 0u;
    const uint32_t mv = 4u * m2, mp = 4u * m2 + 2u, mm_shift = (ieee_m != 0u || ieee_e <= 1u) ? 1u : 0u, mm = 4u * m2 - 1u - mm_shift;
    uint32_t vr, vp, vm, last = 0;
    int32_t e10;
    bool vm_tz = false, vr_tz = false;
    if (e2 >= 0) {
        const uint32_t q = log10pow2(e2);
        e10 = (int32_t)q;
        const int32_t k = 59 + (int32_t)pow5bits((int32_t)q) - 1, i = -e2 + (int32_t)q + k;
        vr = mulshift(mv, kPow5Inv[q], i);
        vp = mulshift(mp, kPow5Inv[q], i);
        vm = mulshift(mm, kPow5Inv[q], i);
        if (q != 0 && (vp - 1u) / 10u <= vm / 10u) {
            const int32_t l = 59 + (int32_t)pow5bits((int32_t)q - 1) - 1;
            last = mulshift(mv, kPow5Inv[q - 1], -e2 + (int32_t)q - 1 + l) % 10u;
        }
        if (q <= 9) {
            if (mv % 5u == 0u) vr_tz = pow5factor(mv) >= q;
            else if (accept) vm_tz = pow5factor(mm) >= q;
            else vp -= pow5factor(mp) >= q ? 1u : 0u;
        }
    } else {
        const uint32_t q = log10pow5(-e2);
        e10 = (int32_t)q + e2;
        const int32_t i = -e2 - (int32_t)q, k = (int32_t)pow5bits(i) - 61;
        int32_t j = (int32_t)q - k;
        vr = mulshift(mv, kPow5[i], j);
        vp = mulshift(mp, kPow5[i], j);
        vm = mulshift(mm, kPow5[i], j);
        if (q != 0 && (vp - 1u) / 10u <= vm / 10u) {
            j = (int32_t)q - 1 - ((int32_t)pow5bits(i + 1) - 61);
            last = mulshift(mv, kPow5[i + 1], j) % 10u;
        }
        if (q <= 1) {
            vr_tz = true;
            if (accept) vm_tz = mm_shift == 1u;
            else --vp;
        } else if (q < 31) {
            vr_tz = (mv & ((1u << (q - 1)) - 1u)) == 0u;
        }
    }
    int32_t removed = 0;
    uint32_t out;
    if (vm_tz || vr_tz) {
        while (vp / 10u > vm / 10u) {
            vm_tz &= vm % 10u == 0u;
            vr_tz &= last == 0u;
            last = vr % 10u;
            vr /= 10u;
            vp /= 10u;
            vm /= 10u;
            ++removed;
        }
        if (vm_tz) {
            while (vm % 10u == 0u) {
                vr_tz &= last == 0u;
                last = vr % 10u;
                vr /= 10u;
                vp /= 10u;
                vm /= 10u;
                ++removed;
            }
        }
        if (vr_tz && last == 5u && vr % 2u == 0u) last = 4u;  // round to even on an exact tie
        out = vr + (((vr == vm && (!accept || !vm_tz)) || last >= 5u) ? 1u : 0u);
    } else {
        while (vp / 10u > vm / 10u) {
            last = vr % 10u;
            vr /= 10u;
            vp /= 10u;
            vm /= 10u;
            ++removed;
        }
        out = vr + ((vr == vm || last >= 5u) ? 1u : 0u);
    }
    *digits = out;
    *exp10 = e10 + removed;
}

// `f32::to_string()`: writes at most kF32DisplayMax bytes to out (NULL: only measures), returns the length
EXON_FD_HD int f32_display(float v, uint8_t *out) {
    uint32_t bits;
    memcpy(&bits, &v, 4);
    const bool neg = (bits >> 31) != 0u;
    const uint32_t ieee_m = bits & 0x7FFFFFu, ieee_e = (bits >> 23) & 0xFFu;
    int n = 0;
    auto put = [&](uint8_t c) {
        if (out) out[n] = c;
        ++n;
    };
    if (ieee_e == 0xFFu) {
        if (ieee_m) {
            put('N'); put('a'); put('N');
            return n;
        }
        if (neg) put('-');
        put('i'); put('n'); put('f');
        return n;
    }
    if (neg) put('-');
    if (ieee_e == 0 && ieee_m == 0) {
        put('0');
        return n;
    }
    uint32_t digits;
    int32_t e10;
    f32_shortest(ieee_m, ieee_e, &digits, &e10);
    uint8_t d[10];
    int nd = 0;
    while (digits) {
        d[nd++] = (uint8_t)('0' + digits % 10u);
        digits /= 10u;
    }  // d[] holds the digits least significant first
    const int point = nd + e10;  // digits before the decimal point
    if (point <= 0) {
        put('0'); put('.');
        for (int k = 0; k < -point; ++k) put('0');
        for (int k = nd - 1; k >= 0; --k) put(d[k]);
    } else if (point >= nd) {
        for (int k = nd - 1; k >= 0; --k) put(d[k]);
        for (int k = 0; k < point - nd; ++k) put('0');
    } else {
        for (int k = nd - 1; k >= 0; --k) {
            if (nd - 1 - k == point) put('.');
            put(d[k]);
        }
    }
    return n;
}

// `i32::to_string()`: at most 11 bytes
EXON_FD_HD int i32_display(int32_t v, uint8_t *out) {
    uint8_t d[10];
    int nd = 0, n = 0;
    uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
    do {
        d[nd++] = (uint8_t)('0' + u % 10u);
        u /= 10u;
    } while (u);
    if (v < 0) {
        if (out) out[n] = '-';
        ++n;
    }
    for (int k = nd - 1; k >= 0; --k) {
        if (out) out[n] = d[k];
        ++n;
    }
    return n;
}

}  // namespace exon
