// mzml.cuh -- what the fused mzML query (mzml.cu) and the mzML column build (mzml_columns.cu) share: the per-spectrum
// descriptor the event walk leaves behind, and the scan that produces the descriptors in file order.
#pragma once
#include <stdint.h>

#include <vector>

#include "internal.h"

namespace exon {

// One <spectrum>, in file order.  Array kinds: 0 = m/z (MS:1000514), 1 = intensity (MS:1000515), 2 = wavelength (MS:1000617).
struct SpecDesc {
    const uint8_t *tag;      // the '<' of "<spectrum"
    const uint8_t *end;      // the '<' of "</spectrum>" (the end of the resident range when the file lacks it)
    const uint8_t *arr[3];   // trimmed base64 payload (NULL: the spectrum has no such array, or its <binary> is empty)
    const uint8_t *raw[3];   // zlib arrays: the inflated little-endian values (set by the inflate detour)
    uint32_t len[3];         // characters of the payload
    uint8_t f32[3], zl[3];   // 32-bit floats; zlib-compressed
    uint8_t pad_[2];
    uint32_t n_default;      // defaultArrayLength of the <spectrum> tag (0: absent or not needed)
    uint32_t seg;            // resident range the spectrum lies in
};

// The scan's result: descriptors in ctx->scratch_b (valid until the next user of that area), plus what keeps them valid.
struct MzScan {
    Ctx *ctx = nullptr;
    unsigned long long n_spec = 0;
    SpecDesc *d_spec = nullptr;
    std::vector<long long> file_spec0;  // rank of the first spectrum of every file in feed order, then n_spec
    void *z_pool[2] = {nullptr, nullptr};  // zlib detour: member table + compressed / inflated bytes (stream-ordered pool)
    unsigned long long *d_out = nullptr;   // device words [0] n_events [1] spare [2] sum (f64) [3] selected [4] flags [5] spectrum events
    ~MzScan();
};
// Events -> sort -> descriptors (+ inflate of zlib arrays).  The caller holds ctx->work_mu; starts the context's kernel timer.
int mzml_scan_spectra(VcfStream *s, MzScan *out);

#ifdef __CUDACC__
// base64 alphabet -> 6-bit values in a 256-entry shared table (0x80: not in the alphabet); blockDim.x >= 256
__device__ __forceinline__ void b64_lut_init(uint8_t *lut) {
    const int c = threadIdx.x;
    if (c < 256) {
        uint8_t v = 0x80;
        if (c >= 'A' && c <= 'Z') v = (uint8_t)(c - 'A');
        else if (c >= 'a' && c <= 'z') v = (uint8_t)(c - 'a' + 26);
        else if (c >= '0' && c <= '9') v = (uint8_t)(c - '0' + 52);
        else if (c == '+') v = 62;
        else if (c == '/') v = 63;
        else if (c == '=') v = 0;   // padding only occurs in the last group, whose padded bytes are never part of a value
        lut[c] = v;
    }
}
// value `i` (w = 4 or 8 bytes, little endian) of a base64 payload; *bad |= 0x80 on a character outside the alphabet
__device__ __forceinline__ unsigned long long b64_value(const uint8_t *p, uint32_t i, int w, const uint8_t *lut, uint32_t &bad) {
    const uint32_t o = i * (uint32_t)w, g0 = o / 3u, s = o - 3u * g0;
    const uint8_t *a = p + 4u * g0;
    const uintptr_t ai = reinterpret_cast<uintptr_t>(a);
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(ai & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(ai & 3u) * 8u;
    uint32_t cw[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) cw[k] = __ldg(wp + k);   // 16 characters at any alignment (the caller guarantees slack)
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = __funnelshift_r(cw[k], cw[k + 1], sh);
        const uint32_t v0 = lut[c & 255u], v1 = lut[(c >> 8) & 255u], v2 = lut[(c >> 16) & 255u], v3 = lut[c >> 24];
        if (k * 3 < (int)s + w) bad |= v0 | v1 | v2 | v3;   // groups past the value may be anything (the next tag)
        const uint32_t t = ((v0 & 63u) << 18) | ((v1 & 63u) << 12) | ((v2 & 63u) << 6) | (v3 & 63u);
        r[k] = __byte_perm(t, 0u, 0x4012u);  // decoded bytes of the group in memory order
    }
    const unsigned long long lo = (unsigned long long)r[0] | ((unsigned long long)r[1] << 24) | ((unsigned long long)r[2] << 48);
    const unsigned long long hi = (unsigned long long)(r[2] >> 16) | ((unsigned long long)r[3] << 8);
    return s == 0u ? lo : (lo >> (8u * s)) | (hi << (64u - 8u * s));
}

// value `i` of an inflated array (16-byte aligned, little endian)
__device__ __forceinline__ unsigned long long raw_value(const uint8_t *p, uint32_t i, int w) {
    return w == 4 ? (unsigned long long)__ldg(reinterpret_cast<const uint32_t *>(p) + i) : __ldg(reinterpret_cast<const unsigned long long *>(p) + i);
}

__device__ __forceinline__ uint32_t b64_bytes(const uint8_t *p, uint32_t len) {
    if (!len) return 0;
    uint32_t n = len / 4u * 3u;
    if (p[len - 1] == '=') --n;
    if (p[len - 2] == '=') --n;
    return n;
}

#endif
}  // namespace exon
