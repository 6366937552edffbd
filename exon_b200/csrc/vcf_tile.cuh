// vcf_tile.cuh -- pieces shared by the warp-private TMA tile pipelines of K1 (vcf_scan.cu) and K2 (vcf_columns.cu):
// the staged-tile geometry, the byte-exact tile view used by the scalar routines, SWAR packing helpers and the
// shared-memory layout of the per-warp rings.
#pragma once
#include "common.cuh"
#include "vcf_scan.cuh"

namespace exon {
namespace {

constexpr int kPre = 16;   // bytes staged before the tile (right-aligned POS fetch may reach back 12 bytes)
constexpr int kHalo = 48;  // bytes staged after the tile (line window + POS digits of a line that starts at the end)

// ===================================================================================================
// Byte-exact view of a staged tile (used by the scalar routines of both kernels)
// ===================================================================================================
struct TileView {
    const uint8_t *sm;  // shared-memory address of tile byte 0
    const uint8_t *g;   // global address of tile byte 0
    int lo;             // smallest tile-relative index inside the segment (<= 0)
    int hi;             // one past the largest (> 0)
    int sm_lo, sm_hi;   // tile-relative index range present in shared memory
};

// Byte at tile-relative index i.  Outside the segment reads as '\n' (a record can neither start before the
// segment nor continue past its end); outside the staged window falls back to a global load.
__device__ __forceinline__ uint32_t ld_byte(const TileView &t, int i) {
    if (i < t.lo || i >= t.hi) return '\n';
    if (i >= t.sm_lo && i < t.sm_hi) return t.sm[i];
    return __ldg(t.g + i);
}


__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t s) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));  // shift amounts >= 32 give 0
    return r;
}

// 0x80 flags in up to 16 bytes -> 16-bit mask, bit i = byte i flagged (IDP.4A: sum of flag * weight, flags are 128)
__device__ __forceinline__ uint32_t pack16(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
    uint32_t a = __dp4a(f0, 0x08040201u, 0u);
    a = __dp4a(f1, 0x80402010u, a);
    uint32_t b = __dp4a(f2, 0x08040201u, 0u);
    b = __dp4a(f3, 0x80402010u, b);
    return (a >> 7) | (b << 1);
}


constexpr int kQueue = 256;  // line starts a warp collects before it parses them (uint16 each)

struct StageMeta {
    const uint8_t *g;  // global address of tile byte 0
    int lo;            // first tile-relative index that is staged and inside the segment (>= 0: first tile of its segment)
    int hi;            // bytes from tile byte 0 to the end of the segment (clamped to 2^30)
};

template <int U, int S, int WARPS>
struct SmemLayout {
    static constexpr int TILE = 512 * U;
    static constexpr int STAGE = ((kPre + TILE + kHalo + 127) / 128) * 128;
    static constexpr size_t ring = 0;
    static constexpr size_t bars = (size_t)WARPS * S * STAGE;
    static constexpr size_t meta = bars + (size_t)WARPS * S * sizeof(uint64_t);
    static constexpr size_t queue = meta + (size_t)WARPS * S * sizeof(StageMeta);
    static constexpr size_t total = queue + (size_t)WARPS * kQueue * sizeof(uint16_t);
};

template <int U, int S, int WARPS>
constexpr int ctas_per_sm() {
    const int by_smem = (int)((227 * 1024) / SmemLayout<U, S, WARPS>::total);
    const int by_warps = 32 / WARPS;  // at most 32 resident warps: the parser wants >= 64 registers per thread
    const int c = by_smem < by_warps ? by_smem : by_warps;
    return c < 1 ? 1 : c;
}


}  // namespace
}  // namespace exon
