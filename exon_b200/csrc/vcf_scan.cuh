// vcf_scan.cuh -- K1: fused VCF text -> region filter -> COUNT.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace exon {

// One contiguous run of VCF body bytes (no header) resident in device memory.  A segment starts at a line
// start; its last line may or may not end in '\n'.  Tiles are cut from the aligned `base`; the 16-byte
// granules that hold the first and the last valid byte are readable in full.
struct ScanSeg {
    const uint8_t *base;  // 16-byte aligned; the first valid byte is base[skip]
    int64_t len;          // valid bytes, starting at base + skip
    int64_t tile0;        // index of this segment's first tile in the launch-wide tile numbering
    int32_t skip;         // 0..15
    int32_t pad_;
};

constexpr int kMaxChrom = 255;

// error bits written to ScanArgs::out_flags
constexpr uint32_t kErrBadPos = 1u;      // POS is not a decimal usize, is 0, or overflows int64
constexpr uint32_t kErrShortLine = 2u;   // the line ended before the field being read

struct ScanArgs {
    const ScanSeg *segs;  // device array, n_segs + 1 entries (the last one carries tile0 = n_tiles, len = 0)
    int32_t n_segs;
    int64_t n_tiles;
    int32_t has_chrom, has_interval;
    int64_t lo, hi;
    int32_t chrom_len;
    int32_t pat_len;             // chrom_len + 2
    uint8_t pat[kMaxChrom + 5];  // '\n' + chrom + '\t'
    unsigned long long *out_count;
    uint32_t *out_flags;
};

enum ScanMode {
    kScanKey3 = 0,   // chrom is 1 byte: SWAR test of the 3-byte pattern "\n<c>\t"
    kScanKey4 = 1,   // chrom >= 2 bytes: 32-bit window compare against the last 4 pattern bytes
    kScanDense = 2,  // every line is examined (no chrom test, or strict POS validation)
    kScanLines = 3   // COUNT(*) with no predicate: count line starts
};

struct ScanConfig {
    int variant;       // 0: tile = 4 KiB/warp, 3 stages, 8 warps/CTA; see vcf_scan.cu for the table
    int ctas;          // 0 = occupancy * SM count
};

// bytes of body covered by one tile for `variant` (needed by the host to number tiles)
int scan_tile_bytes(int variant);
// Enqueue the fused scan on `stream`.  *out_count must have been zeroed on the same stream.
cudaError_t launch_vcf_scan(const ScanArgs &args, ScanMode mode, const ScanConfig &cfg, int sm_count,
                            cudaStream_t stream);
// registers / smem / occupancy report for DESIGN.md and tests
int scan_variant_count();
const char *scan_variant_name(int variant);

}  // namespace exon
