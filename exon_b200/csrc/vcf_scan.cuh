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

// One tile of the launch-wide tile numbering (built on the device from the segment table, vcf_scan.cu).
struct TileDesc {
    const uint8_t *src;  // global address of tile byte 0 (16-byte aligned)
    int32_t hi;          // bytes from tile byte 0 to the end of the segment (clamped to 2^30)
    int16_t lo;          // first tile of its segment: the segment's skip (0..15); otherwise -16 (the pre-halo is staged)
    uint16_t flags;      // kTileInterior
};
constexpr uint16_t kTileInterior = 1;  // every staged byte is segment data and the tile starts at a 16-byte boundary inside it

constexpr int kMaxChrom = 255;

// error bits accumulated in ScanAcc::flags
constexpr uint32_t kErrBadPos = 1u;      // POS is not a decimal usize, is 0, or overflows int64
constexpr uint32_t kErrShortLine = 2u;   // the line ended before the field being read

// Accumulator of one query in device memory.  It is zero before the first launch that adds to it; the launch that
// finalises the query (ScanTail::finalize) leaves it zero again, so a steady-state query needs no memset.
struct ScanAcc {
    unsigned long long count;   // rows selected so far
    unsigned long long flags;   // kErr* bits
    unsigned long long ticket;  // CTAs of the running launch that have added their partials (the last one re-arms both tickets)
    unsigned long long next_tile;  // tile ticket of the running launch: tiles beyond the statically dealt rounds (vcf_scan.cu)
};

// Words of the result record the finalising CTA writes to MAPPED PINNED host memory (the host polls the sequence word:
// no device-to-host copy and no stream synchronisation on the critical path of a query).
enum ScanHostWord { kHostLocal = 0, kHostFlags = 1, kHostGlobal = 2, kHostXchgErr = 3, kHostSeq = 4, kHostWords = 8 };

// What the LAST CTA of a launch does once every CTA has added its partials (vcf_scan.cu: scan_finalize): publish the
// partition's count, exchange it with the other ranks' partials over peer memory (the final aggregate of SURVEY 8e, the
// analogue of AggregateExec(Final) -- fused into the scan's tail, no second launch), and hand the result to the host.
struct ScanTail {
    int32_t finalize;              // 0: this launch only accumulates (pushdown feeds); 1: run the tail
    int32_t n_ranks, rank;         // n_ranks >= 2: peer exchange
    int32_t pad_;
    long long *device_out;         // optional device copy of the local count
    unsigned long long *host_out;  // mapped pinned record, kHostWords words (may be NULL)
    unsigned long long host_seq;   // value written to host_out[kHostSeq] after everything else
    unsigned long long *const *peers;  // nccl.cu: peer p's exchange slots as mapped on this device
    unsigned long long xseq;       // sequence number of this exchange
};

struct ScanArgs {
    const TileDesc *tiles;  // device array, n_tiles entries (launch_build_tile_descs)
    int32_t n_segs;
    int64_t n_tiles;
    int32_t has_chrom, has_interval;
    int64_t lo, hi;
    int32_t chrom_len;
    int32_t pat_len;             // chrom_len + 2
    uint8_t pat[kMaxChrom + 5];  // '\n' + chrom + '\t'
    uint32_t static_rounds;      // set by launch_vcf_scan: rounds of the round-robin tile deal before the ticket takes over
    ScanAcc *acc;
    ScanTail tail;
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
// Enqueue the fused scan on `stream`.  With n_tiles == 0 and tail.finalize only the tail runs (one warp).
cudaError_t launch_vcf_scan(const ScanArgs &args, ScanMode mode, const ScanConfig &cfg, int sm_count,
                            cudaStream_t stream);
// Fills d_out[0 .. n_tiles) from a segment table (n_segs + 1 entries, the last one a sentinel with tile0 = n_tiles).
cudaError_t launch_build_tile_descs(const ScanSeg *d_segs, int n_segs, int64_t n_tiles, int variant, TileDesc *d_out, cudaStream_t stream);
// registers / smem / occupancy report for DESIGN.md and tests
int scan_variant_count();
const char *scan_variant_name(int variant);

}  // namespace exon
