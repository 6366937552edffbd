// bam.cu -- BAM record stream -> flag / MAPQ predicate -> per-reference COUNT, fused (BASELINE.json configs[3]).
//
// Replaces, for `SELECT reference, COUNT(*) FROM bam WHERE <flag tests> AND mapping_quality >= q GROUP BY reference`:
//   BAMOpener::open                 exon/exon-core/src/datasources/bam/file_opener.rs:39 (BGZF reader, header, references)
//   BatchReader::read_batch          exon/exon-bam/src/batch_reader.rs:70-107 (read_record_buf per row)
//   BAMArrayBuilder::append 1, 2, 5  exon/exon-bam/src/array_builder.rs:102-143 (flag u16 -> i32, reference name of
//                                    refID, mapping_quality as a string, NULL when 255)
//   is_unmapped / is_secondary / ... exon/exon-core/src/udfs/sam/samflags.rs:26-47, 111-141 (flag bit tests)
//   FilterExec + AggregateExec(Partial) GROUP BY reference                          (DataFusion 44, third party)
// The file's BGZF members are inflated on the device (bgzf.cu) into one contiguous record stream.  A BAM record is
// found only through the block_size of the record before it, so the chain is serial per file -- but htslib-style
// writers never let a record straddle a member unless it is larger than one, so almost every member starts with a
// record.  The kernel therefore walks every member's chain SPECULATIVELY from the member's start, one thread per
// member (thousands of independent chains in flight), evaluating the predicate and counting per reference as it
// goes, and records where each walk left its member.  A second kernel checks that every walk ended exactly where
// the next one began; if so (the usual case) the speculation was right for every member by induction from the
// first record, and the counts stand.  Otherwise the entry points are corrected from the exits and the pass runs
// again (records that straddle members), and after a few rounds one thread per file walks the chain serially --
// slow, but exact for any input.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <new>
#include <string>

#include "common.cuh"
#include "f32_display.cuh"
#include "internal.h"
#include "scan_i64.cuh"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

struct BamEntry {
    const uint8_t *base;   // first byte of the file's uncompressed stream
    uint64_t start;        // where this walk begins (speculation: the member's first byte)
    uint64_t end;          // first byte of the next member (walks stop at the first record that starts at or after it)
    uint64_t total;        // bytes in the file's stream
    int32_t remap0;        // index of the file's refID -> group table in `remap`
    int32_t n_ref;
    int32_t last_of_file;  // 1: the walk must end exactly at `total`
    int32_t file_idx;      // index of the file in the stream (per-file query parameters)
};

struct BamArgs {
    const BamEntry *entries;
    int32_t n_entries;
    const int32_t *remap;          // per file: n_ref + 1 group ids (the last one for refID -1)
    uint64_t *exits;               // out: where each walk stopped (UINT64_MAX: malformed record)
    unsigned long long *counts;    // out: n_groups selected-record counts
    unsigned long long *rows;      // out: records seen
    int32_t n_groups;
    int32_t has_pred;
    uint32_t flag_exclude, flag_require;
    int32_t min_mapq;
    int32_t has_region;            // bam_region_filter: same reference and [start, end] intersects [lo, hi]
    const int32_t *region_ref;     // per file: refID of the region's reference in that file's header (-2: absent)
    int64_t lo, hi;
    int32_t serial;                // 1: one thread per FILE walks entries[first .. last] as a single chain
    unsigned long long *entry_rows;  // optional out: records each walk saw (column build)
    const uint8_t **rec_ptr;         // optional out (with entry_row0): address of every record, in file order
    const unsigned long long *entry_row0;
};

constexpr int kBamThreads = 128;
constexpr int kSmemGroups = 2048;
constexpr uint64_t kBadExit = ~0ull;

// 32-bit little-endian load at any alignment, from words already in registers: w[i] = aligned word i of the window
__device__ __forceinline__ uint32_t win_u32(const uint32_t *w, int byte_off) {
    const int i = byte_off >> 2, sh = (byte_off & 3) * 8;
    return __funnelshift_r(w[i], w[i + 1], sh);
}

__global__ void __launch_bounds__(kBamThreads) bam_walk_kernel(const __grid_constant__ BamArgs a) {
    __shared__ unsigned int hist[kSmemGroups];
    const bool use_smem = a.n_groups <= kSmemGroups;
    if (use_smem) {
        for (int i = threadIdx.x; i < a.n_groups; i += kBamThreads) hist[i] = 0;
        __syncthreads();
    }
    unsigned long long my_rows = 0;
    const int e = blockIdx.x * kBamThreads + threadIdx.x;
    if (e < a.n_entries) {
        const BamEntry E = a.entries[e];
        const int32_t *remap = a.remap + E.remap0;
        uint64_t p = E.start;
        const uint64_t stop = a.serial ? E.total : E.end;
        uint64_t exit_at = 0;
        bool bad = false;
        while (p < stop) {
            if (p + 36 > E.total) {
                bad = true;
                break;
            }
            // bytes [p, p + 20): block_size, refID, pos, l_read_name / mapq / bin, n_cigar_op / flag
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(E.base + (p & ~(uint64_t)3));
            uint32_t w[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = __ldg(wp + i);
            const int o = (int)(p & 3);
            const int32_t block_size = (int32_t)win_u32(w, o);
            const int32_t ref_id = (int32_t)win_u32(w, o + 4);
            const uint32_t mq = (win_u32(w, o + 12) >> 8) & 0xFFu;   // bytes 12..15 = l_read_name, mapq, bin
            const uint32_t flag = win_u32(w, o + 16) >> 16;           // bytes 16..19 = n_cigar_op, flag
            if (block_size < 32 || p + 4 + (uint64_t)(uint32_t)block_size > E.total || ref_id < -1 || ref_id >= E.n_ref) {
                bad = true;
                break;
            }
            bool sel = true;
            if (a.has_pred) {
                sel = (flag & a.flag_exclude) == 0u && (flag & a.flag_require) == a.flag_require;
                // mapping_quality is a nullable string column: 255 is NULL and NULL fails every comparison
                if (a.min_mapq >= 0) sel = sel && mq != 255u && (int32_t)mq >= a.min_mapq;
            }
            if (sel && a.has_region) {
                // exon-bam/src/indexed_async_batch_stream.rs:66-86: reference, start and end must be present; start = pos + 1,
                // end = start + (reference-consuming CIGAR length) - 1; intersects = lo <= end && start <= hi
                const int32_t pos0 = (int32_t)win_u32(w, o + 8);
                sel = ref_id >= 0 && ref_id == a.region_ref[E.file_idx] && pos0 >= 0 && (int64_t)pos0 + 1 <= a.hi;
                if (sel) {
                    const uint32_t l_read_name = win_u32(w, o + 12) & 0xFFu, n_cigar = win_u32(w, o + 16) & 0xFFFFu;
                    const uint8_t *cg = E.base + p + 36 + l_read_name;
                    if (p + 36 + l_read_name + 4ull * n_cigar > p + 4 + (uint64_t)(uint32_t)block_size) {
                        bad = true;
                        break;
                    }
                    int64_t span = 0;
                    for (uint32_t i = 0; i < n_cigar; ++i) {
                        const uint32_t v = (uint32_t)cg[4 * i] | ((uint32_t)cg[4 * i + 1] << 8) | ((uint32_t)cg[4 * i + 2] << 16) | ((uint32_t)cg[4 * i + 3] << 24);
                        const uint32_t op = v & 15u;
                        if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) span += v >> 4;  // M D N = X
                    }
                    const int64_t start = (int64_t)pos0 + 1, end = start + span - 1;
                    sel = end >= 1 && a.lo <= end;
                }
            }
            if (sel) {
                const int g = remap[ref_id < 0 ? E.n_ref : ref_id];
                if (use_smem) atomicAdd(&hist[g], 1u);
                else atomicAdd(&a.counts[g], 1ull);
            }
            if (a.rec_ptr) a.rec_ptr[a.entry_row0[e] + my_rows] = E.base + p;
            ++my_rows;
            p += 4 + (uint64_t)(uint32_t)block_size;
        }
        exit_at = bad ? kBadExit : p;
        a.exits[e] = exit_at;
        if (a.entry_rows) a.entry_rows[e] = my_rows;
    }
    // rows: warp sum, one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_rows += __shfl_xor_sync(0xFFFFFFFFu, my_rows, d);
    if ((threadIdx.x & 31) == 0 && my_rows) atomicAdd(a.rows, my_rows);
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < a.n_groups; i += kBamThreads)
            if (hist[i]) atomicAdd(&a.counts[i], (unsigned long long)hist[i]);
    }
}

// out[0] = walks that did not end where the next one starts (entry points are corrected in place when `fix` != 0),
// out[1] = malformed walks whose own entry point is known to be right
__global__ void bam_verify_kernel(BamEntry *entries, int n, const uint64_t *exits, int fix, unsigned int *out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint64_t x = exits[e];
    const uint64_t want = entries[e].last_of_file ? entries[e].total : entries[e + 1].start;
    if (x == want) return;
    if (x == kBadExit) {
        atomicAdd(&out[1], 1u);
        return;
    }
    atomicAdd(&out[0], 1u);
    if (fix && !entries[e].last_of_file) entries[e + 1].start = x;  // a record straddles the member boundary
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------

// Called by flush_gz for every inflated file of a BAM stream: header (magic, text, references) from a host copy of the
// file's first bytes; members -> walk entries.
int VcfStream::bam_frame_file(uint8_t *dst, uint64_t total, const uint8_t *probe, size_t probe_len, const BgzfMember *members,
                              size_t n_members) {
    cudaStream_t st = ctx->stream;
    std::vector<uint8_t> big;
    auto need = [&](uint64_t upto) -> int {  // make bytes [0, upto) of the stream available at `probe`
        if (upto <= probe_len) return EXON_GPU_OK;
        if (upto > total) return fail(EXON_GPU_ERR_PARSE, "bam: truncated header");
        const size_t n = (size_t)std::min<uint64_t>(total, std::max<uint64_t>(upto, (uint64_t)probe_len * 4));
        big.resize(n);
        CUDA_TRY(cudaMemcpyAsync(big.data(), dst, n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        probe = big.data();
        probe_len = n;
        return EXON_GPU_OK;
    };
    auto rd32 = [&](uint64_t o) { return (int32_t)((uint32_t)probe[o] | ((uint32_t)probe[o + 1] << 8) | ((uint32_t)probe[o + 2] << 16) | ((uint32_t)probe[o + 3] << 24)); };
    if (int rc = need(12)) return rc;
    if (memcmp(probe, "BAM\1", 4) != 0) return fail(EXON_GPU_ERR_PARSE, "bam: bad magic");
    const int32_t l_text = rd32(4);
    if (l_text < 0) return fail(EXON_GPU_ERR_PARSE, "bam: negative l_text");
    uint64_t p = 8 + (uint64_t)l_text;
    if (int rc = need(p + 4)) return rc;
    const int32_t n_ref = rd32(p);
    p += 4;
    if (n_ref < 0) return fail(EXON_GPU_ERR_PARSE, "bam: negative n_ref");
    BamFile f;
    f.dst = dst;
    f.total = total;
    f.ref_names.reserve((size_t)n_ref);
    for (int32_t i = 0; i < n_ref; ++i) {
        if (int rc = need(p + 4)) return rc;
        const int32_t l_name = rd32(p);
        p += 4;
        if (l_name < 1) return fail(EXON_GPU_ERR_PARSE, "bam: bad reference name length");
        if (int rc = need(p + (uint64_t)l_name + 4)) return rc;
        f.ref_names.emplace_back(reinterpret_cast<const char *>(probe + p), (size_t)l_name - 1);
        p += (uint64_t)l_name + 4;
    }
    f.records_at = p;
    // one walk per member that holds record bytes; the one that holds the first record starts there
    const uint64_t base_addr = (uint64_t)reinterpret_cast<uintptr_t>(dst);
    for (size_t i = 0; i < n_members; ++i) {
        if (!members[i].isize) continue;
        const uint64_t off = members[i].out_addr - base_addr, end = off + members[i].isize;
        if (end <= f.records_at) continue;
        f.walk_starts.push_back(std::max(off, f.records_at));
    }
    bam_files.push_back(std::move(f));
    bam_tables_dirty = true;
    return EXON_GPU_OK;
}

// Device tables of a BAM stream: walk entries | first entry of every file (serial fallback) | refID -> group maps |
// exits | counts | misc.  Built once per set of resident files; the verification kernel corrects entry points in
// place, so later queries start from walks that are already right.
int VcfStream::bam_build_tables() {
    cudaStream_t st = ctx->stream;
    bam_groups.clear();
    std::vector<int32_t> remap;
    std::vector<BamEntry> entries, firsts;
    for (const BamFile &f : bam_files) {
        const int32_t remap0 = (int32_t)remap.size();
        for (const std::string &nm : f.ref_names) {
            auto it = std::find(bam_groups.begin(), bam_groups.end(), nm);
            if (it == bam_groups.end()) {
                bam_groups.push_back(nm);
                remap.push_back((int32_t)bam_groups.size() - 1);
            } else {
                remap.push_back((int32_t)(it - bam_groups.begin()));
            }
        }
        remap.push_back(-1);  // refID -1: patched to the NULL group below
        for (size_t i = 0; i < f.walk_starts.size(); ++i) {
            BamEntry e;
            e.base = f.dst;
            e.start = f.walk_starts[i];
            e.end = i + 1 < f.walk_starts.size() ? f.walk_starts[i + 1] : f.total;
            e.total = f.total;
            e.remap0 = remap0;
            e.n_ref = (int32_t)f.ref_names.size();
            e.last_of_file = i + 1 == f.walk_starts.size();
            e.file_idx = (int32_t)(&f - &bam_files[0]);
            entries.push_back(e);
            if (i == 0) {
                BamEntry s = e;
                s.last_of_file = 1;
                firsts.push_back(s);
            }
        }
    }
    const int32_t n_groups = (int32_t)bam_groups.size() + 1;
    for (int32_t &r : remap)
        if (r < 0) r = n_groups - 1;
    bam_n_entries = entries.size();
    bam_n_firsts = firsts.size();
    bam_n_groups = n_groups;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    bam_o_firsts = al((entries.size() + 1) * sizeof(BamEntry));
    bam_o_remap = bam_o_firsts + al((firsts.size() + 1) * sizeof(BamEntry));
    bam_o_exits = bam_o_remap + al(remap.size() * 4 + 4);
    bam_o_counts = bam_o_exits + al((entries.size() + 1) * 8);
    bam_o_misc = bam_o_counts + al((size_t)n_groups * 8);
    bam_o_region = bam_o_misc + 256;
    const size_t need = bam_o_region + al(bam_files.size() * 4 + 4);
    if (need > d_bam_cap) {
        if (d_bam) {
            CUDA_TRY(cudaStreamSynchronize(st));
            CUDA_TRY(cudaFree(d_bam));
            d_bam = nullptr;
            d_bam_cap = 0;
        }
        CUDA_TRY(cudaMalloc(&d_bam, need * 2));
        d_bam_cap = need * 2;
    }
    uint8_t *d = (uint8_t *)d_bam;
    if (!entries.empty()) {
        CUDA_TRY(cudaMemcpyAsync(d, entries.data(), entries.size() * sizeof(BamEntry), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d + bam_o_firsts, firsts.data(), firsts.size() * sizeof(BamEntry), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d + bam_o_remap, remap.data(), remap.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));  // the sources are locals
    }
    bam_tables_dirty = false;
    return EXON_GPU_OK;
}

int VcfStream::bam_filter_count(const exon_gpu_bam_pred *pred, int64_t *counts, int32_t cap, int32_t *n_groups_out, int64_t *total_rows) {
    if (int rc = flush_gz()) return rc;
    Ctx *c = ctx;
    cudaStream_t st = c->stream;
    if (bam_tables_dirty)
        if (int rc = bam_build_tables()) return rc;
    const int32_t n_groups = bam_n_groups;
    if (n_groups_out) *n_groups_out = n_groups;
    if (total_rows) *total_rows = 0;
    if (counts)
        for (int32_t g = 0; g < std::min(cap, n_groups); ++g) counts[g] = 0;
    if (counts && cap < n_groups) return fail(EXON_GPU_ERR_ARG, "bam_filter_count: %d groups, room for %d", n_groups, cap);
    if (bam_n_entries == 0) return EXON_GPU_OK;
    std::lock_guard<std::recursive_mutex> work(c->work_mu);
    if (int rc = c->ensure_scratch(0, 256 + (size_t)n_groups * 8)) return rc;
    uint8_t *d = (uint8_t *)d_bam;
    BamArgs a;
    memset(&a, 0, sizeof(a));
    a.remap = (const int32_t *)(d + bam_o_remap);
    a.exits = (uint64_t *)(d + bam_o_exits);
    a.counts = (unsigned long long *)(d + bam_o_counts);
    a.rows = (unsigned long long *)(d + bam_o_misc);
    a.n_groups = n_groups;
    a.has_pred = pred != nullptr;
    a.flag_exclude = pred ? pred->flag_exclude : 0;
    a.flag_require = pred ? pred->flag_require : 0;
    a.min_mapq = pred ? pred->min_mapq : -1;
    if (pred && pred->has_region) {
        if (!pred->region_ref || pred->region_ref_len < 0) return fail(EXON_GPU_ERR_ARG, "bam_filter_count: region without a reference name");
        const std::string want(pred->region_ref, (size_t)pred->region_ref_len);
        std::vector<int32_t> ids;
        for (const BamFile &f : bam_files) {
            auto it = std::find(f.ref_names.begin(), f.ref_names.end(), want);
            ids.push_back(it == f.ref_names.end() ? -2 : (int32_t)(it - f.ref_names.begin()));
        }
        CUDA_TRY(cudaMemcpyAsync(d + bam_o_region, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        a.has_region = 1;
        a.region_ref = (const int32_t *)(d + bam_o_region);
        a.lo = pred->region_lo < 1 ? 1 : pred->region_lo;
        a.hi = pred->region_hi;
    }
    unsigned int *d_verify = (unsigned int *)(d + bam_o_misc + 64);
    uint8_t *h = (uint8_t *)c->h_scratch;
    bool ok = false;
    constexpr int kRounds = 6;
    for (int round = 0; round <= kRounds && !ok; ++round) {
        const bool serial = round == kRounds;  // last resort: one thread per file, a single chain from the first record
        BamEntry *ent = (BamEntry *)(serial ? d + bam_o_firsts : d);
        const int launch_n = (int)(serial ? bam_n_firsts : bam_n_entries);
        a.entries = ent;
        a.n_entries = launch_n;
        a.serial = serial;
        CUDA_TRY(cudaMemsetAsync(d + bam_o_counts, 0, (size_t)n_groups * 8, st));
        CUDA_TRY(cudaMemsetAsync(d + bam_o_misc, 0, 128, st));
        if (round == 0) CUDA_TRY(c->timed_begin(st));
        bam_walk_kernel<<<(launch_n + kBamThreads - 1) / kBamThreads, kBamThreads, 0, st>>>(a);
        bam_verify_kernel<<<(launch_n + 127) / 128, 128, 0, st>>>(ent, launch_n, a.exits, serial ? 0 : 1, d_verify);
        if (round == 0) {
            CUDA_TRY(c->timed_end(st));
        }
        c->launches.fetch_add(2);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(h, d + bam_o_misc, 128, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h + 128, d + bam_o_counts, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const unsigned int *v = reinterpret_cast<const unsigned int *>(h + 64);
        if (v[0] == 0 && v[1] == 0) {
            ok = true;
            bam_serial_ok = serial;
        } else if (v[0] == 0 || serial)
            return fail(EXON_GPU_ERR_PARSE, "malformed BAM record (block_size / refID out of range, or a truncated record)");
        // else: some walks began inside a record; the entry points were corrected, go again
    }
    if (total_rows) *total_rows = (int64_t) * reinterpret_cast<const unsigned long long *>(h);
    if (counts) memcpy(counts, h + 128, (size_t)n_groups * 8);
    return EXON_GPU_OK;
}


// =====================================================================================================================
// BAM records -> Arrow columns 0..9 {name, flag, reference, start, end, mapping_quality, cigar, mate_reference, sequence,
// quality_score} in reference-sized batches (exon_gpu_bam_next_batch).
//
// Replaces BatchReader::read_batch (exon/exon-bam/src/batch_reader.rs:88-107), BAMArrayBuilder::{append, finish}
// (exon/exon-bam/src/array_builder.rs:102-218) over noodles' RecordBuf, and SemiLazyRecord::alignment_end
// (exon/exon-bam/src/indexed_async_batch_stream.rs:43-50).  Schema: SAMSchemaBuilder::default,
// exon/exon-sam/src/schema_builder.rs:385-401.  `tags` (column 10) is not built.
//   1. the verified speculative walks of the fused query give every member's first record; a second walk per member
//      writes the address of every record in file order (rows per walk -> host prefix -> rec_ptr[])
//   2. measure   one thread per record: byte lengths of the six string columns and the quality list, the fixed-width
//                columns (flag, start, end) and the validity flags
//   3. 7 exclusive scans (cub)
//   4. emit      one thread per record: batch-relative int32 offsets, bytes at their final place (CIGAR rendered as
//                decimal length + op letter, bases decoded from 4 bits, MAPQ as a decimal string, quality bytes widened
//                i8 -> i64), validity bits
// Batches restart at every file.  name == "*" would be NULL in a column the reference declares non-nullable (its batch
// construction fails): reported as EXON_GPU_ERR_PARSE here too.
// =====================================================================================================================
namespace {

enum { kBName = 0, kBRef = 1, kBMapq = 2, kBCigar = 3, kBMate = 4, kBSeq = 5, kBQual = 6, kBPlain = 7, kBTagN = 7, kBTagB = 8, kBNVar = 9 };
constexpr uint32_t kBErrLayout = 1u;  // the variable part does not fit the record's block_size / bad CIGAR op / bad refID
constexpr uint32_t kBErrName = 2u;    // missing read name ("*")
constexpr uint32_t kBErrTag = 4u;     // an auxiliary field with an unknown type, or one that runs past the record
constexpr uint32_t kBErrTagFloat = 8u;  // a B:f element of 9e13 or more in magnitude: "{:.2}" of it is not printed here

struct BamFileTab {
    long long row0;    // first record of the file (global numbering); the sentinel carries n_records
    long long batch0;  // first batch of the file
    int32_t ref0;      // first entry of the file's reference names in ref_off / ref_len
    int32_t n_ref;
};

struct BamColArgs {
    int64_t n_rows;
    const uint8_t *const *rec_ptr;
    const BamFileTab *files;
    int32_t n_files;
    const int32_t *ref_off, *ref_len;  // per reference name: offset into ref_blob, length
    const uint8_t *ref_blob;
    const long long *brow;  // n_batches + 1
    int64_t n_batches;
    int32_t batch_rows, wpb;
    int32_t want[11];
    int32_t *cnt[kBNVar];
    const long long *pre[kBNVar];
    // tags (column 10): list offsets in the batch layout; per entry (index entry + batch): tag / value string offsets
    int32_t *tags_loff, *tag_off, *tval_off;
    uint8_t *tag_val, *tval_val;
    uint8_t *rowflags;  // bit0 reference valid, bit1 start valid, bit2 end valid, bit3 mapq valid, bit4 mate valid
    int32_t *flag;
    long long *start, *end;
    int32_t *off[kBNVar];  // batch-relative offsets, n_batches * (batch_rows + 1) each
    uint8_t *val[kBNVar];  // kBQual: int64 values
    uint32_t *valid[5];    // reference, start, end, mapq, mate
    uint32_t *flags;
    unsigned long long *first_bad_row;
};

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ int dec_digits(uint32_t v) {
    int n = 1;
    while (v >= 10u) {
        v /= 10u;
        ++n;
    }
    return n;
}
__device__ __forceinline__ int bam_find_file(const BamFileTab *files, int n_files, long long r) {
    int lo = 0, hi = n_files - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (files[mid].row0 <= r) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

struct BamRec {
    const uint8_t *r;  // refID field (the byte after block_size)
    int32_t block_size, ref_id, pos, mate_ref;
    uint32_t l_read_name, mapq, n_cigar, flag;
    int32_t l_seq;
    const uint8_t *name, *cigar, *seq, *qual;
    bool ok;
};
__device__ __forceinline__ BamRec bam_rec(const uint8_t *p, int32_t n_ref) {
    BamRec R;
    R.block_size = (int32_t)ld_u32(p);
    R.r = p + 4;
    R.ref_id = (int32_t)ld_u32(R.r);
    R.pos = (int32_t)ld_u32(R.r + 4);
    R.l_read_name = R.r[8];
    R.mapq = R.r[9];
    R.n_cigar = (uint32_t)R.r[12] | ((uint32_t)R.r[13] << 8);
    R.flag = (uint32_t)R.r[14] | ((uint32_t)R.r[15] << 8);
    R.l_seq = (int32_t)ld_u32(R.r + 16);
    R.mate_ref = (int32_t)ld_u32(R.r + 20);
    R.name = R.r + 32;
    R.cigar = R.name + R.l_read_name;
    R.seq = R.cigar + 4ull * R.n_cigar;
    R.qual = R.seq + (R.l_seq >= 0 ? (R.l_seq + 1) / 2 : 0);
    R.ok = R.l_seq >= 0 && R.l_read_name >= 1 && 32ull + R.l_read_name + 4ull * R.n_cigar + (uint64_t)((R.l_seq + 1) / 2) + (uint64_t)R.l_seq <= (uint64_t)R.block_size &&
           R.ref_id >= -1 && R.ref_id < n_ref && R.mate_ref >= -1 && R.mate_ref < n_ref;
    return R;
}

// ---- tags (column 10) ------------------------------------------------------------------------------------------------
// TagsMapBuilder::append (exon/exon-sam/src/tag_builder.rs:497-741, the default `tags` type List<Struct{tag, value: Utf8}>):
// every auxiliary field in record order, its value as text -- integers through i64 Display (:527-538), A as the character,
// Z / H as their text, f through f32 Display (:566-576), B integer arrays joined by "," (:590-680) and B:f arrays as
// "{:.2}" joined by ", " (:682-698).
struct TagSink {
    uint8_t *dst;
    int32_t n;
    __device__ __forceinline__ void put(uint8_t c) {
        if (dst) dst[n] = c;
        ++n;
    }
    __device__ void put_i64(long long v) {
        uint8_t d[20];
        int nd = 0;
        unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
        do {
            d[nd++] = (uint8_t)('0' + u % 10ull);
            u /= 10ull;
        } while (u);
        if (v < 0) put('-');
        while (nd) put(d[--nd]);
    }
};
// Rust `format!("{:.2}", v)` for f32: the exact value rounded half-to-even at two decimals.  |v| * 100 is exact in a double
// (24 + 7 bits), rint() rounds to nearest-even; false when the magnitude is beyond what is printed here.
__device__ bool put_f32_2(float v, TagSink &o) {
    if (v != v) {
        o.put('N'), o.put('a'), o.put('N');
        return true;
    }
    const bool neg = (__float_as_uint(v) >> 31) != 0u;
    if (neg) o.put('-');
    const float av = fabsf(v);
    if (av > 3.0e38f) {
        o.put('i'), o.put('n'), o.put('f');
        return true;
    }
    const double x = (double)av * 100.0;
    if (x >= 9.0e15) return false;
    const unsigned long long r = (unsigned long long)rint(x);
    o.put_i64((long long)(r / 100ull));
    o.put('.');
    o.put((uint8_t)('0' + (r / 10ull) % 10ull));
    o.put((uint8_t)('0' + r % 10ull));
    return true;
}
__device__ __forceinline__ int bam_aux_width(uint8_t t) {
    return (t == 'c' || t == 'C' || t == 'A') ? 1 : (t == 's' || t == 'S') ? 2 : (t == 'i' || t == 'I' || t == 'f') ? 4 : 0;
}
__device__ __forceinline__ long long bam_aux_int(const uint8_t *p, uint8_t t) {
    switch (t) {
        case 'c': return (long long)(int8_t)p[0];
        case 'C': return (long long)p[0];
        case 's': return (long long)(int16_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8));
        case 'S': return (long long)((uint32_t)p[0] | ((uint32_t)p[1] << 8));
        case 'i': return (long long)(int32_t)ld_u32(p);
        default: return (long long)ld_u32(p);
    }
}
struct TagOut {
    int32_t *tag_off, *tval_off;  // entries of the record's first field
    uint8_t *tag_val, *tval_val;  // batch bases
    int32_t ent0, byte0;          // entries / value bytes of the batch before this record
};
// Walks the auxiliary fields [p, end): *n = fields, *bytes = value text bytes.  Returns error bits.
__device__ uint32_t bam_tags_walk(const uint8_t *p, const uint8_t *end, int32_t *n, int32_t *bytes, const TagOut *o) {
    int32_t ne = 0, nb = 0;
    uint32_t err = 0;
    while (p < end) {
        if (end - p < 4) {
            err |= kBErrTag;
            break;
        }
        const uint8_t t = p[2];
        if (o) {
            o->tag_off[ne] = 2 * (o->ent0 + ne);
            o->tag_val[2 * (o->ent0 + ne)] = p[0];
            o->tag_val[2 * (o->ent0 + ne) + 1] = p[1];
            o->tval_off[ne] = o->byte0 + nb;
        }
        TagSink sk{o ? o->tval_val + o->byte0 + nb : nullptr, 0};
        p += 3;
        if (t == 'A') {  // `*c as char` then to_string: bytes of 0x80 and above become two UTF-8 bytes
            if (p[0] < 0x80) {
                sk.put(p[0]);
            } else {
                sk.put((uint8_t)(0xC0 | (p[0] >> 6)));
                sk.put((uint8_t)(0x80 | (p[0] & 0x3F)));
            }
            p += 1;
        } else if (t == 'c' || t == 'C' || t == 's' || t == 'S' || t == 'i' || t == 'I') {
            const int w = bam_aux_width(t);
            if (end - p < w) {
                err |= kBErrTag;
                break;
            }
            sk.put_i64(bam_aux_int(p, t));
            p += w;
        } else if (t == 'f') {
            if (end - p < 4) {
                err |= kBErrTag;
                break;
            }
            uint8_t buf[kF32DisplayMax];
            const int k = f32_display(__uint_as_float(ld_u32(p)), buf);
            for (int q = 0; q < k; ++q) sk.put(buf[q]);
            p += 4;
        } else if (t == 'Z' || t == 'H') {
            while (p < end && *p) sk.put(*p++);
            if (p >= end) {
                err |= kBErrTag;
                break;
            }
            ++p;
        } else if (t == 'B') {
            if (end - p < 5) {
                err |= kBErrTag;
                break;
            }
            const uint8_t st = p[0];
            const uint32_t cnt = ld_u32(p + 1);
            const int w = st == 'A' ? 0 : bam_aux_width(st);
            p += 5;
            if (!w || (uint64_t)cnt * (uint64_t)w > (uint64_t)(end - p)) {
                err |= kBErrTag;
                break;
            }
            for (uint32_t i = 0; i < cnt; ++i) {
                if (st == 'f') {
                    if (i) sk.put(','), sk.put(' ');
                    if (!put_f32_2(__uint_as_float(ld_u32(p)), sk)) err |= kBErrTagFloat;
                } else {
                    if (i) sk.put(',');
                    sk.put_i64(bam_aux_int(p, st));
                }
                p += w;
            }
        } else {
            err |= kBErrTag;
            break;
        }
        nb += sk.n;
        ++ne;
    }
    *n = ne;
    *bytes = nb;
    return err;
}

__global__ void __launch_bounds__(256) bam_col_measure_kernel(const __grid_constant__ BamColArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= a.n_rows) return;
    const int f = bam_find_file(a.files, a.n_files, r);
    const BamFileTab F = a.files[f];
    const BamRec R = bam_rec(a.rec_ptr[r], F.n_ref);
    int32_t c[kBNVar] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t rf = 0;
    uint32_t err = 0;
    long long start = 0, end = 0;
    if (!R.ok) {
        err = kBErrLayout;
    } else {
        if (a.want[10]) err |= bam_tags_walk(R.qual + R.l_seq, R.r + R.block_size, &c[kBTagN], &c[kBTagB], nullptr);
        c[kBName] = (int32_t)R.l_read_name - 1;
        if (R.l_read_name == 2 && R.name[0] == '*') err |= kBErrName;
        if (R.ref_id >= 0) {
            rf |= 1u;
            c[kBRef] = a.ref_len[F.ref0 + R.ref_id];
        }
        if (R.mate_ref >= 0) {
            rf |= 16u;
            c[kBMate] = a.ref_len[F.ref0 + R.mate_ref];
        }
        if (R.mapq != 255u) {
            rf |= 8u;
            c[kBMapq] = dec_digits(R.mapq);
        }
        long long span = 0;
        int32_t cg = 0;
        for (uint32_t i = 0; i < R.n_cigar; ++i) {
            const uint32_t v = ld_u32(R.cigar + 4 * i), op = v & 15u, ln = v >> 4;
            if (op > 8u) err |= kBErrLayout;
            if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) span += ln;  // M D N = X consume the reference
            cg += dec_digits(ln) + 1;
        }
        c[kBCigar] = cg;
        c[kBSeq] = R.l_seq;
        bool missing = true;
        for (int32_t i = 0; i < R.l_seq; ++i) missing = missing && R.qual[i] == 0xFFu;
        c[kBQual] = missing ? 0 : R.l_seq;
        if (R.pos >= 0) {  // alignment_start = pos + 1; alignment_end = start + reference span - 1 (None when that is 0)
            rf |= 2u;
            start = (long long)R.pos + 1;
            end = start + span - 1;
            if (end >= 1) rf |= 4u;
            else end = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < kBNVar; ++k)
        if (a.cnt[k]) a.cnt[k][r] = c[k];
    a.rowflags[r] = rf;
    if (a.flag) a.flag[r] = (int32_t)R.flag;
    if (a.start) a.start[r] = start;
    if (a.end) a.end[r] = end;
    if (err) {
        atomicOr(a.flags, err);
        atomicMin(a.first_bad_row, (unsigned long long)r);
    }
}

__device__ __forceinline__ int put_dec(uint8_t *dst, uint32_t v) {
    const int n = dec_digits(v);
    for (int i = n - 1; i >= 0; --i) {
        dst[i] = (uint8_t)('0' + v % 10u);
        v /= 10u;
    }
    return n;
}

__global__ void __launch_bounds__(256) bam_col_emit_kernel(const __grid_constant__ BamColArgs a) {
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    __shared__ long long s_b0;
    if (threadIdx.x == 0) {
        const int64_t rb = (int64_t)blockIdx.x * 256;
        int64_t lo = 0, hi = a.n_batches;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(&a.brow[mid]) <= rb) lo = mid;
            else hi = mid;
        }
        s_b0 = lo;
    }
    __syncthreads();
    if (r >= a.n_rows) return;
    int64_t b = s_b0;
    while (__ldg(&a.brow[b + 1]) <= r) ++b;
    const int64_t r0 = __ldg(&a.brow[b]);
    const int in_batch = (int)(r - r0);
    const bool last = r + 1 == __ldg(&a.brow[b + 1]);
    const uint8_t rf = a.rowflags[r];
    {
        const uint32_t bit = 1u << (in_batch & 31);
        const int64_t word = b * a.wpb + (in_batch >> 5);
        const uint32_t peers = __match_any_sync(__activemask(), word);
        const bool leader = (threadIdx.x & 31) == __ffs((int)peers) - 1;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (!a.valid[k]) continue;
            const uint32_t v = __reduce_or_sync(peers, (rf >> k) & 1u ? bit : 0u);
            if (leader && v) atomicOr(a.valid[k] + word, v);
        }
    }
    const int f = bam_find_file(a.files, a.n_files, r);
    const BamFileTab F = a.files[f];
    const BamRec R = bam_rec(a.rec_ptr[r], F.n_ref);
    if (!R.ok) return;  // reported by the measure pass
    const int64_t lrow = b * (int64_t)(a.batch_rows + 1) + in_batch;
    auto open_cell = [&](int k, long long &v_abs) {  // offsets of the cell; returns where its bytes go
        const long long v = a.pre[k][r], v0 = a.pre[k][r0];
        a.off[k][lrow] = (int32_t)(v - v0);
        if (last) a.off[k][lrow + 1] = (int32_t)(a.pre[k][r + 1] - v0);
        v_abs = v;
    };
    long long v;
    if (a.off[kBName]) {
        open_cell(kBName, v);
        for (uint32_t i = 0; i + 1 < R.l_read_name; ++i) a.val[kBName][v + i] = R.name[i];
    }
    if (a.off[kBRef]) {
        open_cell(kBRef, v);
        if (R.ref_id >= 0) {
            const uint8_t *nm = a.ref_blob + a.ref_off[F.ref0 + R.ref_id];
            const int32_t n = a.ref_len[F.ref0 + R.ref_id];
            for (int32_t i = 0; i < n; ++i) a.val[kBRef][v + i] = nm[i];
        }
    }
    if (a.off[kBMate]) {
        open_cell(kBMate, v);
        if (R.mate_ref >= 0) {
            const uint8_t *nm = a.ref_blob + a.ref_off[F.ref0 + R.mate_ref];
            const int32_t n = a.ref_len[F.ref0 + R.mate_ref];
            for (int32_t i = 0; i < n; ++i) a.val[kBMate][v + i] = nm[i];
        }
    }
    if (a.off[kBMapq]) {
        open_cell(kBMapq, v);
        if (R.mapq != 255u) put_dec(a.val[kBMapq] + v, R.mapq);
    }
    if (a.off[kBCigar]) {
        open_cell(kBCigar, v);
        uint8_t *dst = a.val[kBCigar] + v;
        for (uint32_t i = 0; i < R.n_cigar; ++i) {
            const uint32_t x = ld_u32(R.cigar + 4 * i), op = x & 15u;
            dst += put_dec(dst, x >> 4);
            *dst++ = (uint8_t)("MIDNSHP=X"[op > 8u ? 0u : op]);
        }
    }
    // the long cells (bases, qualities) only get their offsets here: their bytes are written by bam_col_emit_long_kernel, one
    // warp per record, so that the stores of a record coalesce
    if (a.off[kBSeq]) open_cell(kBSeq, v);
    if (a.off[kBQual]) open_cell(kBQual, v);
    if (a.tags_loff) {
        const long long e = a.pre[kBTagN][r], e0 = a.pre[kBTagN][r0], y = a.pre[kBTagB][r], y0 = a.pre[kBTagB][r0];
        a.tags_loff[lrow] = (int32_t)(e - e0);
        TagOut o;
        o.tag_off = a.tag_off + e + b;
        o.tval_off = a.tval_off + e + b;
        o.tag_val = a.tag_val + 2 * e0;
        o.tval_val = a.tval_val + y0;
        o.ent0 = (int32_t)(e - e0);
        o.byte0 = (int32_t)(y - y0);
        if (last) {
            const long long e1 = a.pre[kBTagN][r + 1];
            a.tags_loff[lrow + 1] = (int32_t)(e1 - e0);
            a.tag_off[e1 + b] = (int32_t)(2 * (e1 - e0));
            a.tval_off[e1 + b] = (int32_t)(a.pre[kBTagB][r + 1] - y0);
        }
        int32_t n, nb;
        bam_tags_walk(R.qual + R.l_seq, R.r + R.block_size, &n, &nb, &o);
    }
}

// sequence (4-bit -> letters) and quality_score (i8 -> i64) of one record per warp
__global__ void __launch_bounds__(256) bam_col_emit_long_kernel(const __grid_constant__ BamColArgs a) {
    const int64_t r = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n_rows) return;
    const int f = bam_find_file(a.files, a.n_files, r);
    const BamRec R = bam_rec(a.rec_ptr[r], a.files[f].n_ref);
    if (!R.ok) return;
    if (a.off[kBSeq]) {
        uint8_t *dst = a.val[kBSeq] + a.pre[kBSeq][r];
        for (int32_t i = lane; i < R.l_seq; i += 32) dst[i] = (uint8_t)("=ACMGRSVTWYHKDBN"[(R.seq[i >> 1] >> ((i & 1) ? 0 : 4)) & 15]);
    }
    if (a.off[kBQual]) {
        const int32_t n = (int32_t)(a.pre[kBQual][r + 1] - a.pre[kBQual][r]);
        long long *dst = reinterpret_cast<long long *>(a.val[kBQual]) + a.pre[kBQual][r];
        for (int32_t i = lane; i < n; i += 32) dst[i] = (long long)(int8_t)R.qual[i];
    }
}

__global__ void bam_gather_i64(const long long *src, const long long *idx, int64_t n, long long *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

size_t bal256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

struct BamBuf {
    void *d = nullptr, *h = nullptr;
    size_t bytes = 0;
};

// Column store of one BAM stream; batches are views into it.
struct BamColumns {
    std::atomic<int> refs{1};
    int device = 0;
    bool on_device = false;
    int batch_rows = 8192, wpb = 256;
    int64_t n_rows = 0, n_batches = 0, next = 0;
    std::vector<int> projection;
    BamBuf off[kBNVar], val[kBNVar], valid[5], flag, start, end, tags_loff, tag_off, tval_off, tag_val, tval_val;
    std::vector<long long> batch_row0, base[kBNVar];
    template <class T>
    const T *p(const BamBuf &b) const { return static_cast<const T *>(on_device ? b.d : b.h); }
    void each(void (*fn)(BamBuf &)) {
        for (int k = 0; k < kBNVar; ++k) fn(off[k]), fn(val[k]);
        for (int k = 0; k < 5; ++k) fn(valid[k]);
        fn(flag), fn(start), fn(end);
        fn(tags_loff), fn(tag_off), fn(tval_off), fn(tag_val), fn(tval_val);
    }
    void unref() {
        if (refs.fetch_sub(1) == 1) {
            cudaSetDevice(device);
            each([](BamBuf &b) {
                cudaFree(b.d);
                cudaFreeHost(b.h);
            });
            delete this;
        }
    }
};

void bam_columns_free(VcfStream *s) {
    if (s->bam_cols) {
        s->bam_cols->unref();
        s->bam_cols = nullptr;
    }
}

namespace {

// column -> variable-length slot (-1: fixed width)
constexpr int kColVar[10] = {kBName, -1, kBRef, -1, -1, kBMapq, kBCigar, kBMate, kBSeq, kBQual};
// column -> validity slot (-1: never NULL)
constexpr int kColValid[10] = {-1, -1, 0, 1, 2, 3, -1, 4, -1, -1};

int bam_build_columns(VcfStream *s) {
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto *c = new (std::nothrow) BamColumns();
    if (!c) return fail(EXON_GPU_ERR_OOM, "bam_next_batch: out of host memory");
    s->bam_cols = c;
    c->device = ctx->device;
    c->on_device = s->columns_on_device;
    c->batch_rows = s->batch_rows;
    c->wpb = ((s->batch_rows + 63) / 64) * 2;
    c->projection = s->projection;
    c->batch_row0.assign(1, 0);
    bool want[11] = {false, false, false, false, false, false, false, false, false, false, false};
    for (int p : s->projection) want[p] = true;
    if (s->bam_n_entries == 0) return EXON_GPU_OK;
    uint8_t *d = (uint8_t *)s->d_bam;
    const bool serial = s->bam_serial_ok;
    const size_t n_ent = serial ? s->bam_n_firsts : s->bam_n_entries;
    BamEntry *ent = (BamEntry *)(serial ? d + s->bam_o_firsts : d);

    // ---- 1. rows per walk -> record addresses ----
    unsigned long long *d_erows = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&d_erows, 2 * bal256((n_ent + 1) * 8), st));
    struct PoolFree {
        void *p;
        cudaStream_t st;
        ~PoolFree() { cudaFreeAsync(p, st); }
    } g0{d_erows, st};
    unsigned long long *d_erow0 = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(d_erows) + bal256((n_ent + 1) * 8));
    BamArgs a;
    memset(&a, 0, sizeof(a));
    a.entries = ent;
    a.n_entries = (int32_t)n_ent;
    a.remap = (const int32_t *)(d + s->bam_o_remap);
    a.exits = (uint64_t *)(d + s->bam_o_exits);
    a.counts = (unsigned long long *)(d + s->bam_o_counts);
    a.rows = (unsigned long long *)(d + s->bam_o_misc);
    a.n_groups = s->bam_n_groups;
    a.serial = serial;
    a.entry_rows = d_erows;
    CUDA_TRY(cudaMemsetAsync(d + s->bam_o_misc, 0, 128, st));
    bam_walk_kernel<<<(unsigned)((n_ent + kBamThreads - 1) / kBamThreads), kBamThreads, 0, st>>>(a);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    std::vector<unsigned long long> erows(n_ent), erow0(n_ent + 1);
    CUDA_TRY(cudaMemcpyAsync(erows.data(), d_erows, n_ent * 8, cudaMemcpyDeviceToHost, st));
    std::vector<BamEntry> h_ent(n_ent);
    CUDA_TRY(cudaMemcpyAsync(h_ent.data(), ent, n_ent * sizeof(BamEntry), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    erow0[0] = 0;
    for (size_t i = 0; i < n_ent; ++i) erow0[i + 1] = erow0[i] + erows[i];
    const int64_t n_rows = (int64_t)erow0[n_ent];
    c->n_rows = n_rows;
    if (n_rows == 0) return EXON_GPU_OK;
    // files -> batches (batches restart at every file), reference-name tables
    const int n_files = (int)s->bam_files.size();
    std::vector<BamFileTab> ftab((size_t)n_files + 1);
    std::vector<long long> file_rows((size_t)n_files, 0);
    for (size_t i = 0; i < n_ent; ++i) file_rows[(size_t)h_ent[i].file_idx] += (long long)erows[i];
    std::vector<int32_t> ref_off, ref_len;
    std::string blob;
    c->batch_row0.clear();
    long long row0 = 0;
    for (int f = 0; f < n_files; ++f) {
        ftab[(size_t)f] = BamFileTab{row0, (long long)c->batch_row0.size(), (int32_t)ref_off.size(), (int32_t)s->bam_files[(size_t)f].ref_names.size()};
        for (const std::string &nm : s->bam_files[(size_t)f].ref_names) {
            ref_off.push_back((int32_t)blob.size());
            ref_len.push_back((int32_t)nm.size());
            blob += nm;
        }
        for (long long r = 0; r < file_rows[(size_t)f]; r += c->batch_rows) c->batch_row0.push_back(row0 + r);
        row0 += file_rows[(size_t)f];
    }
    ftab[(size_t)n_files] = BamFileTab{row0, (long long)c->batch_row0.size(), (int32_t)ref_off.size(), 0};
    c->n_batches = (int64_t)c->batch_row0.size();
    c->batch_row0.push_back(n_rows);
    const size_t nb1 = (size_t)c->n_batches + 1, nr1 = (size_t)n_rows + 1;

    // scratch_b: rec_ptr | 7 counts | 7 prefixes | rowflags | cub | tables
    size_t cub_bytes = 0;
    CUDA_TRY(exclusive_sum_i32_i64(nullptr, cub_bytes, (const int32_t *)nullptr, (long long *)nullptr, (int)nr1, st));
    const size_t tab_bytes = bal256(ftab.size() * sizeof(BamFileTab)) + 2 * bal256(ref_off.size() * 4 + 4) + bal256(blob.size() + 1) + (kBNVar + 1) * bal256(nb1 * 8) + 256;
    if (int rc = ctx->ensure_scratch_b(bal256(nr1 * 8) + kBNVar * (bal256(nr1 * 4) + bal256(nr1 * 8)) + bal256(nr1) + bal256(cub_bytes) + tab_bytes + 4096)) return rc;
    uint8_t *x = (uint8_t *)ctx->scratch_b;
    auto take = [&](size_t bytes) {
        uint8_t *p = x;
        x += bal256(bytes);
        return p;
    };
    const uint8_t **d_rec = (const uint8_t **)take(nr1 * 8);
    CUDA_TRY(cudaMemcpyAsync(d_erow0, erow0.data(), (n_ent + 1) * 8, cudaMemcpyHostToDevice, st));
    a.entry_rows = nullptr;
    a.rec_ptr = d_rec;
    a.entry_row0 = d_erow0;
    bam_walk_kernel<<<(unsigned)((n_ent + kBamThreads - 1) / kBamThreads), kBamThreads, 0, st>>>(a);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());

    BamColArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.n_rows = n_rows;
    ca.rec_ptr = d_rec;
    ca.n_files = n_files;
    ca.n_batches = c->n_batches;
    ca.batch_rows = c->batch_rows;
    ca.wpb = c->wpb;
    long long *pre[kBNVar];
    bool need[kBNVar];
    for (int k = 0; k < kBNVar; ++k) need[k] = false;
    for (int col = 0; col < 10; ++col) {
        ca.want[col] = want[col];
        if (want[col] && kColVar[col] >= 0) need[kColVar[col]] = true;
    }
    ca.want[10] = want[10];
    need[kBTagN] = need[kBTagB] = want[10];
    for (int k = 0; k < kBNVar; ++k) {
        pre[k] = nullptr;
        if (!need[k]) continue;
        ca.cnt[k] = (int32_t *)take(nr1 * 4);
        pre[k] = (long long *)take(nr1 * 8);
        ca.pre[k] = pre[k];
        CUDA_TRY(cudaMemsetAsync(ca.cnt[k] + n_rows, 0, 4, st));
    }
    ca.rowflags = take(nr1);
    uint8_t *cub_tmp = take(cub_bytes);
    BamFileTab *d_ftab = (BamFileTab *)take(ftab.size() * sizeof(BamFileTab));
    int32_t *d_roff = (int32_t *)take(ref_off.size() * 4 + 4), *d_rlen = (int32_t *)take(ref_len.size() * 4 + 4);
    uint8_t *d_blob = take(blob.size() + 1);
    long long *d_brow = (long long *)take(nb1 * 8);
    long long *d_base[kBNVar];
    for (int k = 0; k < kBNVar; ++k) d_base[k] = (long long *)take(nb1 * 8);
    unsigned long long *d_misc = (unsigned long long *)take(64);
    CUDA_TRY(cudaMemcpyAsync(d_ftab, ftab.data(), ftab.size() * sizeof(BamFileTab), cudaMemcpyHostToDevice, st));
    if (!ref_off.empty()) {
        CUDA_TRY(cudaMemcpyAsync(d_roff, ref_off.data(), ref_off.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_rlen, ref_len.data(), ref_len.size() * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(d_brow, c->batch_row0.data(), nb1 * 8, cudaMemcpyHostToDevice, st));
    const unsigned long long init_misc[2] = {0ull, ~0ull};
    CUDA_TRY(cudaMemcpyAsync(d_misc, init_misc, sizeof(init_misc), cudaMemcpyHostToDevice, st));
    ca.files = d_ftab;
    ca.ref_off = d_roff;
    ca.ref_len = d_rlen;
    ca.ref_blob = d_blob;
    ca.brow = d_brow;
    ca.flags = reinterpret_cast<uint32_t *>(d_misc);
    ca.first_bad_row = d_misc + 1;

    auto dev_alloc = [&](BamBuf &b, size_t bytes, bool zero) -> int {
        b.bytes = std::max<size_t>(bytes, 8);
        CUDA_TRY(cudaMallocAsync(&b.d, b.bytes, st));
        if (zero) CUDA_TRY(cudaMemsetAsync(b.d, 0, b.bytes, st));
        return EXON_GPU_OK;
    };
    if (want[1]) {
        if (int rc = dev_alloc(c->flag, (size_t)n_rows * 4, false)) return rc;
        ca.flag = (int32_t *)c->flag.d;
    }
    if (want[3]) {
        if (int rc = dev_alloc(c->start, (size_t)n_rows * 8, false)) return rc;
        ca.start = (long long *)c->start.d;
    }
    if (want[4]) {
        if (int rc = dev_alloc(c->end, (size_t)n_rows * 8, false)) return rc;
        ca.end = (long long *)c->end.d;
    }
    // ---- 2. measure ----
    const unsigned grid = (unsigned)((n_rows + 255) / 256);
    bam_col_measure_kernel<<<grid, 256, 0, st>>>(ca);
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    // ---- 3. scans ----
    for (int k = 0; k < kBNVar; ++k) {
        if (!need[k]) continue;
        size_t tb = cub_bytes;
        CUDA_TRY(exclusive_sum_i32_i64(cub_tmp, tb, (const int32_t *)ca.cnt[k], pre[k], (int)nr1, st));
        bam_gather_i64<<<(unsigned)((nb1 + 255) / 256), 256, 0, st>>>(pre[k], d_brow, (int64_t)nb1, d_base[k]);
        ctx->launches.fetch_add(2);
        c->base[k].resize(nb1);
        CUDA_TRY(cudaMemcpyAsync(c->base[k].data(), d_base[k], nb1 * 8, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long h_misc[2];
    CUDA_TRY(cudaMemcpyAsync(h_misc, d_misc, sizeof(h_misc), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (const uint32_t e = (uint32_t)h_misc[0])
        return fail((e & kBErrTagFloat) && !(e & ~kBErrTagFloat) ? EXON_GPU_ERR_UNSUPPORTED : EXON_GPU_ERR_PARSE, "malformed BAM record at row %llu:%s%s%s%s", h_misc[1],
                    (e & kBErrLayout) ? " fields do not fit block_size, or an invalid CIGAR op / reference id;" : "",
                    (e & kBErrName) ? " missing read name in the non-nullable name column;" : "",
                    (e & kBErrTag) ? " an auxiliary field with an unknown type or one that runs past the record;" : "",
                    (e & kBErrTagFloat) ? " a B:f element too large for the \"{:.2}\" printer;" : "");
    const size_t off_bytes = (size_t)c->n_batches * (size_t)(c->batch_rows + 1) * 4;
    const size_t valid_bytes = (size_t)c->n_batches * (size_t)c->wpb * 4;
    for (int k = 0; k < kBNVar; ++k) {
        if (!need[k]) continue;
        for (int64_t b = 0; b < c->n_batches; ++b)
            if (c->base[k][(size_t)b + 1] - c->base[k][(size_t)b] > 0x7FFFFFFFll)
                return fail(EXON_GPU_ERR_UNSUPPORTED, "bam_next_batch: batch %lld overflows int32 offsets", (long long)b);
        if (k >= kBPlain) continue;  // the tags column has its own layout (below)
        if (int rc = dev_alloc(c->off[k], off_bytes, false)) return rc;
        if (int rc = dev_alloc(c->val[k], (size_t)c->base[k][nb1 - 1] * (k == kBQual ? 8 : 1), false)) return rc;
        ca.off[k] = (int32_t *)c->off[k].d;
        ca.val[k] = (uint8_t *)c->val[k].d;
    }
    if (want[10]) {
        const size_t n_ent_t = (size_t)c->base[kBTagN][nb1 - 1];
        if (int rc = dev_alloc(c->tags_loff, off_bytes, false)) return rc;
        if (int rc = dev_alloc(c->tag_off, (n_ent_t + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->tval_off, (n_ent_t + nb1) * 4, false)) return rc;
        if (int rc = dev_alloc(c->tag_val, 2 * n_ent_t, false)) return rc;
        if (int rc = dev_alloc(c->tval_val, (size_t)c->base[kBTagB][nb1 - 1], false)) return rc;
        ca.tags_loff = (int32_t *)c->tags_loff.d, ca.tag_off = (int32_t *)c->tag_off.d, ca.tval_off = (int32_t *)c->tval_off.d;
        ca.tag_val = (uint8_t *)c->tag_val.d, ca.tval_val = (uint8_t *)c->tval_val.d;
    }
    for (int col = 0; col < 10; ++col) {
        const int v = kColValid[col];
        if (!want[col] || v < 0) continue;
        if (int rc = dev_alloc(c->valid[v], valid_bytes, true)) return rc;
        ca.valid[v] = (uint32_t *)c->valid[v].d;
    }
    // ---- 4. emit ----
    bam_col_emit_kernel<<<grid, 256, 0, st>>>(ca);
    ctx->launches.fetch_add(1);
    if (need[kBSeq] || need[kBQual]) {
        bam_col_emit_long_kernel<<<(unsigned)((n_rows * 32 + 255) / 256), 256, 0, st>>>(ca);
        ctx->launches.fetch_add(1);
    }
    CUDA_TRY(cudaGetLastError());
    if (!c->on_device) {
        int rc = EXON_GPU_OK;
        auto to_host = [&](BamBuf &b) {
            if (!b.d || rc) return;
            if (cudaHostAlloc(&b.h, b.bytes, cudaHostAllocDefault) != cudaSuccess || cudaMemcpyAsync(b.h, b.d, b.bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                rc = fail(EXON_GPU_ERR_OOM, "bam_next_batch: host copy of the columns failed");
        };
        for (int k = 0; k < kBNVar; ++k) to_host(c->off[k]), to_host(c->val[k]);
        for (int k = 0; k < 5; ++k) to_host(c->valid[k]);
        to_host(c->flag), to_host(c->start), to_host(c->end);
        to_host(c->tags_loff), to_host(c->tag_off), to_host(c->tval_off), to_host(c->tag_val), to_host(c->tval_val);
        if (rc) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return EXON_GPU_OK;
}

struct BamBatchPriv {
    BamColumns *cols;
    int n_children;
    ArrowArray children[11];
    ArrowArray *child_ptrs[11];
    const void *bufs[11][3];
    // tags: list -> struct -> {tag utf8, value utf8}
    ArrowArray tag_struct, tag_name, tag_value;
    ArrowArray *tag_struct_ptr, *tag_kids[2];
    const void *tag_struct_bufs[1], *tag_name_bufs[3], *tag_value_bufs[3];
    ArrowArray item;  // quality_score's int64 child
    ArrowArray *item_ptr;
    const void *item_bufs[2];
    const void *struct_buffers[1];
};
void bam_release_child(ArrowArray *a) { a->release = nullptr; }
void bam_release_batch(ArrowArray *a) {
    auto *p = static_cast<BamBatchPriv *>(a->private_data);
    p->cols->unref();
    delete p;
    a->release = nullptr;
}
struct BamSchemaPriv {
    int n_children;
    ArrowSchema children[11];
    ArrowSchema *child_ptrs[11];
    ArrowSchema item;
    ArrowSchema *item_ptr;
    ArrowSchema tag_struct, tag_name, tag_value;
    ArrowSchema *tag_struct_ptr, *tag_kids[2];
};
void bam_release_schema_child(ArrowSchema *s) { s->release = nullptr; }
void bam_release_schema(ArrowSchema *s) {
    delete static_cast<BamSchemaPriv *>(s->private_data);
    s->release = nullptr;
}
// SAMSchemaBuilder::default, exon/exon-sam/src/schema_builder.rs:385-401
void bam_fill_schema(const std::vector<int> &projection, ArrowSchema *out) {
    static const char *names[11] = {"name", "flag", "reference", "start", "end", "mapping_quality", "cigar", "mate_reference", "sequence", "quality_score", "tags"};
    static const char *formats[11] = {"u", "i", "u", "l", "l", "u", "u", "u", "u", "+l", "+l"};
    static const bool nullable[11] = {false, false, true, true, true, true, false, true, false, false, true};
    auto *p = new BamSchemaPriv();
    p->n_children = (int)projection.size();
    for (int i = 0; i < p->n_children; ++i) {
        const int col = projection[(size_t)i];
        ArrowSchema &c = p->children[i];
        memset(&c, 0, sizeof(c));
        c.format = formats[col];
        c.name = names[col];
        c.flags = nullable[col] ? ARROW_FLAG_NULLABLE : 0;
        c.release = bam_release_schema_child;
        if (col == 9) {
            memset(&p->item, 0, sizeof(p->item));
            p->item.format = "l";
            p->item.name = "item";
            p->item.flags = ARROW_FLAG_NULLABLE;
            p->item.release = bam_release_schema_child;
            p->item_ptr = &p->item;
            c.n_children = 1;
            c.children = &p->item_ptr;
        }
        if (col == 10) {  // TagsMapBuilder::new, exon/exon-sam/src/tag_builder.rs:480-495: List<item: Struct{tag: Utf8 !null, value: Utf8}>
            auto init = [](ArrowSchema &x, const char *fmt, const char *name, bool nullable_) {
                memset(&x, 0, sizeof(x));
                x.format = fmt;
                x.name = name;
                x.flags = nullable_ ? ARROW_FLAG_NULLABLE : 0;
                x.release = bam_release_schema_child;
            };
            init(p->tag_struct, "+s", "item", true);
            init(p->tag_name, "u", "tag", false);
            init(p->tag_value, "u", "value", true);
            p->tag_kids[0] = &p->tag_name;
            p->tag_kids[1] = &p->tag_value;
            p->tag_struct.n_children = 2;
            p->tag_struct.children = p->tag_kids;
            p->tag_struct_ptr = &p->tag_struct;
            c.n_children = 1;
            c.children = &p->tag_struct_ptr;
        }
        p->child_ptrs[i] = &c;
    }
    memset(out, 0, sizeof(*out));
    out->format = "+s";
    out->name = "";
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = bam_release_schema;
    out->private_data = p;
}

}  // namespace

void bam_stream_schema(VcfStream *s, ArrowSchema *out) { bam_fill_schema(s->projection, out); }

int bam_next_batch(VcfStream *s, ArrowArray *out, ArrowSchema *out_schema) {
    if (!s->bam_cols) {
        int64_t rows = 0;
        if (int rc = s->bam_filter_count(nullptr, nullptr, 0, nullptr, &rows)) return rc;  // verified walks (and flush_gz)
        std::lock_guard<std::recursive_mutex> work(s->ctx->work_mu);
        if (int rc = bam_build_columns(s)) {
            bam_columns_free(s);
            return rc;
        }
        s->drained = true;
    }
    BamColumns *c = s->bam_cols;
    if (out_schema) bam_fill_schema(s->projection, out_schema);
    memset(out, 0, sizeof(*out));
    if (c->next >= c->n_batches) return EXON_GPU_OK;  // end of stream: release == NULL
    const int64_t b = c->next++;
    const int64_t row0 = c->batch_row0[(size_t)b], rows = c->batch_row0[(size_t)b + 1] - row0;
    auto *p = new BamBatchPriv();
    memset(static_cast<void *>(p), 0, sizeof(*p));
    p->cols = c;
    c->refs.fetch_add(1);
    p->n_children = (int)s->projection.size();
    const size_t lo = (size_t)b * (size_t)(c->batch_rows + 1), vw = (size_t)b * (size_t)c->wpb;
    for (int i = 0; i < p->n_children; ++i) {
        const int col = s->projection[(size_t)i];
        ArrowArray &a = p->children[i];
        a.length = rows;
        a.buffers = p->bufs[i];
        a.release = bam_release_child;
        const int vs = col < 10 ? kColValid[col] : -1, k = col < 10 ? kColVar[col] : -1;
        a.null_count = vs >= 0 ? -1 : 0;
        p->bufs[i][0] = vs >= 0 ? (const void *)(c->p<uint32_t>(c->valid[vs]) + vw) : nullptr;
        if (col == 10) {
            const long long e0 = c->base[kBTagN][(size_t)b], n_ent_b = c->base[kBTagN][(size_t)b + 1] - e0;
            a.n_buffers = 2;
            p->bufs[i][1] = c->p<int32_t>(c->tags_loff) + lo;
            auto init = [](ArrowArray &x, int64_t len, int nb, const void **bufs) {
                memset(&x, 0, sizeof(x));
                x.length = len;
                x.n_buffers = nb;
                x.buffers = bufs;
                x.release = bam_release_child;
            };
            init(p->tag_struct, n_ent_b, 1, p->tag_struct_bufs);
            init(p->tag_name, n_ent_b, 3, p->tag_name_bufs);
            init(p->tag_value, n_ent_b, 3, p->tag_value_bufs);
            p->tag_struct_bufs[0] = nullptr;
            p->tag_name_bufs[0] = nullptr;
            p->tag_name_bufs[1] = c->p<int32_t>(c->tag_off) + e0 + b;
            p->tag_name_bufs[2] = c->p<uint8_t>(c->tag_val) + 2 * e0;
            p->tag_value_bufs[0] = nullptr;
            p->tag_value_bufs[1] = c->p<int32_t>(c->tval_off) + e0 + b;
            p->tag_value_bufs[2] = c->p<uint8_t>(c->tval_val) + c->base[kBTagB][(size_t)b];
            p->tag_kids[0] = &p->tag_name;
            p->tag_kids[1] = &p->tag_value;
            p->tag_struct.n_children = 2;
            p->tag_struct.children = p->tag_kids;
            p->tag_struct_ptr = &p->tag_struct;
            a.n_children = 1;
            a.children = &p->tag_struct_ptr;
        } else if (col == 1) {
            a.n_buffers = 2;
            p->bufs[i][1] = c->p<int32_t>(c->flag) + row0;
        } else if (col == 3 || col == 4) {
            a.n_buffers = 2;
            p->bufs[i][1] = c->p<long long>(col == 3 ? c->start : c->end) + row0;
        } else if (col == 9) {
            a.n_buffers = 2;
            p->bufs[i][1] = c->p<int32_t>(c->off[k]) + lo;
            p->item.length = c->base[k][(size_t)b + 1] - c->base[k][(size_t)b];
            p->item.n_buffers = 2;
            p->item_bufs[0] = nullptr;
            p->item_bufs[1] = c->p<long long>(c->val[k]) + c->base[k][(size_t)b];
            p->item.buffers = p->item_bufs;
            p->item.release = bam_release_child;
            p->item_ptr = &p->item;
            a.n_children = 1;
            a.children = &p->item_ptr;
        } else {
            a.n_buffers = 3;
            p->bufs[i][1] = c->p<int32_t>(c->off[k]) + lo;
            p->bufs[i][2] = c->p<uint8_t>(c->val[k]) + c->base[k][(size_t)b];
        }
        p->child_ptrs[i] = &a;
    }
    p->struct_buffers[0] = nullptr;
    out->length = rows;
    out->n_buffers = 1;
    out->buffers = p->struct_buffers;
    out->n_children = p->n_children;
    out->children = p->child_ptrs;
    out->release = bam_release_batch;
    out->private_data = p;
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_bam_open(exon_gpu_ctx *c, exon_gpu_stream **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "bam_open: NULL argument");
    exon_gpu_vcf_opts vo;
    memset(&vo, 0, sizeof(vo));
    if (int rc = exon_gpu_vcf_open(c, &vo, out)) return rc;
    (*out)->fmt = kFmtBam;
    (*out)->hdr = VcfStream::kBody;
    return EXON_GPU_OK;
}

int exon_gpu_bam_open_columns(exon_gpu_ctx *c, const exon_gpu_bam_opts *o, exon_gpu_stream **out) {
    if (!c || !o || !out) return fail(EXON_GPU_ERR_ARG, "bam_open_columns: NULL argument");
    if (o->batch_rows < 0 || o->n_projection < 0 || o->n_projection > 11 || (o->n_projection > 0 && !o->projection))
        return fail(EXON_GPU_ERR_ARG, "bam_open_columns: bad batch_rows / projection");
    for (int i = 0; i < o->n_projection; ++i) {
        if (o->projection[i] < 0 || o->projection[i] > 10) return fail(EXON_GPU_ERR_ARG, "bam_open_columns: projection index %d is not a BAM file-schema column", o->projection[i]);
        for (int j = 0; j < i; ++j)
            if (o->projection[j] == o->projection[i]) return fail(EXON_GPU_ERR_ARG, "bam_open_columns: column %d is projected twice", o->projection[i]);
    }
    if (int rc = exon_gpu_bam_open(c, out)) return rc;
    if (o->batch_rows > 0) (*out)->batch_rows = o->batch_rows;
    (*out)->projection.assign(o->projection, o->projection + o->n_projection);
    (*out)->columns_on_device = o->columns_on_device != 0;
    return EXON_GPU_OK;
}

int exon_gpu_bam_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema) {
    if (!s || !out || s->fmt != kFmtBam) return fail(EXON_GPU_ERR_ARG, "bam_next_batch: not a BAM stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return bam_next_batch(s, out, out_schema);
}

int exon_gpu_bam_feed(exon_gpu_stream *s, const uint8_t *data, size_t len, int is_last) {
    if (!s || s->fmt != kFmtBam) return fail(EXON_GPU_ERR_ARG, "bam_feed: not a BAM stream");
    return exon_gpu_stream_feed_gzip(s, data, len, is_last);
}

int exon_gpu_bam_filter_count_by_reference(exon_gpu_stream *s, const exon_gpu_bam_pred *pred, int64_t *counts, int32_t cap,
                                           int32_t *n_groups, int64_t *total_rows) {
    if (!s || s->fmt != kFmtBam) return fail(EXON_GPU_ERR_ARG, "bam_filter_count_by_reference: not a BAM stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return s->bam_filter_count(pred, counts, cap, n_groups, total_rows);
}

int exon_gpu_bam_group_name(exon_gpu_stream *s, int32_t group, const char **name) {
    if (!s || !name || s->fmt != kFmtBam) return fail(EXON_GPU_ERR_ARG, "bam_group_name: bad argument");
    if (group < 0 || group > (int32_t)s->bam_groups.size()) return fail(EXON_GPU_ERR_ARG, "bam_group_name: group %d out of range", group);
    *name = group == (int32_t)s->bam_groups.size() ? nullptr : s->bam_groups[(size_t)group].c_str();  // the last group is the NULL reference
    return EXON_GPU_OK;
}

}  // extern "C"
