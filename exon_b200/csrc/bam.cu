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
#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"
#include "internal.h"

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
            ++my_rows;
            p += 4 + (uint64_t)(uint32_t)block_size;
        }
        exit_at = bad ? kBadExit : p;
        a.exits[e] = exit_at;
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
    std::lock_guard<std::mutex> work(c->work_mu);
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
        if (round == 0) CUDA_TRY(cudaEventRecord(c->ev0, st));
        bam_walk_kernel<<<(launch_n + kBamThreads - 1) / kBamThreads, kBamThreads, 0, st>>>(a);
        bam_verify_kernel<<<(launch_n + 127) / 128, 128, 0, st>>>(ent, launch_n, a.exits, serial ? 0 : 1, d_verify);
        if (round == 0) {
            CUDA_TRY(cudaEventRecord(c->ev1, st));
            c->timed = true;
        }
        c->launches.fetch_add(2);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(h, d + bam_o_misc, 128, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h + 128, d + bam_o_counts, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const unsigned int *v = reinterpret_cast<const unsigned int *>(h + 64);
        if (v[0] == 0 && v[1] == 0) ok = true;
        else if (v[0] == 0 || serial)
            return fail(EXON_GPU_ERR_PARSE, "malformed BAM record (block_size / refID out of range, or a truncated record)");
        // else: some walks began inside a record; the entry points were corrected, go again
    }
    if (total_rows) *total_rows = (int64_t) * reinterpret_cast<const unsigned long long *>(h);
    if (counts) memcpy(counts, h + 128, (size_t)n_groups * 8);
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
