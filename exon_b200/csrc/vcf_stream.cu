// vcf_stream.cu -- partition-stream state machine: header skipping, the device arena, segment tables and the
// launch of the fused scan (K1).  Host-side work here is framing only (what VCFOpener::open / read_header do,
// exon/exon-core/src/datasources/vcf/file_opener/unindex_file_opener.rs:48-92); records are never parsed on
// the host.
#include <algorithm>
#include <atomic>
#include <cstring>

#include "internal.h"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

int OwnedRegion::assign(const exon_gpu_region *r) {
    has_chrom = has_interval = false;
    chrom.clear();
    lo = 1;
    hi = INT64_MAX;
    if (!r) return EXON_GPU_OK;
    if (r->has_chrom) {
        if (!r->chrom || r->chrom_len < 0) return fail(EXON_GPU_ERR_ARG, "region: chrom is NULL");
        has_chrom = true;
        chrom.assign(r->chrom, (size_t)r->chrom_len);
    }
    if (r->has_interval) {
        has_interval = true;
        lo = r->lo;
        hi = r->hi;
    }
    return EXON_GPU_OK;
}

static bool same_region(const OwnedRegion &a, const OwnedRegion &b) {
    return a.has_chrom == b.has_chrom && a.has_interval == b.has_interval && a.chrom == b.chrom &&
           (!a.has_interval || (a.lo == b.lo && a.hi == b.hi));
}

// A chrom literal no CHROM field can equal: empty, longer than the kernel's pattern buffer, or holding a
// field/line separator.  The count is 0 by construction (FilterExec's `eq` would be false on every row).
static bool unmatchable(const OwnedRegion &r) {
    if (!r.has_chrom) return false;
    if (r.chrom.empty() || r.chrom.size() > (size_t)kMaxChrom) return true;
    return r.chrom.find('\t') != std::string::npos || r.chrom.find('\n') != std::string::npos;
}

void VcfStream::release_all() {
    for (auto &b : blocks) ctx->put_block(b);
    blocks.clear();
    runs.clear();
    file_marks.clear();
    gz_pending.clear();
    gz_files.clear();
    if (gz_copy_stream) cudaStreamSynchronize(gz_copy_stream);
    for (auto &g : gz_inflight) cudaFree(g.d_tab);  // the caller has synchronised the context stream
    gz_inflight.clear();
    bam_files.clear();
    bam_tables_dirty = true;
    bam_groups.clear();
    gz_members.clear();
    gz_staged = 0;
    cur_run_open = false;
    tail_len = 0;
    body_bytes = 0;
    eager_scanned = 0;
    eager_runs_done = 0;
    hdr = fmt == kFmtVcf ? kAtLineStart : kBody;
    file_open = false;
    last_byte_newline = true;
    segs_dirty = true;
    drained = false;
    columns_free(this);
    fq_columns_free(this);
    bam_columns_free(this);
    gff_columns_free(this);
    mzml_columns_free(this);
}

// Append body bytes that live on the host to the arena (one async H2D copy on the stream).
int VcfStream::append_host(const uint8_t *p, size_t n) {
    if (n == 0) return EXON_GPU_OK;
    if (!cur_run_open || blocks.back().used + n > blocks.back().cap) {
        DevBlock nb;
        const size_t carry = cur_run_open ? (size_t)tail_len : 0;
        if (int rc = ctx->get_block(n + carry, &nb)) return rc;
        if (carry) {
            // the partial last line moves to the new block so that every run holds whole lines up to its tail
            Run &old = runs.back();
            CUDA_TRY(cudaMemcpyAsync(nb.ptr, old.base + old.len - carry, carry, cudaMemcpyDeviceToDevice, ctx->stream));
            old.len -= (int64_t)carry;
            old.ends_with_newline = true;
            blocks.back().used -= carry;
            if (old.len == 0) runs.pop_back();
            nb.used = carry;
        }
        blocks.push_back(nb);
        runs.push_back(Run{nb.ptr, (int64_t)carry, false});
        cur_run_open = true;
    }
    DevBlock &b = blocks.back();
    CUDA_TRY(cudaMemcpyAsync(b.ptr + b.used, p, n, cudaMemcpyHostToDevice, ctx->stream));
    b.used += n;
    runs.back().len += (int64_t)n;
    const void *nl = memrchr(p, '\n', n);
    if (nl) tail_len = (int64_t)(p + n - ((const uint8_t *)nl + 1));
    else tail_len += (int64_t)n;
    last_byte_newline = p[n - 1] == '\n';
    runs.back().ends_with_newline = last_byte_newline;
    segs_dirty = true;
    return EXON_GPU_OK;
}

int VcfStream::end_file() {
    if (cur_run_open && !last_byte_newline) {
        // normalise: the file's last record gets its '\n', so the next file's body can follow in the same run
        static const uint8_t nl = '\n';
        const int64_t before = body_bytes;
        if (int rc = append_host(&nl, 1)) return rc;
        body_bytes = before;
    }
    if (!runs.empty()) file_marks.push_back(FileMark{runs.size() - 1, runs.back().len});
    hdr = fmt == kFmtVcf ? kAtLineStart : kBody;
    file_open = false;
    return EXON_GPU_OK;
}

// Advance the header state machine over [p, end); returns the first body byte (or end).
static const uint8_t *skip_header(VcfStream::HdrState &st, const uint8_t *p, const uint8_t *end) {
    while (st != VcfStream::kBody && p < end) {
        if (st == VcfStream::kAtLineStart) {
            if (*p == '#') st = VcfStream::kInHeaderLine;
            else { st = VcfStream::kBody; break; }
        }
        const void *nl = memchr(p, '\n', (size_t)(end - p));
        if (!nl) return end;
        p = (const uint8_t *)nl + 1;
        st = VcfStream::kAtLineStart;
    }
    return p;
}

int VcfStream::feed_host(const uint8_t *text, size_t len, bool is_last) {
    const uint8_t *end = text + len;
    const uint8_t *p = skip_header(hdr, text, end);
    file_open = true;
    if (p < end) {
        body_bytes += (int64_t)(end - p);
        if (int rc = append_host(p, (size_t)(end - p))) return rc;
    }
    if (is_last)
        if (int rc = end_file()) return rc;
    if (has_pushdown) return eager_scan(false);
    return EXON_GPU_OK;
}

int VcfStream::feed_device(const uint8_t *text, size_t len, bool is_last) { return frame_device_range(text, len, is_last, -1, -1); }

// Frames a device-resident byte range as (part of) a file: header skipped, run recorded, file end marked.
// known_body_off >= 0: the caller already knows where the body starts inside the range (header state machine run on
// a host copy); known_last_byte >= 0: the caller already knows the last byte of the range.
int VcfStream::frame_device_range(const uint8_t *text, size_t len, bool is_last, int64_t known_body_off, int known_last_byte) {
    size_t off = 0;
    if (known_body_off >= 0) {
        off = (size_t)known_body_off;
        hdr = kBody;
    } else if (hdr != kBody) {
        // locate the end of the header by bouncing prefixes through pinned memory (records are not touched)
        std::lock_guard<std::recursive_mutex> work(ctx->work_mu);  // h_scratch is a context-wide area
        const size_t kProbe = (size_t)1 << 20;
        if (int rc = ctx->ensure_scratch(0, kProbe)) return rc;
        while (off < len && hdr != kBody) {
            const size_t n = std::min(kProbe, len - off);
            CUDA_TRY(cudaMemcpyAsync(ctx->h_scratch, text + off, n, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            const uint8_t *h = (const uint8_t *)ctx->h_scratch;
            const uint8_t *q = skip_header(hdr, h, h + n);
            off += (size_t)(q - h);
        }
    }
    file_open = true;
    const size_t n = len - off;
    if (n) {
        if (cur_run_open && tail_len > 0)
            return fail(EXON_GPU_ERR_STATE, "vcf_feed: a device range cannot follow a host range that ended mid-line");
        uint8_t last;
        if (known_last_byte >= 0) {
            last = (uint8_t)known_last_byte;
        } else {
            CUDA_TRY(cudaMemcpyAsync(h_res + 12, text + len - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            last = *reinterpret_cast<const uint8_t *>(h_res + 12);
        }
        if (!is_last && last != '\n')
            return fail(EXON_GPU_ERR_ARG, "vcf_feed: a non-final device range must end on a line boundary");
        cur_run_open = false;
        tail_len = 0;
        runs.push_back(Run{text + off, (int64_t)n, last == '\n'});
        body_bytes += (int64_t)n;
        last_byte_newline = true;  // nothing pending in the arena
        segs_dirty = true;
    }
    if (is_last) {
        if (!runs.empty()) file_marks.push_back(FileMark{runs.size() - 1, runs.back().len});
        hdr = fmt == kFmtVcf ? kAtLineStart : kBody;
        file_open = false;
    }
    if (has_pushdown) return eager_scan(false);
    return EXON_GPU_OK;
}

// header state machine over a host copy of a file's first bytes; returns the body offset, or -1 when the header
// does not end inside [p, p + n) (and n is not the whole file)
int64_t VcfStream::probe_body_offset(const uint8_t *p, size_t n, bool whole_file) const {
    HdrState st = fmt == kFmtVcf ? kAtLineStart : kBody;
    const uint8_t *q = skip_header(st, p, p + n);
    if (st == kBody || whole_file) return (int64_t)(q - p);
    return -1;
}

static void fill_segs(const std::vector<Run> &runs, size_t first, size_t last, int64_t first_skip_bytes,
                      int64_t last_len_override, int tile, std::vector<ScanSeg> &out, int64_t *n_tiles) {
    out.clear();
    int64_t t = 0;
    for (size_t i = first; i < last; ++i) {
        const uint8_t *b = runs[i].base;
        int64_t len = runs[i].len;
        if (i + 1 == last && last_len_override >= 0) len = last_len_override;
        if (i == first) { b += first_skip_bytes; len -= first_skip_bytes; }
        if (len <= 0) continue;
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)b & 15);
        sg.base = b - sg.skip;
        sg.len = len;
        sg.tile0 = t;
        sg.pad_ = 0;
        t += (sg.skip + len + tile - 1) / tile;
        out.push_back(sg);
    }
    ScanSeg sentinel;
    sentinel.base = nullptr;
    sentinel.len = 0;
    sentinel.tile0 = t;
    sentinel.skip = 0;
    sentinel.pad_ = 0;
    out.push_back(sentinel);
    // the kernel's cursor reads segs[c + 1].tile0 after the last tile of the last segment
    *n_tiles = t;
}

// Uploads a segment table and derives the tile descriptors the scan kernel's producers read (vcf_scan.cu).
static int upload_segs(VcfStream *s, const std::vector<ScanSeg> &h, int64_t n_tiles) {
    if ((size_t)n_tiles > s->d_tiles_cap) {
        if (s->d_tiles) {
            CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
            CUDA_TRY(cudaFree(s->d_tiles));
            s->d_tiles = nullptr;
        }
        s->d_tiles_cap = std::max<size_t>((size_t)n_tiles + (size_t)n_tiles / 4, 1024);
        CUDA_TRY(cudaMalloc((void **)&s->d_tiles, s->d_tiles_cap * sizeof(TileDesc)));
    }
    if (h.size() > s->d_segs_cap) {
        if (s->d_segs) {
            CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
            CUDA_TRY(cudaFree(s->d_segs));
            s->d_segs = nullptr;
        }
        s->d_segs_cap = std::max<size_t>(h.size() * 2, 256);
        CUDA_TRY(cudaMalloc((void **)&s->d_segs, s->d_segs_cap * sizeof(ScanSeg)));
    }
    CUDA_TRY(cudaMemcpyAsync(s->d_segs, h.data(), h.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, s->ctx->stream));
    CUDA_TRY(launch_build_tile_descs(s->d_segs, (int)h.size() - 1, n_tiles, s->variant, s->d_tiles, s->ctx->stream));
    if (n_tiles > 0) s->ctx->launches.fetch_add(1);
    return EXON_GPU_OK;
}

int VcfStream::build_seg_table() {
    const int tile = scan_tile_bytes(variant);
    if (!segs_dirty && seg_variant == variant) return EXON_GPU_OK;
    fill_segs(runs, 0, runs.size(), 0, -1, tile, h_segs, &n_tiles);
    if (int rc = upload_segs(this, h_segs, n_tiles)) return rc;
    segs_dirty = false;
    seg_variant = variant;
    return EXON_GPU_OK;
}

int VcfStream::launch_scan(const OwnedRegion &r, const TileDesc *d_table, int n_segs, int64_t tiles, ScanAcc *acc,
                           const ScanTail &tail) {
    if (tiles <= 0 && !tail.finalize) return EXON_GPU_OK;
    ScanArgs a;
    memset(&a, 0, sizeof(a));
    a.tiles = d_table;
    a.n_segs = n_segs;
    a.n_tiles = tiles;
    a.has_chrom = r.has_chrom;
    a.has_interval = r.has_interval;
    a.lo = r.lo;
    a.hi = r.hi;
    a.chrom_len = (int32_t)r.chrom.size();
    a.pat_len = a.chrom_len + 2;
    a.pat[0] = '\n';
    memcpy(a.pat + 1, r.chrom.data(), r.chrom.size());
    a.pat[1 + r.chrom.size()] = '\t';
    a.acc = acc;
    a.tail = tail;
    ScanMode mode;
    if (strict) mode = kScanDense;
    else if (r.has_chrom) mode = r.chrom.size() == 1 ? kScanKey3 : kScanKey4;
    else mode = r.has_interval ? kScanDense : kScanLines;
    ScanConfig cfg;
    cfg.variant = variant;
    cfg.ctas = 0;
    CUDA_TRY(ctx->timed_begin(ctx->stream));
    CUDA_TRY(launch_vcf_scan(a, mode, cfg, ctx->sm_count, ctx->stream));
    ctx->launches.fetch_add(1);
    CUDA_TRY(ctx->timed_end(ctx->stream));
    return EXON_GPU_OK;
}

// Pushdown mode: scan what arrived since the last call, so that the scan of feed k runs while the caller
// prepares feed k+1.  Only whole lines are covered; a partial last line waits for its continuation.
int VcfStream::eager_scan(bool final_flush) {
    if (unmatchable(pushdown)) return EXON_GPU_OK;
    const int tile = scan_tile_bytes(variant);
    ScanTail accumulate_only;
    memset(&accumulate_only, 0, sizeof(accumulate_only));
    while (eager_runs_done < runs.size()) {
        const size_t i = eager_runs_done;
        const bool open = cur_run_open && i + 1 == runs.size();
        int64_t upto = runs[i].len;
        if (open && !final_flush) upto -= tail_len;
        if (upto > eager_scanned) {
            std::vector<ScanSeg> h;
            int64_t tiles = 0;
            fill_segs(runs, i, i + 1, eager_scanned, upto, tile, h, &tiles);
            // table slots for eager launches live behind the lazy table: reuse d_segs' tail via a private buffer
            if (int rc = upload_segs(this, h, tiles)) return rc;
            segs_dirty = true;  // the lazy table was overwritten
            if (int rc = launch_scan(pushdown, d_tiles, (int)h.size() - 1, tiles, reinterpret_cast<ScanAcc *>(d_res + 4), accumulate_only))
                return rc;
            eager_scanned = upto;
        }
        if (open) break;  // may still grow
        ++eager_runs_done;
        eager_scanned = 0;
    }
    return EXON_GPU_OK;
}

// Waits for the record the scan tail publishes in mapped pinned memory.  The host polls the sequence word instead of
// synchronising the stream (the wake-up of a blocking synchronise costs more than the D2H hop of one 8-byte store); the
// stream is queried from time to time so that a failed launch turns into an error instead of a hang.
int VcfStream::wait_published() {
    volatile unsigned long long *h = h_res;
    for (uint64_t spins = 1;; ++spins) {
        if (h[kHostSeq] == host_seq) break;
        if ((spins & 0x3FFF) == 0) {
            const cudaError_t e = cudaStreamQuery(ctx->stream);
            if (e == cudaSuccess) {
                if (h[kHostSeq] == host_seq) break;
                return fail(EXON_GPU_ERR_CUDA, "fused scan: the stream drained without publishing a result");
            }
            if (e != cudaErrorNotReady) return fail(EXON_GPU_ERR_CUDA, "fused scan: %s", cudaGetErrorString(e));
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return EXON_GPU_OK;
}

// One query = one launch: the scan, and in its tail the publication of the result (and, for a global query, the
// exchange with the peer ranks).  Queries that have nothing to scan launch the tail alone, so that a rank ALWAYS takes
// part in a global query's exchange -- a rank that skipped it would leave its peers waiting (and the sequence numbers of
// the exchange out of step for good).
int VcfStream::run_query(const exon_gpu_region *region, int64_t *device_out, bool want_host, bool global) {
    ScanTail tail;
    memset(&tail, 0, sizeof(tail));
    tail.finalize = 1;
    tail.device_out = reinterpret_cast<long long *>(device_out);
    OwnedRegion r;
    int rc = flush_gz();
    if (rc == EXON_GPU_OK) rc = r.assign(region);
    if (r.has_interval && r.lo > r.hi) {
        // empty interval: arrow's gt_eq AND lt_eq selects nothing
        r.has_chrom = true;
        r.chrom.clear();
    }
    const bool eager = rc == EXON_GPU_OK && has_pushdown && same_region(r, pushdown);
    last_eager = eager;
    ScanAcc *acc = reinterpret_cast<ScanAcc *>(eager ? d_res + 4 : d_res);
    bool scan = false;
    if (rc == EXON_GPU_OK) {
        if (eager) {
            rc = eager_scan(true);
            tail.finalize = 2;  // publish, keep accumulating
        } else if (!unmatchable(r)) {
            rc = build_seg_table();
            scan = rc == EXON_GPU_OK;
        }
    }
    bool nccl_fallback = false;
    if (global) {
        if (ctx->nccl_ranks > 1 && !peer_xchg_arm(ctx, &tail)) {
            nccl_fallback = true;
            tail.device_out = reinterpret_cast<long long *>(d_res + 8);
        }
    }
    if (want_host || global) {
        tail.host_out = h_res_dev;
        tail.host_seq = ++host_seq;
    }
    if (rc != EXON_GPU_OK) {
        // the local part failed before its launch: contribute whatever the accumulator holds (0) and report rc afterwards
        if (global && tail.n_ranks > 1) {
            const std::string keep = exon_gpu_last_error();
            ScanTail t2 = tail;
            t2.host_out = nullptr;
            launch_scan(r, nullptr, 0, 0, acc, t2);
            fail(rc, "%s", keep.c_str());
        }
        return rc;
    }
    if (int rc2 = launch_scan(r, scan ? d_tiles : nullptr, scan ? (int)h_segs.size() - 1 : 0, scan ? n_tiles : 0, acc, tail)) return rc2;
    if (nccl_fallback) {
        CUDA_TRY(cudaMemcpyAsync(d_res + 9, d_res + 8, sizeof(int64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        if (int rc2 = nccl_allreduce_i64(ctx, reinterpret_cast<int64_t *>(d_res + 9), 1)) return rc2;
        CUDA_TRY(cudaMemcpyAsync(h_res + 8, d_res + 9, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    if (!(want_host || global)) return EXON_GPU_OK;
    if (int rc2 = wait_published()) return rc2;
    if (nccl_fallback) h_res[kHostGlobal] = h_res[8];
    if (h_res[kHostXchgErr]) return fail(EXON_GPU_ERR_NCCL, "peer exchange timed out: a rank did not deliver its partial");
    const uint32_t flags = (uint32_t)h_res[kHostFlags];
    if (flags)
        return fail(EXON_GPU_ERR_PARSE, "malformed VCF record:%s%s", (flags & kErrBadPos) ? " POS is not a positive decimal integer;" : "",
                    (flags & kErrShortLine) ? " line ended before the field being read;" : "");
    return EXON_GPU_OK;
}

int VcfStream::filter_count(const exon_gpu_region *region, int64_t *device_out, int64_t *host_out) {
    if (int rc = run_query(region, device_out, host_out != nullptr, false)) return rc;
    if (host_out) *host_out = (int64_t)h_res[kHostLocal];
    return EXON_GPU_OK;
}

int VcfStream::filter_count_global(const exon_gpu_region *region, int64_t *out_local, int64_t *out_global) {
    if (int rc = run_query(region, nullptr, true, true)) return rc;
    if (out_local) *out_local = (int64_t)h_res[kHostLocal];
    if (out_global) *out_global = (int64_t)h_res[kHostGlobal];
    return EXON_GPU_OK;
}

int Ctx::ensure_scratch(size_t dev_bytes, size_t host_bytes) {
    if (dev_bytes > scratch_cap) {
        if (scratch) {
            CUDA_TRY(cudaStreamSynchronize(stream));
            CUDA_TRY(cudaFree(scratch));
            scratch = nullptr;
        }
        CUDA_TRY(cudaMalloc(&scratch, dev_bytes));
        scratch_cap = dev_bytes;
    }
    if (host_bytes > h_scratch_cap) {
        if (h_scratch) {
            CUDA_TRY(cudaStreamSynchronize(stream));
            CUDA_TRY(cudaFreeHost(h_scratch));
            h_scratch = nullptr;
        }
        CUDA_TRY(cudaHostAlloc(&h_scratch, host_bytes, cudaHostAllocDefault));
        h_scratch_cap = host_bytes;
    }
    return EXON_GPU_OK;
}

// Cuts the resident runs at the recorded file ends (a run may hold the tail of one file and the head of the
// next when the host fed them back to back).
void VcfStream::cut_pieces(std::vector<Piece> &out) const {
    out.clear();
    bool pending = true;
    size_t mi = 0;
    const auto &marks = file_marks;
    for (size_t r = 0; r < runs.size(); ++r) {
        const Run &run = runs[r];
        int64_t at = 0;
        while (mi < marks.size() && marks[mi].run <= r) {
            if (marks[mi].run == r) {
                const int64_t end = std::min<int64_t>(marks[mi].len, run.len);
                if (end > at) {
                    out.push_back(Piece{run.base + at, end - at, pending});
                    at = end;
                }
                pending = true;  // whatever follows belongs to the next file
            }
            ++mi;
        }
        if (run.len > at) {
            out.push_back(Piece{run.base + at, run.len - at, pending});
            pending = false;
        }
    }
}

int Ctx::ensure_scratch_b(size_t dev_bytes) {
    if (dev_bytes > scratch_b_cap) {
        if (scratch_b) {
            CUDA_TRY(cudaStreamSynchronize(stream));
            CUDA_TRY(cudaFree(scratch_b));
            scratch_b = nullptr;
            scratch_b_cap = 0;
        }
        CUDA_TRY(cudaMalloc(&scratch_b, dev_bytes));
        scratch_b_cap = dev_bytes;
    }
    return EXON_GPU_OK;
}

}  // namespace exon
