// bgzf.cu -- BGZF / gzip framing and the host side of the device inflate.
//
// Replaces the CPU DEFLATE the reference runs in front of every parser: noodles-bgzf 0.34 `AsyncReader` /
// async-compression `GzipDecoder` at exon/exon-core/src/datasources/vcf/file_opener/unindex_file_opener.rs:59-73,
// exon/exon-core/src/streaming_bgzf.rs:22-118 (block framing), fastq/file_opener.rs:51.  A BGZF file is a series
// of independent gzip members of <= 64 KiB uncompressed data each (SAM spec 4.1), so members decode in parallel:
// the host walks the member headers (18 bytes per member, no payload byte is touched; bgzf_walk), the compressed bytes
// travel to HBM as they are on a copy stream into double-buffered staging (gz_stage), and the members of many files are
// inflated by ONE launch of the two kernels of inflate.cu straight into the arena (launch_gz) while the next group's bytes
// are still on their way; launched groups are checked (stream errors, ISIZE) and framed at the next query (harvest_gz).
// Plain single-member gzip (not BGZF) takes the same path as one member.
// The first cut of this file also held a one-kernel inflate (16 lanes per member: 44 ms per 2.75 GB); inflate.cu replaced
// it (19 ms) and it was removed.
// Output is bit-exact DEFLATE (RFC 1951): tests compare with zlib on the reference's .gz fixtures and on
// synthetic shards; ISIZE of every member is checked, CRC32 is not (documented in DESIGN.md).
#include <algorithm>
#include <cstring>

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

// Parses the gzip member that starts at data[p]: payload range, ISIZE, and where the next member starts.  BGZF members
// carry their size in the 'BC' extra subfield; a member without it (plain gzip) is taken to extend to the end of the data.
static int bgzf_parse_member(const uint8_t *data, size_t len, size_t p, BgzfMember *m, size_t *next) {
    if (len - p < 18 || data[p] != 0x1f || data[p + 1] != 0x8b || data[p + 2] != 8)
        return fail(EXON_GPU_ERR_PARSE, "bgzf: bad gzip magic at byte %zu", p);
    const uint8_t flg = data[p + 3];
    size_t q = p + 10;
    long bsize = -1;
    if (flg & 4) {
        const size_t xlen = (size_t)data[q] | ((size_t)data[q + 1] << 8);
        q += 2;
        if (q + xlen > len) return fail(EXON_GPU_ERR_PARSE, "bgzf: truncated extra field at byte %zu", p);
        size_t x = q;
        while (x + 4 <= q + xlen) {
            const size_t slen = (size_t)data[x + 2] | ((size_t)data[x + 3] << 8);
            if (data[x] == 'B' && data[x + 1] == 'C' && slen == 2 && x + 6 <= q + xlen) bsize = (long)data[x + 4] | ((long)data[x + 5] << 8);
            x += 4 + slen;
        }
        q += xlen;
    }
    if (flg & 8) { while (q < len && data[q]) ++q; ++q; }   // FNAME
    if (flg & 16) { while (q < len && data[q]) ++q; ++q; }  // FCOMMENT
    if (flg & 2) q += 2;                                    // FHCRC
    const size_t end = bsize >= 0 ? p + (size_t)bsize + 1 : len;
    if (end > len || q + 8 > end) return fail(EXON_GPU_ERR_PARSE, "bgzf: truncated member at byte %zu", p);
    m->in_off = q;
    m->in_len = (uint32_t)(end - 8 - q);
    m->isize = (uint32_t)data[end - 4] | ((uint32_t)data[end - 3] << 8) | ((uint32_t)data[end - 2] << 16) | ((uint32_t)data[end - 1] << 24);
    m->out_addr = 0;
    m->tok_off = 0;
    m->pad_ = 0;
    if (bsize >= 0 && m->isize > 65536u) return fail(EXON_GPU_ERR_PARSE, "bgzf: member at byte %zu claims %u bytes (> 64 KiB)", p, m->isize);
    *next = end;
    return EXON_GPU_OK;
}

// Walks the gzip members of `data` (a whole file); out_addr = offset of each member in the uncompressed stream.
int bgzf_walk(const uint8_t *data, size_t len, std::vector<BgzfMember> &out, uint64_t *total_out) {
    out.clear();
    size_t p = 0;
    uint64_t uo = 0;
    while (p < len) {
        BgzfMember m;
        size_t next = 0;
        if (int rc = bgzf_parse_member(data, len, p, &m, &next)) return rc;
        m.out_addr = uo;  // the caller rebases it to a device address
        uo += m.isize;
        out.push_back(m);
        p = next;
    }
    *total_out = uo;
    return EXON_GPU_OK;
}

// One whole BGZF / gzip file: the compressed bytes start their way to HBM at once; the inflate itself is deferred so
// that the members of several files share one launch (flush_gz).
int VcfStream::feed_gzip(const uint8_t *data, size_t len, bool is_last) {
    if (!is_last || !gz_pending.empty()) {
        gz_pending.insert(gz_pending.end(), data, data + len);
        if (!is_last) {
            file_open = true;
            return EXON_GPU_OK;
        }
        data = gz_pending.data();
        len = gz_pending.size();
    }
    struct Clear {
        std::vector<uint8_t> &v;
        ~Clear() { v.clear(); }
    } clear{gz_pending};
    if (cur_run_open && tail_len > 0) return fail(EXON_GPU_ERR_STATE, "feed_gzip: the previous plain-text range ended mid-line");
    std::vector<BgzfMember> members;
    uint64_t total = 0;
    if (len)
        if (int rc = bgzf_walk(data, len, members, &total)) return rc;
    uint8_t *dst = nullptr;
    // a group that would outgrow one wave of decoder lanes goes first: the members beyond the wave would cost a second pass
    if (!gz_members.empty() && gz_members.size() + members.size() > gz_wave_members())
        if (int rc = launch_gz()) return rc;
    if (total > 0) {
        const size_t need = ((len + 16 + 255) & ~(size_t)255);
        uint8_t *dz = nullptr;
        if (int rc = gz_stage(need, &dz)) return rc;
        CUDA_TRY(cudaMemcpyAsync(dz, data, len, cudaMemcpyHostToDevice, gz_copy_stream));
        if (!gz_pending.empty()) CUDA_TRY(cudaStreamSynchronize(gz_copy_stream));  // the source is our own buffer, cleared on return
        // arena space: the tail of the current block if the file fits, else a fresh block (+1 for a missing final '\n')
        if (blocks.empty() || blocks.back().used + total + 1 > blocks.back().cap) {
            DevBlock nb;
            if (int rc = ctx->get_block((size_t)total + 1, &nb)) return rc;
            blocks.push_back(nb);
        }
        DevBlock &b = blocks.back();
        dst = b.ptr + b.used;
        b.used = std::min(b.cap, b.used + (((size_t)total + 1 + 15) & ~(size_t)15));  // the next file starts 16-byte aligned
        cur_run_open = false;  // plain-text feeds that follow start their own block
        tail_len = 0;
        for (BgzfMember m : members) {
            m.in_off += gz_staged;
            m.out_addr = (uint64_t)reinterpret_cast<uintptr_t>(dst) + m.out_addr;
            gz_members.push_back(m);
        }
        gz_staged += need;
    }
    gz_files.push_back(GzFile{dst, total, gz_members.size() - (total > 0 ? members.size() : 0), total > 0 ? members.size() : 0});
    file_open = false;
    if (gz_members.size() >= gz_wave_members()) return launch_gz();
    return EXON_GPU_OK;
}

// Indexed scan (IndexedVCFOpener::open, exon/exon-core/src/datasources/vcf/file_opener/indexed_file_opener.rs:53-214): only
// the members a tabix chunk covers are inflated.  `data` holds file bytes [file_off, file_off + len) and starts at a
// member boundary at or before the chunk's first member; it must reach through the member at the chunk's end.
int VcfStream::feed_gzip_chunk(const uint8_t *data, size_t len, uint64_t file_off, uint64_t vstart, uint64_t vend) {
    if (cur_run_open && tail_len > 0) return fail(EXON_GPU_ERR_STATE, "feed_bgzf_chunk: the previous plain-text range ended mid-line");
    const uint64_t c0 = vstart >> 16, c1 = vend >> 16;
    const uint32_t u0 = (uint32_t)(vstart & 0xFFFFu), u1 = (uint32_t)(vend & 0xFFFFu);
    if (vend < vstart || c0 < file_off) return fail(EXON_GPU_ERR_ARG, "feed_bgzf_chunk: the chunk does not lie inside the given byte range");
    std::vector<BgzfMember> members;
    uint64_t total = 0, end_pos = 0;
    bool saw_end = false;
    size_t p = 0;
    while (p < len) {
        BgzfMember m;
        size_t next = 0;
        if (int rc = bgzf_parse_member(data, len, p, &m, &next)) return rc;
        const uint64_t at = file_off + p;
        if (at >= c0 && at <= c1) {
            if (at == c1) {
                saw_end = true;
                end_pos = total + u1;  // the chunk ends u1 bytes into this member
                if (u1 == 0) break;     // nothing of it is needed
                if (u1 > m.isize) return fail(EXON_GPU_ERR_ARG, "feed_bgzf_chunk: chunk end beyond its member");
            }
            m.out_addr = total;
            total += m.isize;
            members.push_back(m);
            if (at == c1) break;
        } else if (at > c1) {
            break;
        }
        p = next;
    }
    if (!saw_end) {
        if (c0 == c1 || p >= len) end_pos = total;  // the reference reads to the end of the data in this case (indexed_file_opener.rs:113-117)
        else return fail(EXON_GPU_ERR_ARG, "feed_bgzf_chunk: no member starts at the chunk's end offset");
    }
    if (members.empty() || end_pos <= u0) {
        if (!runs.empty()) file_marks.push_back(FileMark{runs.size() - 1, runs.back().len});
        return EXON_GPU_OK;
    }
    if (u0 >= members[0].isize && members[0].isize) return fail(EXON_GPU_ERR_ARG, "feed_bgzf_chunk: chunk start beyond its member");
    // stage the compressed bytes of the selected members (one contiguous range of `data`)
    const size_t lo = (size_t)(members.front().in_off >= 18 ? members.front().in_off - 18 : 0), hi = (size_t)(members.back().in_off + members.back().in_len + 8);
    const size_t need = ((hi - lo + 16 + 255) & ~(size_t)255);
    uint8_t *dz = nullptr;
    if (int rc = gz_stage(need, &dz)) return rc;
    CUDA_TRY(cudaMemcpyAsync(dz, data + lo, hi - lo, cudaMemcpyHostToDevice, gz_copy_stream));
    if (blocks.empty() || blocks.back().used + total + 1 > blocks.back().cap) {
        DevBlock nb;
        if (int rc = ctx->get_block((size_t)total + 1, &nb)) return rc;
        blocks.push_back(nb);
    }
    DevBlock &b = blocks.back();
    uint8_t *dst = b.ptr + b.used;
    b.used = std::min(b.cap, b.used + (((size_t)total + 1 + 15) & ~(size_t)15));
    cur_run_open = false;
    tail_len = 0;
    const size_t first = gz_members.size();
    for (BgzfMember m : members) {
        m.in_off = m.in_off - lo + gz_staged;
        m.out_addr = (uint64_t)reinterpret_cast<uintptr_t>(dst) + m.out_addr;
        gz_members.push_back(m);
    }
    gz_staged += need;
    GzFile f{dst, total, first, members.size()};
    f.range_lo = (int64_t)u0;
    f.range_hi = (int64_t)end_pos;
    gz_files.push_back(f);
    if (gz_members.size() >= gz_wave_members()) return launch_gz();
    return EXON_GPU_OK;
}

// Staging space for `need` compressed bytes.  When the current buffer is full its group is launched and filling moves to
// the other buffer -- after the copy stream has been told to wait for the inflate that last read that buffer.
constexpr size_t kGzStageMax = (size_t)4 << 30;  // a staging buffer stops growing here (two of them exist)

// members one launch of the decoder holds in flight at full occupancy: 12 warps of 32 lanes per SM (inflate.cu, 7/6-bit tables)
size_t VcfStream::gz_wave_members() const { return (size_t)ctx->sm_count * 384; }

int VcfStream::gz_stage(size_t need, uint8_t **out) {
    if (!gz_copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&gz_copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&gz_copied_ev, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreateWithFlags(&gz_done_ev[i], cudaEventDisableTiming));
    }
    if (gz_staged + need > d_gz_buf_cap[gz_cur] && gz_staged > 0) {
        // The decoder runs one lane per member and every member is a serial chain of the same length, so a launch takes about
        // as long for 10 000 members as for a full wave of them (inflate.cu): a group is worth launching only when it is a
        // wave, and until then a full staging buffer grows instead (the staged bytes move with a device copy on the copy
        // stream; this happens a few times in a stream's life, the buffers are kept).
        if (gz_members.size() < gz_wave_members() && d_gz_buf_cap[gz_cur] < kGzStageMax) {
            const size_t cap = std::min(std::max(d_gz_buf_cap[gz_cur] * 2, gz_staged + need), std::max(kGzStageMax, gz_staged + need));
            void *nb = nullptr;
            CUDA_TRY(cudaMalloc(&nb, cap));
            cudaError_t e = cudaMemcpyAsync(nb, d_gz_buf[gz_cur], gz_staged, cudaMemcpyDeviceToDevice, gz_copy_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(gz_copy_stream);
            if (e != cudaSuccess) {
                cudaFree(nb);
                return fail(EXON_GPU_ERR_CUDA, "feed_gzip: growing the staging buffer: %s", cudaGetErrorString(e));
            }
            CUDA_TRY(cudaFree(d_gz_buf[gz_cur]));
            d_gz_buf[gz_cur] = nb;
            d_gz_buf_cap[gz_cur] = cap;
        } else {
            if (int rc = launch_gz()) return rc;  // moves on to the other buffer
        }
    }
    if (need > d_gz_buf_cap[gz_cur]) {
        // grow: nothing is staged in this buffer now; an earlier inflate may still be reading it
        if (d_gz_buf[gz_cur]) {
            if (gz_done_armed[gz_cur]) CUDA_TRY(cudaEventSynchronize(gz_done_ev[gz_cur]));
            CUDA_TRY(cudaFree(d_gz_buf[gz_cur]));
            d_gz_buf[gz_cur] = nullptr;
            d_gz_buf_cap[gz_cur] = 0;
            gz_done_armed[gz_cur] = false;
        }
        const size_t cap = std::max(need * 2, (size_t)256 << 20);
        CUDA_TRY(cudaMalloc(&d_gz_buf[gz_cur], cap));
        d_gz_buf_cap[gz_cur] = cap;
    }
    if (gz_staged == 0 && gz_done_armed[gz_cur]) {
        CUDA_TRY(cudaStreamWaitEvent(gz_copy_stream, gz_done_ev[gz_cur], 0));
        gz_done_armed[gz_cur] = false;
    }
    *out = (uint8_t *)d_gz_buf[gz_cur] + gz_staged;
    return EXON_GPU_OK;
}

void VcfStream::gz_teardown() {
    if (gz_copy_stream) cudaStreamSynchronize(gz_copy_stream);
    for (auto &g : gz_inflight) cudaFree(g.d_tab);
    gz_inflight.clear();
    for (int i = 0; i < 2; ++i) {
        cudaFree(d_gz_buf[i]);
        d_gz_buf[i] = nullptr;
        if (gz_done_ev[i]) cudaEventDestroy(gz_done_ev[i]);
        gz_done_ev[i] = nullptr;
    }
    if (gz_copied_ev) cudaEventDestroy(gz_copied_ev);
    gz_copied_ev = nullptr;
    if (gz_copy_stream) cudaStreamDestroy(gz_copy_stream);
    gz_copy_stream = nullptr;
}

// Enqueues the inflate of every pending member as one launch on the context's stream, behind the arrival of the group's
// compressed bytes on the copy stream.  No host synchronisation: the next group's bytes start to travel while this one
// inflates.  harvest_gz() checks and frames the launched groups.
int VcfStream::launch_gz() {
    if (gz_files.empty()) return EXON_GPU_OK;
    cudaStream_t st = ctx->stream;
    GzGroup g;
    g.files.swap(gz_files);
    g.members.swap(gz_members);
    const int buf = gz_cur;
    if (!g.members.empty()) {
        g.tab_bytes = (g.members.size() * sizeof(BgzfMember) + 255) & ~(size_t)255;
        CUDA_TRY(cudaMallocAsync(&g.d_tab, g.tab_bytes + 256, st));
        uint8_t *dt = (uint8_t *)g.d_tab;
        const size_t tok_units = bgzf_assign_tokens(g.members.data(), g.members.size());
        size_t comp_bytes = 0;
        for (const BgzfMember &m : g.members) comp_bytes += m.in_len;
        CUDA_TRY(cudaMemcpyAsync(dt, g.members.data(), g.members.size() * sizeof(BgzfMember), cudaMemcpyHostToDevice, st));
        const int init_flags[2] = {0, 0x7FFFFFFF};
        CUDA_TRY(cudaMemcpyAsync(dt + g.tab_bytes, init_flags, sizeof(init_flags), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(gz_copied_ev, gz_copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, gz_copied_ev, 0));
        std::lock_guard<std::recursive_mutex> work(ctx->work_mu);  // the token scratch is a context-wide area
        if (int rc = bgzf_inflate_launch(ctx, (const uint8_t *)d_gz_buf[buf], (const BgzfMember *)dt, (int)g.members.size(), (uint32_t *)(dt + g.tab_bytes),
                                         tok_units, comp_bytes)) {
            cudaFreeAsync(g.d_tab, st);
            return rc;
        }
        CUDA_TRY(cudaEventRecord(gz_done_ev[buf], st));
        gz_done_armed[buf] = true;
        gz_cur ^= 1;
    }
    gz_staged = 0;
    gz_inflight.push_back(std::move(g));
    return EXON_GPU_OK;
}

// Launches what is pending, then waits for every launched group, checks it and frames its files in feed order exactly like
// device-resident ranges (header skipped on a host copy of each file's first bytes, last record normalised to end in '\n').
int VcfStream::flush_gz() {
    if (int rc = launch_gz()) return rc;
    return harvest_gz();
}

int VcfStream::harvest_gz() {
    if (gz_inflight.empty()) return EXON_GPU_OK;
    std::lock_guard<std::recursive_mutex> work(ctx->work_mu);  // the pinned probe area (h_scratch) is context-wide
    cudaStream_t st = ctx->stream;
    std::vector<GzGroup> groups;
    groups.swap(gz_inflight);
    struct FreeTabs {
        std::vector<GzGroup> &g;
        cudaStream_t st;
        ~FreeTabs() {
            for (auto &x : g)
                if (x.d_tab) cudaFreeAsync(x.d_tab, st);
        }
    } free_tabs{groups, st};
    constexpr size_t kProbe = 64 << 10;
    // feed order across groups
    std::vector<GzFile> files;
    std::vector<BgzfMember> members;
    for (GzGroup &g : groups) {
        for (GzFile f : g.files) {
            f.first_member += members.size();
            files.push_back(f);
        }
        members.insert(members.end(), g.members.begin(), g.members.end());
    }
    {
        // flags of every group + per file: first bytes (header probe) and the last byte, all in one round trip
        if (int rc = ctx->ensure_scratch(0, 64 + 16 * groups.size() + files.size() * (kProbe + 16))) return rc;
        uint8_t *h = (uint8_t *)ctx->h_scratch;
        uint8_t *h_files = h + 64 + 16 * groups.size();
        for (size_t gi = 0; gi < groups.size(); ++gi)
            if (groups[gi].d_tab) CUDA_TRY(cudaMemcpyAsync(h + 64 + 16 * gi, (uint8_t *)groups[gi].d_tab + groups[gi].tab_bytes, 8, cudaMemcpyDeviceToHost, st));
        for (size_t i = 0; i < files.size(); ++i) {
            if (!files[i].total) continue;
            uint8_t *slot = h_files + i * (kProbe + 16);
            CUDA_TRY(cudaMemcpyAsync(slot, files[i].dst, (size_t)std::min<uint64_t>(files[i].total, kProbe), cudaMemcpyDeviceToHost, st));
            const uint64_t last_at = files[i].range_lo >= 0 ? (uint64_t)files[i].range_hi - 1 : files[i].total - 1;
            CUDA_TRY(cudaMemcpyAsync(slot + kProbe, files[i].dst + last_at, 1, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        size_t member0 = 0;
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            const uint32_t *fl = reinterpret_cast<const uint32_t *>(h + 64 + 16 * gi);
            if (groups[gi].d_tab && fl[0])
                return fail(EXON_GPU_ERR_PARSE, "bgzf: member %d does not inflate:%s%s", (int)(member0 + fl[1]), (fl[0] & 1u) ? " invalid DEFLATE data;" : "",
                            (fl[0] & 2u) ? " size differs from ISIZE;" : "");
            member0 += groups[gi].members.size();
        }
    }
    const uint8_t *h_files = (const uint8_t *)ctx->h_scratch + 64 + 16 * groups.size();
    // the probe area is reused by frame_device_range's slow path: copy what we need out of it first
    struct Probe { int64_t body_off; int last; };
    std::vector<Probe> probes(files.size(), Probe{-1, -1});
    for (size_t i = 0; i < files.size(); ++i) {
        if (!files[i].total) continue;
        const uint8_t *slot = h_files + i * (kProbe + 16);
        probes[i].body_off = probe_body_offset(slot, (size_t)std::min<uint64_t>(files[i].total, kProbe), files[i].total <= kProbe);
        probes[i].last = slot[kProbe];
    }
    if (fmt == kFmtBam) {
        for (size_t i = 0; i < files.size(); ++i) {
            const GzFile &f = files[i];
            if (!f.total) continue;
            // the probe slot is read before bam_frame_file may overwrite the pinned area: copy it
            const uint8_t *slot = h_files + i * (kProbe + 16);
            std::vector<uint8_t> head(slot, slot + (size_t)std::min<uint64_t>(f.total, kProbe));
            if (int rc = bam_frame_file(f.dst, f.total, head.data(), head.size(), members.data() + f.first_member, f.n_members)) return rc;
        }
        return EXON_GPU_OK;
    }
    for (size_t i = 0; i < files.size(); ++i) {
        const GzFile &f = files[i];
        if (!f.total) {  // an empty file (no members, or only empty members such as the BGZF EOF marker)
            if (!runs.empty()) file_marks.push_back(FileMark{runs.size() - 1, runs.back().len});
            continue;
        }
        if (f.range_lo >= 0) {
            // a tabix chunk: whole records from range_lo to range_hi of the inflated members, no header inside
            if (probes[i].last != '\n') return fail(EXON_GPU_ERR_PARSE, "feed_bgzf_chunk: the chunk does not end at a record boundary");
            hdr = kBody;
            if (int rc = frame_device_range(f.dst + f.range_lo, (size_t)(f.range_hi - f.range_lo), true, 0, '\n')) return rc;
            continue;
        }
        uint64_t n = f.total;
        if (probes[i].last != '\n') {
            static const uint8_t nl = '\n';
            CUDA_TRY(cudaMemcpyAsync(f.dst + n, &nl, 1, cudaMemcpyHostToDevice, st));
            n += 1;
        }
        hdr = fmt == kFmtVcf ? kAtLineStart : kBody;
        const int64_t before = body_bytes;
        if (int rc = frame_device_range(f.dst, (size_t)n, true, probes[i].body_off, '\n')) return rc;
        if (n != f.total && body_bytes > before) body_bytes -= 1;  // the added '\n' is not a fed byte
    }
    return EXON_GPU_OK;
}

}  // namespace exon

// Whole-file inflate into caller memory (host or device): used by hosts that need the uncompressed bytes themselves
// (e.g. a BAM header) and by the byte-exact parity tests against zlib.
extern "C" int exon_gpu_gzip_inflate(exon_gpu_ctx *c, const uint8_t *data, size_t len, uint8_t *out, size_t out_cap, int out_is_device,
                                     size_t *out_len) {
    using namespace exon;
    if (!c || (!data && len) || !out_len) return fail(EXON_GPU_ERR_ARG, "gzip_inflate: NULL argument");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    std::vector<BgzfMember> members;
    uint64_t total = 0;
    if (len)
        if (int rc = bgzf_walk(data, len, members, &total)) return rc;
    *out_len = (size_t)total;
    if (total == 0) return EXON_GPU_OK;
    if (!out || out_cap < total) return fail(EXON_GPU_ERR_ARG, "gzip_inflate: output buffer too small (%llu bytes needed)", (unsigned long long)total);
    std::lock_guard<std::recursive_mutex> work(c->work_mu);
    const size_t o_tab = (len + 16 + 255) & ~(size_t)255;
    const size_t o_flags = o_tab + ((members.size() * sizeof(BgzfMember) + 255) & ~(size_t)255);
    const size_t o_out = o_flags + 256;
    if (int rc = c->ensure_scratch(o_out + (out_is_device ? 0 : (size_t)total + 16), 64)) return rc;
    uint8_t *scr = (uint8_t *)c->scratch;
    uint8_t *d_out = out_is_device ? out : scr + o_out;
    cudaStream_t st = c->stream;
    for (BgzfMember &m : members) m.out_addr += (uint64_t)reinterpret_cast<uintptr_t>(d_out);
    const size_t tok_units = bgzf_assign_tokens(members.data(), members.size());
    CUDA_TRY(cudaMemcpyAsync(scr, data, len, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_tab, members.data(), members.size() * sizeof(BgzfMember), cudaMemcpyHostToDevice, st));
    const int init_flags[2] = {0, 0x7FFFFFFF};
    CUDA_TRY(cudaMemcpyAsync(scr + o_flags, init_flags, sizeof(init_flags), cudaMemcpyHostToDevice, st));
    CUDA_TRY(c->timed_begin(st));
    if (int rc = bgzf_inflate_launch(c, scr, (const BgzfMember *)(scr + o_tab), (int)members.size(), (uint32_t *)(scr + o_flags), tok_units, len)) return rc;
    CUDA_TRY(c->timed_end(st));
    CUDA_TRY(cudaMemcpyAsync(c->h_scratch, scr + o_flags, 8, cudaMemcpyDeviceToHost, st));
    if (!out_is_device) CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint32_t *fl = (const uint32_t *)c->h_scratch;
    if (fl[0])
        return fail(EXON_GPU_ERR_PARSE, "gzip: member %d does not inflate:%s%s", (int)fl[1], (fl[0] & 1u) ? " invalid DEFLATE data;" : "",
                    (fl[0] & 2u) ? " size differs from ISIZE;" : "");
    return EXON_GPU_OK;
}

extern "C" int exon_gpu_stream_feed_gzip(exon_gpu_stream *s, const uint8_t *data, size_t len, int is_last) {
    if (!s || (!data && len)) return exon::fail(EXON_GPU_ERR_ARG, "feed_gzip: NULL argument");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return exon::fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (s->drained) return exon::fail(EXON_GPU_ERR_STATE, "feed_gzip: the stream has already produced batches");
    return s->feed_gzip(data, len, is_last != 0);
}
