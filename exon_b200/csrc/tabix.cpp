// tabix.cpp -- tabix (.tbi) region query: which BGZF chunks of a .vcf.gz can hold records of a region.
//
// Replaces `noodles::tabix::Reader::read_index` + `index.query(id, region.interval())` at
// exon/exon-core/src/datasources/indexed_file/indexed_bgzf_file.rs:52-83 (noodles-tabix 0.47 / noodles-csi 0.41, not
// vendored; the format is the tabix specification, the query is the UCSC binning scheme with the linear index).
// Host-side planning only: the .tbi (itself BGZF) is inflated on the device through exon_gpu_gzip_inflate, the few KB of
// bins are walked here, and the chunks select the members exon_gpu_stream_feed_bgzf_chunk inflates and scans.
// Known answer (indexed_bgzf_file.rs:167-187): chr1:1-3388930 on bigger-index/test.vcf.gz -> one chunk,
// 621346816 .. 3014113427456.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "internal.h"

using namespace exon;

namespace {

struct Cursor {
    const uint8_t *p, *end;
    bool ok = true;
    template <class T>
    T get() {
        T v{};
        if ((size_t)(end - p) < sizeof(T)) { ok = false; return v; }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
};

// bins that may overlap [beg, end) (0-based, half open), tabix spec section "reg2bins" (min_shift 14, depth 5)
void reg2bins(int64_t beg, int64_t end, std::vector<uint32_t> &out) {
    --end;
    out.push_back(0);
    for (int64_t k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) out.push_back((uint32_t)k);
    for (int64_t k = 4681 + (beg >> 14); k <= 4681 + (end >> 14); ++k) out.push_back((uint32_t)k);
}

}  // namespace

extern "C" int exon_gpu_tabix_query(exon_gpu_ctx *c, const uint8_t *tbi, size_t len, const exon_gpu_region *region, exon_gpu_chunk *out,
                                    int32_t cap, int32_t *n_chunks) {
    if (!c || !tbi || !region || !n_chunks || !region->has_chrom || !region->chrom) return fail(EXON_GPU_ERR_ARG, "tabix_query: bad argument");
    *n_chunks = 0;
    size_t raw_len = 0;
    if (int rc = exon_gpu_gzip_inflate(c, tbi, len, nullptr, 0, 0, &raw_len)) {
        if (raw_len == 0) return rc;  // not a gzip stream at all
    }
    std::vector<uint8_t> raw(raw_len + 16);
    if (int rc = exon_gpu_gzip_inflate(c, tbi, len, raw.data(), raw.size(), 0, &raw_len)) return rc;
    Cursor cur{raw.data(), raw.data() + raw_len};
    if (raw_len < 36 || memcmp(cur.p, "TBI\1", 4) != 0) return fail(EXON_GPU_ERR_PARSE, "tabix: bad magic");
    cur.p += 4;
    const int32_t n_ref = cur.get<int32_t>();
    for (int i = 0; i < 6; ++i) cur.get<int32_t>();  // format, col_seq, col_beg, col_end, meta, skip
    const int32_t l_nm = cur.get<int32_t>();
    if (!cur.ok || n_ref < 0 || l_nm < 0 || (size_t)(cur.end - cur.p) < (size_t)l_nm) return fail(EXON_GPU_ERR_PARSE, "tabix: truncated header");
    // reference names: NUL-terminated, concatenated
    int32_t id = -1;
    {
        const char *nm = reinterpret_cast<const char *>(cur.p), *nm_end = nm + l_nm;
        for (int32_t i = 0; i < n_ref && nm < nm_end; ++i) {
            const size_t n = strnlen(nm, (size_t)(nm_end - nm));
            if ((int32_t)n == region->chrom_len && memcmp(nm, region->chrom, n) == 0) id = i;
            nm += n + 1;
        }
        cur.p += l_nm;
    }
    if (id < 0) return EXON_GPU_OK;  // the contig is not in the file: no chunks (indexed_bgzf_file.rs:79-83)
    // interval: 1-based inclusive -> 0-based half open, clamped to what the binning scheme addresses (2^29)
    const int64_t kMax = (int64_t)1 << 29;
    int64_t beg = region->has_interval ? std::max<int64_t>(region->lo, 1) - 1 : 0;
    int64_t end = region->has_interval ? std::min<int64_t>(region->hi, kMax) : kMax;
    if (beg >= end) return EXON_GPU_OK;
    std::vector<uint32_t> want;
    reg2bins(beg, end, want);
    std::sort(want.begin(), want.end());
    std::vector<exon_gpu_chunk> chunks;
    uint64_t min_off = 0;
    for (int32_t r = 0; r < n_ref; ++r) {
        const int32_t n_bin = cur.get<int32_t>();
        if (!cur.ok || n_bin < 0) return fail(EXON_GPU_ERR_PARSE, "tabix: truncated bins");
        for (int32_t b = 0; b < n_bin; ++b) {
            const uint32_t bin = cur.get<uint32_t>();
            const int32_t n_chunk = cur.get<int32_t>();
            if (!cur.ok || n_chunk < 0) return fail(EXON_GPU_ERR_PARSE, "tabix: truncated chunks");
            const bool take = r == id && bin != 37450u && std::binary_search(want.begin(), want.end(), bin);  // 37450: metadata pseudo-bin
            for (int32_t k = 0; k < n_chunk; ++k) {
                exon_gpu_chunk ch;
                ch.start = cur.get<uint64_t>();
                ch.end = cur.get<uint64_t>();
                if (take) chunks.push_back(ch);
            }
        }
        const int32_t n_intv = cur.get<int32_t>();
        if (!cur.ok || n_intv < 0 || (size_t)(cur.end - cur.p) < (size_t)n_intv * 8) return fail(EXON_GPU_ERR_PARSE, "tabix: truncated linear index");
        if (r == id) {
            const int64_t w = beg >> 14;
            if (n_intv > 0) {
                uint64_t v;
                memcpy(&v, cur.p + 8 * (size_t)std::min<int64_t>(w, n_intv - 1), 8);
                min_off = v;
            }
        }
        cur.p += (size_t)n_intv * 8;
    }
    // optimize_chunks: drop what ends before the linear-index offset, sort, merge overlapping / touching chunks
    chunks.erase(std::remove_if(chunks.begin(), chunks.end(), [&](const exon_gpu_chunk &ch) { return ch.end <= min_off; }), chunks.end());
    std::sort(chunks.begin(), chunks.end(), [](const exon_gpu_chunk &a, const exon_gpu_chunk &b) { return a.start < b.start; });
    std::vector<exon_gpu_chunk> merged;
    for (const exon_gpu_chunk &ch : chunks) {
        if (!merged.empty() && ch.start <= merged.back().end) merged.back().end = std::max(merged.back().end, ch.end);
        else merged.push_back(ch);
    }
    *n_chunks = (int32_t)merged.size();
    if (out) {
        if (cap < (int32_t)merged.size()) return fail(EXON_GPU_ERR_ARG, "tabix_query: %d chunks, room for %d", (int)merged.size(), cap);
        for (size_t i = 0; i < merged.size(); ++i) out[i] = merged[i];
    }
    return EXON_GPU_OK;
}

extern "C" int exon_gpu_stream_feed_bgzf_chunk(exon_gpu_stream *s, const uint8_t *data, size_t len, uint64_t file_offset, const exon_gpu_chunk *chunk) {
    if (!s || !data || !chunk) return fail(EXON_GPU_ERR_ARG, "feed_bgzf_chunk: NULL argument");
    if (s->fmt != kFmtVcf) return fail(EXON_GPU_ERR_ARG, "feed_bgzf_chunk: not a VCF stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (s->drained) return fail(EXON_GPU_ERR_STATE, "feed_bgzf_chunk: the stream has already produced batches");
    return s->feed_gzip_chunk(data, len, file_offset, chunk->start, chunk->end);
}
