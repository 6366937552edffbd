// fasta_scan.cu -- FASTA text -> COUNT(*) (BASELINE.json configs[0]: SELECT COUNT(*) FROM fasta_scan('small.fa')).
//
// Replaces, for the row count, FASTAScan::execute / BatchReader::read_batch (exon/exon-core/src/datasources/fasta/
// scanner.rs, exon/exon-fasta/src/batch_reader.rs: noodles-fasta `read_definition` + `read_sequence` per record) followed
// by AggregateExec count(*): a record is a definition line that starts with '>' and the sequence lines up to the next
// one, so COUNT(*) is the number of '>' at line starts.  A non-empty file that does not begin with '>' is an error
// (noodles: "invalid definition" on the first read).  One pass of the warp-private TMA tile pipeline (tile_ring.cuh):
// '\n' and '>' masks per 16-byte chunk, popc((newline << 1) & greater-than); the byte after a chunk's last '\n' is
// fetched from the next chunk.  HBM-bound by construction (each byte read once, 8 bytes out).
#include <cstring>

#include "common.cuh"
#include "internal.h"
#include "tile_ring.cuh"

namespace exon {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? EXON_GPU_ERR_OOM : EXON_GPU_ERR_CUDA, "%s: %s", \
                        #expr, cudaGetErrorString(_e));                                                  \
    } while (0)

namespace {

using FaRing = TileRing<4096, 3, 8, 16, 48, 16>;
constexpr int kFaU = FaRing::TILE / 512;
constexpr uint32_t kGT4 = 0x3E3E3E3Eu;

struct FastaArgs {
    const ScanSeg *segs;
    int64_t n_tiles;
    const uint8_t *file_start;    // per segment: 1 = the segment starts a file (its first byte must be '>')
    unsigned long long *out;      // [0] records
    uint32_t *flags;              // bit 0: a file does not start with '>'
};

__global__ void __launch_bounds__(FaRing::WARPS * 32, 2) fasta_count_kernel(const __grid_constant__ FastaArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FaRing ring;
    ring.init(smem_raw, a.segs, a.n_tiles);
    const int lane = ring.lane;
    uint32_t cnt = 0, err = 0;
#pragma unroll 1
    for (int64_t T = ring.first_tile(); T < a.n_tiles; T += ring.nw) {
        const FaRing::View v = ring.acquire();
        if (v.first && v.hi > v.seg_lo && lane == 0) {  // the segment's first line has no '\n' before it
            const bool gt = view_byte(v, v.seg_lo) == '>';
            cnt += gt;
            if (!gt && a.file_start[v.seg]) err |= 1u;
        }
#pragma unroll
        for (int u = 0; u < kFaU; ++u) {
            const int c0 = (u * 32 + lane) * 16;
            if (!(v.interior || c0 < v.sm_hi)) continue;
            const uint4 w = lds128(v.sa + (uint32_t)c0);
            uint32_t nl = newline_mask16(w);
            if (!v.interior) nl = clip_mask16(nl, c0, v.seg_lo, v.hi);
            if (!nl) continue;
            const uint32_t gt = pack_flags16(zero_bytes_exact(w.x ^ kGT4), zero_bytes_exact(w.y ^ kGT4), zero_bytes_exact(w.z ^ kGT4),
                                             zero_bytes_exact(w.w ^ kGT4));
            cnt += (uint32_t)__popc((nl << 1) & gt & 0xFFFFu);
            if ((nl >> 15) & 1u) cnt += lds8(v.sa + (uint32_t)(c0 + 16)) == '>';  // the line starts in the next chunk
        }
        ring.release(T);
    }
    cnt = warp_sum(cnt);
    err = __reduce_or_sync(0xFFFFFFFFu, err);
    if (lane == 0) {
        if (cnt) atomicAdd(a.out, (unsigned long long)cnt);
        if (err) atomicOr(a.flags, err);
    }
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int fasta_rows(VcfStream *s, int64_t *out_rows) {
    if (int rc = s->flush_gz()) return rc;
    Ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    std::lock_guard<std::recursive_mutex> work(ctx->work_mu);
    *out_rows = 0;
    std::vector<Piece> pieces;
    s->cut_pieces(pieces);
    if (pieces.empty()) return EXON_GPU_OK;
    std::vector<ScanSeg> h_segs;
    std::vector<uint8_t> starts;
    int64_t n_tiles = 0;
    for (const Piece &p : pieces) {
        ScanSeg sg;
        sg.skip = (int32_t)((uintptr_t)p.base & 15);
        sg.base = p.base - sg.skip;
        sg.len = p.len;
        sg.tile0 = n_tiles;
        sg.pad_ = 0;
        n_tiles += (sg.skip + p.len + FaRing::TILE - 1) / FaRing::TILE;
        starts.push_back(p.starts_file || h_segs.empty() ? 1 : 0);
        h_segs.push_back(sg);
    }
    ScanSeg sentinel;
    memset(&sentinel, 0, sizeof(sentinel));
    sentinel.tile0 = n_tiles;
    h_segs.push_back(sentinel);
    const size_t o_segs = 0, o_starts = al256(h_segs.size() * sizeof(ScanSeg)), o_out = o_starts + al256(starts.size());
    if (int rc = ctx->ensure_scratch(o_out + 256, 64)) return rc;
    uint8_t *scr = (uint8_t *)ctx->scratch;
    CUDA_TRY(cudaMemcpyAsync(scr + o_segs, h_segs.data(), h_segs.size() * sizeof(ScanSeg), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(scr + o_starts, starts.data(), starts.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(scr + o_out, 0, 64, st));
    FastaArgs a;
    a.segs = (const ScanSeg *)(scr + o_segs);
    a.n_tiles = n_tiles;
    a.file_start = scr + o_starts;
    a.out = (unsigned long long *)(scr + o_out);
    a.flags = (uint32_t *)(scr + o_out + 16);
    static int occ = 0;
    if (!occ) {
        CUDA_TRY(cudaFuncSetAttribute(fasta_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FaRing::smem_bytes));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fasta_count_kernel, FaRing::WARPS * 32, FaRing::smem_bytes));
        if (occ < 1) occ = 1;
    }
    int64_t grid = std::min<int64_t>((int64_t)occ * ctx->sm_count, (n_tiles + FaRing::WARPS - 1) / FaRing::WARPS);
    if (grid < 1) grid = 1;
    CUDA_TRY(ctx->timed_begin(st));
    fasta_count_kernel<<<(unsigned)grid, FaRing::WARPS * 32, FaRing::smem_bytes, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(ctx->timed_end(st));
    ctx->launches.fetch_add(1);
    CUDA_TRY(cudaMemcpyAsync(s->h_res, a.out, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((uint32_t)s->h_res[2]) return fail(EXON_GPU_ERR_PARSE, "malformed FASTA: a file does not start with a '>' definition line");
    *out_rows = (int64_t)s->h_res[0];
    return EXON_GPU_OK;
}

}  // namespace exon

using namespace exon;

extern "C" {

int exon_gpu_fasta_open(exon_gpu_ctx *c, exon_gpu_stream **out) {
    if (!c || !out) return fail(EXON_GPU_ERR_ARG, "fasta_open: NULL argument");
    exon_gpu_vcf_opts vo;
    memset(&vo, 0, sizeof(vo));
    if (int rc = exon_gpu_vcf_open(c, &vo, out)) return rc;
    (*out)->fmt = kFmtFasta;
    (*out)->hdr = VcfStream::kBody;
    return EXON_GPU_OK;
}

int exon_gpu_fasta_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last) {
    if (!s || s->fmt != kFmtFasta) return fail(EXON_GPU_ERR_ARG, "fasta_feed: not a FASTA stream");
    return exon_gpu_vcf_feed(s, text, len, is_device_ptr, is_last);
}

int exon_gpu_fasta_rows(exon_gpu_stream *s, int64_t *out_rows) {
    if (!s || !out_rows) return fail(EXON_GPU_ERR_ARG, "fasta_rows: NULL argument");
    if (s->fmt != kFmtFasta) return fail(EXON_GPU_ERR_ARG, "fasta_rows: not a FASTA stream");
    cudaError_t e = cudaSetDevice(s->ctx->device);
    if (e != cudaSuccess) return fail(EXON_GPU_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return fasta_rows(s, out_rows);
}

}  // extern "C"
