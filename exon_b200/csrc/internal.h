// internal.h -- library-internal state behind the opaque handles of include/exon_gpu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/exon_gpu.h"
#include "vcf_scan.cuh"

namespace exon {

int fail(int code, const char *fmt, ...);

constexpr size_t kArenaBlock = (size_t)256 << 20;  // arena granularity (HBM is 180 GB; blocks are recycled)

struct DevBlock {
    uint8_t *ptr = nullptr;
    size_t cap = 0, used = 0;
};

struct Ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::atomic<int64_t> launches{0};
    // event pairs that bracket the most recent fused-scan launches (a ring: benches read the durations of a whole timed
    // region afterwards instead of synchronising on every step)
    static constexpr int kEvRing = 64;
    cudaEvent_t ev[kEvRing][2] = {};
    int64_t ev_count = 0;  // launches bracketed so far (guarded by mu)
    cudaError_t timed_begin(cudaStream_t st) {
        int slot;
        {
            std::lock_guard<std::mutex> g(mu);
            slot = (int)(ev_count % kEvRing);
        }
        return cudaEventRecord(ev[slot][0], st);
    }
    cudaError_t timed_end(cudaStream_t st) {
        std::lock_guard<std::mutex> g(mu);
        const cudaError_t e = cudaEventRecord(ev[ev_count % kEvRing][1], st);
        ++ev_count;
        return e;
    }
    std::mutex mu;
    std::recursive_mutex work_mu;  // serialises users of the context-wide scratch areas (scratch, h_scratch, scratch_b, inf_tokens); recursive: framing helpers that need it are reached both with and without it held
    std::vector<DevBlock> free_blocks;
    // scratch for filter_agg / allreduce
    void *scratch = nullptr;
    size_t scratch_cap = 0;
    void *h_scratch = nullptr;
    size_t h_scratch_cap = 0;
    void *scratch_b = nullptr;  // second device scratch area (K2 keeps per-row temporaries here while `scratch` holds tile tables)
    size_t scratch_b_cap = 0;
    void *inf_tokens = nullptr;  // inflate.cu: literal / token streams of the launch in flight (about 4/3 of its output)
    size_t inf_tokens_cap = 0;
    // NCCL (dlopen'ed lazily; see nccl.cu)
    void *nccl_comm = nullptr;
    int nccl_ranks = 0;
    void *peer_xchg = nullptr;  // scalar all-reduce over peer memory (nccl.cu), NULL: NCCL only

    int get_block(size_t min_bytes, DevBlock *out);
    void put_block(DevBlock b);
    int ensure_scratch(size_t dev_bytes, size_t host_bytes);
    int ensure_scratch_b(size_t dev_bytes);
};

void nccl_teardown(Ctx *c);
// in-place sum all-reduce of n int64 in device memory, enqueued on the context's stream
int nccl_allreduce_i64(Ctx *c, int64_t *device_buf, size_t n);
// Scalar exchange over peer memory (nccl.cu): fills the peer fields of a scan tail and consumes one sequence number.
// Returns false when the exchange is not active (single rank, setup failed somewhere, EXON_GPU_PEER_XCHG=0).
bool peer_xchg_arm(Ctx *c, ScanTail *tail);

// A region whose chrom bytes are owned (the caller's exon_gpu_region may go away).
struct OwnedRegion {
    std::string chrom;
    bool has_chrom = false, has_interval = false;
    int64_t lo = 1, hi = INT64_MAX;
    int assign(const exon_gpu_region *r);
};

// One resident run of body bytes.
struct Run {
    const uint8_t *base;  // first valid byte (any alignment)
    int64_t len;
    bool ends_with_newline;
};

// ---- BGZF inflate (bgzf.cu) ----
struct BgzfMember {
    uint64_t in_off;   // byte offset of the DEFLATE payload inside the compressed file
    uint32_t in_len;   // payload bytes
    uint32_t isize;    // uncompressed bytes (gzip trailer)
    uint64_t out_addr; // bgzf_walk: byte offset in the uncompressed stream; at launch: absolute device address of the member's output
    uint32_t tok_off;  // bgzf_assign_tokens: first 16-byte unit of the member's token area (inflate.cu)
    uint32_t pad_;
};

// 16-byte units of token scratch a member of `isize` output bytes needs (inflate.cu), a whole number of 32-byte sectors: a
// header sector, the literal bytes and four bytes per token -- a match is worth at least three output bytes, so 4/3 of ISIZE
// bounds the two together -- plus the padding of the last literal sector and token group.
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t bgzf_token_units(uint32_t isize) { return ((isize + isize / 3u + 160u) / 32u + 1u) * 2u; }

// One piece of a run that belongs to a single file (runs are cut at the recorded file ends).
struct Piece {
    const uint8_t *base;
    int64_t len;
    bool starts_file;
};

// INFO definitions of a VCF header (##INFO=<ID=..,Number=..,Type=..>): what the reference's builder takes from noodles' Header
struct InfoDefs {
    std::vector<std::string> ids;
    std::vector<uint8_t> types;   // 0 Integer, 1 Float, 2 Flag, 3 Character, 4 String
    std::vector<uint8_t> single;  // Number=1: the value is one element, not a ','-separated array
    bool set = false;
};

enum StreamFormat { kFmtVcf = 0, kFmtFastq = 1, kFmtBam = 2, kFmtMzml = 3, kFmtFasta = 4, kFmtGff = 5 };

struct VcfStream {
    int fmt = kFmtVcf;  // FASTQ streams share the arena / run / file-mark machinery; they have no header to skip
    Ctx *ctx = nullptr;
    int batch_rows = 8192;
    std::vector<int> projection;
    bool columns_on_device = false, strict = false, has_pushdown = false, drained = false;
    int variant = 0;
    OwnedRegion pushdown;
    InfoDefs info_defs, format_defs;  // exon_gpu_vcf_set_header: ##INFO / ##FORMAT lines

    // ---- file framing state (what read_header + the line reader keep between feeds) ----
    enum HdrState { kAtLineStart, kInHeaderLine, kBody };
    HdrState hdr = kAtLineStart;
    bool file_open = false;       // a file has received bytes and no is_last yet
    bool last_byte_newline = true;  // last body byte appended so far was '\n' (or nothing appended yet)

    // ---- arena ----
    std::vector<DevBlock> blocks;   // owned blocks, in fill order
    std::vector<Run> runs;          // resident body bytes in scan order (arena runs and zero-copy device ranges)
    bool cur_run_open = false;      // runs.back() lives at the tail of blocks.back() and may still grow
    int64_t tail_len = 0;           // bytes after the last '\n' of the open arena run (a partial line)
    int64_t body_bytes = 0;
    // end of each finished file as (run index, run length at that moment): record batches never span files
    // (FileStream opens one AsyncBatchStream per file, exon/exon-vcf/src/async_batch_stream.rs:70-109)
    struct FileMark { size_t run; int64_t len; };
    std::vector<FileMark> file_marks;

    // ---- eager (pushdown) scan bookkeeping ----
    int64_t eager_scanned = 0;      // bytes of the open run already covered by eager launches
    size_t eager_runs_done = 0;     // runs [0, eager_runs_done) fully covered
    bool last_eager = false;        // the most recent filter_count was answered by the eager accumulator

    // ---- device-side tables / results ----
    ScanSeg *d_segs = nullptr;
    size_t d_segs_cap = 0;
    TileDesc *d_tiles = nullptr;  // one descriptor per tile of the current table (built on the device from d_segs)
    size_t d_tiles_cap = 0;
    std::vector<ScanSeg> h_segs;
    bool segs_dirty = true;
    int seg_variant = -1;
    int64_t n_tiles = 0;
    // device words: [0..3] ScanAcc of plain queries, [4..7] ScanAcc of the pushdown (eager) scans, [8] local count of a
    // global query, [9] its all-reduced value (NCCL fallback), [10..15] spare
    unsigned long long *d_res = nullptr;
    unsigned long long *h_res = nullptr;  // MAPPED pinned: [0..7] the record the scan tail publishes (ScanHostWord), [8..15] copy targets
    unsigned long long *h_res_dev = nullptr;  // device address of h_res
    unsigned long long host_seq = 0;      // sequence number of the last published record

    // ---- column build (K2) state lives in vcf_columns.cu ----
    struct Columns *cols = nullptr;
    struct FqColumns *fq_cols = nullptr;  // FASTQ column store (fastq_scan.cu)

    // ---- compressed feeds (bgzf.cu): bytes of the current .gz file that arrived before its last range ----
    std::vector<uint8_t> gz_pending;
    // whole files whose compressed bytes are on their way to HBM but whose inflate has not been launched yet: members of
    // several files go into ONE launch so that thousands of members are in flight (bgzf.cu)
    struct GzFile {
        uint8_t *dst;
        uint64_t total;
        size_t first_member, n_members;
        int64_t range_lo = -1, range_hi = -1;  // >= 0: only bytes [range_lo, range_hi) of the inflated members are records (tabix chunk)
    };
    std::vector<GzFile> gz_files;
    std::vector<BgzfMember> gz_members;
    // Compressed bytes travel on their own copy stream into one of two staging buffers while the inflate of the previous
    // group of files runs on the context's stream out of the other (bgzf.cu).
    size_t gz_staged = 0;                        // bytes of the current staging buffer in use
    void *d_gz_buf[2] = {nullptr, nullptr};      // device staging for compressed bytes
    size_t d_gz_buf_cap[2] = {0, 0};
    int gz_cur = 0;                              // staging buffer being filled
    cudaStream_t gz_copy_stream = nullptr;
    cudaEvent_t gz_copied_ev = nullptr;          // copy stream: the current group's bytes have arrived
    cudaEvent_t gz_done_ev[2] = {nullptr, nullptr};  // context stream: the inflate that read staging buffer i has finished
    bool gz_done_armed[2] = {false, false};
    // groups whose inflate has been launched but not yet checked / framed
    struct GzGroup {
        std::vector<GzFile> files;
        std::vector<BgzfMember> members;
        void *d_tab = nullptr;  // member table | flags (stream-ordered pool)
        size_t tab_bytes = 0;
    };
    std::vector<GzGroup> gz_inflight;
    size_t gz_wave_members() const;
    int gz_stage(size_t need, uint8_t **out);  // `need` bytes of staging (launches the pending group / grows the buffer as needed)
    int launch_gz();                           // enqueue the inflate of the pending files; no host synchronisation
    int harvest_gz();                          // wait for every launched group, check it, frame its files in feed order
    void gz_teardown();
    int feed_gzip(const uint8_t *data, size_t len, bool is_last);
    // ---- BAM streams (bam.cu): one entry per inflated file ----
    struct BamFile {
        uint8_t *dst = nullptr;
        uint64_t total = 0, records_at = 0;
        std::vector<std::string> ref_names;
        std::vector<uint64_t> walk_starts;  // stream offsets where the speculative record walks begin (member starts)
    };
    std::vector<BamFile> bam_files;
    std::vector<std::string> bam_groups;  // reference names in group order (valid after a query)
    bool bam_tables_dirty = true;
    bool bam_serial_ok = false;           // the last verified pass was the serial one (one chain per file)
    struct BamColumns *bam_cols = nullptr;  // column store of exon_gpu_bam_next_batch (bam.cu)
    struct GffColumns *gff_cols = nullptr;  // column store of exon_gpu_gff_next_batch (gff_columns.cu)
    struct MzColumns *mz_cols = nullptr;    // column store of exon_gpu_mzml_next_batch (mzml_columns.cu)
    void *d_bam = nullptr;                // walk entries | per-file first entries | remap | exits | counts | misc
    size_t d_bam_cap = 0, bam_n_entries = 0, bam_n_firsts = 0;
    size_t bam_o_firsts = 0, bam_o_remap = 0, bam_o_exits = 0, bam_o_counts = 0, bam_o_misc = 0, bam_o_region = 0;
    int32_t bam_n_groups = 1;
    int bam_build_tables();
    int bam_frame_file(uint8_t *dst, uint64_t total, const uint8_t *probe, size_t probe_len, const BgzfMember *members, size_t n_members);
    int bam_filter_count(const exon_gpu_bam_pred *pred, int64_t *counts, int32_t cap, int32_t *n_groups, int64_t *total_rows);
    int flush_gz();
    int feed_gzip_chunk(const uint8_t *data, size_t len, uint64_t file_off, uint64_t vstart, uint64_t vend);
    int frame_device_range(const uint8_t *text, size_t len, bool is_last, int64_t known_body_off, int known_last_byte);
    int64_t probe_body_offset(const uint8_t *p, size_t n, bool whole_file) const;
    int feed_host(const uint8_t *text, size_t len, bool is_last);
    int feed_device(const uint8_t *text, size_t len, bool is_last);
    int append_host(const uint8_t *p, size_t n);
    int end_file();
    int filter_count(const exon_gpu_region *region, int64_t *device_out, int64_t *host_out);
    int filter_count_global(const exon_gpu_region *region, int64_t *out_local, int64_t *out_global);
    int launch_scan(const OwnedRegion &r, const TileDesc *d_table, int n_segs, int64_t tiles, ScanAcc *acc, const ScanTail &tail);
    int run_query(const exon_gpu_region *region, int64_t *device_out, bool want_host, bool global);
    int wait_published();
    int build_seg_table();
    void cut_pieces(std::vector<Piece> &out) const;
    int eager_scan(bool final_flush);
    void release_all();
};

// ---- K3 over many device-resident batches (filter_agg.cu) ----
enum ValueType { kValNone = 0, kValI64 = 1, kValF64 = 2, kValF32 = 3, kValI32 = 4 };
struct FaBatchDesc {
    const uint8_t *chrom_valid;  // validity bitmaps may be NULL
    const int32_t *chrom_offsets;
    const uint8_t *chrom_values;
    int64_t chrom_off;  // logical offset of row 0 in the chrom arrays
    const uint8_t *pos_valid;
    const int64_t *pos;
    int64_t pos_off;
    const uint8_t *val_valid;
    const void *val;
    int64_t val_off;
    int64_t n_rows;
};
struct FaCommon {
    int32_t has_chrom, lit_len;
    uint8_t lit[kMaxChrom + 1];
    int32_t has_pos;
    int64_t lo, hi;
    int32_t val_type, agg_kind;
    int32_t has_nulls;  // some batch carries a validity bitmap
};
int fa_common_from(const exon_gpu_pred *pred, const exon_gpu_agg *agg, FaCommon &k);
int filter_agg_multi_launch(Ctx *c, const FaBatchDesc *d_descs, int n_batches, int64_t max_rows, const FaCommon &k,
                            unsigned long long *d_out, bool timed);

int bgzf_walk(const uint8_t *data, size_t len, std::vector<BgzfMember> &out, uint64_t *total_out);
size_t bgzf_assign_tokens(BgzfMember *m, size_t n);
int bgzf_inflate_launch(Ctx *c, const uint8_t *d_comp, const BgzfMember *d_table, int n_members, uint32_t *d_flags, size_t token_units,
                        size_t comp_bytes);

// defined in bam.cu
void bam_columns_free(VcfStream *s);
int bam_next_batch(VcfStream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);

// defined in mzml_columns.cu
void mzml_columns_free(VcfStream *s);

// defined in gff_columns.cu
void gff_columns_free(VcfStream *s);
int gff_next_batch(VcfStream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);

// defined in gff_scan.cu
int gff_filter_count(VcfStream *s, const exon_gpu_region *region, int64_t *out_count, int64_t *out_rows);

// defined in fasta_scan.cu
int fasta_rows(VcfStream *s, int64_t *out_rows);
// defined in fasta_columns.cu: fills s->fq_cols with {id, description, sequence}
int fasta_build_columns(VcfStream *s);

// defined in mzml.cu
int mzml_filter_sum(VcfStream *s, const exon_gpu_mzml_pred *pred, double *out_sum, int64_t *out_selected, int64_t *out_spectra);

// defined in fastq_scan.cu
void fq_columns_free(VcfStream *s);
int fastq_next_batch(VcfStream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);
int fastq_filter_count(VcfStream *s, const exon_gpu_fastq_pred *pred, int64_t *out_count, int64_t *out_rows);

// Column store of one FASTQ or FASTA stream (up to four utf8 columns, the second one nullable); batches are views into it.
struct FqColumns {
    std::atomic<int> refs{1};
    bool on_device = false;
    int device = 0;
    int64_t n_rows = 0, n_batches = 0, next = 0;
    int batch_rows = 8192, words_per_batch = 256;
    std::vector<int> projection;
    uint8_t *d_values[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t *d_offsets[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *d_valid = nullptr;
    uint8_t *h_values[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t *h_offsets[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t *h_valid = nullptr;
    std::vector<long long> batch_row0;
    std::vector<long long> batch_v0[4];
    void unref() {
        if (refs.fetch_sub(1) == 1) {
            cudaSetDevice(device);
            for (int k = 0; k < 4; ++k) {
                cudaFree(d_values[k]);
                cudaFree(d_offsets[k]);
                cudaFreeHost(h_values[k]);
                cudaFreeHost(h_offsets[k]);
            }
            cudaFree(d_valid);
            cudaFreeHost(h_valid);
            delete this;
        }
    }
};


// Line index of a resident partition (fastq_scan.cu), used by the FASTQ and the wide VCF column builds.
struct LineIndex {
    const uint8_t **line_start = nullptr;  // device (scratch_b), n_lines + 1 entries
    const uint8_t **line_end = nullptr;
    int64_t n_lines = 0;
    std::vector<long long> file_line0;  // first line of every file in feed order, then n_lines
    uint8_t *extra = nullptr;           // caller's share of scratch_b
};
int build_line_index(VcfStream *s, size_t extra_per_line, size_t extra_fixed, LineIndex *out);

// defined in vcf_wide.cu: VCF columns 2..6 (id, ref, alt, qual, filter)
struct WideStore;
struct WideChildSlot {
    ArrowArray item;  // the utf8 child of a list column
    ArrowArray *item_ptr;
    const void *bufs[3];
    const void *item_bufs[3];
};
bool wide_wanted(const std::vector<int> &projection);
int wide_build(VcfStream *s, std::vector<long long> *batch_row0, int64_t *n_rows_io, WideStore **out);
void wide_free(WideStore *w);
void wide_export(const WideStore *w, int col, int64_t b, int64_t rows, ArrowArray *a, WideChildSlot *slot);

// schema of the batches a stream produces (arrow_stream.cpp: exon_gpu_stream_schema / the Arrow C stream's get_schema)
void vcf_stream_schema(VcfStream *s, struct ArrowSchema *out);
void fastq_stream_schema(VcfStream *s, struct ArrowSchema *out);
void bam_stream_schema(VcfStream *s, struct ArrowSchema *out);
void gff_stream_schema(VcfStream *s, struct ArrowSchema *out);
int mzml_stream_schema(VcfStream *s, struct ArrowSchema *out);

// defined in vcf_columns.cu
int columns_filter_agg(VcfStream *s, const exon_gpu_pred *pred, const exon_gpu_agg *agg, exon_gpu_partial *out);
void columns_free(VcfStream *s);
int columns_next_batch(VcfStream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);

}  // namespace exon

struct exon_gpu_ctx : public exon::Ctx {};
struct exon_gpu_stream : public exon::VcfStream {};
