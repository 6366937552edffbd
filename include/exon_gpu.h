/*
 * exon_gpu.h -- C ABI of the B200-native scan -> filter -> aggregate path for Exon.
 *
 * This is the drop-in boundary.  Every entry point names the reference interface it
 * replaces (paths relative to the reference repo root, wheretrue/exon v0.32.4).  The Rust side binds these
 * with bindgen (INTEGRATION.md shows the stub); in this repository the same symbols are bound with ctypes by
 * exon_b200/_abi.py, and the host-side mirror of the reference operators lives in exon_b200/host/.
 *
 * Conventions
 *   - every function returns an int status (EXON_GPU_OK == 0); on failure a thread-local message is readable
 *     through exon_gpu_last_error() -- the same contract as ArrowArrayStream.get_last_error, which is the
 *     reference's own FFI convention (exon/exon-core/src/ffi/mod.rs:58-73);
 *   - handles are opaque; ONE exon_gpu_stream per DataFusion partition stream, one CUDA stream per handle;
 *     calls are thread-safe across handles and not re-entrant on one handle (the threading contract of
 *     ExecutionPlan::execute, exon/exon-core/src/datasources/vcf/scanner.rs:142-162);
 *   - columns cross the boundary as Arrow C Data Interface structs whose buffers stay owned by the library
 *     until the consumer calls release();
 *   - there is NO CPU fallback: every compute entry point fails with EXON_GPU_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef EXON_GPU_H
#define EXON_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---------------- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
    const char *format;
    const char *name;
    const char *metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema **children;
    struct ArrowSchema *dictionary;
    void (*release)(struct ArrowSchema *);
    void *private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void **buffers;
    struct ArrowArray **children;
    struct ArrowArray *dictionary;
    void (*release)(struct ArrowArray *);
    void *private_data;
};
#endif
/* Arrow C Stream Interface (https://arrow.apache.org/docs/format/CStreamInterface.html) */
#ifndef ARROW_C_STREAM_INTERFACE
#define ARROW_C_STREAM_INTERFACE
struct ArrowArrayStream {
    int (*get_schema)(struct ArrowArrayStream *, struct ArrowSchema *out);
    int (*get_next)(struct ArrowArrayStream *, struct ArrowArray *out);
    const char *(*get_last_error)(struct ArrowArrayStream *);
    void (*release)(struct ArrowArrayStream *);
    void *private_data;
};
#endif

/* ---- status codes ---------------------------------------------------------------------------------- */
enum {
    EXON_GPU_OK = 0,
    EXON_GPU_ERR_ARG = 1,         /* bad argument / NULL pointer */
    EXON_GPU_ERR_CUDA = 2,        /* CUDA runtime failure, or no sm_100 device */
    EXON_GPU_ERR_PARSE = 3,       /* malformed record (the reference surfaces these as ArrowError::ExternalError) */
    EXON_GPU_ERR_STATE = 4,       /* call sequence violated (e.g. feed after the stream was drained) */
    EXON_GPU_ERR_OOM = 5,
    EXON_GPU_ERR_UNSUPPORTED = 6, /* valid request outside what this build implements */
    EXON_GPU_ERR_NCCL = 7
};

typedef struct exon_gpu_ctx exon_gpu_ctx;
typedef struct exon_gpu_stream exon_gpu_stream;

/* Thread-local message of the last failing call on this thread ("" if none). */
const char *exon_gpu_last_error(void);
/* "exon_gpu <semver> sm_100a" */
const char *exon_gpu_version(void);

/* ---- context --------------------------------------------------------------------------------------- */
/* One context per (process, device).  `cuda_stream` may carry a caller-owned cudaStream_t / CUstream on which
 * every kernel and copy of streams opened from this context is enqueued (so the caller can bracket work with
 * its own events); NULL lets the library create a private non-blocking stream. */
int exon_gpu_ctx_create(int device, void *cuda_stream, exon_gpu_ctx **out);
int exon_gpu_ctx_destroy(exon_gpu_ctx *ctx);
/* Kernels launched by this library through `ctx` since creation (bench.py's `gpu_launches`). */
int exon_gpu_ctx_launch_count(exon_gpu_ctx *ctx, int64_t *out);
/* Device time of the most recent fused-scan kernel launch, measured with CUDA events on the launching stream. */
int exon_gpu_ctx_last_kernel_ms(exon_gpu_ctx *ctx, float *out);
/* Device times of the most recent timed launches (oldest first, at most 64 are kept): a bench reads the launches of a
 * whole timed region afterwards, so that no step pays for an event synchronisation.  *out_n <= cap entries are written. */
int exon_gpu_ctx_kernel_ms_history(exon_gpu_ctx *ctx, float *out, int32_t cap, int32_t *out_n);
int exon_gpu_ctx_synchronize(exon_gpu_ctx *ctx);

/* Pinned host / device buffers for callers that want zero-staging feeds. */
int exon_gpu_host_alloc(exon_gpu_ctx *ctx, size_t bytes, void **out);
int exon_gpu_host_free(exon_gpu_ctx *ctx, void *p);
int exon_gpu_device_alloc(exon_gpu_ctx *ctx, size_t bytes, void **out);
int exon_gpu_device_free(exon_gpu_ctx *ctx, void *p);
int exon_gpu_memcpy_h2d(exon_gpu_ctx *ctx, void *dst_device, const void *src_host, size_t bytes);
/* The same copy enqueued on the context's stream without waiting (src_host should be pinned and must stay valid until
 * exon_gpu_ctx_synchronize); bench.py uses it to measure the bare host->device ceiling next to the end-to-end number. */
int exon_gpu_memcpy_h2d_async(exon_gpu_ctx *ctx, void *dst_device, const void *src_host, size_t bytes);

/* ---- region predicate (a10/a11) --------------------------------------------------------------------- */
/* `chrom = <name> AND pos BETWEEN lo AND hi` -- 1-based, both ends inclusive
 * (exon/exon-core/src/physical_plan/pos_interval_physical_expr.rs:79-98, region_physical_expr.rs:220-240,
 * exon/exon-core/src/udfs/vcf/mod.rs:65-131).  has_chrom == 0 drops the name test (interval_match, :232-274);
 * has_interval == 0 drops the position test (chrom_match, :167-196). */
typedef struct {
    const char *chrom;
    int32_t chrom_len;
    int32_t has_chrom;
    int32_t has_interval;
    int64_t lo; /* >= 1 */
    int64_t hi; /* INT64_MAX = open end */
} exon_gpu_region;

/* Parses "name", "name:start", "name:start-end" the way noodles-core Region::from_str does at its reference
 * call sites (exon/exon-core/src/physical_plan/infer_region.rs:25-42, udfs/vcf/mod.rs:85-95).
 * `name_buf` (>= 256 bytes) receives the contig name and out->chrom points into it. */
int exon_gpu_region_parse(const char *s, char *name_buf, size_t name_buf_len, exon_gpu_region *out);
/* Interval literal of interval_match ("a-b", "a", "a-", "-b"; udfs/vcf/mod.rs:246-252): has_chrom = 0. */
int exon_gpu_interval_parse(const char *s, exon_gpu_region *out);

/* QUAL text -> f32 exactly as the reference's builder obtains it: Rust `f32::from_str` applied by noodles-vcf 0.70
 * `Record::quality_score` (call site exon/exon-vcf/src/array_builder/lazy_array_builder.rs:205-208): correctly rounded,
 * "inf" / "infinity" / "nan" in any case, no hex, at least one mantissa digit.  The same routine runs inside the column
 * kernel; this host entry exists so that it can be pinned against exact arithmetic without a device.
 * EXON_GPU_ERR_PARSE: not a float literal; EXON_GPU_ERR_UNSUPPORTED: more than 36 significant digits. */
int exon_gpu_parse_f32(const char *s, size_t len, float *out);

/* ---- file -> partition assignment (a3) ---------------------------------------------------------------- */
/* ExonFileScanConfig::regroup_files_by_size (exon/exon-core/src/datasources/exon_file_scan_config.rs:79-110):
 * stable sort by size ascending, partitions = min(target, n_files), file i of the sorted order -> i % partitions.
 * out_partition[i] is the partition of INPUT file i; *out_n_partitions the number of non-empty partitions. */
int exon_gpu_regroup_files_by_size(const int64_t *sizes, int32_t n_files, int32_t target_partitions,
                                   int32_t *out_partition, int32_t *out_n_partitions);

/* ---- VCF partition stream (a4-a9) --------------------------------------------------------------------- */
typedef struct {
    int32_t batch_rows;        /* session batch size; reference default 8192 (exon/exon-common/src/lib.rs:27) */
    int32_t n_projection;      /* file-schema column indices to materialise, in output order ... */
    const int32_t *projection; /* ... VCFConfig.projection (exon/exon-vcf/src/config.rs:23-64); cols 0..8 (chrom pos id ref alt qual filter info formats) */
    int32_t columns_on_device; /* 0: next_batch buffers are pinned host memory; 1: device memory */
    /* Optional predicate declared up front so that every feed() can be scanned while the next one is still
     * copying (fused a5-a9).  NULL = none declared; filter_count() then scans what is resident. */
    const exon_gpu_region *pushdown;
    int32_t strict;            /* 1: validate POS of every row like the reference does; 0: only rows the predicate reads */
    int32_t kernel_variant;    /* 0 = default (TMA-staged); other values select experimental kernels */
} exon_gpu_vcf_opts;

/* VCFScan::execute + VCFOpener::open (exon/exon-core/src/datasources/vcf/scanner.rs:142-162,
 * vcf/file_opener/unindex_file_opener.rs:48-92): opens one partition stream. */
int exon_gpu_vcf_open(exon_gpu_ctx *ctx, const exon_gpu_vcf_opts *opts, exon_gpu_stream **out);
int exon_gpu_vcf_close(exon_gpu_stream *s);
/* Forget everything fed so far but keep the device arena for the next query on this partition. */
int exon_gpu_vcf_reset(exon_gpu_stream *s);

/* The header text of the partition's files (only its ##INFO lines are read: ID and Type), as the reference's builder gets
 * them from the noodles Header its opener parsed (unindex_file_opener.rs:74-88, lazy_array_builder.rs:70-75).  Required
 * before exon_gpu_vcf_next_batch when the projection holds column 7 (info); every file of the partition is taken to
 * share these definitions. */
int exon_gpu_vcf_set_header(exon_gpu_stream *s, const char *text, size_t len);

/* Bytes in.  Consecutive calls deliver consecutive byte ranges of ONE file (uncompressed VCF text, header
 * included -- the library skips it like `read_header`, unindex_file_opener.rs:74-88); is_last != 0 ends the
 * file, and the next feed() starts the next file of the partition's file group (what FileStream does).
 * is_device_ptr != 0: `text` is device memory, 16-byte aligned, used in place (zero copy) and must stay
 * valid until the stream is closed; a non-final device range must end on a line boundary. */
int exon_gpu_vcf_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last);

/* Columns out: AsyncBatchStream::read_batch + LazyVCFArrayBuilder::{append,finish} +
 * ExonArrayBuilder::try_into_record_batch (exon/exon-vcf/src/async_batch_stream.rs:80-109,
 * exon/exon-vcf/src/array_builder/lazy_array_builder.rs:153-484, exon/exon-common/src/array_builder.rs:25-36).
 * Fills a struct array (format "+s") of <= batch_rows rows whose children are the projected columns in
 * projection order: chrom = utf8 ("u": validity NULL, int32 offsets starting at 0, bytes), pos = int64 ("l"),
 * id / alt / filter = list<item: utf8> ("+l"), ref = utf8, qual = float32 ("f") -- with the lazy builder's own
 * semantics (lazy_array_builder.rs:169-216): id "." -> NULL; alt "." -> NULL, anything else -> a valid EMPTY list
 * (the builder never appends the alleles); qual "." -> NULL, else Rust f32::from_str (correctly rounded);
 * filter always valid, "." -> [].  A record with fewer than 8 fields or a malformed QUAL fails the call
 * (EXON_GPU_ERR_PARSE).  info (7, utf8, string mode): the reference does not copy the field, it prints noodles' typed view
 * of it again (lazy_array_builder.rs:217-298: `key=value` joined by ';', a flag as `key=true`, numbers through Rust's
 * Display).  The column holds exactly that string whenever every number of the field already has the form Display gives
 * it (the reference's own golden rows do, slt/vcf-select-tests.slt:6-10); a value for which that does not hold, a key
 * the header does not define, a '%' escape in a string, a flag with a value: EXON_GPU_ERR_UNSUPPORTED, never an
 * approximation; a non-flag key without a value fails like the reference's unwrap does (EXON_GPU_ERR_PARSE).
 * formats (8) is not built: exon_gpu_vcf_open rejects it with EXON_GPU_ERR_UNSUPPORTED.
 * End of stream: returns EXON_GPU_OK with out->release == NULL (ArrowArrayStream.get_next convention). */
int exon_gpu_vcf_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);

/* Fused a5-a9: the count FilterExec + AggregateExec(Partial) would produce for this partition.
 * region == NULL counts every record (COUNT(*) with an empty projection).
 * Synchronises the stream and returns the count on the host. */
int exon_gpu_vcf_filter_count(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *out_count);
/* Same, but leaves the int64 partial in caller-provided DEVICE memory and does not synchronise (used when the
 * partial feeds an all-reduce on the same CUDA stream). */
int exon_gpu_vcf_filter_count_async(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *device_out);
/* Records (lines after the header) fed so far. */
int exon_gpu_vcf_rows(exon_gpu_stream *s, int64_t *out_rows);
/* Body bytes (text minus headers) resident for this stream: the algorithmic bytes of the fused scan. */
int exon_gpu_vcf_body_bytes(exon_gpu_stream *s, int64_t *out_bytes);

/* ---- FASTQ partition stream (BASELINE configs[1]) ------------------------------------ */
/* FASTQScan::execute + FASTQOpener::open + BatchReader (exon/exon-core/src/datasources/fastq/scanner.rs:126,
 * fastq/file_opener.rs:51, exon/exon-fastq/src/batch_reader.rs:56-82).  A FASTQ stream is an exon_gpu_stream: it
 * shares the arena, the host/device feeds and the file framing with the VCF stream; there is no header. */
typedef struct {
    int32_t batch_rows;        /* session batch size (8192) */
    int32_t n_projection;      /* columns exon_gpu_fastq_next_batch materialises, in output order (0 for the fused query alone) */
    const int32_t *projection; /* FASTQ file schema: 0 name, 1 description, 2 sequence, 3 quality_scores (exon-fastq/src/config.rs:79-88) */
    int32_t columns_on_device;
} exon_gpu_fastq_opts;
/* `mean(quality) > min_mean_num / min_mean_den` with Phred score = byte - phred_offset (33:
 * exon/exon-core/src/udfs/sequence/quality_score_string_to_list.rs:80-93), evaluated over integers:
 * (sum(byte) - phred_offset * len) * den > num * len; a record with an empty quality string is not selected. */
typedef struct {
    int32_t phred_offset;
    int32_t pad_;
    int64_t min_mean_num;
    int64_t min_mean_den; /* > 0 */
} exon_gpu_fastq_pred;
int exon_gpu_fastq_open(exon_gpu_ctx *ctx, const exon_gpu_fastq_opts *opts, exon_gpu_stream **out);
/* Same contract as exon_gpu_vcf_feed: consecutive byte ranges of one file, is_last ends the file. */
int exon_gpu_fastq_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last);
/* Fused scan -> filter -> COUNT over everything fed so far (pred == NULL: COUNT(*) = number of records).  Fails with
 * EXON_GPU_ERR_PARSE where noodles' reader would: a definition line that does not start with '@', a third line
 * that does not start with '+', a file that ends after the first or second line of a record. */
int exon_gpu_fastq_filter_count(exon_gpu_stream *s, const exon_gpu_fastq_pred *pred, int64_t *out_count);
int exon_gpu_fastq_rows(exon_gpu_stream *s, int64_t *out_rows);
/* Columns out: BatchReader::read_batch + FASTQArrayBuilder::{append, finish} (exon/exon-fastq/src/batch_reader.rs:63-82,
 * exon/exon-fastq/src/array_builder.rs:68-118).  A struct array of <= batch_rows rows whose children are the projected
 * utf8 columns; `description` carries a validity bitmap (NULL when the definition line has nothing after the name).
 * Same end-of-stream and ownership conventions as exon_gpu_vcf_next_batch. */
int exon_gpu_fastq_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);
/* Compressed bytes in: consecutive byte ranges of ONE BGZF file (or a plain single-member .gz; slower, one warp) --
 * what VCFOpener::open / FASTQOpener::open wrap in a BGZF / gzip decoder for FileCompressionType::GZIP
 * (exon/exon-core/src/datasources/vcf/file_opener/unindex_file_opener.rs:59-73, exon-core/src/streaming_bgzf.rs:22-118).
 * The compressed members are copied to HBM unchanged and inflated there (one warp per member) into the stream's
 * arena; the result is framed exactly like exon_gpu_vcf_feed / exon_gpu_fastq_feed of the uncompressed text.
 * Ranges are buffered on the host until is_last != 0 completes the file.  Works on VCF and FASTQ streams. */
int exon_gpu_stream_feed_gzip(exon_gpu_stream *s, const uint8_t *data, size_t len, int is_last);
/* Inflates a whole BGZF / gzip file on the device into `out` (host memory, or device memory when out_is_device != 0).
 * *out_len receives the uncompressed size even when `out` is too small (EXON_GPU_ERR_ARG then). */
int exon_gpu_gzip_inflate(exon_gpu_ctx *ctx, const uint8_t *data, size_t len, uint8_t *out, size_t out_cap, int out_is_device,
                          size_t *out_len);
/* ---- indexed scan (a12): tabix chunks -> only the covered BGZF members are inflated --------------------------- */
typedef struct {
    uint64_t start, end; /* BGZF virtual positions: compressed offset << 16 | offset inside the inflated member */
} exon_gpu_chunk;
/* noodles::tabix::Reader::read_index + Index::query as called by get_byte_range_for_file
 * (exon/exon-core/src/datasources/indexed_file/indexed_bgzf_file.rs:52-83): chunks of `tbi` (the bytes of the .tbi file)
 * that can hold records of `region` (name required; 1-based inclusive interval optional), merged and sorted.
 * A contig that is not in the index gives 0 chunks.  `out` may be NULL to ask for the count. */
int exon_gpu_tabix_query(exon_gpu_ctx *ctx, const uint8_t *tbi, size_t len, const exon_gpu_region *region, exon_gpu_chunk *out,
                         int32_t cap, int32_t *n_chunks);
/* IndexedVCFOpener::open (exon/exon-core/src/datasources/vcf/file_opener/indexed_file_opener.rs:53-214): `data` holds the
 * bytes [file_offset, file_offset + len) of a .vcf.gz -- what a ranged object-store GET returns -- starting at a member
 * boundary at or before the chunk's first member and reaching through the member in which the chunk ends.  Only the
 * chunk's members are inflated (on the device); the records between the two virtual positions become one file of the
 * partition.  The region predicate itself is applied by exon_gpu_vcf_filter_count / the batches' consumer, to every
 * record (the reference's IndexedAsyncBatchStream stops filtering after a full batch, SURVEY 2.2 #6: not reproduced). */
int exon_gpu_stream_feed_bgzf_chunk(exon_gpu_stream *s, const uint8_t *data, size_t len, uint64_t file_offset, const exon_gpu_chunk *chunk);
/* Format-independent stream calls (the exon_gpu_vcf_* spellings remain valid for VCF streams). */
int exon_gpu_stream_close(exon_gpu_stream *s);

/* ---- Arrow C stream export (seam B4) --------------------------------------------------------------------------
 * The reference's own FFI hands batches out as an FFI_ArrowArrayStream (create_dataset_stream_from_table_provider,
 * exon/exon-core/src/ffi/mod.rs:58-73; stream object :25-49) and exon-r / exon-py read that struct.  These two calls give
 * the same object for any stream opened with a projection (VCF, FASTQ, FASTA, BAM, GFF, mzML; host-resident columns):
 * get_schema / get_next (release == NULL in the array = end of stream) / get_last_error / release.  get_next returns an
 * errno value (EINVAL, ENOMEM, ENOTSUP, EIO) and get_last_error the library's message.  take_ownership != 0: releasing the
 * Arrow stream closes the exon_gpu_stream too.  A Rust host wraps it with ArrowArrayStreamReader::from_raw. */
int exon_gpu_stream_schema(exon_gpu_stream *s, struct ArrowSchema *out);
int exon_gpu_stream_export(exon_gpu_stream *s, struct ArrowArrayStream *out, int take_ownership);
int exon_gpu_stream_reset(exon_gpu_stream *s);
int exon_gpu_stream_body_bytes(exon_gpu_stream *s, int64_t *out_bytes);

/* ---- FASTA partition stream (BASELINE configs[0]: SELECT COUNT(*) FROM fasta_scan(...)) ---------------------------- */
/* FASTAScan::execute + BatchReader::read_batch (exon/exon-core/src/datasources/fasta/scanner.rs,
 * exon/exon-fasta/src/batch_reader.rs) reduced to the row count: records = '>' definition lines.  Fed like the other
 * text formats (exon_gpu_fasta_feed, or exon_gpu_stream_feed_gzip for .gz).  Fails with EXON_GPU_ERR_PARSE when a
 * non-empty file does not start with '>'. */
int exon_gpu_fasta_open(exon_gpu_ctx *ctx, exon_gpu_stream **out);
int exon_gpu_fasta_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last);
int exon_gpu_fasta_rows(exon_gpu_stream *s, int64_t *out_rows);
/* Columns out: BatchReader::read_batch + FASTAArrayBuilder::{append, finish} for the Utf8 sequence type
 * (exon/exon-fasta/src/batch_reader.rs:52-103, exon/exon-fasta/src/array_builder.rs:108-160; schema exon/exon-fasta/src/config.rs:162-226):
 * 0 id utf8 !null (up to the first ASCII whitespace of the definition line), 1 description utf8 (the trimmed rest, NULL when
 * there is none), 2 sequence utf8 (all sequence lines, terminators removed).  A file that does not begin with '>', a
 * definition without a name or without any sequence line: EXON_GPU_ERR_PARSE.  A batch whose sequence bytes exceed 2^31 - 1
 * (the reference's LargeUtf8 option) is EXON_GPU_ERR_UNSUPPORTED.  exon_gpu_fastq_opts carries batch_rows / projection. */
int exon_gpu_fasta_open_columns(exon_gpu_ctx *ctx, const exon_gpu_fastq_opts *opts, exon_gpu_stream **out);
int exon_gpu_fasta_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);

/* ---- GFF partition stream ----------------------------------------------------------------------------------------- */
/* GFFScan + BatchReader::{read_line, filter, read_batch} (exon/exon-gff/src/batch_reader.rs:56-130) under
 * AggregateExec count(*).  region == NULL: COUNT(*) = record lines (directives "##" and comments "#" are skipped).
 * region: gff_region_filter semantics -- seqname == region->chrom, and the interval (1-based, inclusive), if any,
 * contains the record's START (batch_reader.rs:70-96).  Only the fields the predicate reads are validated. */
int exon_gpu_gff_open(exon_gpu_ctx *ctx, exon_gpu_stream **out);
int exon_gpu_gff_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last);
int exon_gpu_gff_filter_count(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *out_count);
/* Columns out: BatchReader::read_batch + GFFArrayBuilder::{append, finish} (exon/exon-gff/src/batch_reader.rs:99-130,
 * exon/exon-gff/src/array_builder.rs:84-200; schema exon/exon-gff/src/config.rs:81-108), columns 0..7:
 *   0 seqname utf8 | 1 source utf8 | 2 type utf8 | 3 start int64 | 4 end int64 | 5 score float32 ("." -> NULL, else Rust
 *   f32::from_str) | 6 strand utf8 ("+" / "-") | 7 phase utf8 ("." -> NULL, else "0" / "1" / "2")
 * Like the reference, ONE batch per file whatever batch_rows says (read_batch has no row limit).  A strand of "." or "?" is
 * NULL in a column the reference declares non-nullable -- its batch construction fails, and so does this call
 * (EXON_GPU_ERR_PARSE) -- as do an empty line, fewer than 9 fields, a start / end of 0.
 *   8 attributes map<utf8, list<utf8>> ("+m"; entries struct{keys, values: list<item>}): `key=v1,v2;...`, keys and values
 *   percent-decoded, one map entry per attribute in line order.  The reference's builder (array_builder.rs:150-176) closes a
 *   single-valued attribute's list BEFORE appending its string, so that string lands in the NEXT entry's list (the last
 *   one is left for the next row's first entry, the file's last one is dropped); batches come out exactly like that.
 * exon_gpu_fastq_opts carries the projection. */
int exon_gpu_gff_open_columns(exon_gpu_ctx *ctx, const exon_gpu_fastq_opts *opts, exon_gpu_stream **out);
int exon_gpu_gff_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);

/* ---- BAM partition stream (BASELINE configs[3]) ---------------------------------------- */
/* BAMScan::execute + BAMOpener::open + BatchReader (exon/exon-core/src/datasources/bam/scanner.rs:138,
 * bam/file_opener.rs:39, exon/exon-bam/src/batch_reader.rs:70-107).  Fed with the BGZF bytes of whole .bam files;
 * the members are inflated on the device and the records never leave HBM. */
int exon_gpu_bam_open(exon_gpu_ctx *ctx, exon_gpu_stream **out);
int exon_gpu_bam_feed(exon_gpu_stream *s, const uint8_t *data, size_t len, int is_last);
/* `(flag & flag_exclude) == 0 AND (flag & flag_require) == flag_require AND CAST(mapping_quality AS INT) >= min_mapq`:
 * the flag tests are is_unmapped / is_secondary / ... (exon/exon-core/src/udfs/sam/samflags.rs:111-141) combined;
 * mapping_quality is a nullable STRING column, NULL for MAPQ 255 (exon/exon-bam/src/array_builder.rs:136-143), so a
 * record with MAPQ 255 never passes when min_mapq >= 0.  min_mapq < 0 drops the MAPQ term. */
typedef struct {
    uint32_t flag_exclude;
    uint32_t flag_require;
    int32_t min_mapq;
    /* bam_region_filter('name:lo-hi', reference, start, end) -- SemiLazyRecord::intersects,
     * exon/exon-bam/src/indexed_async_batch_stream.rs:66-86: same reference (by NAME, looked up in every file's header) and
     * [start, end] intersects [region_lo, region_hi] (1-based, inclusive; end = start + reference-consuming CIGAR length - 1);
     * records without a reference or a position never match. */
    int32_t has_region;
    const char *region_ref;
    int32_t region_ref_len;
    int32_t pad_;
    int64_t region_lo, region_hi;
} exon_gpu_bam_pred;
/* SELECT reference, COUNT(*) ... GROUP BY reference over everything fed so far.  Groups are reference NAMES in
 * header order (first file first), followed by one group for the NULL reference (refID -1): counts[g] for
 * g < *n_groups, `cap` = room in counts.  pred == NULL selects every record; *total_rows = records scanned.
 * Fails with EXON_GPU_ERR_PARSE on a record chain that does not end at the end of the file. */
int exon_gpu_bam_filter_count_by_reference(exon_gpu_stream *s, const exon_gpu_bam_pred *pred, int64_t *counts, int32_t cap,
                                           int32_t *n_groups, int64_t *total_rows);
/* Name of group g after a query (*name == NULL for the last group, the NULL reference). */
int exon_gpu_bam_group_name(exon_gpu_stream *s, int32_t group, const char **name);

/* Columns out: BatchReader::read_batch + BAMArrayBuilder::{append, finish} (exon/exon-bam/src/batch_reader.rs:88-107,
 * exon/exon-bam/src/array_builder.rs:102-218) with SemiLazyRecord::alignment_end (indexed_async_batch_stream.rs:43-50).
 * File schema (SAMSchemaBuilder::default, exon/exon-sam/src/schema_builder.rs:385-401), projected by index:
 *   0 name utf8 !null | 1 flag int32 !null | 2 reference utf8 (NULL for refID -1) | 3 start int64 (pos + 1, NULL for -1) |
 *   4 end int64 (start + reference-consuming CIGAR length - 1; NULL without a start or when that is 0) |
 *   5 mapping_quality utf8 (decimal string, NULL for 255) | 6 cigar utf8 ("55M13394N21M") | 7 mate_reference utf8 |
 *   8 sequence utf8 (bases decoded from 4 bits) | 9 quality_score list<item: int64> (raw bytes as i8, [] when missing)
 *   10 tags list<item: struct{tag utf8 !null, value utf8}> (TagsMapBuilder, exon/exon-sam/src/tag_builder.rs:480-741, the
 *   default bam_parse_tags = false form): every auxiliary field in record order, its value as text -- integers in
 *   decimal, A as the character, Z / H as written, f through Rust's f32 Display, B integer arrays joined by ",", B:f
 *   arrays as "{:.2}" joined by ", " (an element of 9e13 or more in magnitude: EXON_GPU_ERR_UNSUPPORTED).
 * A record whose read name is missing ("*") fails the call, as the reference's batch construction does (NULL in a
 * non-nullable column); so does an auxiliary field of unknown type or one that overruns its record.  Batches never span
 * files. */
typedef struct {
    int32_t batch_rows;        /* session batch size; reference default 8192 */
    int32_t n_projection;
    const int32_t *projection;
    int32_t columns_on_device; /* 0: batch buffers are pinned host memory; 1: device memory */
} exon_gpu_bam_opts;
int exon_gpu_bam_open_columns(exon_gpu_ctx *ctx, const exon_gpu_bam_opts *opts, exon_gpu_stream **out);
/* Same contract as exon_gpu_vcf_next_batch: struct array of <= batch_rows rows, release == NULL at the end of the stream. */
int exon_gpu_bam_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);

/* ---- mzML partition stream (BASELINE configs[4]) --------------------------------------- */
/* MzMLScan::execute + MzMLOpener::open + BatchReader (exon/exon-core/src/datasources/mzml/scanner.rs:139,
 * mzml/file_opener.rs:48, exon/exon-mzml/src/batch_reader.rs:54-82).  Fed like a VCF stream (plain text through
 * exon_gpu_mzml_feed, .gz through exon_gpu_stream_feed_gzip); a range must end on a line boundary unless it is the last. */
int exon_gpu_mzml_open(exon_gpu_ctx *ctx, exon_gpu_stream **out);
int exon_gpu_mzml_feed(exon_gpu_stream *s, const uint8_t *text, size_t len, int is_device_ptr, int is_last);
typedef struct {
    double mz_lo, mz_hi; /* mz BETWEEN mz_lo AND mz_hi, both ends inclusive */
} exon_gpu_mzml_pred;
/* SELECT SUM(i) FROM (SELECT unnest(mz.mz) m, unnest(intensity.intensity) i FROM mzml) WHERE m BETWEEN lo AND hi over
 * everything fed so far (pred == NULL: every zipped peak).  *out_sum is an f64 sum accumulated in device order (1e-6
 * relative parity, north_star); *out_selected the number of peaks summed; *out_spectra the number of <spectrum>
 * elements (COUNT(*)).  Arrays are decoded as exon/exon-mzml/src/mzml_reader/binary_conversion.rs:26-95 does
 * (base64, optional zlib -- inflated on the device --, little-endian f32 / f64).  A zlib-compressed array is sized by the
 * defaultArrayLength of its <spectrum>; without it, or when the stream inflates to another size: EXON_GPU_ERR_PARSE. */
/* Record batches of an mzML stream (MzMLArrayBuilder, exon/exon-mzml/src/array_builder.rs:236-439; schema config.rs:92-147). */
int exon_gpu_mzml_open_columns(exon_gpu_ctx *ctx, const exon_gpu_fastq_opts *opts, exon_gpu_stream **out);
int exon_gpu_mzml_next_batch(exon_gpu_stream *s, struct ArrowArray *out, struct ArrowSchema *out_schema);
int exon_gpu_mzml_filter_sum(exon_gpu_stream *s, const exon_gpu_mzml_pred *pred, double *out_sum, int64_t *out_selected,
                             int64_t *out_spectra);

/* ---- columnar filter + aggregate over Arrow buffers (a8-a9) ---------------------------------------------- */
enum { EXON_GPU_AGG_COUNT_STAR = 0, EXON_GPU_AGG_COUNT = 1, EXON_GPU_AGG_SUM = 2, EXON_GPU_AGG_AVG = 3 };

typedef struct {
    int32_t chrom_col;  /* child index of the utf8 column compared with region.chrom, -1 = none */
    int32_t pos_col;    /* child index of the int64 column compared with [lo, hi], -1 = none */
    exon_gpu_region region;
} exon_gpu_pred;

typedef struct {
    int32_t kind;       /* EXON_GPU_AGG_* */
    int32_t value_col;  /* child index of the aggregated column (int64 "l", float32 "f" or float64 "g"); -1 for COUNT(*) */
} exon_gpu_agg;

/* AggregateExec(Partial) state: COUNT -> count; SUM -> sum_*; AVG -> (sum, count) (datafusion-functions-aggregate 44). */
typedef struct {
    int64_t count;
    int64_t sum_i64;
    double sum_f64;
} exon_gpu_partial;

/* FilterExec (arrow eq / gt_eq / lt_eq / and_kleene: a NULL operand makes the row unselected) followed by the
 * partial aggregate, over ONE record batch given as an Arrow struct array.  Buffers may be host or device
 * memory (`buffers_on_device`); host buffers are staged through the context's pinned ring. */
int exon_gpu_filter_agg(exon_gpu_ctx *ctx, const struct ArrowArray *batch, const struct ArrowSchema *schema,
                        int buffers_on_device, const exon_gpu_pred *pred, const exon_gpu_agg *agg,
                        exon_gpu_partial *out);

/* The same operators for a stream of batches whose buffers already live in DEVICE memory (what
 * exon_gpu_vcf_next_batch yields with columns_on_device = 1): each call enqueues one kernel that adds into the
 * caller's device-resident accumulator (zero it with exon_gpu_device_alloc + exon_gpu_memset, or reuse), nothing
 * synchronises until exon_gpu_partial_read.  This is AggregateExec(Partial)'s accumulator living across batches. */
int exon_gpu_filter_agg_accumulate(exon_gpu_ctx *ctx, const struct ArrowArray *batch, const struct ArrowSchema *schema,
                                   const exon_gpu_pred *pred, const exon_gpu_agg *agg, exon_gpu_partial *device_acc);
/* The same operators over MANY device-resident batches in one kernel launch (all batches share `schema`); the
 * partial comes back on the host.  One launch per partition instead of one per 8192-row batch. */
int exon_gpu_filter_agg_batches(exon_gpu_ctx *ctx, const struct ArrowArray *const *batches, int32_t n_batches,
                                const struct ArrowSchema *schema, const exon_gpu_pred *pred, const exon_gpu_agg *agg,
                                exon_gpu_partial *out);
/* VCFScan -> FilterExec -> AggregateExec(Partial) with the record batches kept in HBM: builds the projected columns
 * of everything fed so far (the batches exon_gpu_vcf_next_batch would hand out) if that has not happened yet, then
 * filters and aggregates all of them in one launch.  pred->chrom_col / pos_col and agg->value_col are indices into
 * the stream's projection. */
int exon_gpu_vcf_filter_agg(exon_gpu_stream *s, const exon_gpu_pred *pred, const exon_gpu_agg *agg, exon_gpu_partial *out);
int exon_gpu_partial_read(exon_gpu_ctx *ctx, const exon_gpu_partial *device_acc, int sum_is_integer, exon_gpu_partial *out);
int exon_gpu_memset(exon_gpu_ctx *ctx, void *device_ptr, int value, size_t bytes);

/* Evaluated region UDFs (exon/exon-core/src/udfs/vcf/mod.rs): region_match(chrom, pos, 'name:lo-hi') :65-131,
 * chrom_match(chrom, 'name') :167-196, interval_match(pos, 'lo-hi') :232-274.  One byte per row in HOST memory:
 * out_values[i] in {0,1}; out_valid[i] == 0 marks a NULL result (chrom_match of a NULL chrom); may be NULL.
 * region_match fails on a NULL operand, region_match / interval_match fail on pos < 1 (Position::try_from),
 * interval_match maps a NULL pos to false -- all as the reference does. */
enum { EXON_GPU_UDF_REGION_MATCH = 0, EXON_GPU_UDF_CHROM_MATCH = 1, EXON_GPU_UDF_INTERVAL_MATCH = 2 };
int exon_gpu_region_udf(exon_gpu_ctx *ctx, int kind, const struct ArrowArray *batch, const struct ArrowSchema *schema,
                        int buffers_on_device, const exon_gpu_pred *pred, uint8_t *out_values, uint8_t *out_valid);

/* ---- multi-GPU final aggregate ---------------------------------------------------------------- */
/* CoalescePartitionsExec + AggregateExec(Final) across GPUs: one ncclAllReduce(sum) of the partial over
 * NVLink.  The communicator is created from an id produced on rank 0 and distributed by the host. */
#define EXON_GPU_NCCL_ID_BYTES 128
int exon_gpu_nccl_unique_id(uint8_t id[EXON_GPU_NCCL_ID_BYTES]);
int exon_gpu_nccl_init(exon_gpu_ctx *ctx, const uint8_t id[EXON_GPU_NCCL_ID_BYTES], int n_ranks, int rank);
int exon_gpu_allreduce_partial(exon_gpu_ctx *ctx, exon_gpu_partial *inout);
/* The same for a GROUP BY whose groups are the same on every rank (per-reference counts): one
 * ncclAllReduce(sum, int64, n) of the host vector `inout`. */
int exon_gpu_allreduce_counts(exon_gpu_ctx *ctx, int64_t *inout, int32_t n);
/* AggregateExec(Partial) -> CoalescePartitionsExec -> AggregateExec(Final) for the fused COUNT in one call: the
 * local scan leaves its int64 partial in device memory, one ncclAllReduce(sum, int64, 1) runs on the same CUDA
 * stream, and both the local and the global count come back with a single synchronisation.  Collective: every
 * rank of the communicator must call it (a rank whose input is malformed still takes part, then fails). */
int exon_gpu_vcf_filter_count_global(exon_gpu_stream *s, const exon_gpu_region *region, int64_t *out_local,
                                     int64_t *out_global);

#ifdef __cplusplus
}
#endif
#endif /* EXON_GPU_H */
