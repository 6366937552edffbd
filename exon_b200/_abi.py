"""ctypes binding of include/exon_gpu.h -- the same symbols a Rust host binds with bindgen.

Loading is strict: a missing or stale libexon_gpu.so raises, and every compute call surfaces the library's
error (there is no CPU fallback anywhere in this package).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libexon_gpu.so")

OK, ERR_ARG, ERR_CUDA, ERR_PARSE, ERR_STATE, ERR_OOM, ERR_UNSUPPORTED, ERR_NCCL = range(8)
AGG_COUNT_STAR, AGG_COUNT, AGG_SUM, AGG_AVG = range(4)
UDF_REGION_MATCH, UDF_CHROM_MATCH, UDF_INTERVAL_MATCH = range(3)
INT64_MAX = (1 << 63) - 1
NCCL_ID_BYTES = 128

# Every symbol include/exon_gpu.h declares (tests/test_abi_symbols.py checks the header against this list).
SYMBOLS = [
    "exon_gpu_last_error", "exon_gpu_version", "exon_gpu_ctx_create", "exon_gpu_ctx_destroy",
    "exon_gpu_ctx_launch_count", "exon_gpu_ctx_last_kernel_ms", "exon_gpu_ctx_kernel_ms_history", "exon_gpu_ctx_synchronize", "exon_gpu_host_alloc",
    "exon_gpu_host_free", "exon_gpu_device_alloc", "exon_gpu_device_free", "exon_gpu_memcpy_h2d", "exon_gpu_memcpy_h2d_async",
    "exon_gpu_region_parse", "exon_gpu_interval_parse", "exon_gpu_parse_f32", "exon_gpu_regroup_files_by_size", "exon_gpu_vcf_open",
    "exon_gpu_vcf_close", "exon_gpu_vcf_reset", "exon_gpu_vcf_set_header", "exon_gpu_vcf_feed", "exon_gpu_vcf_next_batch",
    "exon_gpu_vcf_filter_count", "exon_gpu_vcf_filter_count_async", "exon_gpu_vcf_rows", "exon_gpu_vcf_body_bytes",
    "exon_gpu_filter_agg", "exon_gpu_nccl_unique_id", "exon_gpu_nccl_init", "exon_gpu_allreduce_partial",
    "exon_gpu_vcf_filter_count_global", "exon_gpu_filter_agg_accumulate", "exon_gpu_partial_read", "exon_gpu_memset",
    "exon_gpu_region_udf", "exon_gpu_filter_agg_batches", "exon_gpu_vcf_filter_agg",
    "exon_gpu_fastq_open", "exon_gpu_fastq_feed", "exon_gpu_fastq_filter_count", "exon_gpu_fastq_rows", "exon_gpu_fastq_next_batch",
    "exon_gpu_stream_close", "exon_gpu_stream_schema", "exon_gpu_stream_export", "exon_gpu_stream_reset", "exon_gpu_stream_body_bytes", "exon_gpu_stream_feed_gzip", "exon_gpu_gzip_inflate", "exon_gpu_bam_open", "exon_gpu_bam_feed",
    "exon_gpu_bam_filter_count_by_reference", "exon_gpu_bam_open_columns", "exon_gpu_bam_next_batch", "exon_gpu_bam_group_name", "exon_gpu_allreduce_counts",
    "exon_gpu_mzml_open", "exon_gpu_mzml_feed", "exon_gpu_mzml_filter_sum", "exon_gpu_mzml_open_columns", "exon_gpu_mzml_next_batch", "exon_gpu_tabix_query", "exon_gpu_stream_feed_bgzf_chunk",
    "exon_gpu_fasta_open", "exon_gpu_fasta_feed", "exon_gpu_fasta_rows", "exon_gpu_fasta_open_columns", "exon_gpu_fasta_next_batch", "exon_gpu_gff_open", "exon_gpu_gff_feed",
    "exon_gpu_gff_filter_count", "exon_gpu_gff_open_columns", "exon_gpu_gff_next_batch",
]


class ExonGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"exon_gpu error {code}: {message}")
        self.code = code
        self.message = message


class Region(C.Structure):
    _fields_ = [("chrom", C.c_char_p), ("chrom_len", C.c_int32), ("has_chrom", C.c_int32),
                ("has_interval", C.c_int32), ("lo", C.c_int64), ("hi", C.c_int64)]


class VcfOpts(C.Structure):
    _fields_ = [("batch_rows", C.c_int32), ("n_projection", C.c_int32), ("projection", C.POINTER(C.c_int32)),
                ("columns_on_device", C.c_int32), ("pushdown", C.POINTER(Region)), ("strict", C.c_int32),
                ("kernel_variant", C.c_int32)]


class FastqOpts(C.Structure):
    _fields_ = [("batch_rows", C.c_int32), ("n_projection", C.c_int32), ("projection", C.POINTER(C.c_int32)),
                ("columns_on_device", C.c_int32)]


class FastqPred(C.Structure):
    _fields_ = [("phred_offset", C.c_int32), ("pad_", C.c_int32), ("min_mean_num", C.c_int64), ("min_mean_den", C.c_int64)]


class BamPred(C.Structure):
    _fields_ = [("flag_exclude", C.c_uint32), ("flag_require", C.c_uint32), ("min_mapq", C.c_int32), ("has_region", C.c_int32),
                ("region_ref", C.c_char_p), ("region_ref_len", C.c_int32), ("pad_", C.c_int32), ("region_lo", C.c_int64),
                ("region_hi", C.c_int64)]


class MzmlPred(C.Structure):
    _fields_ = [("mz_lo", C.c_double), ("mz_hi", C.c_double)]


class Chunk(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64)]


class ArrowSchema(C.Structure):
    pass


ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64),
                        ("n_children", C.c_int64), ("children", C.POINTER(C.POINTER(ArrowSchema))),
                        ("dictionary", C.POINTER(ArrowSchema)),
                        ("release", C.CFUNCTYPE(None, C.POINTER(ArrowSchema))), ("private_data", C.c_void_p)]


class ArrowArray(C.Structure):
    pass


ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64),
                       ("n_buffers", C.c_int64), ("n_children", C.c_int64), ("buffers", C.POINTER(C.c_void_p)),
                       ("children", C.POINTER(C.POINTER(ArrowArray))), ("dictionary", C.POINTER(ArrowArray)),
                       ("release", C.CFUNCTYPE(None, C.POINTER(ArrowArray))), ("private_data", C.c_void_p)]


class Pred(C.Structure):
    _fields_ = [("chrom_col", C.c_int32), ("pos_col", C.c_int32), ("region", Region)]


class Agg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("value_col", C.c_int32)]


class Partial(C.Structure):
    _fields_ = [("count", C.c_int64), ("sum_i64", C.c_int64), ("sum_f64", C.c_double)]


_lib = None


def load():
    """Load libexon_gpu.so; raises if it has not been built (python -m exon_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("EXON_GPU_LIB", LIB_PATH)  # experiment builds of the same library (tools/sweep.py)
    if not os.path.exists(path):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m exon_b200.build` "
                          "(nvcc, sm_100a). exon_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.exon_gpu_last_error.restype = C.c_char_p
    L.exon_gpu_version.restype = C.c_char_p
    sigs = {
        "exon_gpu_ctx_create": [C.c_int, vp, C.POINTER(vp)],
        "exon_gpu_ctx_destroy": [vp],
        "exon_gpu_ctx_launch_count": [vp, C.POINTER(i64)],
        "exon_gpu_ctx_last_kernel_ms": [vp, C.POINTER(C.c_float)],
        "exon_gpu_ctx_kernel_ms_history": [vp, C.POINTER(C.c_float), i32, C.POINTER(i32)],
        "exon_gpu_ctx_synchronize": [vp],
        "exon_gpu_host_alloc": [vp, C.c_size_t, C.POINTER(vp)],
        "exon_gpu_host_free": [vp, vp],
        "exon_gpu_device_alloc": [vp, C.c_size_t, C.POINTER(vp)],
        "exon_gpu_device_free": [vp, vp],
        "exon_gpu_memcpy_h2d": [vp, vp, vp, C.c_size_t],
        "exon_gpu_memcpy_h2d_async": [vp, vp, vp, C.c_size_t],
        "exon_gpu_region_parse": [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(Region)],
        "exon_gpu_interval_parse": [C.c_char_p, C.POINTER(Region)],
        "exon_gpu_parse_f32": [C.c_char_p, C.c_size_t, C.POINTER(C.c_float)],
        "exon_gpu_regroup_files_by_size": [C.POINTER(i64), i32, i32, C.POINTER(i32), C.POINTER(i32)],
        "exon_gpu_vcf_open": [vp, C.POINTER(VcfOpts), C.POINTER(vp)],
        "exon_gpu_vcf_close": [vp],
        "exon_gpu_vcf_reset": [vp],
        "exon_gpu_vcf_feed": [vp, vp, C.c_size_t, C.c_int, C.c_int],
        "exon_gpu_vcf_next_batch": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)],
        "exon_gpu_vcf_filter_count": [vp, C.POINTER(Region), C.POINTER(i64)],
        "exon_gpu_vcf_filter_count_async": [vp, C.POINTER(Region), vp],
        "exon_gpu_vcf_filter_count_global": [vp, C.POINTER(Region), C.POINTER(i64), C.POINTER(i64)],
        "exon_gpu_vcf_rows": [vp, C.POINTER(i64)],
        "exon_gpu_vcf_body_bytes": [vp, C.POINTER(i64)],
        "exon_gpu_filter_agg": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema), C.c_int, C.POINTER(Pred),
                                C.POINTER(Agg), C.POINTER(Partial)],
        "exon_gpu_filter_agg_accumulate": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema), C.POINTER(Pred),
                                           C.POINTER(Agg), vp],
        "exon_gpu_filter_agg_batches": [vp, C.POINTER(C.POINTER(ArrowArray)), i32, C.POINTER(ArrowSchema), C.POINTER(Pred),
                                        C.POINTER(Agg), C.POINTER(Partial)],
        "exon_gpu_vcf_filter_agg": [vp, C.POINTER(Pred), C.POINTER(Agg), C.POINTER(Partial)],
        "exon_gpu_fastq_open": [vp, C.POINTER(FastqOpts), C.POINTER(vp)],
        "exon_gpu_fastq_feed": [vp, vp, C.c_size_t, C.c_int, C.c_int],
        "exon_gpu_fastq_filter_count": [vp, C.POINTER(FastqPred), C.POINTER(i64)],
        "exon_gpu_fastq_rows": [vp, C.POINTER(i64)],
        "exon_gpu_fastq_next_batch": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)],
        "exon_gpu_stream_feed_gzip": [vp, vp, C.c_size_t, C.c_int],
        "exon_gpu_gzip_inflate": [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_int, C.POINTER(C.c_size_t)],
        "exon_gpu_fasta_open_columns": [vp, C.POINTER(FastqOpts), C.POINTER(vp)],
        "exon_gpu_fasta_next_batch": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)],
        "exon_gpu_gff_open_columns": [vp, C.POINTER(FastqOpts), C.POINTER(vp)],
        "exon_gpu_gff_next_batch": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)],
        "exon_gpu_vcf_set_header": [vp, C.c_char_p, C.c_size_t],
        "exon_gpu_bam_open": [vp, C.POINTER(vp)],
        "exon_gpu_bam_open_columns": [vp, C.POINTER(FastqOpts), C.POINTER(vp)],  # exon_gpu_bam_opts has the layout of exon_gpu_fastq_opts
        "exon_gpu_bam_next_batch": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)],
        "exon_gpu_bam_feed": [vp, vp, C.c_size_t, C.c_int],
        "exon_gpu_bam_filter_count_by_reference": [vp, C.POINTER(BamPred), C.POINTER(i64), i32, C.POINTER(i32), C.POINTER(i64)],
        "exon_gpu_bam_group_name": [vp, i32, C.POINTER(C.c_char_p)],
        "exon_gpu_allreduce_counts": [vp, C.POINTER(i64), i32],
        "exon_gpu_mzml_open": [vp, C.POINTER(vp)],
        "exon_gpu_mzml_open_columns": [vp, C.POINTER(FastqOpts), C.POINTER(vp)],
        "exon_gpu_mzml_next_batch": [vp, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)],
        "exon_gpu_mzml_feed": [vp, vp, C.c_size_t, C.c_int, C.c_int],
        "exon_gpu_mzml_filter_sum": [vp, C.POINTER(MzmlPred), C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i64)],
        "exon_gpu_tabix_query": [vp, vp, C.c_size_t, C.POINTER(Region), C.POINTER(Chunk), i32, C.POINTER(i32)],
        "exon_gpu_stream_feed_bgzf_chunk": [vp, vp, C.c_size_t, C.c_uint64, C.POINTER(Chunk)],
        "exon_gpu_fasta_open": [vp, C.POINTER(vp)],
        "exon_gpu_fasta_feed": [vp, vp, C.c_size_t, C.c_int, C.c_int],
        "exon_gpu_fasta_rows": [vp, C.POINTER(i64)],
        "exon_gpu_gff_open": [vp, C.POINTER(vp)],
        "exon_gpu_gff_feed": [vp, vp, C.c_size_t, C.c_int, C.c_int],
        "exon_gpu_gff_filter_count": [vp, C.POINTER(Region), C.POINTER(i64)],
        "exon_gpu_stream_close": [vp],
        "exon_gpu_stream_schema": [vp, C.POINTER(ArrowSchema)],
        "exon_gpu_stream_export": [vp, vp, C.c_int],
        "exon_gpu_stream_reset": [vp],
        "exon_gpu_stream_body_bytes": [vp, C.POINTER(i64)],
        "exon_gpu_partial_read": [vp, vp, C.c_int, C.POINTER(Partial)],
        "exon_gpu_memset": [vp, vp, C.c_int, C.c_size_t],
        "exon_gpu_region_udf": [vp, C.c_int, C.POINTER(ArrowArray), C.POINTER(ArrowSchema), C.c_int, C.POINTER(Pred), vp, vp],
        "exon_gpu_nccl_unique_id": [C.c_char_p],
        "exon_gpu_nccl_init": [vp, C.c_char_p, C.c_int, C.c_int],
        "exon_gpu_allreduce_partial": [vp, C.POINTER(Partial)],
    }
    for name, argtypes in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        raise ExonGpuError(rc, load().exon_gpu_last_error().decode(errors="replace"))


def make_region(chrom=None, lo=None, hi=None) -> Region | None:
    """exon_gpu_region for `chrom = <chrom> AND pos BETWEEN lo AND hi` (None drops a term; all None -> NULL)."""
    if chrom is None and lo is None and hi is None:
        return None
    r = Region()
    if chrom is not None:
        b = chrom.encode() if isinstance(chrom, str) else bytes(chrom)
        r.chrom = b
        r._keep = b
        r.chrom_len = len(b)
        r.has_chrom = 1
    if lo is not None or hi is not None:
        r.has_interval = 1
        r.lo = 1 if lo is None else int(lo)
        r.hi = INT64_MAX if hi is None else int(hi)
    else:
        r.lo, r.hi = 1, INT64_MAX
    return r
