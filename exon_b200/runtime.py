"""Thin object layer over the C ABI: contexts, pinned buffers, partition streams, Arrow batch import.

These classes add no logic of their own -- each method is one exon_gpu_* call (plus buffer bookkeeping), so
the tests that go through them are tests of the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._abi import ExonGpuError, check  # noqa: F401


class PinnedBuffer:
    """Page-locked host memory from exon_gpu_host_alloc, exposed as a numpy uint8 array."""

    def __init__(self, ctx: "Context", nbytes: int):
        self._ctx = ctx
        self._ptr = C.c_void_p()
        check(ctx.lib.exon_gpu_host_alloc(ctx.handle, max(int(nbytes), 1), C.byref(self._ptr)))
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(C.c_uint8)), (max(self.nbytes, 1),))[: self.nbytes]

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def free(self):
        if self._ptr:
            self.array = None
            check(self._ctx.lib.exon_gpu_host_free(self._ctx.handle, self._ptr))
            self._ptr = C.c_void_p()


class DeviceBuffer:
    def __init__(self, ctx: "Context", nbytes: int):
        self._ctx = ctx
        self._ptr = C.c_void_p()
        check(ctx.lib.exon_gpu_device_alloc(ctx.handle, max(int(nbytes), 1), C.byref(self._ptr)))
        self.nbytes = int(nbytes)

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def upload(self, host: np.ndarray, offset: int = 0):
        assert host.dtype == np.uint8 and host.flags.c_contiguous and offset + host.size <= self.nbytes
        check(self._ctx.lib.exon_gpu_memcpy_h2d(self._ctx.handle, self.ptr + offset, host.ctypes.data, host.size))

    def upload_async(self, host: np.ndarray, offset: int = 0):
        """Enqueue the copy without waiting (pinned source; Context.synchronize() completes it)."""
        assert host.dtype == np.uint8 and host.flags.c_contiguous and offset + host.size <= self.nbytes
        check(self._ctx.lib.exon_gpu_memcpy_h2d_async(self._ctx.handle, self.ptr + offset, host.ctypes.data, host.size))

    def free(self):
        if self._ptr:
            check(self._ctx.lib.exon_gpu_device_free(self._ctx.handle, self._ptr))
            self._ptr = C.c_void_p()


class Context:
    """exon_gpu_ctx: one per (process, device)."""

    def __init__(self, device: int = 0, cuda_stream: int | None = None):
        self.lib = _abi.load()
        self.handle = C.c_void_p()
        check(self.lib.exon_gpu_ctx_create(int(device), C.c_void_p(cuda_stream) if cuda_stream else None,
                                           C.byref(self.handle)))
        self.device = device

    def close(self):
        if self.handle:
            self.lib.exon_gpu_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def launch_count(self) -> int:
        v = C.c_int64()
        check(self.lib.exon_gpu_ctx_launch_count(self.handle, C.byref(v)))
        return v.value

    def last_kernel_ms(self) -> float:
        v = C.c_float()
        check(self.lib.exon_gpu_ctx_last_kernel_ms(self.handle, C.byref(v)))
        return v.value

    def kernel_ms_history(self, n: int = 64) -> list[float]:
        """Device times of the most recent timed launches, oldest first (the library keeps 64)."""
        buf = (C.c_float * max(n, 1))()
        got = C.c_int32()
        check(self.lib.exon_gpu_ctx_kernel_ms_history(self.handle, buf, n, C.byref(got)))
        return list(buf)[: got.value]

    def synchronize(self):
        check(self.lib.exon_gpu_ctx_synchronize(self.handle))

    def pinned(self, nbytes: int) -> PinnedBuffer:
        return PinnedBuffer(self, nbytes)

    def device_buffer(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def open_vcf(self, **kw) -> "VcfStream":
        return VcfStream(self, **kw)

    def gzip_inflate(self, data) -> np.ndarray:
        """exon_gpu_gzip_inflate: the uncompressed bytes of a whole BGZF / gzip file, inflated on the device."""
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        n = C.c_size_t()
        rc = self.lib.exon_gpu_gzip_inflate(self.handle, C.c_void_p(data.ctypes.data), data.size, None, 0, 0, C.byref(n))
        if n.value == 0:
            check(rc)
            return np.zeros(0, np.uint8)
        out = np.empty(n.value, dtype=np.uint8)
        check(self.lib.exon_gpu_gzip_inflate(self.handle, C.c_void_p(data.ctypes.data), data.size, C.c_void_p(out.ctypes.data),
                                             out.size, 0, C.byref(n)))
        return out

    def tabix_query(self, tbi, region: "_abi.Region"):
        """exon_gpu_tabix_query: [(start vpos, end vpos)] of the chunks of a .tbi that can hold records of `region`."""
        if isinstance(tbi, (bytes, bytearray, memoryview)):
            tbi = np.frombuffer(tbi, dtype=np.uint8)
        n = C.c_int32()
        check(self.lib.exon_gpu_tabix_query(self.handle, C.c_void_p(tbi.ctypes.data), tbi.size, C.byref(region), None, 0, C.byref(n)))
        out = (_abi.Chunk * max(n.value, 1))()
        check(self.lib.exon_gpu_tabix_query(self.handle, C.c_void_p(tbi.ctypes.data), tbi.size, C.byref(region), out, n.value, C.byref(n)))
        return [(int(out[i].start), int(out[i].end)) for i in range(n.value)]

    def open_gff(self, **kw) -> "GffStream":
        return GffStream(self, **kw)

    def open_fasta(self, **kw) -> "FastaStream":
        return FastaStream(self, **kw)

    def open_mzml(self, **kw) -> "MzmlStream":
        return MzmlStream(self, **kw)

    def open_bam(self, **kw) -> "BamStream":
        return BamStream(self, **kw)

    def allreduce_counts(self, counts):
        """exon_gpu_allreduce_counts: element-wise sum of an int64 vector over the NCCL communicator."""
        n = len(counts)
        buf = (C.c_int64 * max(n, 1))(*[int(x) for x in counts])
        check(self.lib.exon_gpu_allreduce_counts(self.handle, buf, n))
        return list(buf)[:n]

    def open_fastq(self, **kw) -> "FastqStream":
        return FastqStream(self, **kw)

    # ---- multi-GPU final aggregate ----
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(_abi.NCCL_ID_BYTES)
        check(self.lib.exon_gpu_nccl_unique_id(buf))
        return buf.raw

    def nccl_init(self, unique_id: bytes, n_ranks: int, rank: int):
        check(self.lib.exon_gpu_nccl_init(self.handle, unique_id, n_ranks, rank))

    def allreduce_partial(self, count: int = 0, sum_i64: int = 0, sum_f64: float = 0.0):
        p = _abi.Partial(count, sum_i64, sum_f64)
        check(self.lib.exon_gpu_allreduce_partial(self.handle, C.byref(p)))
        return p.count, p.sum_i64, p.sum_f64

    # ---- K3 ----
    def filter_agg(self, arrow_array: "_abi.ArrowArray", arrow_schema: "_abi.ArrowSchema", *, on_device: bool,
                   chrom_col: int = -1, pos_col: int = -1, region: "_abi.Region | None" = None,
                   kind: int = _abi.AGG_COUNT_STAR, value_col: int = -1):
        pred = _abi.Pred(chrom_col, pos_col, region if region is not None else _abi.Region())
        agg = _abi.Agg(kind, value_col)
        out = _abi.Partial()
        check(self.lib.exon_gpu_filter_agg(self.handle, C.byref(arrow_array), C.byref(arrow_schema), int(on_device),
                                           C.byref(pred), C.byref(agg), C.byref(out)))
        return out.count, out.sum_i64, out.sum_f64


    def filter_agg_accumulate(self, arrow_array, arrow_schema, device_acc_ptr: int, *, chrom_col: int = -1,
                              pos_col: int = -1, region=None, kind: int = _abi.AGG_COUNT_STAR, value_col: int = -1):
        pred = _abi.Pred(chrom_col, pos_col, region if region is not None else _abi.Region())
        agg = _abi.Agg(kind, value_col)
        check(self.lib.exon_gpu_filter_agg_accumulate(self.handle, C.byref(arrow_array), C.byref(arrow_schema),
                                                      C.byref(pred), C.byref(agg), C.c_void_p(device_acc_ptr)))

    def filter_agg_batches(self, batches, *, chrom_col: int = -1, pos_col: int = -1, region=None,
                           kind: int = _abi.AGG_COUNT_STAR, value_col: int = -1):
        """One launch over a list of device-resident VcfBatch objects (they share the first batch's schema)."""
        n = len(batches)
        arr = (C.POINTER(_abi.ArrowArray) * max(n, 1))(*[C.pointer(b.c_array) for b in batches])
        pred = _abi.Pred(chrom_col, pos_col, region if region is not None else _abi.Region())
        agg = _abi.Agg(kind, value_col)
        out = _abi.Partial()
        check(self.lib.exon_gpu_filter_agg_batches(self.handle, arr, n, C.byref(batches[0].c_schema), C.byref(pred),
                                                   C.byref(agg), C.byref(out)))
        return out.count, out.sum_i64, out.sum_f64

    def partial_read(self, device_acc_ptr: int, sum_is_integer: bool = True):
        out = _abi.Partial()
        check(self.lib.exon_gpu_partial_read(self.handle, C.c_void_p(device_acc_ptr), int(sum_is_integer), C.byref(out)))
        return out.count, out.sum_i64, out.sum_f64

    def memset(self, device_ptr: int, value: int, nbytes: int):
        check(self.lib.exon_gpu_memset(self.handle, C.c_void_p(device_ptr), value, nbytes))

    def region_udf(self, kind: int, arrow_array, arrow_schema, *, on_device: bool, chrom_col: int = -1, pos_col: int = -1,
                   region=None):
        """(values bool[n], valid bool[n]) of region_match / chrom_match / interval_match over one batch."""
        n = int(arrow_array.length)
        vals, valid = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.uint8)
        pred = _abi.Pred(chrom_col, pos_col, region if region is not None else _abi.Region())
        check(self.lib.exon_gpu_region_udf(self.handle, kind, C.byref(arrow_array), C.byref(arrow_schema), int(on_device),
                                           C.byref(pred), C.c_void_p(vals.ctypes.data), C.c_void_p(valid.ctypes.data)))
        return vals[:n].astype(bool), valid[:n].astype(bool)


class VcfBatch:
    """One record batch from exon_gpu_vcf_next_batch, imported into numpy (host columns only)."""

    def __init__(self, arr: _abi.ArrowArray, schema: _abi.ArrowSchema, on_device: bool):
        self._arr, self._schema = arr, schema
        self.on_device = on_device
        self.num_rows = int(arr.length)
        self.names = [schema.children[i].contents.name.decode() for i in range(schema.n_children)]
        self.formats = [schema.children[i].contents.format.decode() for i in range(schema.n_children)]

    def buffer_ptrs(self, name: str):
        ch = self._arr.children[self.names.index(name)].contents
        return [ch.buffers[i] for i in range(ch.n_buffers)]

    def column(self, name: str):
        """chrom -> (offsets int32[rows+1], values uint8[...]); pos -> int64[rows]  (copies)."""
        assert not self.on_device, "device-resident batch: read it with a CUDA consumer"
        i = self.names.index(name)
        ch = self._arr.children[i].contents
        n = self.num_rows
        if self.formats[i] == "u":
            off = np.ctypeslib.as_array(C.cast(ch.buffers[1], C.POINTER(C.c_int32)), (n + 1,)).copy()
            nv = int(off[-1])
            val = (np.ctypeslib.as_array(C.cast(ch.buffers[2], C.POINTER(C.c_uint8)), (max(nv, 1),))[:nv].copy()
                   if nv else np.zeros(0, np.uint8))
            return off, val
        if self.formats[i] == "l":
            return np.ctypeslib.as_array(C.cast(ch.buffers[1], C.POINTER(C.c_int64)), (max(n, 1),))[:n].copy()
        raise NotImplementedError(self.formats[i])

    def strings(self, name: str):
        """utf8 column as a list of bytes (None where the validity bitmap, if any, marks NULL)."""
        off, val = self.column(name)
        ch = self._arr.children[self.names.index(name)].contents
        b = val.tobytes()
        out = [b[off[i]:off[i + 1]] for i in range(self.num_rows)]
        if ch.buffers[0]:
            bits = np.ctypeslib.as_array(C.cast(ch.buffers[0], C.POINTER(C.c_uint8)), ((self.num_rows + 7) // 8,))
            out = [x if (bits[i >> 3] >> (i & 7)) & 1 else None for i, x in enumerate(out)]
        return out

    def to_pyarrow(self):
        """Import the batch into pyarrow through the Arrow C Data Interface (host batches only).  Ownership moves to
        pyarrow: the library's release callback runs when the pyarrow batch is dropped."""
        assert not self.on_device, "device-resident batch: read it with a CUDA consumer"
        import pyarrow as pa

        rb = pa.RecordBatch._import_from_c(C.addressof(self._arr), C.addressof(self._schema))
        rb.validate(full=True)
        return rb

    def chrom_strings(self):
        off, val = self.column("chrom")
        b = val.tobytes()
        return [b[off[i]:off[i + 1]].decode() for i in range(self.num_rows)]

    def release(self):
        if self._arr is not None and self._arr.release:
            self._arr.release(C.byref(self._arr))
        if self._schema is not None and self._schema.release:
            self._schema.release(C.byref(self._schema))
        self._arr = self._schema = None

    @property
    def c_array(self):
        return self._arr

    @property
    def c_schema(self):
        return self._schema


class ArrowArrayStream(C.Structure):
    _fields_ = [("get_schema", C.c_void_p), ("get_next", C.c_void_p), ("get_last_error", C.c_void_p), ("release", C.c_void_p),
                ("private_data", C.c_void_p)]


def export_reader(stream, take_ownership: bool = False):
    """exon_gpu_stream_export -> pyarrow.RecordBatchReader (the Arrow C stream interface, the reference's own FFI shape)."""
    import pyarrow as pa

    c = ArrowArrayStream()
    check(stream.lib.exon_gpu_stream_export(stream.handle, C.byref(c), int(take_ownership)))
    if take_ownership:
        stream.handle = C.c_void_p()
    return pa.RecordBatchReader._import_from_c(C.addressof(c))


class VcfStream:
    """exon_gpu_stream: one DataFusion partition stream over a group of VCF files."""

    def __init__(self, ctx: Context, *, batch_rows: int = 8192, projection=(0, 1), columns_on_device: bool = False,
                 pushdown: "_abi.Region | None" = None, strict: bool = False, kernel_variant: int = 0):
        self.ctx = ctx
        self.lib = ctx.lib
        self._proj = (C.c_int32 * len(projection))(*projection)
        self._pushdown = pushdown
        opts = _abi.VcfOpts(batch_rows, len(projection), self._proj, int(columns_on_device),
                            C.pointer(pushdown) if pushdown is not None else None, int(strict), kernel_variant)
        self.handle = C.c_void_p()
        check(self.lib.exon_gpu_vcf_open(ctx.handle, C.byref(opts), C.byref(self.handle)))
        self.columns_on_device = columns_on_device

    def close(self):
        if self.handle:
            self.lib.exon_gpu_vcf_close(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        check(self.lib.exon_gpu_vcf_reset(self.handle))

    def set_header(self, text: bytes):
        """exon_gpu_vcf_set_header: the header text (its ##INFO lines give the info column its types)."""
        check(self.lib.exon_gpu_vcf_set_header(self.handle, bytes(text), len(text)))

    def feed(self, data, *, is_last: bool = True, device_ptr: int | None = None, nbytes: int | None = None):
        """Feed host bytes (bytes / numpy uint8 / PinnedBuffer) or a raw device range (device_ptr, nbytes)."""
        if device_ptr is not None:
            check(self.lib.exon_gpu_vcf_feed(self.handle, C.c_void_p(device_ptr), int(nbytes), 1, int(is_last)))
            return
        if isinstance(data, PinnedBuffer):
            data = data.array
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data  # keep alive until the next synchronising call
        check(self.lib.exon_gpu_vcf_feed(self.handle, C.c_void_p(data.ctypes.data), data.size, 0, int(is_last)))

    def feed_gzip(self, data, *, is_last: bool = True):
        """Feed BGZF / gzip bytes of one file (exon_gpu_stream_feed_gzip): inflated on the device."""
        if isinstance(data, PinnedBuffer):
            data = data.array
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data
        check(self.lib.exon_gpu_stream_feed_gzip(self.handle, C.c_void_p(data.ctypes.data), data.size, int(is_last)))

    def feed_bgzf_chunk(self, data, chunk, file_offset: int = 0):
        """exon_gpu_stream_feed_bgzf_chunk: bytes [file_offset, ...) of a .vcf.gz and one (start, end) virtual-position chunk."""
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        self._last_host = data
        ch = _abi.Chunk(int(chunk[0]), int(chunk[1]))
        check(self.lib.exon_gpu_stream_feed_bgzf_chunk(self.handle, C.c_void_p(data.ctypes.data), data.size, int(file_offset), C.byref(ch)))

    def filter_count(self, region: "_abi.Region | None" = None) -> int:
        out = C.c_int64()
        check(self.lib.exon_gpu_vcf_filter_count(self.handle, C.byref(region) if region is not None else None,
                                                 C.byref(out)))
        return out.value

    def filter_count_global(self, region: "_abi.Region | None" = None):
        """(local, global) counts; collective over the context's NCCL communicator."""
        loc, glob = C.c_int64(), C.c_int64()
        check(self.lib.exon_gpu_vcf_filter_count_global(self.handle, C.byref(region) if region is not None else None,
                                                        C.byref(loc), C.byref(glob)))
        return loc.value, glob.value

    def filter_count_async(self, region, device_out_ptr: int):
        check(self.lib.exon_gpu_vcf_filter_count_async(self.handle, C.byref(region) if region is not None else None,
                                                       C.c_void_p(device_out_ptr)))

    def rows(self) -> int:
        out = C.c_int64()
        check(self.lib.exon_gpu_vcf_rows(self.handle, C.byref(out)))
        return out.value

    def body_bytes(self) -> int:
        out = C.c_int64()
        check(self.lib.exon_gpu_vcf_body_bytes(self.handle, C.byref(out)))
        return out.value

    def filter_agg(self, *, chrom_col: int = -1, pos_col: int = -1, region=None, kind: int = _abi.AGG_COUNT_STAR,
                   value_col: int = -1):
        """exon_gpu_vcf_filter_agg: (count, sum_i64, sum_f64) over every batch of this stream, columns kept in HBM."""
        pred = _abi.Pred(chrom_col, pos_col, region if region is not None else _abi.Region())
        agg = _abi.Agg(kind, value_col)
        out = _abi.Partial()
        check(self.lib.exon_gpu_vcf_filter_agg(self.handle, C.byref(pred), C.byref(agg), C.byref(out)))
        return out.count, out.sum_i64, out.sum_f64

    def next_batch(self) -> VcfBatch | None:
        arr, sch = _abi.ArrowArray(), _abi.ArrowSchema()
        check(self.lib.exon_gpu_vcf_next_batch(self.handle, C.byref(arr), C.byref(sch)))
        if not arr.release:
            if sch.release:
                sch.release(C.byref(sch))
            return None
        return VcfBatch(arr, sch, self.columns_on_device)

    def batches(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b


class FastqStream:
    """exon_gpu_stream opened with exon_gpu_fastq_open: one partition stream over a group of FASTQ files."""

    def __init__(self, ctx: Context, *, batch_rows: int = 8192, projection=(), columns_on_device: bool = False):
        self.ctx = ctx
        self.lib = ctx.lib
        self._proj = (C.c_int32 * max(len(projection), 1))(*projection)
        opts = _abi.FastqOpts(batch_rows, len(projection), self._proj, int(columns_on_device))
        self.handle = C.c_void_p()
        check(self.lib.exon_gpu_fastq_open(ctx.handle, C.byref(opts), C.byref(self.handle)))
        self.columns_on_device = columns_on_device

    def next_batch(self):
        arr, sch = _abi.ArrowArray(), _abi.ArrowSchema()
        check(self.lib.exon_gpu_fastq_next_batch(self.handle, C.byref(arr), C.byref(sch)))
        if not arr.release:
            if sch.release:
                sch.release(C.byref(sch))
            return None
        return VcfBatch(arr, sch, self.columns_on_device)

    def batches(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b

    def close(self):
        if self.handle:
            self.lib.exon_gpu_stream_close(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        check(self.lib.exon_gpu_stream_reset(self.handle))

    def feed(self, data, *, is_last: bool = True, device_ptr: int | None = None, nbytes: int | None = None):
        if device_ptr is not None:
            check(self.lib.exon_gpu_fastq_feed(self.handle, C.c_void_p(device_ptr), int(nbytes), 1, int(is_last)))
            return
        if isinstance(data, PinnedBuffer):
            data = data.array
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data
        check(self.lib.exon_gpu_fastq_feed(self.handle, C.c_void_p(data.ctypes.data), data.size, 0, int(is_last)))

    def feed_gzip(self, data, *, is_last: bool = True):
        """Feed BGZF / gzip bytes of one file (exon_gpu_stream_feed_gzip): inflated on the device."""
        if isinstance(data, PinnedBuffer):
            data = data.array
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data
        check(self.lib.exon_gpu_stream_feed_gzip(self.handle, C.c_void_p(data.ctypes.data), data.size, int(is_last)))

    def filter_count(self, min_mean=None, phred_offset: int = 33) -> int:
        """records with mean(quality) > min_mean (an int, or a (num, den) pair); None -> COUNT(*)."""
        out = C.c_int64()
        if min_mean is None:
            check(self.lib.exon_gpu_fastq_filter_count(self.handle, None, C.byref(out)))
            return out.value
        num, den = min_mean if isinstance(min_mean, tuple) else (int(min_mean), 1)
        pred = _abi.FastqPred(phred_offset, 0, num, den)
        check(self.lib.exon_gpu_fastq_filter_count(self.handle, C.byref(pred), C.byref(out)))
        return out.value

    def rows(self) -> int:
        out = C.c_int64()
        check(self.lib.exon_gpu_fastq_rows(self.handle, C.byref(out)))
        return out.value

    def body_bytes(self) -> int:
        out = C.c_int64()
        check(self.lib.exon_gpu_stream_body_bytes(self.handle, C.byref(out)))
        return out.value


class BamStream:
    """exon_gpu_stream opened with exon_gpu_bam_open: one partition stream over a group of .bam files."""

    # SAM flag bits (exon/exon-core/src/udfs/sam/samflags.rs:111-141)
    UNMAPPED, SECONDARY, SUPPLEMENTARY = 0x4, 0x100, 0x800

    def __init__(self, ctx: Context, *, batch_rows: int = 8192, projection=None, columns_on_device: bool = False):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = C.c_void_p()
        self.columns_on_device = columns_on_device
        if projection is None:
            check(self.lib.exon_gpu_bam_open(ctx.handle, C.byref(self.handle)))
        else:
            self._proj = (C.c_int32 * max(len(projection), 1))(*projection)
            opts = _abi.FastqOpts(batch_rows, len(projection), self._proj, int(columns_on_device))
            check(self.lib.exon_gpu_bam_open_columns(ctx.handle, C.byref(opts), C.byref(self.handle)))

    def next_batch(self):
        """exon_gpu_bam_next_batch -> VcfBatch (the Arrow import helper is format-agnostic) or None at the end."""
        arr, sch = _abi.ArrowArray(), _abi.ArrowSchema()
        check(self.lib.exon_gpu_bam_next_batch(self.handle, C.byref(arr), C.byref(sch)))
        if not arr.release:
            if sch.release:
                sch.release(C.byref(sch))
            return None
        return VcfBatch(arr, sch, self.columns_on_device)

    def batches(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b

    def close(self):
        if self.handle:
            self.lib.exon_gpu_stream_close(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        check(self.lib.exon_gpu_stream_reset(self.handle))

    def feed(self, data, *, is_last: bool = True):
        if isinstance(data, PinnedBuffer):
            data = data.array
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data
        check(self.lib.exon_gpu_bam_feed(self.handle, C.c_void_p(data.ctypes.data), data.size, int(is_last)))

    def count_by_reference(self, *, flag_exclude: int = 0, flag_require: int = 0, min_mapq: int = -1, all_rows: bool = False,
                           region=None):
        """({reference name | None: count}, rows scanned): SELECT reference, COUNT(*) ... GROUP BY reference."""
        n = C.c_int32()
        rows = C.c_int64()
        pred = None
        if not all_rows:
            pred = _abi.BamPred(flag_exclude, flag_require, min_mapq, 0, None, 0, 0, 1, _abi.INT64_MAX)
            if region is not None:  # (reference name, lo, hi): bam_region_filter
                name = region[0].encode()
                pred.has_region, pred.region_ref, pred.region_ref_len = 1, name, len(name)
                pred.region_lo = 1 if region[1] is None else int(region[1])
                pred.region_hi = _abi.INT64_MAX if region[2] is None else int(region[2])
        pp = C.byref(pred) if pred is not None else None
        rc = self.lib.exon_gpu_bam_filter_count_by_reference(self.handle, pp, None, 0, C.byref(n), C.byref(rows))
        check(rc)
        counts = (C.c_int64 * max(n.value, 1))()
        check(self.lib.exon_gpu_bam_filter_count_by_reference(self.handle, pp, counts, n.value, C.byref(n), C.byref(rows)))
        out = {}
        for g in range(n.value):
            nm = C.c_char_p()
            check(self.lib.exon_gpu_bam_group_name(self.handle, g, C.byref(nm)))
            out[nm.value.decode() if nm.value is not None else None] = int(counts[g])
        return out, int(rows.value)


class MzmlStream(FastqStream):
    """exon_gpu_stream opened with exon_gpu_mzml_open (feeds / reset / close as for the text formats)."""

    def __init__(self, ctx: Context, *, batch_rows: int = 8192, projection=None, columns_on_device: bool = False):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = C.c_void_p()
        self.columns_on_device = columns_on_device
        if projection is None:
            check(self.lib.exon_gpu_mzml_open(ctx.handle, C.byref(self.handle)))
        else:
            self._proj = (C.c_int32 * max(len(projection), 1))(*projection)
            opts = _abi.FastqOpts(batch_rows, len(projection), self._proj, int(columns_on_device))
            check(self.lib.exon_gpu_mzml_open_columns(ctx.handle, C.byref(opts), C.byref(self.handle)))

    def next_batch(self):
        """exon_gpu_mzml_next_batch -> VcfBatch (import it with to_pyarrow(): the columns are nested) or None at the end."""
        arr, sch = _abi.ArrowArray(), _abi.ArrowSchema()
        check(self.lib.exon_gpu_mzml_next_batch(self.handle, C.byref(arr), C.byref(sch)))
        if not arr.release:
            if sch.release:
                sch.release(C.byref(sch))
            return None
        return VcfBatch(arr, sch, self.columns_on_device)

    def batches(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b

    def feed(self, data, *, is_last: bool = True, device_ptr: int | None = None, nbytes: int | None = None):
        if device_ptr is not None:
            check(self.lib.exon_gpu_mzml_feed(self.handle, C.c_void_p(device_ptr), int(nbytes), 1, int(is_last)))
            return
        if isinstance(data, PinnedBuffer):
            data = data.array
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data
        check(self.lib.exon_gpu_mzml_feed(self.handle, C.c_void_p(data.ctypes.data), data.size, 0, int(is_last)))

    def filter_sum(self, lo=None, hi=None):
        """(sum, peaks selected, spectra): SUM(intensity) over zipped peaks with lo <= mz <= hi (None: all)."""
        s, n, sp = C.c_double(), C.c_int64(), C.c_int64()
        pred = _abi.MzmlPred(float(lo), float(hi)) if lo is not None else None
        check(self.lib.exon_gpu_mzml_filter_sum(self.handle, C.byref(pred) if pred is not None else None, C.byref(s), C.byref(n),
                                                C.byref(sp)))
        return s.value, n.value, sp.value

    def filter_count(self, *a, **k):
        raise NotImplementedError

    def rows(self):
        return self.filter_sum()[2]


class FastaStream(MzmlStream):
    """exon_gpu_stream opened with exon_gpu_fasta_open: COUNT(*) of FASTA records."""

    def __init__(self, ctx: Context, *, batch_rows: int = 8192, projection=None, columns_on_device: bool = False):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = C.c_void_p()
        self.columns_on_device = columns_on_device
        if projection is None:
            check(self.lib.exon_gpu_fasta_open(ctx.handle, C.byref(self.handle)))
        else:
            self._proj = (C.c_int32 * max(len(projection), 1))(*projection)
            opts = _abi.FastqOpts(batch_rows, len(projection), self._proj, int(columns_on_device))
            check(self.lib.exon_gpu_fasta_open_columns(ctx.handle, C.byref(opts), C.byref(self.handle)))

    def next_batch(self):
        arr, sch = _abi.ArrowArray(), _abi.ArrowSchema()
        check(self.lib.exon_gpu_fasta_next_batch(self.handle, C.byref(arr), C.byref(sch)))
        if not arr.release:
            if sch.release:
                sch.release(C.byref(sch))
            return None
        return VcfBatch(arr, sch, self.columns_on_device)

    def batches(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b

    def feed_gzip(self, data, *, is_last: bool = True):
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        self._last_host = data
        check(self.lib.exon_gpu_stream_feed_gzip(self.handle, C.c_void_p(data.ctypes.data), data.size, int(is_last)))

    def feed(self, data, *, is_last: bool = True, device_ptr: int | None = None, nbytes: int | None = None):
        if device_ptr is not None:
            check(self.lib.exon_gpu_fasta_feed(self.handle, C.c_void_p(device_ptr), int(nbytes), 1, int(is_last)))
            return
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data
        check(self.lib.exon_gpu_fasta_feed(self.handle, C.c_void_p(data.ctypes.data), data.size, 0, int(is_last)))

    def filter_sum(self, *a, **k):
        raise NotImplementedError

    def rows(self) -> int:
        out = C.c_int64()
        check(self.lib.exon_gpu_fasta_rows(self.handle, C.byref(out)))
        return out.value


class GffStream(FastaStream):
    """exon_gpu_stream opened with exon_gpu_gff_open: COUNT(*) / gff_region_filter counts of GFF records."""

    def __init__(self, ctx: Context, *, projection=None, columns_on_device: bool = False):
        self.ctx = ctx
        self.lib = ctx.lib
        self.handle = C.c_void_p()
        self.columns_on_device = columns_on_device
        if projection is None:
            check(self.lib.exon_gpu_gff_open(ctx.handle, C.byref(self.handle)))
        else:
            self._proj = (C.c_int32 * max(len(projection), 1))(*projection)
            opts = _abi.FastqOpts(0, len(projection), self._proj, int(columns_on_device))
            check(self.lib.exon_gpu_gff_open_columns(ctx.handle, C.byref(opts), C.byref(self.handle)))

    def next_batch(self):
        arr, sch = _abi.ArrowArray(), _abi.ArrowSchema()
        check(self.lib.exon_gpu_gff_next_batch(self.handle, C.byref(arr), C.byref(sch)))
        if not arr.release:
            if sch.release:
                sch.release(C.byref(sch))
            return None
        return VcfBatch(arr, sch, self.columns_on_device)

    def feed(self, data, *, is_last: bool = True, device_ptr: int | None = None, nbytes: int | None = None):
        if device_ptr is not None:
            check(self.lib.exon_gpu_gff_feed(self.handle, C.c_void_p(device_ptr), int(nbytes), 1, int(is_last)))
            return
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        self._last_host = data
        check(self.lib.exon_gpu_gff_feed(self.handle, C.c_void_p(data.ctypes.data), data.size, 0, int(is_last)))

    def filter_count(self, region: "_abi.Region | None" = None) -> int:
        out = C.c_int64()
        check(self.lib.exon_gpu_gff_filter_count(self.handle, C.byref(region) if region is not None else None, C.byref(out)))
        return out.value

    def rows(self) -> int:
        return self.filter_count(None)
