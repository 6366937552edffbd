"""ctypes wrapper over oracle/liboracle.so -- the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg.  The product package (exon_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
INT64_MAX = (1 << 63) - 1


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class Region(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("name_len", C.c_int32), ("has_interval", C.c_int32),
                ("lo", C.c_int64), ("hi", C.c_int64)]


class VcfBatch(C.Structure):
    _fields_ = [("rows", C.c_int64), ("chrom_offsets", C.POINTER(C.c_int32)), ("chrom_values", C.POINTER(C.c_uint8)),
                ("chrom_values_len", C.c_int64), ("pos", C.POINTER(C.c_int64))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.exo_vcf_header_len.restype = C.c_int64
        L.exo_vcf_header_len.argtypes = [C.c_void_p, C.c_int64]
        L.exo_region_parse.restype = C.c_int
        L.exo_region_parse.argtypes = [C.c_char_p, C.POINTER(Region)]
        L.exo_interval_parse.restype = C.c_int
        L.exo_interval_parse.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.exo_vcf_reader_open.restype = C.c_void_p
        L.exo_vcf_reader_open.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.c_int32]
        L.exo_vcf_reader_next.restype = C.c_int
        L.exo_vcf_reader_next.argtypes = [C.c_void_p, C.POINTER(VcfBatch)]
        L.exo_vcf_reader_close.argtypes = [C.c_void_p]
        L.exo_vcf_reader_err_row.restype = C.c_int64
        L.exo_vcf_reader_err_row.argtypes = [C.c_void_p]
        L.exo_vcf_filter_count.restype = C.c_int64
        L.exo_vcf_filter_count.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_char_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        L.exo_vcf_filter_count_files.restype = C.c_int64
        L.exo_vcf_filter_count_files.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_int32,
                                                 C.c_int64, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                                 C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.exo_regroup_files_by_size.restype = C.c_int32
        L.exo_regroup_files_by_size.argtypes = [C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.exo_region_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Region), C.c_void_p]
        L.exo_chrom_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int32, C.c_void_p]
        L.exo_interval_match.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    return _lib


def _buf(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        return data
    return np.frombuffer(bytes(data), dtype=np.uint8)


def header_len(data) -> int:
    a = _buf(data)
    return int(lib().exo_vcf_header_len(a.ctypes.data, a.size))


def parse_region(s: str) -> Region:
    r = Region()
    rc = lib().exo_region_parse(s.encode(), C.byref(r))
    if rc != 0:
        raise ValueError(f"bad region {s!r}")
    return r


def parse_interval(s: str):
    lo, hi = C.c_int64(), C.c_int64()
    b = s.encode()
    if not lib().exo_interval_parse(b, len(b), C.byref(lo), C.byref(hi)):
        raise ValueError(f"bad interval {s!r}")
    return lo.value, hi.value


def read_batches(data, batch_size: int = 8192, projection=(0, 1)):
    """Yield dicts {rows, chrom_offsets, chrom_values, pos} (numpy copies), one per reference batch."""
    a = _buf(data)
    proj = (C.c_int32 * len(projection))(*projection)
    r = lib().exo_vcf_reader_open(a.ctypes.data, a.size, batch_size, proj, len(projection))
    try:
        b = VcfBatch()
        while True:
            rc = lib().exo_vcf_reader_next(r, C.byref(b))
            if rc == 0:
                return
            if rc < 0:
                raise ValueError(f"malformed VCF record at row {lib().exo_vcf_reader_err_row(r)}")
            out = {"rows": int(b.rows)}
            if 0 in projection:
                out["chrom_offsets"] = np.ctypeslib.as_array(b.chrom_offsets, (b.rows + 1,)).copy()
                out["chrom_values"] = np.ctypeslib.as_array(b.chrom_values, (max(b.chrom_values_len, 1),))[
                    : b.chrom_values_len].copy()
            if 1 in projection:
                out["pos"] = np.ctypeslib.as_array(b.pos, (b.rows,)).copy()
            yield out
    finally:
        lib().exo_vcf_reader_close(r)


class VcfWide(C.Structure):
    _fields_ = [("rows", C.c_int64), ("id_valid", C.POINTER(C.c_uint8)), ("alt_valid", C.POINTER(C.c_uint8)),
                ("qual_valid", C.POINTER(C.c_uint8)), ("id_count", C.POINTER(C.c_int32)), ("filter_count", C.POINTER(C.c_int32)),
                ("ref_len", C.POINTER(C.c_int32)), ("qual", C.POINTER(C.c_float)), ("id_items", C.c_int64),
                ("filter_items", C.c_int64), ("id_item_len", C.POINTER(C.c_int32)), ("filter_item_len", C.POINTER(C.c_int32)),
                ("id_bytes", C.POINTER(C.c_uint8)), ("filter_bytes", C.POINTER(C.c_uint8)), ("ref_bytes", C.POINTER(C.c_uint8)),
                ("id_bytes_len", C.c_int64), ("filter_bytes_len", C.c_int64), ("ref_bytes_len", C.c_int64), ("err_row", C.c_int64)]


def vcf_wide_rows(data):
    """Columns 2..6 of every record of one VCF text as the reference's lazy builder materialises them:
    dict of per-row Python lists  id (None | [bytes]), ref (bytes), alt (None | []), qual (None | float32 bits), filter ([bytes])."""
    a = _buf(data)
    L = lib()
    L.exo_vcf_wide_scan.restype = C.POINTER(VcfWide)
    L.exo_vcf_wide_scan.argtypes = [C.c_void_p, C.c_int64]
    L.exo_vcf_wide_free.argtypes = [C.POINTER(VcfWide)]
    wp = L.exo_vcf_wide_scan(a.ctypes.data, a.size)
    try:
        w = wp.contents
        if w.err_row >= 0:
            raise ValueError(f"malformed VCF record at row {w.err_row}")
        n = int(w.rows)

        def arr(p, m, dt):
            return np.ctypeslib.as_array(p, (max(int(m), 1),))[: int(m)].astype(dt) if m else np.zeros(0, dt)

        def items(lens_p, bytes_p, n_items, n_bytes):
            lens = arr(lens_p, n_items, np.int64)
            b = arr(bytes_p, n_bytes, np.uint8).tobytes()
            offs = np.concatenate([[0], np.cumsum(lens)])
            return [b[offs[i]:offs[i + 1]] for i in range(len(lens))]

        id_items = items(w.id_item_len, w.id_bytes, w.id_items, w.id_bytes_len)
        fi_items = items(w.filter_item_len, w.filter_bytes, w.filter_items, w.filter_bytes_len)
        refs = items(w.ref_len, w.ref_bytes, n, w.ref_bytes_len)
        id_valid, alt_valid, qual_valid = arr(w.id_valid, n, bool), arr(w.alt_valid, n, bool), arr(w.qual_valid, n, bool)
        id_count, fi_count = arr(w.id_count, n, np.int64), arr(w.filter_count, n, np.int64)
        qual = arr(w.qual, n, np.float32).view(np.uint32)
        out = {"id": [], "ref": refs, "alt": [[] if v else None for v in alt_valid],
               "qual": [int(q) if v else None for q, v in zip(qual, qual_valid)], "filter": []}
        i0 = f0 = 0
        for r in range(n):
            out["id"].append(id_items[i0:i0 + id_count[r]] if id_valid[r] else None)
            i0 += int(id_count[r])
            out["filter"].append(fi_items[f0:f0 + fi_count[r]])
            f0 += int(fi_count[r])
        return out
    finally:
        L.exo_vcf_wide_free(wp)


# ---- INFO / FORMAT re-serialisation (string mode), pure Python ---------------------------------------------------------------
# Reserved keys noodles falls back to when the header does not define a key (VCF 4.3 / 4.4 tables 1 and 2; the subset whose
# definition does not differ between the versions): key -> (type, Number == 1)
_RESERVED_INFO = {b"AA": (b"String", True), b"AC": (b"Integer", False), b"AD": (b"Integer", False), b"ADF": (b"Integer", False),
                  b"ADR": (b"Integer", False), b"AF": (b"Float", False), b"AN": (b"Integer", True), b"BQ": (b"Float", True),
                  b"CIGAR": (b"String", False), b"DB": (b"Flag", False), b"DP": (b"Integer", True), b"END": (b"Integer", True),
                  b"H2": (b"Flag", False), b"H3": (b"Flag", False), b"MQ": (b"Float", True), b"MQ0": (b"Integer", True),
                  b"NS": (b"Integer", True), b"SB": (b"Integer", False), b"SOMATIC": (b"Flag", False), b"VALIDATED": (b"Flag", False),
                  b"1000G": (b"Flag", False), b"IMPRECISE": (b"Flag", False), b"NOVEL": (b"Flag", False), b"SVTYPE": (b"String", True)}
_RESERVED_FORMAT = {b"AD": (b"Integer", False), b"ADF": (b"Integer", False), b"ADR": (b"Integer", False), b"DP": (b"Integer", True),
                    b"EC": (b"Integer", False), b"FT": (b"String", True), b"GL": (b"Float", False), b"GP": (b"Float", False),
                    b"GQ": (b"Integer", True), b"HQ": (b"Integer", False), b"MQ": (b"Integer", True), b"PL": (b"Integer", False),
                    b"PP": (b"Integer", False), b"PQ": (b"Integer", True), b"PS": (b"Integer", True)}


def _header_defs(text: bytes, kind: bytes):
    import re

    out = {}
    for line in text.split(b"\n"):
        if not line.startswith(b"##" + kind + b"=<"):
            continue
        ident = re.search(rb"[<,]ID=([^,>]+)", line).group(1)
        num = re.search(rb",Number=([^,>]+)", line)
        out[ident] = (re.search(rb",Type=([^,>]+)", line).group(1), num is not None and num.group(1) == b"1")
    return out


def rust_f32_display(x) -> bytes:
    """`f32::to_string()`: shortest digits that round-trip, positional, "NaN" / "inf" / "-inf", "-0"."""
    v = np.float32(x)
    if np.isnan(v):
        return b"NaN"
    if np.isinf(v):
        return b"inf" if v > 0 else b"-inf"
    if v == 0:
        return b"-0" if np.signbit(v) else b"0"
    return np.format_float_positional(v, unique=True, trim="-").encode()


def _rust_f32_parse(e: bytes) -> float:
    import re

    t = e.decode("ascii")
    low = t.lower().lstrip("+-")
    if low in ("inf", "infinity", "nan"):
        return float(t)
    if not re.fullmatch(r"[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?", t):
        raise ValueError(f"not an f32 literal: {t!r}")
    return float(t)


def _percent_decode(e: bytes) -> bytes:
    out, i = bytearray(), 0
    while i < len(e):
        if e[i] == 0x25 and i + 2 < len(e) and all(c in b"0123456789abcdefABCDEF" for c in e[i + 1:i + 3]):
            out.append(int(e[i + 1:i + 3], 16))
            i += 3
        else:
            out.append(e[i])
            i += 1
    return bytes(out)


def _elem(ty: bytes, e: bytes) -> bytes:
    import re

    if not e:
        raise ValueError("empty element")
    if ty == b"Integer":
        if not re.fullmatch(rb"[+-]?\d+", e) or not -2**31 <= int(e) < 2**31:
            raise ValueError(f"not an i32: {e!r}")
        return str(int(e)).encode()
    if ty == b"Float":
        return rust_f32_display(_rust_f32_parse(e))
    d = _percent_decode(e)
    if ty == b"Character" and len(d) != 1:
        raise ValueError("not a character")
    return d


def _value(ty: bytes, single: bool, val: bytes, skip_missing_chars: bool = False) -> bytes:
    if val in (b"", b"."):
        raise ValueError("missing value: the reference's builder unwraps a None")
    if single or ty == b"String":
        return _elem(ty, val)
    elems = []
    for e in val.split(b","):
        if e == b".":
            if not (skip_missing_chars and ty == b"Character"):
                elems.append(b".")
        else:
            elems.append(_elem(ty, e))
    return b",".join(elems)


def vcf_info_strings(data):
    """Column 7 (info, string mode) of every record as LazyVCFArrayBuilder::append prints it
    (/root/reference/exon/exon-vcf/src/array_builder/lazy_array_builder.rs:217-298): noodles' typed view of the field -- types from
    the header's ##INFO lines, else the specification's reserved keys, else one String -- re-serialised as `key=value` joined by
    ';', a flag as `key=true`, integers through i32 Display, floats through f32 Display (shortest digits that round-trip, never
    an exponent), strings percent-decoded, array elements joined by ',' with '.' for a missing one.  Pure Python (small inputs
    only); pinned by slt/vcf-select-tests.slt:6-10.  A missing value (`key=.`, a non-flag key without `=`) raises: the
    reference's builder unwraps a None there."""
    text = bytes(_buf(data))
    types = _header_defs(text, b"INFO")
    out = []
    for line in text[header_len(text):].split(b"\n"):
        if not line:
            continue
        field = line.split(b"\t")[7]
        if field in (b".", b""):
            out.append(b"")
            continue
        parts = []
        for entry in field.split(b";"):
            key, eq, val = entry.partition(b"=")
            ty, single = types.get(key) or _RESERVED_INFO.get(key) or (b"String", True)
            if ty == b"Flag":
                if val:
                    raise ValueError("flag with a value")
                parts.append(key + b"=true")
                continue
            if not eq:
                raise ValueError("missing value")
            parts.append(key + b"=" + _value(ty, single, val))
        out.append(b";".join(parts))
    return out


def _genotype(val: bytes) -> bytes:
    import re

    if val in (b"", b"."):
        raise ValueError("missing genotype")
    alleles = re.split(rb"[/|]", val)
    seps = [bytes([c]) for c in val if c in b"/|"]
    out = []
    prev = b"|" if all(s == b"|" for s in seps) else b"/"   # noodles' inferred phasing of the first allele
    for k, a in enumerate(alleles):
        if a != b"." and not a.isdigit():
            raise ValueError(f"bad allele {a!r}")
        txt = b"." if a == b"." else str(int(a)).encode()
        if k:
            out.append(prev + txt)
            prev = seps[k - 1]
        else:
            out.append(txt)
    return b"".join(out)


def vcf_formats_strings(data):
    """Column 8 (formats, string mode) as LazyVCFArrayBuilder::append prints it (lazy_array_builder.rs:310-432): the FORMAT keys
    joined by ':', a tab, and every sample with its values re-serialised (genotype allele by allele, numbers through Rust's
    Display) and joined by ':', samples joined by tabs.  A sites-only record gives "\t".  Pinned by the first row of
    slt/vcf-select-tests.slt:12-15 (`GT:PL:PG\t0/0:0,3,26:0`).  A missing sample value ('.') raises: the builder unwraps a None."""
    text = bytes(_buf(data))
    types = _header_defs(text, b"FORMAT")
    out = []
    for line in text[header_len(text):].split(b"\n"):
        if not line:
            continue
        f = line.split(b"\t")
        if len(f) < 9 or not f[8]:
            out.append(b"\t")
            continue
        keys = f[8].split(b":")
        samples = []
        for smp in f[9:]:
            vals = []
            for key, val in zip(keys, smp.split(b":")):
                if key == b"GT":
                    vals.append(_genotype(val))
                    continue
                ty, single = types.get(key) or _RESERVED_FORMAT.get(key) or (b"String", True)
                vals.append(_value(ty, single, val, skip_missing_chars=True))
            samples.append(b":".join(vals))
        out.append(f[8] + b"\t" + b"\t".join(samples))
    return out


def filter_count(data, chrom=None, lo=None, hi=None, batch_size: int = 8192):
    """(count, rows) for `chrom = <chrom> AND pos BETWEEN lo AND hi` over one VCF text."""
    a = _buf(data)
    has_chrom = chrom is not None
    has_iv = lo is not None or hi is not None
    cb = chrom.encode() if isinstance(chrom, str) else (chrom or b"")
    rows = C.c_int64()
    c = lib().exo_vcf_filter_count(a.ctypes.data, a.size, batch_size, cb, len(cb), int(has_chrom), int(has_iv),
                                   1 if lo is None else lo, INT64_MAX if hi is None else hi, C.byref(rows))
    if c < 0:
        raise ValueError(f"oracle error {c}")
    return int(c), int(rows.value)


def filter_count_files(datas, chrom=None, lo=None, hi=None, target_partitions: int = 1, batch_size: int = 8192):
    """(count, rows, partitions) over several in-memory VCF files, one worker thread per partition."""
    arrs = [_buf(d) for d in datas]
    n = len(arrs)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * n)(*[a.size for a in arrs])
    has_chrom = chrom is not None
    has_iv = lo is not None or hi is not None
    cb = chrom.encode() if isinstance(chrom, str) else (chrom or b"")
    rows, parts = C.c_int64(), C.c_int32()
    c = lib().exo_vcf_filter_count_files(ptrs, lens, n, target_partitions, batch_size, cb, len(cb), int(has_chrom),
                                         int(has_iv), 1 if lo is None else lo, INT64_MAX if hi is None else hi,
                                         C.byref(rows), C.byref(parts))
    if c < 0:
        raise ValueError(f"oracle error {c}")
    return int(c), int(rows.value), int(parts.value)


def regroup_files_by_size(sizes, target_partitions: int):
    n = len(sizes)
    s = (C.c_int64 * n)(*sizes)
    g = (C.c_int32 * n)()
    parts = lib().exo_regroup_files_by_size(s, n, target_partitions, g)
    return parts, list(g)


def region_match(offsets, values, pos, region: str):
    rg = parse_region(region)
    n = len(pos)
    out = np.zeros(n, dtype=np.uint8)
    lib().exo_region_match(offsets.ctypes.data, values.ctypes.data, pos.ctypes.data, n, C.byref(rg), out.ctypes.data)
    return out.astype(bool)


def chrom_match(offsets, values, lit: str):
    n = len(offsets) - 1
    out = np.zeros(n, dtype=np.uint8)
    b = lit.encode()
    lib().exo_chrom_match(offsets.ctypes.data, values.ctypes.data, n, b, len(b), out.ctypes.data)
    return out.astype(bool)


def interval_match(pos, interval: str):
    lo, hi = parse_interval(interval)
    out = np.zeros(len(pos), dtype=np.uint8)
    lib().exo_interval_match(pos.ctypes.data, len(pos), lo, hi, out.ctypes.data)
    return out.astype(bool)


# ---- FASTQ (oracle/fastq_oracle.c) ---------------------------------------------------------------------------

class _Utf8Col(C.Structure):
    _fields_ = [("offsets", C.POINTER(C.c_int32)), ("values", C.POINTER(C.c_uint8)), ("valid", C.POINTER(C.c_uint8)),
                ("values_len", C.c_int64), ("values_cap", C.c_int64)]


class FastqBatch(C.Structure):
    _fields_ = [("rows", C.c_int64), ("name", _Utf8Col), ("description", _Utf8Col), ("sequence", _Utf8Col),
                ("quality", _Utf8Col)]


_fq_ready = False


def _fq():
    global _fq_ready
    L = lib()
    if not _fq_ready:
        L.exo_fastq_reader_open.restype = C.c_void_p
        L.exo_fastq_reader_open.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.exo_fastq_reader_next.restype = C.c_int
        L.exo_fastq_reader_next.argtypes = [C.c_void_p, C.POINTER(FastqBatch)]
        L.exo_fastq_reader_close.argtypes = [C.c_void_p]
        L.exo_fastq_reader_err_record.restype = C.c_int64
        L.exo_fastq_reader_err_record.argtypes = [C.c_void_p]
        L.exo_fastq_filter_count.restype = C.c_int64
        L.exo_fastq_filter_count.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                             C.POINTER(C.c_int64)]
        L.exo_fastq_filter_count_files.restype = C.c_int64
        L.exo_fastq_filter_count_files.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int64,
                                                   C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        L.exo_quality_scores_to_list.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        _fq_ready = True
    return L


def _utf8_rows(col: _Utf8Col, rows: int, nullable: bool = False):
    off = np.ctypeslib.as_array(col.offsets, (rows + 1,))
    val = bytes(np.ctypeslib.as_array(col.values, (max(int(col.values_len), 1),))[: int(col.values_len)])
    valid = np.ctypeslib.as_array(col.valid, (rows,))
    return [None if (nullable and not valid[i]) else val[off[i]:off[i + 1]] for i in range(rows)]


def fastq_read_batches(data, batch_size: int = 8192):
    """Yield one dict per reference batch: rows, name / description (None = NULL) / sequence / quality as bytes lists."""
    a = _buf(data)
    L = _fq()
    r = L.exo_fastq_reader_open(a.ctypes.data, a.size, batch_size)
    try:
        b = FastqBatch()
        while True:
            rc = L.exo_fastq_reader_next(r, C.byref(b))
            if rc == 0:
                return
            if rc < 0:
                raise ValueError(f"malformed FASTQ record {L.exo_fastq_reader_err_record(r)}")
            n = int(b.rows)
            yield {"rows": n, "name": _utf8_rows(b.name, n), "description": _utf8_rows(b.description, n, True),
                   "sequence": _utf8_rows(b.sequence, n), "quality": _utf8_rows(b.quality, n)}
    finally:
        L.exo_fastq_reader_close(r)


def fastq_filter_count(data, min_mean=None, phred_offset: int = 33, batch_size: int = 8192):
    """(count, rows): records with mean(quality) > min_mean (int or (num, den)); None -> COUNT(*)."""
    a = _buf(data)
    num, den = (0, 1) if min_mean is None else (min_mean if isinstance(min_mean, tuple) else (int(min_mean), 1))
    rows = C.c_int64()
    c = _fq().exo_fastq_filter_count(a.ctypes.data, a.size, batch_size, int(min_mean is not None), phred_offset, num, den,
                                     C.byref(rows))
    if c < 0:
        raise ValueError("malformed FASTQ record")
    return int(c), int(rows.value)


def fastq_filter_count_files(files, min_mean=None, phred_offset: int = 33, target_partitions: int = 8, batch_size: int = 8192):
    bufs = [_buf(f) for f in files]
    n = len(bufs)
    ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
    lens = (C.c_int64 * max(n, 1))(*[b.size for b in bufs])
    num, den = (0, 1) if min_mean is None else (min_mean if isinstance(min_mean, tuple) else (int(min_mean), 1))
    rows = C.c_int64()
    c = _fq().exo_fastq_filter_count_files(ptrs, lens, n, target_partitions, batch_size, int(min_mean is not None),
                                           phred_offset, num, den, C.byref(rows))
    if c < 0:
        raise ValueError("malformed FASTQ record")
    return int(c), int(rows.value)


def quality_scores_to_list(s: bytes):
    a = _buf(s)
    out = np.empty(a.size, dtype=np.int32)
    _fq().exo_quality_scores_to_list(a.ctypes.data, a.size, out.ctypes.data)
    return out.tolist()


def filter_count_gz_files(datas, chrom=None, lo=None, hi=None, target_partitions: int = 1, batch_size: int = 8192):
    """(count, rows) over several in-memory .vcf.gz (BGZF / gzip) files: zlib inflate + the same record path,
    one worker thread per file partition."""
    arrs = [_buf(d) for d in datas]
    n = len(arrs)
    L = lib()
    L.exo_vcf_gz_filter_count_files.restype = C.c_int64
    L.exo_vcf_gz_filter_count_files.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int64,
                                                C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                                C.POINTER(C.c_int64)]
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * max(n, 1))(*[a.size for a in arrs])
    cb = chrom.encode() if isinstance(chrom, str) else (chrom or b"")
    rows = C.c_int64()
    c = L.exo_vcf_gz_filter_count_files(ptrs, lens, n, target_partitions, batch_size, cb, len(cb), int(chrom is not None),
                                        int(lo is not None or hi is not None), 1 if lo is None else lo,
                                        INT64_MAX if hi is None else hi, C.byref(rows))
    if c < 0:
        raise ValueError(f"oracle error {c}")
    return int(c), int(rows.value)


def gunzip_all(data) -> np.ndarray:
    """Every member of a gzip / BGZF file inflated with zlib in C (exo_gunzip_all): the CPU arm's decompression step."""
    a = _buf(data)
    L = lib()
    L.exo_gunzip_all.restype = C.c_void_p
    L.exo_gunzip_all.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    n = C.c_int64()
    p = L.exo_gunzip_all(a.ctypes.data, a.size, C.byref(n))
    if not p:
        raise ValueError("corrupt gzip stream")
    try:
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (max(n.value, 1),))[: n.value].copy()
    finally:
        C.CDLL(None).free(C.c_void_p(p))


# ---- BAM (oracle/bam_oracle.c) -------------------------------------------------------------------------------

class BamRow(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("flag", C.c_int32), ("ref_id", C.c_int32), ("mate_ref_id", C.c_int32), ("mapq", C.c_int32),
                ("start", C.c_int64), ("end", C.c_int64), ("cigar", C.c_char * 1024), ("l_seq", C.c_int32), ("first_quals", C.c_int32 * 8)]


class Bam:
    """One .bam file opened by the oracle (zlib inflate + header)."""

    def __init__(self, data):
        L = lib()
        L.exo_bam_open.restype = C.c_void_p
        L.exo_bam_open.argtypes = [C.c_void_p, C.c_int64]
        L.exo_bam_close.argtypes = [C.c_void_p]
        L.exo_bam_n_ref.restype = C.c_int32
        L.exo_bam_n_ref.argtypes = [C.c_void_p]
        L.exo_bam_ref_name.restype = C.c_char_p
        L.exo_bam_ref_name.argtypes = [C.c_void_p, C.c_int32]
        L.exo_bam_scan.restype = C.c_int64
        L.exo_bam_scan.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_int64), C.c_int64,
                                   C.POINTER(BamRow)]
        self._L = L
        a = _buf(data)
        self._h = L.exo_bam_open(a.ctypes.data, a.size)
        if not self._h:
            raise ValueError("not a BAM file")
        self.refs = [L.exo_bam_ref_name(self._h, i).decode() for i in range(L.exo_bam_n_ref(self._h))]

    def close(self):
        if self._h:
            self._L.exo_bam_close(self._h)
            self._h = None

    def count_by_reference(self, flag_exclude=0, flag_require=0, min_mapq=-1, all_rows=False, region=None):
        self._L.exo_bam_set_region.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64]
        if region is None:
            self._L.exo_bam_set_region(self._h, None, 1, INT64_MAX)
        else:
            self._L.exo_bam_set_region(self._h, region[0].encode(), 1 if region[1] is None else region[1],
                                       INT64_MAX if region[2] is None else region[2])
        counts = (C.c_int64 * (len(self.refs) + 1))()
        n = self._L.exo_bam_scan(self._h, int(not all_rows), flag_exclude, flag_require, min_mapq, counts, -1, None)
        if n < 0:
            raise ValueError("malformed BAM record chain")
        out = {nm: int(counts[i]) for i, nm in enumerate(self.refs)}
        out[None] = int(counts[len(self.refs)])
        return out, int(n)

    def seq_qual(self, i: int):
        """(sequence str, quality_score list[int]) of record i: columns 8 and 9 of the reference's BAM batches."""
        L = self._L
        L.exo_bam_seq_qual.restype = C.c_int32
        L.exo_bam_seq_qual.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int32, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int32)]
        cap = 1 << 16
        seq = C.create_string_buffer(cap + 1)
        qual = (C.c_int64 * cap)()
        nq = C.c_int32()
        n = L.exo_bam_seq_qual(self._h, i, seq, cap + 1, qual, cap, C.byref(nq))
        if n < 0:
            raise IndexError(i)
        return seq.value.decode(), list(qual[: nq.value])

    def row(self, i: int):
        r = BamRow()
        n = self._L.exo_bam_scan(self._h, 0, 0, 0, -1, None, i, C.byref(r))
        if n < 0 or i >= n:
            raise IndexError(i)
        return {"name": r.name.decode() or None, "flag": r.flag, "reference": self.refs[r.ref_id] if r.ref_id >= 0 else None,
                "start": r.start or None, "end": r.end or None, "mapping_quality": None if r.mapq == 255 else str(r.mapq),
                "cigar": r.cigar.decode(), "mate_reference": self.refs[r.mate_ref_id] if r.mate_ref_id >= 0 else None,
                "l_seq": r.l_seq, "first_quals": list(r.first_quals)}


def _rust_f32_fixed2(x) -> str:
    """Rust `format!("{:.2}", x)` for an f32: the exact value, round-half-even at two decimals (core::fmt float_to_decimal_exact)."""
    x = float(np.float32(x))
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    return format(x, ".2f")  # CPython formats the exact binary value, ties to even, and keeps the sign of -0.0


def bam_tags(data):
    """Per record, the `tags` column (column 10) of the reference's BAM / SAM batches in its default List<Struct{tag, value}> form:
    [(tag, value text), ...] in record order.  Follows TagsMapBuilder::append (/root/reference/exon/exon-sam/src/tag_builder.rs:497-741):
    integers through i64 Display, A as its character, Z / H as text, f through f32 Display, B integer arrays joined by ",",
    B:f arrays as "{:.2}" joined by ", ".  Pinned by sam-select-tests.slt:47-53 (same builder, same data model)."""
    import struct

    raw = bytes(gunzip_all(data))
    if raw[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    p = 8 + struct.unpack_from("<i", raw, 4)[0]
    n_ref = struct.unpack_from("<i", raw, p)[0]
    p += 4
    for _ in range(n_ref):
        p += 4 + struct.unpack_from("<i", raw, p)[0] + 4
    ints = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}
    out = []
    while p + 4 <= len(raw):
        bs = struct.unpack_from("<i", raw, p)[0]
        rec = raw[p + 4: p + 4 + bs]
        p += 4 + bs
        l_name, n_cig, l_seq = rec[8], struct.unpack_from("<H", rec, 12)[0], struct.unpack_from("<i", rec, 16)[0]
        q = 32 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq
        row = []
        while q < len(rec):
            tag, ty = rec[q: q + 2].decode("latin-1"), chr(rec[q + 2])
            q += 3
            if ty == "A":
                val, q = chr(rec[q]), q + 1
            elif ty in ints:
                val = str(struct.unpack_from(ints[ty], rec, q)[0])
                q += struct.calcsize(ints[ty])
            elif ty == "f":
                val = rust_f32_display(struct.unpack_from("<f", rec, q)[0]).decode()
                q += 4
            elif ty in "ZH":
                e = rec.index(b"\x00", q)
                val, q = rec[q:e].decode("latin-1"), e + 1
            elif ty == "B":
                st, cnt = chr(rec[q]), struct.unpack_from("<I", rec, q + 1)[0]
                q += 5
                if st == "f":
                    val = ", ".join(_rust_f32_fixed2(v) for v in struct.unpack_from("<%df" % cnt, rec, q))
                    q += 4 * cnt
                else:
                    val = ",".join(str(v) for v in struct.unpack_from("<%d%s" % (cnt, ints[st][1]), rec, q))
                    q += cnt * struct.calcsize(ints[st])
            else:
                raise ValueError("unknown auxiliary field type %r" % ty)
            row.append((tag, val))
        out.append(row)
    return out


def bam_count_by_reference_files(files, **kw):
    """Sum of per-file group counts keyed by reference NAME (the GROUP BY merges equal names across files)."""
    total, rows = {}, 0
    for f in files:
        b = Bam(f)
        try:
            c, n = b.count_by_reference(**kw)
        finally:
            b.close()
        rows += n
        for k, v in c.items():
            total[k] = total.get(k, 0) + v
    return total, rows


# ---- mzML (oracle/mzml_oracle.c) -----------------------------------------------------------------------------

class MzmlResult(C.Structure):
    _fields_ = [("n_spectra", C.c_int64), ("n_selected", C.c_int64), ("sum", C.c_double), ("kind_sum", C.c_double * 3),
                ("kind_count", C.c_int64 * 3)]


def mzml_scan(data, lo=None, hi=None, spectrum: int = -1) -> MzmlResult:
    """SUM(intensity) over peaks with lo <= mz <= hi (None: all zipped peaks), spectra count, per-kind totals."""
    a = _buf(data)
    L = lib()
    L.exo_mzml_scan.restype = C.c_int
    L.exo_mzml_scan.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_int64, C.POINTER(MzmlResult)]
    r = MzmlResult()
    rc = L.exo_mzml_scan(a.ctypes.data, a.size, int(lo is not None), float(lo or 0.0), float(hi or 0.0), spectrum, C.byref(r))
    if rc != 0:
        raise ValueError("malformed mzML")
    return r


def mzml_decode_binary(b64: bytes, zlib_compressed: bool, f32: bool):
    L = lib()
    L.exo_mzml_decode_binary.restype = C.c_int64
    L.exo_mzml_decode_binary.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.POINTER(C.c_double))]
    out = C.POINTER(C.c_double)()
    n = L.exo_mzml_decode_binary(b64, len(b64), int(zlib_compressed), int(f32), C.byref(out))
    if n < 0:
        raise ValueError("bad binary array")
    vals = [out[i] for i in range(n)]
    C.CDLL(None).free(out)
    return vals


def gff_attributes(data, reference_quirk: bool = True):
    """Column 8 (attributes, Map<Utf8, List<Utf8>>) of every record of ONE GFF file (= one batch) as GFFArrayBuilder::append
    builds it (/root/reference/exon/exon-gff/src/array_builder.rs:142-160) from noodles-gff 0.41 lazy attributes: fields split
    at ';', `key=value`, a value with ',' is an array, keys and values percent-decoded.  Returns, per record, [(key, [values])].
    reference_quirk: for a plain string value the builder calls `values().append(true)` BEFORE it appends the string, so the
    string lands in the NEXT entry's list (simulated here with the builder's own pending-values state); False gives the
    lists one would expect, for comparison."""
    text = bytes(_buf(data))
    rows, pending = [], []
    for line in text.split(b"\n"):
        if not line or line.startswith(b"#"):
            continue
        f = line.split(b"\t")
        if len(f) < 9:
            raise ValueError("fewer than 9 fields")
        row = []
        field = f[8]
        if field != b".":
            parts = field.split(b";")
            if parts and parts[-1] == b"":
                parts.pop()
            for part in parts:
                if b"=" not in part:
                    raise ValueError("attribute without '='")
                k, v = part.split(b"=", 1)
                key = _percent_decode(k).decode()
                vals = [_percent_decode(x).decode() for x in v.split(b",")]
                if not reference_quirk:
                    row.append((key, vals))
                elif b"," in v:           # Array: values appended, then the list is closed
                    pending += vals
                    row.append((key, pending))
                    pending = []
                else:                      # String: the list is closed first, then the value is appended
                    row.append((key, pending))
                    pending = list(vals)
        rows.append(row)
    return rows


def mzml_rows(data):
    """Every <spectrum> as MzMLArrayBuilder::append materialises it (/root/reference/exon/exon-mzml/src/array_builder.rs:330-416):
    {id, mz, intensity, wavelength (list of f64 | None), cv_params [(accession, name, value | None)], precursor_mz, precursor_charge}.
    Pure Python on xml.etree (small inputs only): the spectrum's own cvParam children; arrays decoded as
    binary_conversion.rs:26-95 does (base64 -> optional zlib -> LE f32 / f64 -> f64), an empty <binary> gives [];
    MS:1000744 / MS:1000041 of the first selected ion of the first precursor."""
    import base64
    import struct
    import xml.etree.ElementTree as ET
    import zlib

    ns = "{http://psi.hupo.org/ms/mzml}"
    text = bytes(_buf(data))
    root = ET.fromstring(text)
    if not root.tag.startswith(ns):
        ns = ""
    out = []
    for sp in root.iter(ns + "spectrum"):
        row = {"id": sp.get("id"), "mz": None, "intensity": None, "wavelength": None, "precursor_mz": None, "precursor_charge": None}
        row["cv_params"] = [(c.get("accession"), c.get("name"), c.get("value") or None) for c in sp.findall(ns + "cvParam")]
        for bda in sp.iter(ns + "binaryDataArray"):
            acc = [c.get("accession") for c in bda.findall(ns + "cvParam")]
            kind = next((k for a in acc for k, code in (("mz", "MS:1000514"), ("intensity", "MS:1000515"), ("wavelength", "MS:1000617")) if a == code), None)
            b = bda.find(ns + "binary")
            if kind is None or b is None:
                continue
            txt = (b.text or "").strip()
            if not txt:
                row[kind] = []
                continue
            raw = base64.b64decode(txt)
            if "MS:1000574" in acc:
                raw = zlib.decompress(raw)
            w, f = (4, "f") if ("MS:1000521" in acc and "MS:1000523" not in acc) else (8, "d")
            row[kind] = [float(x) for x in struct.unpack("<%d%s" % (len(raw) // w, f), raw[: len(raw) // w * w])]
        pl = sp.find(ns + "precursorList")
        if pl is not None:
            ion = pl.findall(ns + "precursor")[0].find(ns + "selectedIonList").findall(ns + "selectedIon")[0]
            for c in ion.findall(ns + "cvParam"):
                if c.get("accession") == "MS:1000744" and c.get("value") is not None and row["precursor_mz"] is None:
                    row["precursor_mz"] = float(c.get("value"))
                if c.get("accession") == "MS:1000041" and c.get("value") is not None and row["precursor_charge"] is None:
                    row["precursor_charge"] = int(c.get("value"))
        out.append(row)
    return out


def fasta_count(data) -> int:
    """COUNT(*) of a FASTA file: definition lines (oracle/fastq_oracle.c: exo_fasta_count)."""
    a = _buf(data)
    L = lib()
    L.exo_fasta_count.restype = C.c_int64
    L.exo_fasta_count.argtypes = [C.c_void_p, C.c_int64]
    n = L.exo_fasta_count(a.ctypes.data, a.size)
    if n < 0:
        raise ValueError("malformed FASTA")
    return int(n)


class FastaRecords(C.Structure):
    _fields_ = [("rows", C.c_int64), ("name_len", C.POINTER(C.c_int32)), ("desc_len", C.POINTER(C.c_int32)), ("seq_len", C.POINTER(C.c_int64)),
                ("names", C.POINTER(C.c_uint8)), ("descs", C.POINTER(C.c_uint8)), ("seqs", C.POINTER(C.c_uint8)),
                ("names_len", C.c_int64), ("descs_len", C.c_int64), ("seqs_len", C.c_int64), ("err", C.c_int32)]


def fasta_records(data):
    """[(id bytes, description bytes | None, sequence bytes)] of one FASTA text, as the reference's batches hold them."""
    a = _buf(data)
    L = _fq()
    L.exo_fasta_read.restype = C.POINTER(FastaRecords)
    L.exo_fasta_read.argtypes = [C.c_void_p, C.c_int64]
    L.exo_fasta_records_free.argtypes = [C.POINTER(FastaRecords)]
    rp = L.exo_fasta_read(a.ctypes.data, a.size)
    try:
        r = rp.contents
        if r.err:
            raise ValueError({1: "invalid definition", 2: "missing name", 3: "invalid sequence"}[r.err])
        names = bytes(np.ctypeslib.as_array(r.names, (max(r.names_len, 1),))[: r.names_len]) if r.names_len else b""
        descs = bytes(np.ctypeslib.as_array(r.descs, (max(r.descs_len, 1),))[: r.descs_len]) if r.descs_len else b""
        seqs = bytes(np.ctypeslib.as_array(r.seqs, (max(r.seqs_len, 1),))[: r.seqs_len]) if r.seqs_len else b""
        out, n0, d0, s0 = [], 0, 0, 0
        for i in range(r.rows):
            nl, dl, sl = r.name_len[i], r.desc_len[i], r.seq_len[i]
            out.append((names[n0:n0 + nl], None if dl < 0 else descs[d0:d0 + dl], seqs[s0:s0 + sl]))
            n0 += nl
            d0 += max(dl, 0)
            s0 += sl
        return out
    finally:
        L.exo_fasta_records_free(rp)


def gff_filter_count(data, name=None, lo=None, hi=None):
    """(count, rows): GFF records with seqname == name and lo <= start <= hi (None drops a term)."""
    a = _buf(data)
    L = lib()
    L.exo_gff_filter_count.restype = C.c_int64
    L.exo_gff_filter_count.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                       C.POINTER(C.c_int64)]
    nb = name.encode() if isinstance(name, str) else (name or b"")
    rows = C.c_int64()
    c = L.exo_gff_filter_count(a.ctypes.data, a.size, nb, len(nb), int(name is not None), int(lo is not None or hi is not None),
                               1 if lo is None else lo, INT64_MAX if hi is None else hi, C.byref(rows))
    if c < 0:
        raise ValueError("malformed GFF record")
    return int(c), int(rows.value)


class GffRow(C.Structure):
    _fields_ = [("seqname", C.c_char * 256), ("source", C.c_char * 256), ("type", C.c_char * 256), ("start", C.c_int64), ("end", C.c_int64),
                ("score", C.c_float), ("score_valid", C.c_int32), ("strand", C.c_char * 4), ("phase", C.c_char * 4)]


def gff_rows(data, limit=None):
    """[(seqname, source, type, start, end, score f32 bits | None, strand, phase | None)] of one GFF text (columns 0..7)."""
    a = _buf(data)
    L = _fq()
    L.exo_gff_row_at.restype = C.c_int32
    L.exo_gff_row_at.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(GffRow)]
    out, r, i = [], GffRow(), 0
    while limit is None or i < limit:
        rc = L.exo_gff_row_at(a.ctypes.data, a.size, i, C.byref(r))
        if rc < 0:
            raise ValueError("malformed GFF")
        if rc == 0:
            break
        bits = int(np.array([r.score], np.float32).view(np.uint32)[0]) if r.score_valid else None
        out.append((r.seqname, r.source, r.type, r.start, r.end, bits, r.strand, r.phase or None))
        i += 1
    return out
