"""Synthetic sites-only VCF 4.2 shards.

N records over the 24 primary GRCh37 contigs (lengths as in the reference fixture
exon/exon-core/test-data/datasources/vcf/index.vcf:6-29), rows per contig proportional to contig length
(largest-remainder rounding), sorted positions within a contig, written as K shard files split at record
boundaries, each with the full header.  rng = numpy.random.default_rng(20241017).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import numpy as np

SEED = 20241017
CONTIGS = [
    ("1", 249250621), ("2", 243199373), ("3", 198022430), ("4", 191154276), ("5", 180915260), ("6", 171115067),
    ("7", 159138663), ("8", 146364022), ("9", 141213431), ("10", 135534747), ("11", 135006516), ("12", 133851895),
    ("13", 115169878), ("14", 107349540), ("15", 102531392), ("16", 90354753), ("17", 81195210), ("18", 78077248),
    ("19", 59128983), ("20", 63025520), ("21", 48129895), ("22", 51304566), ("X", 155270560), ("Y", 59373566),
]
MAX_LINE = 32  # "22\t249250621\t.\tA\tC\t99\tPASS\t.\n" is 30 bytes

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsynth.so")
        src = os.path.join(_HERE, "vcf_format.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", so, src], check=True)
        _LIB = C.CDLL(so)
        _LIB.synth_vcf_format.restype = C.c_int64
        _LIB.synth_vcf_format.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_char_p, C.c_void_p]
    return _LIB


def header_text(contigs=CONTIGS) -> bytes:
    lines = ["##fileformat=VCFv4.2", '##FILTER=<ID=PASS,Description="All filters passed">']
    lines += [f"##contig=<ID={n},length={l}>" for n, l in contigs]
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO")
    return ("\n".join(lines) + "\n").encode()


@dataclass
class VcfColumns:
    contig: np.ndarray  # uint8 index into CONTIGS
    pos: np.ndarray     # int64
    ref: np.ndarray     # uint8 0..3
    alt: np.ndarray     # uint8 0..3, != ref
    qual: np.ndarray    # int8, -1 = '.'
    contigs: list = field(default_factory=lambda: CONTIGS)

    @property
    def n(self) -> int:
        return int(self.pos.size)

    def truth_count(self, chrom: str | None, lo: int | None, hi: int | None) -> int:
        m = np.ones(self.n, dtype=bool)
        if chrom is not None:
            names = [c for c, _ in self.contigs]
            m &= (self.contig == names.index(chrom)) if chrom in names else False
        if lo is not None:
            m &= self.pos >= lo
        if hi is not None:
            m &= self.pos <= hi
        return int(m.sum())


def columns(n: int, seed: int = SEED, contigs=CONTIGS) -> VcfColumns:
    rng = np.random.default_rng(seed)
    lengths = np.array([l for _, l in contigs], dtype=np.float64)
    quota = lengths / lengths.sum() * n
    per = np.floor(quota).astype(np.int64)
    rem = n - int(per.sum())
    order = np.argsort(-(quota - per), kind="stable")
    per[order[:rem]] += 1
    contig = np.repeat(np.arange(len(contigs), dtype=np.uint8), per)
    pos = np.empty(n, dtype=np.int64)
    o = 0
    for (_, length), k in zip(contigs, per):
        pos[o:o + k] = np.sort(rng.integers(1, length + 1, int(k), dtype=np.int64))
        o += int(k)
    ref = rng.integers(0, 4, n, dtype=np.uint8)
    alt = ((ref + rng.integers(1, 4, n, dtype=np.uint8)) & 3).astype(np.uint8)
    qual = rng.integers(0, 100, n, dtype=np.int8)
    qual[rng.integers(0, 16, n, dtype=np.uint8) == 0] = -1
    return VcfColumns(contig, pos, ref, alt, qual, list(contigs))


def _names_blob(contigs) -> bytes:
    return b"".join(n.encode().ljust(8, b"\0") for n, _ in contigs)


def format_rows(cols: VcfColumns, lo: int, hi: int, out: np.ndarray) -> int:
    """Serialise rows [lo, hi) into `out` (uint8); returns the bytes written."""
    assert out.dtype == np.uint8 and out.size >= (hi - lo) * MAX_LINE
    L = _lib()
    return int(L.synth_vcf_format(cols.contig[lo:hi].ctypes.data, cols.pos[lo:hi].ctypes.data,
                                  cols.ref[lo:hi].ctypes.data, cols.alt[lo:hi].ctypes.data,
                                  cols.qual[lo:hi].ctypes.data, hi - lo, _names_blob(cols.contigs), out.ctypes.data))


def shard_bounds(n: int, k: int):
    """Row ranges of the K shards: split at record boundaries in file order, sizes differ by at most one row."""
    edges = [(n * i) // k for i in range(k + 1)]
    return [(edges[i], edges[i + 1]) for i in range(k)]


def shards(cols: VcfColumns, k: int = 64, with_header: bool = True, threads: int | None = None, alloc=None):
    """Return K uint8 arrays, each a complete VCF file (header + its rows).

    `alloc(nbytes) -> np.ndarray[uint8]` lets the caller place the text in pinned memory.
    """
    hdr = header_text(cols.contigs) if with_header else b""
    bounds = shard_bounds(cols.n, k)
    alloc = alloc or (lambda nb: np.empty(nb, dtype=np.uint8))

    def one(b):
        lo, hi = b
        buf = alloc(len(hdr) + (hi - lo) * MAX_LINE)
        buf[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
        w = format_rows(cols, lo, hi, buf[len(hdr):])
        return buf[:len(hdr) + w]

    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        return list(ex.map(one, bounds))
