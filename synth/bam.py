"""Synthetic BAM shards for BASELINE configs[3]: fixed-length alignments over 25 references,
written the way htslib writes BAM (header in its own BGZF block, records never split across blocks)."""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from .vcf import CONTIGS

SEED = 20241019
REFS = CONTIGS + [("MT", 16569)]
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "tests"))


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsynth_bam.so")
        src = os.path.join(_HERE, "bam_format.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", so, src], check=True)
        _LIB = C.CDLL(so)
        _LIB.synth_bam_record_bytes.restype = C.c_int64
        _LIB.synth_bam_record_bytes.argtypes = [C.c_int32]
        _LIB.synth_bam_format.restype = C.c_int64
        _LIB.synth_bam_format.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def header_bytes(refs=REFS) -> bytes:
    text = "@HD\tVN:1.6\tSO:unsorted\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    out = [b"BAM\x01", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(refs))]
    for n, l in refs:
        out += [struct.pack("<i", len(n) + 1), n.encode() + b"\0", struct.pack("<i", l)]
    return b"".join(out)


@dataclass
class BamShards:
    files: list            # bytes objects: complete BGZF-compressed .bam files
    raw_bytes: int         # uncompressed bytes over all files (headers included)
    ref_id: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    refs: list

    @property
    def n(self) -> int:
        return int(self.ref_id.size)

    def truth(self, flag_exclude=0, flag_require=0, min_mapq=-1):
        """{reference name | None: count} of the selected records (NULL MAPQ = 255 fails a MAPQ comparison)."""
        f = self.flag.astype(np.int64)
        sel = ((f & flag_exclude) == 0) & ((f & flag_require) == flag_require)
        if min_mapq >= 0:
            sel &= (self.mapq != 255) & (self.mapq.astype(np.int64) >= min_mapq)
        ids = self.ref_id[sel]
        out = {n: int((ids == i).sum()) for i, (n, _) in enumerate(self.refs)}
        out[None] = int((ids < 0).sum())
        return out


def shards(n: int, k: int = 8, seed: int = SEED, l_seq: int = 100, level: int = 1, refs=REFS, threads: int | None = None) -> BamShards:
    from bgzf_util import EOF_MARKER, bgzf_member

    L = _lib()
    rb = int(L.synth_bam_record_bytes(l_seq))
    lens = np.array([l for _, l in refs], dtype=np.float64)
    cdf = np.minimum(np.cumsum(lens / lens.sum()) * 2.0**32, 2.0**32 - 1).astype(np.uint32)
    ref_len = np.array([l for _, l in refs], dtype=np.int32)
    edges = [(n * i) // k for i in range(k + 1)]
    ref_id = np.empty(n, np.int32)
    flag = np.empty(n, np.uint16)
    mapq = np.empty(n, np.uint8)
    hdr = header_bytes(refs)
    per_block = max(1, 0xFF00 // rb)  # whole records per BGZF block, as htslib's bgzf_flush_try keeps them

    def one(i):
        lo, hi = edges[i], edges[i + 1]
        buf = np.empty((hi - lo) * rb, dtype=np.uint8)
        w = L.synth_bam_format(seed, lo, hi - lo, l_seq, len(refs), cdf.ctypes.data, ref_len.ctypes.data, buf.ctypes.data,
                               ref_id[lo:hi].ctypes.data, flag[lo:hi].ctypes.data, mapq[lo:hi].ctypes.data)
        assert w == buf.size
        raw = buf.tobytes()
        parts = [bgzf_member(hdr[o:o + 0xFF00], level) for o in range(0, len(hdr), 0xFF00)]
        step = per_block * rb
        parts += [bgzf_member(raw[o:o + step], level) for o in range(0, len(raw), step)]
        parts.append(EOF_MARKER)
        return b"".join(parts), len(hdr) + len(raw)

    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        res = list(ex.map(one, range(k)))
    return BamShards([r[0] for r in res], sum(r[1] for r in res), ref_id, flag, mapq, list(refs))
