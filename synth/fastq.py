"""Synthetic FASTQ shards for BASELINE configs[1]: N reads of fixed length, name r%09d,
Phred+33 qualities around a per-read mean.  `truth_count` is computed from the generated integer quality sums."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

SEED = 20241018
READ_LEN = 150
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsynth_fastq.so")
        src = os.path.join(_HERE, "fastq_format.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", so, src, "-lm"], check=True)
        _LIB = C.CDLL(so)
        _LIB.synth_fastq_format.restype = C.c_int64
        _LIB.synth_fastq_format.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    return _LIB


@dataclass
class FastqShards:
    files: list          # uint8 arrays, one complete FASTQ file each
    sum_q: np.ndarray    # int32 per read: sum of Phred scores
    read_len: int

    @property
    def n(self) -> int:
        return int(self.sum_q.size)

    def truth_count(self, num: int = 30, den: int = 1) -> int:
        """reads with mean(quality) > num / den, over integers"""
        return int((self.sum_q.astype(np.int64) * den > num * self.read_len).sum())


def record_bytes(read_len: int = READ_LEN) -> int:
    return 1 + 10 + 1 + read_len + 1 + 2 + read_len + 1


def shards(n: int, k: int = 16, seed: int = SEED, read_len: int = READ_LEN, alloc=None, threads: int | None = None) -> FastqShards:
    edges = [(n * i) // k for i in range(k + 1)]
    alloc = alloc or (lambda nb: np.empty(nb, dtype=np.uint8))
    sum_q = np.empty(n, dtype=np.int32)
    rb = record_bytes(read_len)
    L = _lib()

    def one(i):
        lo, hi = edges[i], edges[i + 1]
        buf = alloc((hi - lo) * rb)
        w = L.synth_fastq_format(seed, lo, hi - lo, read_len, buf.ctypes.data, sum_q[lo:hi].ctypes.data)
        assert w == (hi - lo) * rb
        return buf[:w]

    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        files = list(ex.map(one, range(k)))
    return FastqShards(files, sum_q, read_len)
