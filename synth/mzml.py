"""Synthetic mzML shards for BASELINE configs[4]: spectra x peaks, m/z sorted uniform [100, 2000),
intensity lognormal(8, 2), 64-bit uncompressed base64 arrays; truth of `SUM(intensity) WHERE mz BETWEEN lo AND hi`."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

SEED = 20241020
HEADER = (b'<?xml version="1.0" encoding="utf-8"?>\n<mzML xmlns="http://psi.hupo.org/ms/mzml" version="1.1.0">\n'
          b'  <cvList count="1">\n    <cv id="MS" fullName="Proteomics Standards Initiative Mass Spectrometry Ontology" version="4.1.0"/>\n'
          b'  </cvList>\n  <run id="synthetic">\n    <spectrumList count="%d">\n')
FOOTER = b"    </spectrumList>\n  </run>\n</mzML>\n"
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsynth_mzml.so")
        src = os.path.join(_HERE, "mzml_format.c")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", so, src, "-lm"], check=True)
        _LIB = C.CDLL(so)
        _LIB.synth_mzml_spectrum_max_bytes.restype = C.c_int64
        _LIB.synth_mzml_spectrum_max_bytes.argtypes = [C.c_int32]
        _LIB.synth_mzml_format.restype = C.c_int64
        _LIB.synth_mzml_format.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
    return _LIB


@dataclass
class MzmlShards:
    files: list            # uint8 arrays, one complete mzML document each
    sel_sum: np.ndarray    # per spectrum: sum of intensities with lo <= mz <= hi
    sel_cnt: np.ndarray
    peaks: int
    lo: float
    hi: float

    @property
    def n(self) -> int:
        return int(self.sel_sum.size)

    @property
    def truth_sum(self) -> float:
        import math

        return math.fsum(self.sel_sum.tolist())

    @property
    def truth_count(self) -> int:
        return int(self.sel_cnt.sum())


def shards(n: int, k: int = 8, peaks: int = 200, lo: float = 500.0, hi: float = 600.0, seed: int = SEED, alloc=None,
           threads: int | None = None) -> MzmlShards:
    L = _lib()
    edges = [(n * i) // k for i in range(k + 1)]
    alloc = alloc or (lambda nb: np.empty(nb, dtype=np.uint8))
    sel_sum = np.empty(n, np.float64)
    sel_cnt = np.empty(n, np.int64)
    mx = int(L.synth_mzml_spectrum_max_bytes(peaks))

    def one(i):
        a, b = edges[i], edges[i + 1]
        hdr = HEADER % (b - a)
        buf = alloc(len(hdr) + (b - a) * mx + len(FOOTER))
        buf[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
        w = L.synth_mzml_format(seed, a, b - a, peaks, lo, hi, buf[len(hdr):].ctypes.data, sel_sum[a:b].ctypes.data,
                                sel_cnt[a:b].ctypes.data)
        o = len(hdr) + int(w)
        buf[o:o + len(FOOTER)] = np.frombuffer(FOOTER, dtype=np.uint8)
        return buf[:o + len(FOOTER)]

    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        files = list(ex.map(one, range(k)))
    return MzmlShards(files, sel_sum, sel_cnt, peaks, lo, hi)
