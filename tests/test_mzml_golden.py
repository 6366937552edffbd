"""The mzML oracle (oracle/mzml_oracle.c) against the reference's known answers: the decode vectors of
exon-mzml/src/mzml_reader/binary_conversion.rs:126-135, the spectrum counts of slt/mzml-functions.slt:41-49, the reader
test of parser.rs:121-141, contains_peak of slt/mzml-functions.slt:9-17 -- and an independent xml.etree reading."""
import base64
import gzip
import math
import os
import struct
import xml.etree.ElementTree as ET
import zlib

import numpy as np

import oracle
from conftest import GOLDEN

NS = "{http://psi.hupo.org/ms/mzml}"


def fixture(name):
    p = os.path.join(GOLDEN, name)
    with (gzip.open(p) if name.endswith(".gz") else open(p, "rb")) as f:
        return f.read()


def py_spectra(text: bytes):
    """[(mz | None, intensity | None, wavelength | None)] via ElementTree + base64 + zlib + struct."""
    out = []
    for sp in ET.fromstring(text).iter(NS + "spectrum"):
        arrs = {}
        for bda in sp.iter(NS + "binaryDataArray"):
            acc = {c.get("accession") for c in bda.iter(NS + "cvParam")}
            kind = "mz" if "MS:1000514" in acc else "intensity" if "MS:1000515" in acc else "wavelength" if "MS:1000617" in acc else None
            b = bda.find(NS + "binary")
            if kind is None or b is None or not (b.text or "").strip():
                continue
            raw = base64.b64decode(b.text.strip())
            if "MS:1000574" in acc:
                raw = zlib.decompress(raw)
            w, f = (4, "f") if "MS:1000521" in acc else (8, "d")
            arrs[kind] = [float(x) for x in struct.unpack("<%d%s" % (len(raw) // w, f), raw[: len(raw) // w * w])]
        out.append((arrs.get("mz"), arrs.get("intensity"), arrs.get("wavelength")))
    return out


def test_reference_decode_vectors():
    a = oracle.mzml_decode_binary(b"AAAAAAAALkAAAAAAAAAsQAAAAAAAACpAAAAAAAAAKEAAAAAAAAAmQAAAAAAAACRAAAAAAAAAIkAAAAAAAAAgQAAAAAAAABxAAAAAAAAAGEAAAAAAAAAUQAAAAAAAABBAAAAAAAAACEAAAAAAAAAAQAAAAAAAAPA/", False, False)
    assert a == [15.0, 14.0, 13.0, 12.0, 11.0, 10.0, 9.0, 8.0, 7.0, 6.0, 5.0, 4.0, 3.0, 2.0, 1.0]
    b = oracle.mzml_decode_binary(b"eJxjYEABDhBKAEpLQGkFKK0CpTWgtA6UNoDSRg4AZlQDYw==", True, False)
    assert b == [0.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0, 14.0, 16.0, 18.0]


def test_reference_fixtures():
    t = fixture("test.mzML")
    r = oracle.mzml_scan(t)
    assert r.n_spectra == 2                                              # mzml-functions.slt:46-49 (the .gz twin)
    r0 = oracle.mzml_scan(t, spectrum=0)
    assert r0.kind_count[2] == 15 and r0.kind_sum[2] == 105.0             # parser.rs:121-141: wavelength 0..14
    assert r0.kind_count[1] == 15 and r0.kind_sum[1] == 120.0 and r0.kind_count[0] == 0
    p = fixture("pyoteomics.mzML.gz")
    rp = oracle.mzml_scan(p)
    assert rp.n_spectra == 2                                             # mzml-functions.slt:41-44
    want = py_spectra(p)
    for i, (mz, inten, _) in enumerate(want):
        ri = oracle.mzml_scan(p, spectrum=i)
        assert ri.kind_count[0] == len(mz) == 19914 and ri.kind_count[1] == len(inten)
        assert math.isclose(ri.kind_sum[0], math.fsum(mz), rel_tol=1e-12) and math.isclose(ri.kind_sum[1], math.fsum(inten), rel_tol=1e-12)
    mz0 = np.array(want[0][0])
    assert ((mz0 >= 199.0) & (mz0 <= 201.0)).any() and not ((mz0 >= -1.0) & (mz0 <= 1.0)).any()   # contains_peak(mz, 200, 1) / (0, 1)
    for lo, hi in [(500.0, 600.0), (200.0, 200.5), (0.0, 1.0)]:
        r = oracle.mzml_scan(p, lo, hi)
        ws = math.fsum(i for mz, inten, _ in want for m, i in zip(mz, inten) if lo <= m <= hi)
        wc = sum(1 for mz, inten, _ in want for m in mz[: len(inten)] if lo <= m <= hi)
        assert r.n_selected == wc and math.isclose(r.sum, ws, rel_tol=1e-12, abs_tol=1e-9)


def test_synthetic_truth():
    from synth import mzml

    sh = mzml.shards(3000, 3, peaks=57)
    tot, cnt, n = 0.0, 0, 0
    for f in sh.files:
        r = oracle.mzml_scan(f, sh.lo, sh.hi)
        tot += r.sum
        cnt += r.n_selected
        n += r.n_spectra
    assert n == sh.n and cnt == sh.truth_count and math.isclose(tot, sh.truth_sum, rel_tol=1e-12)
    assert py_spectra(bytes(sh.files[0]))[0][0] == sorted(py_spectra(bytes(sh.files[0]))[0][0])
