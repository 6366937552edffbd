"""GPU parity of the fused mzML scan -> m/z range filter -> SUM(intensity) (BASELINE configs[4]) through the C ABI:
selected-peak counts bit-exact, f64 sums within 1e-6 relative (north_star) of the oracle and of the generator's truth."""
import gzip
import math
import os

import numpy as np
import pytest

import oracle
from bgzf_util import bgzf_compress
from conftest import GOLDEN
from exon_b200 import _abi
from exon_b200._abi import ExonGpuError

pytestmark = pytest.mark.gpu
RANGES = [(500.0, 600.0), (None, None), (100.0, 100.5), (0.0, 1.0), (1999.0, 5000.0), (600.0, 500.0)]


def fixture(name):
    p = os.path.join(GOLDEN, name)
    with (gzip.open(p) if name.endswith(".gz") else open(p, "rb")) as f:
        return f.read()


def close(a, b):
    return math.isclose(a, b, rel_tol=1e-6, abs_tol=1e-9)


def gpu(ctx, feeds, lo=None, hi=None, gz=False):
    with ctx.open_mzml() as s:
        for f in feeds:
            (s.feed_gzip if gz else s.feed)(f)
        return s.filter_sum(lo, hi)


def test_reference_fixtures(gpu_ctx):
    t = fixture("test.mzML")
    s, n, sp = gpu(gpu_ctx, [t])
    assert sp == 2 and n == 0 and s == 0.0        # 2 spectra (slt/mzml-functions.slt:46-49); neither has an m/z array
    assert gpu(gpu_ctx, [bgzf_compress(t)], gz=True)[2] == 2 and gpu(gpu_ctx, [gzip.compress(t)], gz=True)[2] == 2
    assert gpu(gpu_ctx, [t, t])[2] == 4
    # pyoteomics: spectrum 0 is uncompressed (f64 m/z, f32 intensity), spectrum 1 is zlib-compressed f32 / f32: base64 ->
    # device inflate -> values.  Its per-spectrum sums are the fixture facts of SURVEY appendix A (17 648 193.83 / 69 381 842.12)
    p = fixture("pyoteomics.mzML.gz")
    for lo, hi in RANGES + [(None, None)]:
        r = oracle.mzml_scan(p, lo, hi)
        s, n, sp = gpu(gpu_ctx, [p], lo, hi)
        assert sp == 2 and n == r.n_selected and close(s, r.sum), (lo, hi)
    s, n, sp = gpu(gpu_ctx, [p])
    assert n == 2 * 19914 and close(s, 2 * 69381842.11895752)
    assert gpu(gpu_ctx, [bgzf_compress(p)], 500.0, 600.0, gz=True)[1] == oracle.mzml_scan(p, 500.0, 600.0).n_selected
    cut = p.index(b"<spectrum ", p.index(b"<spectrum ") + 10)
    first = p[:cut] + b"</spectrumList></run></mzML>\n"   # the document with its first spectrum only
    for lo, hi in RANGES:
        r = oracle.mzml_scan(first, lo, hi)
        s, n, sp = gpu(gpu_ctx, [first], lo, hi)
        assert sp == 1 and n == r.n_selected and close(s, r.sum), (lo, hi)
    assert gpu(gpu_ctx, [first], 199.0, 201.0)[1] > 0 and gpu(gpu_ctx, [first], -1.0, 1.0)[1] == 0   # contains_peak goldens


@pytest.mark.parametrize("peaks", [1, 2, 3, 57, 200, 1001])
def test_synthetic(gpu_ctx, peaks):
    from synth import mzml

    sh = mzml.shards(4000 if peaks < 500 else 300, 3, peaks=peaks)
    with gpu_ctx.open_mzml() as st:
        for f in sh.files:
            st.feed(f)
        for lo, hi in RANGES:
            want_s, want_n = 0.0, 0
            for f in sh.files:
                r = oracle.mzml_scan(f, lo, hi)
                want_s += r.sum
                want_n += r.n_selected
            s, n, sp = st.filter_sum(lo, hi)
            assert sp == sh.n and n == want_n and close(s, want_s), (peaks, lo, hi)
        s, n, _ = st.filter_sum(sh.lo, sh.hi)
        assert n == sh.truth_count and close(s, sh.truth_sum)


@pytest.mark.parametrize("n_files", [1, 2, 4, 5, 8, 16, 22, 33])
def test_file_counts(gpu_ctx, n_files):
    """Any number of files per stream (a rank's share of a sharded set): 4..21 files used to put the per-file spectrum table on
    top of the query's result slots."""
    from synth import mzml

    sh = mzml.shards(40 * n_files, n_files, peaks=31)
    with gpu_ctx.open_mzml() as st:
        for f in sh.files:
            st.feed(f)
        for _ in range(2):
            s, n, sp = st.filter_sum(sh.lo, sh.hi)
            assert sp == sh.n and n == sh.truth_count and close(s, sh.truth_sum), (n_files, n, sh.truth_count)


def test_feeds_and_f32(gpu_ctx):
    import base64
    import struct
    from synth import mzml

    sh = mzml.shards(2000, 2, peaks=33)
    want = sum(oracle.mzml_scan(f, 500.0, 600.0).sum for f in sh.files)
    # device-resident ranges at odd alignments, ragged host feeds cut at line boundaries, BGZF
    for shift in (0, 5):
        bufs = []
        with gpu_ctx.open_mzml() as st:
            for f in sh.files:
                d = gpu_ctx.device_buffer(f.size + shift + 64)
                d.upload(np.ascontiguousarray(f), offset=shift)
                bufs.append(d)
                st.feed(None, device_ptr=d.ptr + shift, nbytes=f.size)
            assert close(st.filter_sum(500.0, 600.0)[0], want)
        for d in bufs:
            d.free()
    with gpu_ctx.open_mzml() as st:
        for f in sh.files:
            b = f.tobytes()
            cuts = [0] + [b.index(b"\n", o) + 1 for o in range(50_000, len(b) - 1, 50_000)] + [len(b)]
            for a, c in zip(cuts[:-1], cuts[1:]):
                st.feed(b[a:c], is_last=c == len(b))
        assert close(st.filter_sum(500.0, 600.0)[0], want)
    assert close(gpu(gpu_ctx, [bgzf_compress(f.tobytes()) for f in sh.files], 500.0, 600.0, gz=True)[0], want)
    # 32-bit arrays of unequal length, whitespace around the payload, a spectrum with an empty <binary/>
    mz = [100.5, 550.25, 560.0, 900.0, 1500.0]
    inten = [1.0, 2.0, 4.0, 8.0]
    def arr(vals, acc, f32):
        raw = struct.pack("<%d%s" % (len(vals), "f" if f32 else "d"), *vals)
        return ('<binaryDataArray encodedLength="0"><cvParam cvRef="MS" accession="MS:%s" name="x" value=""/>'
                '<cvParam cvRef="MS" accession="MS:1000576" name="no compression" value=""/>'
                '<cvParam cvRef="MS" accession="MS:%s" name="y" value=""/><binary>\n   %s \n</binary></binaryDataArray>'
                % ("1000521" if f32 else "1000523", acc, base64.b64encode(raw).decode()))
    doc = ('<mzML><run><spectrumList count="2"><spectrum index="0" id="a"><binaryDataArrayList count="2">' + arr(mz, "1000514", True) +
           arr(inten, "1000515", True) + '</binaryDataArrayList></spectrum>\n<spectrum index="1" id="b"><binaryDataArrayList count="2">'
           '<binaryDataArray encodedLength="0"><cvParam cvRef="MS" accession="MS:1000514" name="m/z array" value=""/><binary></binary>'
           '</binaryDataArray></binaryDataArrayList></spectrum></spectrumList></run></mzML>\n').encode()
    r = oracle.mzml_scan(doc, 500.0, 1000.0)
    assert (r.sum, r.n_selected, r.n_spectra) == (14.0, 3, 2)
    assert gpu(gpu_ctx, [doc], 500.0, 1000.0) == (14.0, 3, 2)
    bad = doc.replace(b"<binary>\n   AA", b"<binary>\n   A*")
    if bad != doc:
        with pytest.raises(ExonGpuError):
            gpu(gpu_ctx, [bad], 0.0, 5000.0)


def zlib_variant(text: bytes, every: int = 2) -> bytes:
    """Re-encode every `every`-th binary array of a synthetic document with zlib (MS:1000576 -> MS:1000574), as pyteomics /
    ProteoWizard writers do; sizes stay declared by defaultArrayLength."""
    import base64
    import re
    import zlib

    out, pos, k = [], 0, 0
    for m in re.finditer(rb'<binaryDataArray [^>]*>.*?</binaryDataArray>', text, re.S):
        out.append(text[pos:m.start()])
        pos = m.end()
        blk = m.group(0)
        k += 1
        if k % every == 0:
            b = re.search(rb"<binary>(.*?)</binary>", blk, re.S)
            payload = base64.b64encode(zlib.compress(base64.b64decode(b.group(1).strip()), 6))
            blk = blk[:b.start(1)] + payload + blk[b.end(1):]
            blk = blk.replace(b"MS:1000576", b"MS:1000574")
            blk = re.sub(rb'encodedLength="\d+"', b'encodedLength="%d"' % len(payload), blk)
        out.append(blk)
    out.append(text[pos:])
    return b"".join(out)


@pytest.mark.parametrize("every", [1, 2, 3])
def test_zlib_arrays_synthetic(gpu_ctx, every):
    from synth import mzml

    sh = mzml.shards(600, 2, peaks=150)
    files = [zlib_variant(bytes(f), every) for f in sh.files]
    assert b"MS:1000574" in files[0]
    for lo, hi in RANGES[:3] + [(None, None)]:
        want = [oracle.mzml_scan(f, lo, hi) for f in files]
        plain = [oracle.mzml_scan(bytes(f), lo, hi) for f in sh.files]
        assert [w.n_selected for w in want] == [w.n_selected for w in plain]
        s, n, sp = gpu(gpu_ctx, files, lo, hi)
        assert sp == sh.n and n == sum(w.n_selected for w in want) and close(s, sum(w.sum for w in want))


def test_zlib_array_errors(gpu_ctx):
    from synth import mzml

    f = zlib_variant(bytes(mzml.shards(20, 1, peaks=30).files[0]), 1)
    bad = f.replace(b'defaultArrayLength="30"', b'defaultArrayLength="31"')   # declared size differs from the stream
    with pytest.raises(ExonGpuError) as e:
        gpu(gpu_ctx, [bad])
    assert e.value.code == _abi.ERR_PARSE
    import re
    nodefault = re.sub(rb' defaultArrayLength="\d+"', b"", f)
    with pytest.raises(ExonGpuError) as e:
        gpu(gpu_ctx, [nodefault])
    assert e.value.code == _abi.ERR_PARSE
