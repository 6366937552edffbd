"""mzML record batches (exon_gpu_mzml_next_batch) against the pure-Python reading of the same files (oracle.mzml_rows) and the
reference's own known answers: 2 + 2 spectra (slt/mzml-functions.slt:41-49), contains_peak(mz.mz, 200, 1) / (0, 1) (:9-17),
bin_vectors(mz.mz, intensity.intensity, 200, 10, 1) = [0, 0, 0, 0, 203667.40002441406, 0, ...] (:22-25), wavelength 0..14 of the
reader test (exon-mzml/src/mzml_reader/parser.rs:121-141)."""
import gzip
import math
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from exon_b200 import _abi
from exon_b200._abi import ExonGpuError
from exon_b200.runtime import export_reader

pytestmark = pytest.mark.gpu
ALL = (0, 1, 2, 3, 4, 5, 6)


def fixture(name):
    p = os.path.join(GOLDEN, name)
    with (gzip.open(p) if name.endswith(".gz") else open(p, "rb")) as f:
        return f.read()


def gpu_rows(ctx, files, projection=ALL, batch_rows=8192):
    rows, sizes = [], []
    with ctx.open_mzml(projection=projection, batch_rows=batch_rows) as s:
        for f in files:
            s.feed(f)
        for b in s.batches():
            rb = b.to_pyarrow()
            sizes.append(rb.num_rows)
            rows += rb.to_pylist()
    return rows, sizes


def as_want(r, projection=ALL):
    names = ["id", "mz", "intensity", "wavelength", "cv_params", "precursor_mz", "precusor_charge"]
    w = {"id": r["id"], "precursor_mz": r["precursor_mz"], "precusor_charge": r["precursor_charge"],
         "cv_params": [{"accession": a, "name": n, "value": v} for a, n, v in r["cv_params"]]}
    for k in ("mz", "intensity", "wavelength"):
        w[k] = None if r[k] is None else {k: r[k]}
    return {names[p]: w[names[p]] for p in projection}


@pytest.mark.parametrize("name", ["test.mzML", "pyoteomics.mzML.gz"])
def test_reference_fixtures(gpu_ctx, name):
    text = fixture(name)
    want = oracle.mzml_rows(text)
    got, sizes = gpu_rows(gpu_ctx, [text])
    assert sizes == [2] and len(got) == 2                                  # slt/mzml-functions.slt:41-49
    assert got == [as_want(r) for r in want]
    if name == "test.mzML":
        assert got[0]["wavelength"]["wavelength"] == [float(i) for i in range(15)] and got[0]["mz"] is None   # parser.rs:121-141
        assert got[0]["precursor_mz"] == 643.034396630915 and got[0]["precusor_charge"] == 3
    else:
        mz, inten = np.array(got[0]["mz"]["mz"]), np.array(got[0]["intensity"]["intensity"])
        assert ((mz >= 199.0) & (mz <= 201.0)).any() and not ((mz >= -1.0) & (mz <= 1.0)).any()              # contains_peak, :9-17
        bins = [math.fsum(inten[(mz >= 200.0 + k) & (mz < 201.0 + k)]) for k in range(10)]
        assert bins[4] == 203667.40002441406 and sum(1 for b in bins if b) == 1                               # bin_vectors, :22-25


def test_projections_batches_and_files(gpu_ctx):
    from synth import mzml

    sh = mzml.shards(700, 3, peaks=23)
    files = [bytes(f) for f in sh.files] + [fixture("test.mzML")]
    want = [r for f in files for r in oracle.mzml_rows(f)]
    got, sizes = gpu_rows(gpu_ctx, files, batch_rows=100)
    assert got == [as_want(r) for r in want]
    per_file = [len(oracle.mzml_rows(f)) for f in files]
    assert sizes == [min(100, n - o) for n in per_file for o in range(0, n, 100)]   # batches never span files
    for proj in [(2,), (0,), (4, 0), (6, 5, 1), (3, 2)]:
        got, _ = gpu_rows(gpu_ctx, files, projection=proj, batch_rows=64)
        assert got == [as_want(r, proj) for r in want], proj
    # the fused query still agrees with the columns
    with gpu_ctx.open_mzml() as s:
        for f in sh.files:
            s.feed(f)
        ssum, n_sel, n_sp = s.filter_sum(sh.lo, sh.hi)
    assert n_sp == sh.n and n_sel == sh.truth_count and math.isclose(ssum, sh.truth_sum, rel_tol=1e-9)


def test_xml_details(gpu_ctx):
    doc = (b'<?xml version="1.0"?><mzML><run><spectrumList count="3">'
           b'<spectrum index="0" id="a &amp; b &#65;&#x42;" defaultArrayLength="0"><cvParam cvRef="MS" accession="MS:1" name="n&lt;1" value=""/>'
           b'<cvParam cvRef="MS" accession="MS:2" name="two" value="v&quot;2"/><scanList count="1"><cvParam accession="MS:9" name="nested"/></scanList>'
           b'<binaryDataArrayList count="1"><binaryDataArray encodedLength="0"><cvParam accession="MS:1000514" name="m/z array"/>'
           b'<cvParam accession="MS:1000523" name="64-bit float"/><cvParam accession="MS:1000576" name="no compression"/><binary></binary>'
           b'</binaryDataArray></binaryDataArrayList></spectrum>'
           b'<spectrum index="1" id="s2" defaultArrayLength="2"><precursorList count="2"><precursor><selectedIonList count="2"><selectedIon>'
           b'<cvParam accession="MS:1000744" name="selected ion m/z"/><cvParam accession="MS:1000744" name="selected ion m/z" value="445.25"/>'
           b'<cvParam accession="MS:1000041" name="charge state" value="-2"/></selectedIon><selectedIon><cvParam accession="MS:1000744" name="x" value="1"/>'
           b'</selectedIon></selectedIonList><activation/></precursor><precursor><selectedIonList count="1"><selectedIon>'
           b'<cvParam accession="MS:1000041" name="charge state" value="9"/></selectedIon></selectedIonList><activation/></precursor></precursorList>'
           b'<binaryDataArrayList count="1"><binaryDataArray encodedLength="12"><cvParam accession="MS:1000515" name="intensity array"/>'
           b'<cvParam accession="MS:1000521" name="32-bit float"/><cvParam accession="MS:1000576" name="no compression"/><binary>AACAPwAAAEA=</binary>'
           b'</binaryDataArray></binaryDataArrayList></spectrum>'
           b'<spectrum index="2" id="s3" defaultArrayLength="0"><binaryDataArrayList count="0"></binaryDataArrayList></spectrum>'
           b'</spectrumList></run></mzML>')
    want = oracle.mzml_rows(doc)
    assert want[0]["id"] == "a & b AB" and want[0]["cv_params"] == [("MS:1", "n<1", None), ("MS:2", "two", 'v"2')] and want[0]["mz"] == []
    assert want[1]["precursor_mz"] == 445.25 and want[1]["precursor_charge"] == -2 and want[1]["intensity"] == [1.0, 2.0]
    got, _ = gpu_rows(gpu_ctx, [doc])
    assert got == [as_want(r) for r in want]
    with gpu_ctx.open_mzml(projection=(0, 1, 2)) as s:      # and through the Arrow C stream
        s.feed(doc)
        tbl = export_reader(s).read_all()
        assert tbl.schema.names == ["id", "mz", "intensity"] and tbl.num_rows == 3
    bad = doc.replace(b'value="445.25"', b'value="abc"')
    with pytest.raises(ExonGpuError) as e:
        gpu_rows(gpu_ctx, [bad])
    assert e.value.code == _abi.ERR_PARSE                   # the reference unwraps the f64 parse
    with pytest.raises(ExonGpuError):
        gpu_ctx.open_mzml(projection=(7,))
