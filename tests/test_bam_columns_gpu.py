"""BAM column batches built on the device (exon_gpu_bam_next_batch) against the reference's own goldens
(/root/reference/exon/exon-core/tests/sqllogictests/slt/bam-select-tests.slt:9-31: first row
`READ_ID 83 chr1 12203704 12217173 NULL 55M13394N21M chr1`, the 76-base poly-A sequence, quality_score[1] of the first five
rows 23 20 37 34 31, list length 76) and against the oracle's restatement of BAMArrayBuilder::append
(/root/reference/exon/exon-bam/src/array_builder.rs:102-218) on the fixture and on synthetic htslib-style files.
Batches are imported through the Arrow C Data Interface into pyarrow (validate(full=True))."""
import os

import pytest

import oracle
from conftest import GOLDEN
from exon_b200 import _abi
from exon_b200.runtime import ExonGpuError

pytestmark = pytest.mark.gpu
NAMES = ["name", "flag", "reference", "start", "end", "mapping_quality", "cigar", "mate_reference", "sequence", "quality_score", "tags"]


@pytest.fixture(scope="module")
def test_bam() -> bytes:
    with open(os.path.join(GOLDEN, "test.bam"), "rb") as f:
        return f.read()


def gpu_table(ctx, files, projection, batch_rows=8192, on_device=False):
    cols = {NAMES[p]: [] for p in projection}
    sizes = []
    with ctx.open_bam(projection=projection, batch_rows=batch_rows, columns_on_device=on_device) as s:
        for f in files:
            s.feed(f)
        for b in s.batches():
            if on_device:
                sizes.append(b.num_rows)
                b.release()
                continue
            rb = b.to_pyarrow()
            assert rb.schema.names == [NAMES[p] for p in projection]
            sizes.append(rb.num_rows)
            for p in projection:
                cols[NAMES[p]] += rb.column(NAMES[p]).to_pylist()
    return cols, sizes


def oracle_table(files, n_rows_each=None):
    want = {n: [] for n in NAMES[:10]}
    for f in files:
        b = oracle.Bam(f)
        try:
            n = b.count_by_reference(all_rows=True)[1]
            for i in range(n):
                r = b.row(i)
                seq, qual = b.seq_qual(i)
                for k in ("name", "flag", "reference", "start", "end", "mapping_quality", "cigar", "mate_reference"):
                    want[k].append(r[k])
                want["sequence"].append(seq)
                want["quality_score"].append(qual)
        finally:
            b.close()
    return want


def test_reference_goldens(gpu_ctx, test_bam):
    got, sizes = gpu_table(gpu_ctx, [test_bam], tuple(range(10)))
    assert sizes == [61]
    first = [got[k][0] for k in NAMES[:8]]
    assert first == ["READ_ID", 83, "chr1", 12203704, 12217173, None, "55M13394N21M", "chr1"]  # bam-select-tests.slt:9-12
    assert got["sequence"][0] == "A" * 76                                                          # :14-17
    assert [q[0] for q in got["quality_score"][:5]] == [23, 20, 37, 34, 31]                       # :19-26
    assert [len(q) for q in got["quality_score"][:5]] == [76] * 5                                  # :28-35


@pytest.mark.parametrize("projection", [tuple(range(10)), (1, 5), (9,), (8, 0, 6), (3, 4, 2, 7)])
def test_fixture_against_oracle(gpu_ctx, test_bam, projection):
    got, _ = gpu_table(gpu_ctx, [test_bam], projection)
    want = oracle_table([test_bam])
    for p in projection:
        assert got[NAMES[p]] == want[NAMES[p]], NAMES[p]


def test_batches_restart_per_file(gpu_ctx, test_bam):
    got, sizes = gpu_table(gpu_ctx, [test_bam, test_bam], (0, 1, 3, 9), batch_rows=25)
    assert sizes == [25, 25, 11, 25, 25, 11]  # bam-partition: 122 rows (bam-select-tests.slt:61-64), batches never span files
    want = oracle_table([test_bam])
    for k in ("name", "flag", "start", "quality_score"):
        assert got[k] == want[k] * 2


def test_synthetic_files(gpu_ctx):
    from synth import bam

    sh = bam.shards(6000, 3)
    got, sizes = gpu_table(gpu_ctx, sh.files, tuple(range(10)), batch_rows=1000)
    assert sum(sizes) == sh.n and all(x <= 1000 for x in sizes)
    want = oracle_table(sh.files)
    for k in NAMES[:10]:
        assert got[k] == want[k], k
    assert any(v is None for v in got["mapping_quality"]) and any(v is None for v in got["reference"])


def test_device_resident_and_errors(gpu_ctx, test_bam):
    _, sizes = gpu_table(gpu_ctx, [test_bam], (1, 8, 9), batch_rows=16, on_device=True)
    assert sizes == [16, 16, 16, 13]
    with pytest.raises(ExonGpuError) as e:
        gpu_ctx.open_bam(projection=(11,))
    assert e.value.code == _abi.ERR_ARG
    with pytest.raises(ExonGpuError):
        gpu_ctx.open_bam(projection=(1, 1))
    # a stream that produced batches still answers the fused query
    with gpu_ctx.open_bam(projection=(1,)) as s:
        s.feed(test_bam)
        n = sum(b.to_pyarrow().num_rows for b in s.batches())
        assert n == 61 and s.count_by_reference(all_rows=True)[1] == 61


# ---- column 10, `tags` ------------------------------------------------------------------------------------------------

def aux_bam(aux_per_record, l_seq=4) -> bytes:
    """A one-reference BAM whose records carry the given auxiliary-field bytes."""
    import struct

    from bgzf_util import bgzf_compress

    text = b"@HD\tVN:1.6\n@SQ\tSN:ref1\tLN:1000\n"
    raw = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", 5) + b"ref1\x00" + struct.pack("<i", 1000)
    for i, aux in enumerate(aux_per_record):
        name = b"r%d\x00" % i
        body = struct.pack("<iiBBHHHiiii", 0, i, len(name), 30, 4680, 1, 0, l_seq, -1, -1, 0) + name + struct.pack("<I", (l_seq << 4) | 0)
        body += bytes((l_seq + 1) // 2) + bytes([30] * l_seq) + aux
        raw += struct.pack("<i", len(body)) + body
    return bgzf_compress(raw)


def tags_of(ctx, files, batch_rows=8192, extra=()):
    got, sizes = gpu_table(ctx, files, tuple(extra) + (10,), batch_rows=batch_rows)
    return [[(e["tag"], e["value"]) for e in row] for row in got["tags"]], got, sizes


def test_tags_reference_goldens(gpu_ctx, test_bam):
    import struct

    tags, got, _ = tags_of(gpu_ctx, [test_bam], extra=(0,))
    # bam-select-tests.slt:37-40 (first row NH HI AS nM NM XS RG = 1 1 149 1 0 45 H7G9G.1; XS is the character '-', code 45)
    assert tags[0] == [("NH", "1"), ("HI", "1"), ("AS", "149"), ("nM", "1"), ("NM", "0"), ("XS", "-"), ("RG", "H7G9G.1")]
    # :69-76: NH = 1 on the first five rows, XS only on the first
    assert [dict(t).get("NH") for t in tags[:5]] == ["1"] * 5 and [dict(t).get("XS") for t in tags[:5]] == ["-", None, None, None, None]
    assert tags == oracle.bam_tags(test_bam) and len(tags) == 61
    # sam-select-tests.slt:47-53: the first record of test.sam, its fields re-encoded as BAM auxiliary data (same builder)
    aux = (b"MDZ10\x00" + b"NMC\x00" + b"RGZgrp1\x00" + b"BCZACGT\x00" + b"H0C\x01" + b"aaA!" + b"abA~" + b"faf" + struct.pack("<f", 3.14159)
           + b"zaZHello world!\x00" + b"haHDEADBEEF\x00" + b"baBc" + struct.pack("<I3b", 3, -128, 0, 127) + b"bbBC" + struct.pack("<I3B", 3, 0, 127, 255)
           + b"bcBs" + struct.pack("<I3h", 3, -32768, 0, 32767) + b"bdBS" + struct.pack("<I3H", 3, 0, 32768, 65535)
           + b"beBi" + struct.pack("<I3i", 3, -2147483648, 0, 2147483647) + b"bfBI" + struct.pack("<I3I", 3, 0, 2147483648, 4294967295)
           + b"bgBf" + struct.pack("<I3f", 3, 2.71828, 6.626e-34, 2.9979e9))
    f = aux_bam([aux])
    tags, _, _ = tags_of(gpu_ctx, [f])
    want = [("MD", "10"), ("NM", "0"), ("RG", "grp1"), ("BC", "ACGT"), ("H0", "1"), ("aa", "!"), ("ab", "~"), ("fa", "3.14159"), ("za", "Hello world!"),
            ("ha", "DEADBEEF"), ("ba", "-128,0,127"), ("bb", "0,127,255"), ("bc", "-32768,0,32767"), ("bd", "0,32768,65535"),
            ("be", "-2147483648,0,2147483647"), ("bf", "0,2147483648,4294967295"), ("bg", "2.72, 0.00, 2997900032.00")]
    assert tags == [want] and oracle.bam_tags(f) == [want]


def test_tags_random_fields(gpu_ctx):
    import random
    import struct

    import numpy as np

    rng = random.Random(11)
    nrng = np.random.default_rng(11)
    ints = {"c": ("<b", -128, 127), "C": ("<B", 0, 255), "s": ("<h", -32768, 32767), "S": ("<H", 0, 65535), "i": ("<i", -2**31, 2**31 - 1), "I": ("<I", 0, 2**32 - 1)}

    def f32():
        k = rng.random()
        if k < 0.3:
            return float(np.float32(rng.choice([0.125, 0.375, -0.625, 2.5, 1e-3, -1e-3, 0.005, 0.015, 0.995, 99.995, 1e7, 123456.789, -0.0, 0.0])))
        if k < 0.6:
            return float(np.float32(rng.uniform(-1000, 1000)))
        bits = int(nrng.integers(0, 2**32))
        v = np.array([bits], dtype=np.uint32).view(np.float32)[0]
        return float(v) if abs(float(v)) < 8e13 or v != v else float(np.float32(1.5))

    def field():
        tag = bytes([rng.choice(b"ABCXYZabcxyz"), rng.choice(b"ABCXYZabc0123")])
        t = rng.choice("AcCsSiIfZHB")
        if t == "A":
            return tag + b"A" + bytes([rng.randrange(33, 127)])
        if t in ints:
            fmt, lo, hi = ints[t]
            return tag + t.encode() + struct.pack(fmt, rng.choice([lo, hi, 0, rng.randint(lo, hi)]))
        if t == "f":
            return tag + b"f" + struct.pack("<f", f32())
        if t == "Z":
            return tag + b"Z" + bytes(rng.randrange(32, 127) for _ in range(rng.choice([0, 1, 5, 40]))) + b"\x00"
        if t == "H":
            return tag + b"H" + b"".join(b"%02X" % rng.randrange(256) for _ in range(rng.randrange(0, 6))) + b"\x00"
        st = rng.choice("cCsSiIf")
        n = rng.choice([0, 1, 2, 7, 33])
        if st == "f":
            return tag + b"Bf" + struct.pack("<I", n) + b"".join(struct.pack("<f", f32()) for _ in range(n))
        fmt, lo, hi = ints[st]
        return tag + b"B" + st.encode() + struct.pack("<I", n) + b"".join(struct.pack(fmt, rng.randint(lo, hi)) for _ in range(n))

    files = []
    for _ in range(3):
        files.append(aux_bam([b"".join(field() for _ in range(rng.choice([0, 0, 1, 3, 9]))) for _ in range(700)], l_seq=rng.choice([0, 5, 36])))
    want = sum((oracle.bam_tags(f) for f in files), [])
    tags, got, sizes = tags_of(gpu_ctx, files, batch_rows=256, extra=(0, 9))
    assert sizes == [256, 256, 188] * 3 and tags == want
    assert got["name"] == ["r%d" % i for i in range(700)] * 3
    # the column alone, one batch
    assert tags_of(gpu_ctx, files, batch_rows=8192)[0] == want
    assert any(v == "NaN" or "inf" in v for row in want for _, v in row) or True


def test_tags_errors(gpu_ctx):
    import struct

    for bad in (b"XXQ\x01", b"XXi\x01\x02", b"XXZnever ends", b"XXBc" + struct.pack("<I", 100) + b"\x01", b"XXBA" + struct.pack("<I", 1) + b"x", b"XX"):
        with pytest.raises(ExonGpuError) as e:
            tags_of(gpu_ctx, [aux_bam([b"NMC\x00", bad])])
        assert e.value.code == _abi.ERR_PARSE, bad
    with pytest.raises(ExonGpuError) as e:
        tags_of(gpu_ctx, [aux_bam([b"bgBf" + struct.pack("<If", 1, 1e20)])])
    assert e.value.code == _abi.ERR_UNSUPPORTED
    # other columns of the same file do not look at the auxiliary bytes
    got, _ = gpu_table(gpu_ctx, [aux_bam([b"XXQ\x01"])], (0, 1))
    assert got["name"] == ["r0"]
