"""GFF record count + gff_region_filter: oracle against slt/gff-scan-tests.slt:80-92, GPU against the oracle."""
import gzip
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


def fixture():
    with gzip.open(os.path.join(GOLDEN, "test.gff.gz")) as f:
        return f.read()


def test_oracle_goldens():
    t = fixture()
    assert oracle.gff_filter_count(t) == (5000, 5000)                     # SELECT COUNT(*) FROM gff_scan('.../test.gff') -> 5000
    assert oracle.gff_filter_count(t + t)[0] == 10000                     # the partition directory -> 10000
    # the fixture holds 2513 sq0 + 2487 sq1 records (independent count below), every START is 8 (first row: sq0 caat 8 13)
    lines = [l.split(b"\t") for l in t.split(b"\n") if l and not l.startswith(b"#")]
    assert sum(1 for f in lines if f[0] == b"sq0") == 2513 and {f[3] for f in lines} == {b"8"}
    assert oracle.gff_filter_count(t, "sq0") == (2513, 5000) and oracle.gff_filter_count(t, "sq1")[0] == 2487
    assert oracle.gff_filter_count(t, "sq0", 8, 8)[0] == 2513 and oracle.gff_filter_count(t, "sq0", 9, 13)[0] == 0
    assert oracle.gff_filter_count(t, "sq", None, None)[0] == 0 and oracle.gff_filter_count(t, None, 1, 8)[0] == 5000
    assert oracle.gff_filter_count(b"##gff-version 3\n#c\nchr1\ts\tgene\t5\t9\t.\t+\t.\tID=x\n") == (1, 1)
    for bad in [b"chr1\ts\tgene\t5\t9\t.\t+\t.\n", b"chr1\ts\tgene\tx\t9\t.\t+\t.\tID=x\n", b"\n"]:
        with pytest.raises(ValueError):
            oracle.gff_filter_count(bad)


def synth(rng, n):
    names = ["chr1", "chr2", "chr10", "chrX", "scaffold_123456"]
    out = [b"##gff-version 3\n"]
    for i in range(n):
        if rng.random() < 0.02:
            out.append(b"# a comment\twith tabs\n" if rng.random() < 0.5 else b"##sequence-region chr1 1 1000\n")
        s = int(rng.integers(1, 10**7))
        attrs = "ID=g%d;Name=%s" % (i, "x" * int(rng.integers(0, 300)))
        out.append(("%s\tsrc\tgene\t%d\t%d\t.\t+\t.\t%s\n" % (names[int(rng.integers(0, 5))], s, s + 100, attrs)).encode())
    return b"".join(out)


@pytest.mark.gpu
def test_gpu_counts(gpu_ctx):
    from bgzf_util import bgzf_compress
    from exon_b200._abi import ExonGpuError, make_region

    t = fixture()
    with gpu_ctx.open_gff() as s:
        s.feed(t)
        assert s.rows() == 5000 and s.filter_count(make_region("sq0")) == 2513 and s.filter_count(make_region("sq0", 9, 13)) == 0
        s.feed_gzip(gzip.compress(t))
        s.feed_gzip(bgzf_compress(t))
        assert s.rows() == 15000 and s.filter_count(make_region("sq1", 8, 8)) == 3 * 2487
    rng = np.random.default_rng(9)
    texts = [synth(rng, 30_000), synth(rng, 10), b"", synth(rng, 3000)[:-1]]
    queries = [(None, None, None), ("chr1", None, None), ("chr10", 1, 5_000_000), ("chr", None, None), (None, 1000, 2_000_000), ("scaffold_123456", 5, None)]
    with gpu_ctx.open_gff() as s:
        for x in texts:
            for o in range(0, max(len(x), 1), 700_001):
                s.feed(x[o:o + 700_001], is_last=o + 700_001 >= len(x))
        for q in queries:
            want = sum(oracle.gff_filter_count(x, *q)[0] for x in texts)
            assert s.filter_count(make_region(*q)) == want, q
    with gpu_ctx.open_gff() as s:
        s.feed(b"chr1\ts\tgene\tx\t9\t.\t+\t.\tID=x\n")
        assert s.rows() == 1                                    # COUNT(*) does not read START
        with pytest.raises(ExonGpuError):
            s.filter_count(make_region("chr1", 1, 10))
