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


# ---- record batches, columns 0..7 ---------------------------------------------------------------------------------------

def test_oracle_row_golden():
    t = fixture()
    # slt/gff-scan-tests.slt:6-10: seqname, source, start, end, score, strand, phase = sq0 caat 8 13 NULL + NULL
    r = oracle.gff_rows(t, 1)[0]
    assert (r[0], r[1], r[3], r[4], r[5], r[6], r[7]) == (b"sq0", b"caat", 8, 13, None, b"+", None)


def synth_gff(rng, n):
    scores = [".", "0", "50", "0.95", "1e-5", "12.5", "3.4028235677973366e38"]
    lines = ["##gff-version 3"]
    for i in range(n):
        if i % 97 == 5:
            lines.append("# a comment")
        s = int(rng.integers(1, 10 ** 6))
        lines.append("\t".join(["sq%d" % (i % 7), ["caat", "x", "ensembl havana"][i % 3], ["gene", "exon", "CDS"][i % 3], str(s), str(s + int(rng.integers(0, 5000))),
                                scores[i % len(scores)], "+-"[i % 2], ".012"[i % 4], "ID=g%d;Name=n%d,m%d" % (i, i, i)]))
    return ("\n".join(lines) + "\n").encode()


def gpu_rows(ctx, files, projection=tuple(range(8)), gz=False):
    names = ["seqname", "source", "type", "start", "end", "score", "strand", "phase"]
    rows, sizes = [], []
    with ctx.open_gff(projection=projection) as s:
        for f in files:
            (s.feed_gzip if gz else s.feed)(f)
        while True:
            b = s.next_batch()
            if b is None:
                break
            rb = b.to_pyarrow()
            assert rb.schema.names == [names[p] for p in projection]
            sizes.append(rb.num_rows)
            cols = []
            for p in projection:
                col = rb.column(names[p])
                if p == 5:
                    bits = col.to_numpy(zero_copy_only=False).astype(np.float32).view(np.uint32)
                    cols.append([int(v) if ok else None for v, ok in zip(bits, np.asarray(col.is_valid()))])
                elif p in (3, 4):
                    cols.append(col.to_pylist())
                else:
                    cols.append([None if x is None else x.encode() for x in col.to_pylist()])
            rows += list(zip(*cols))
    return rows, sizes


@pytest.mark.gpu
def test_gpu_record_batches(gpu_ctx):
    from exon_b200 import _abi
    from exon_b200._abi import ExonGpuError

    t = fixture()
    rows, sizes = gpu_rows(gpu_ctx, [t])
    assert sizes == [5000]                                  # ONE batch per file: read_batch has no row limit
    assert rows[0] == (b"sq0", b"caat", b"gene", 8, 13, None, b"+", None)   # the slt row
    assert rows[:300] == oracle.gff_rows(t, 300)
    rng = np.random.default_rng(3)
    files = [synth_gff(rng, 1500), synth_gff(rng, 3), b"##only directives\n# and comments\n", synth_gff(rng, 400)]
    want = [r for f in files for r in oracle.gff_rows(f)]
    rows, sizes = gpu_rows(gpu_ctx, files)
    assert sizes == [1500, 3, 400] and rows == want
    rows, sizes = gpu_rows(gpu_ctx, [gzip.compress(f) for f in files], gz=True)
    assert sizes == [1500, 3, 400] and rows == want
    for projection in ((5, 0), (7, 6, 3), (1,), (4, 2)):
        rows, _ = gpu_rows(gpu_ctx, files, projection)
        assert rows == [tuple(r[p] for p in projection) for r in want]
    ok_line = "sq0\tcaat\tgene\t8\t13\t.\t+\t.\tID=a\n"
    for bad, proj in ((ok_line.replace("\t+\t", "\t.\t"), (6,)), (ok_line.replace("\t8\t", "\t0\t"), (3,)), (ok_line.replace("\t.\t+", "\tx\t+"), (5,)),
                      ("sq0\tcaat\tgene\t8\t13\t.\t+\t.\n", (0,)), (ok_line + "\n" + ok_line, (0,)), (ok_line.replace("+\t.", "+\t3"), (7,))):
        with pytest.raises(ExonGpuError) as e:
            gpu_rows(gpu_ctx, [bad.encode()], proj)
        assert e.value.code == _abi.ERR_PARSE, bad
    assert gpu_rows(gpu_ctx, [ok_line.replace("\t+\t", "\t.\t").encode()], (0, 3))[0] == [(b"sq0", 8)]   # strand not projected: not read
    with pytest.raises(ExonGpuError) as e:
        gpu_ctx.open_gff(projection=(9,))
    assert e.value.code == _abi.ERR_ARG


def gpu_attributes(ctx, files, projection=(8,)):
    out = []
    with ctx.open_gff(projection=projection) as s:
        for f in files:
            s.feed(f)
        while True:
            b = s.next_batch()
            if b is None:
                break
            rb = b.to_pyarrow()
            out.append([[(k, v) for k, v in row] for row in rb.column("attributes").to_pylist()])
    return out


@pytest.mark.gpu
def test_gpu_attributes_map(gpu_ctx):
    """Column 8: Map<Utf8, List<Utf8>> built like GFFArrayBuilder::append (array_builder.rs:142-160) -- with the builder's
    off-by-one for plain string values (the list is closed before the value is appended), which oracle.gff_attributes
    simulates with the builder's own state.  The fixture's rows are all `gene_id=caat1;gene_name=gene0`."""
    from exon_b200 import _abi
    from exon_b200._abi import ExonGpuError

    t = fixture()
    want = oracle.gff_attributes(t)
    assert want[0] == [("gene_id", []), ("gene_name", ["caat1"])] and want[1] == [("gene_id", ["gene0"]), ("gene_name", ["caat1"])]
    assert oracle.gff_attributes(t, reference_quirk=False)[0] == [("gene_id", ["caat1"]), ("gene_name", ["gene0"])]
    got = gpu_attributes(gpu_ctx, [t])
    assert len(got) == 1 and got[0] == want
    lines = ["ID=a;Parent=p1,p2;Note=x%3By%2Cz", ".", "ID=b;Dbxref=GO:1,GO:2,GO:3;empty=", "k%3D1=v;Alias=one,two;Name=last;", "ID=c"]
    f1 = "##gff-version 3\n" + "".join(f"sq{i}\tsrc\tgene\t{i + 1}\t{i + 9}\t.\t+\t.\t{x}\n" for i, x in enumerate(lines))
    f2 = "sq9\tsrc\tgene\t1\t2\t.\t-\t.\tID=second;Tags=a,b\n"
    files = [f1.encode(), f2.encode()]
    want = [oracle.gff_attributes(f) for f in files]          # one batch per file; the builder (and its quirk) restarts per batch
    assert want[0][0] == [("ID", []), ("Parent", ["a", "p1", "p2"]), ("Note", [])] and want[0][1] == []
    assert want[0][2][0] == ("ID", ["x;y,z"]) and want[0][3][0] == ("k=1", [""])
    got = gpu_attributes(gpu_ctx, files)
    assert got == want
    rb_rows = []
    with gpu_ctx.open_gff(projection=(0, 8, 3)) as s:          # next to other columns
        s.feed(files[0])
        rb = s.next_batch().to_pyarrow()
        rb_rows = rb.to_pylist()
    assert [r["seqname"] for r in rb_rows] == [f"sq{i}" for i in range(5)] and [[tuple(kv) for kv in r["attributes"]] for r in rb_rows] == want[0]
    with pytest.raises(ExonGpuError) as e:
        gpu_attributes(gpu_ctx, [b"sq0\tsrc\tgene\t1\t2\t.\t+\t.\tID=a;novalue\n"])
    assert e.value.code == _abi.ERR_PARSE
