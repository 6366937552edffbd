"""GPU parity tests of the VCF hot path, all through the C ABI (exon_b200.runtime is a 1:1 ctypes layer).

K1 fused scan->filter->COUNT and K2 column build are compared with the CPU oracle on the reference's own
fixtures (tests/golden/), on seeded synthetic shards, and on hand-made edge cases; bit-exact everywhere.
"""
import numpy as np
import pytest

import oracle
from conftest import make_vcf
from exon_b200 import _abi
from exon_b200._abi import ExonGpuError, make_region

pytestmark = pytest.mark.gpu

QUERIES = [("1", 1_000_000, 2_000_000), ("1", None, None), ("X", None, None), ("22", 5, 30_000_000),
           ("10", 1, None), (None, 1, 1_000_000), (None, None, None), ("a", None, None), ("1", 2_000_000, 1_000_000)]


def gpu_count(ctx, files, chrom=None, lo=None, hi=None, **kw):
    with ctx.open_vcf(**kw) as s:
        for f in files:
            s.feed(f, is_last=True)
        return s.filter_count(make_region(chrom, lo, hi))


# ---- reference fixtures --------------------------------------------------------------------

@pytest.mark.parametrize("strict", [False, True])
def test_reference_goldens(gpu_ctx, index_vcf, index_vcf_gz_twin, biobear_vcf, common_all_vcf, goldens, strict):
    g = goldens["index.vcf"]
    with gpu_ctx.open_vcf(strict=strict) as s:
        s.feed(index_vcf)
        assert s.filter_count(None) == g["reference_pinned"]["count_star"]["value"] == 621
        assert s.rows() == 621
        assert s.filter_count(make_region("1")) == 191
        assert s.filter_count(make_region("a")) == 0
        assert s.filter_count(make_region("2")) == g["derived"]["chrom_2"]
        assert s.filter_count(make_region("10")) == g["derived"]["chrom_10"]
        assert s.filter_count(make_region("1", 9999919, 10000000)) == 82
        assert s.filter_count(make_region("1", 1000000, 2000000)) == 0
        assert s.filter_count(make_region(None, 10000000, None)) == g["derived"]["pos_ge_10000000"]
    # two copies of the data in one partition (vcf-partition: 382) and the gzip twin
    assert gpu_count(gpu_ctx, [index_vcf, index_vcf], "1", strict=strict) == 382
    assert gpu_count(gpu_ctx, [index_vcf_gz_twin], strict=strict) == 621
    b = goldens["biobear_vcf_file.vcf"]
    assert gpu_count(gpu_ctx, [biobear_vcf], "1", strict=strict) == b["reference_pinned"]["chrom_1"]["value"] == 11
    assert gpu_count(gpu_ctx, [biobear_vcf], "1000", strict=strict) == 0
    assert gpu_count(gpu_ctx, [biobear_vcf], strict=strict) == 15
    assert gpu_count(gpu_ctx, [common_all_vcf], strict=strict) == goldens["common_all_head.vcf"]["derived"]["count_star"]


def test_reference_golden_columns(gpu_ctx, index_vcf, biobear_vcf, common_all_vcf, goldens):
    for text, key in [(index_vcf, None), (biobear_vcf, "biobear_vcf_file.vcf"), (common_all_vcf, "common_all_head.vcf")]:
        for batch_rows in (8192, 100, 7):
            want = list(oracle.read_batches(text, batch_size=batch_rows))
            with gpu_ctx.open_vcf(batch_rows=batch_rows) as s:
                s.feed(text)
                got = list(s.batches())
                assert [b.num_rows for b in got] == [w["rows"] for w in want]
                for b, w in zip(got, want):
                    off, val = b.column("chrom")
                    assert np.array_equal(off, w["chrom_offsets"]) and np.array_equal(val, w["chrom_values"])
                    assert np.array_equal(b.column("pos"), w["pos"])
                    b.release()
        if key:
            with gpu_ctx.open_vcf() as s:
                s.feed(text)
                (b,) = list(s.batches())
                rows = [[c, int(p)] for c, p in zip(b.chrom_strings(), b.column("pos"))]
                assert rows == goldens[key]["derived"]["rows"]


# ---- synthetic shards vs oracle and vs the generator's integer truth -----------------------------------------

@pytest.fixture(scope="module")
def synth_small():
    from synth import vcf

    cols = vcf.columns(300_000)
    return cols, vcf.shards(cols, 8)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("strict", [False, True])
def test_synthetic_counts(gpu_ctx, synth_small, variant, strict):
    cols, files = synth_small
    with gpu_ctx.open_vcf(kernel_variant=variant, strict=strict) as s:
        for f in files:
            s.feed(f, is_last=True)
        for chrom, lo, hi in QUERIES:
            got = s.filter_count(make_region(chrom, lo, hi))
            want, rows, _ = oracle.filter_count_files(files, chrom, lo, hi, target_partitions=4)
            assert got == want == cols.truth_count(chrom, lo, hi), (chrom, lo, hi)
        assert s.rows() == cols.n
        assert s.body_bytes() == sum(f.size - oracle.header_len(f) for f in files)


@pytest.mark.parametrize("chunk", [1, 13, 4096, 65536, 1 << 20])
def test_ragged_feeds(gpu_ctx, synth_small, chunk):
    """Feeds cut at arbitrary byte positions (mid-line, mid-header) give the same answer as whole files."""
    cols, files = synth_small
    files = files[:2] if chunk < 4096 else files
    if chunk < 4096:
        files = [f[: 40_000] if f[39_999] == 10 else f[: 40_000 - int(np.argmax(f[39_999::-1] == 10))] for f in files]
    want = {q: oracle.filter_count_files(files, *q, target_partitions=2)[0] for q in QUERIES}
    for pushdown in (None, make_region("1", 1_000_000, 2_000_000)):
        with gpu_ctx.open_vcf(pushdown=pushdown) as s:
            for f in files:
                for o in range(0, f.size, chunk):
                    s.feed(f[o:o + chunk], is_last=o + chunk >= f.size)
            for q in QUERIES:
                assert s.filter_count(make_region(*q)) == want[q], (chunk, q)


def test_device_resident_feed(gpu_ctx, synth_small):
    """Zero-copy device ranges (how bench.py's HBM-resident workload is fed), at unaligned addresses."""
    cols, files = synth_small
    for shift in (0, 1, 7, 15):
        bufs = []
        with gpu_ctx.open_vcf() as s:
            for f in files[:4]:
                d = gpu_ctx.device_buffer(f.size + shift)
                d.upload(np.ascontiguousarray(f), offset=shift)
                bufs.append(d)
                s.feed(None, device_ptr=d.ptr + shift, nbytes=f.size, is_last=True)
            for q in QUERIES:
                assert s.filter_count(make_region(*q)) == oracle.filter_count_files(files[:4], *q)[0], (shift, q)
            pos = np.concatenate([b.column("pos") for b in s.batches()])
            want = np.concatenate([b["pos"] for f in files[:4] for b in oracle.read_batches(f)])
            assert np.array_equal(pos, want)
        for d in bufs:
            d.free()


def test_synthetic_columns(gpu_ctx, synth_small):
    cols, files = synth_small
    for batch_rows in (8192, 1000):
        with gpu_ctx.open_vcf(batch_rows=batch_rows) as s:
            for f in files:
                s.feed(f, is_last=True)
            got = list(s.batches())
        want = [b for f in files for b in oracle.read_batches(f, batch_size=batch_rows)]
        assert [b.num_rows for b in got] == [w["rows"] for w in want]
        for b, w in zip(got, want):
            off, val = b.column("chrom")
            assert np.array_equal(off, w["chrom_offsets"]) and np.array_equal(val, w["chrom_values"])
            assert np.array_equal(b.column("pos"), w["pos"])
    # projection order and single-column projections
    with gpu_ctx.open_vcf(projection=(1, 0)) as s:
        s.feed(files[0])
        b = s.next_batch()
        assert b.names == ["pos", "chrom"] and b.formats == ["l", "u"]
    with gpu_ctx.open_vcf(projection=(1,)) as s:
        s.feed(files[0])
        b = s.next_batch()
        assert b.names == ["pos"]
    with gpu_ctx.open_vcf(projection=()) as s:  # COUNT(*): zero-column batches that carry only a row count
        s.feed(files[0])
        assert sum(b.num_rows for b in s.batches()) == next(iter(oracle.filter_count_files([files[0]])[1:2]))


# ---- edge cases -----------------------------------------------------------------------------------------

LONG = "chrUn_KI270742v1_decoy"


@pytest.mark.parametrize("strict", [False, True])
def test_edge_cases(gpu_ctx, strict):
    def both(text, chrom=None, lo=None, hi=None):
        got = gpu_count(gpu_ctx, [text], chrom, lo, hi, strict=strict)
        assert got == oracle.filter_count(text, chrom, lo, hi)[0]
        return got

    assert both(b"", "1") == 0
    assert both(make_vcf([]), "1") == 0                                  # header only
    assert both(make_vcf([("1", "7")], trailing_newline=False), "1", 1, 10) == 1
    assert both(make_vcf([("1", "7")], header=False), "1", 7, 7) == 1    # no header at all
    assert both(make_vcf([("1", "+7")]), "1", 1, 10) == 1                # Rust usize::from_str takes '+'
    assert both(make_vcf([("1", "007")]), "1", 7, 7) == 1
    assert both(make_vcf([("1", str(2**63 - 1))]), "1", 1, None) == 1
    # names that are prefixes / suffixes of each other, 1..24-byte names
    rows = [("1", "5"), ("11", "5"), ("10", "5"), ("21", "5"), ("1", "50"), (LONG, "5"), (LONG + "x", "5"),
            ("x" + LONG, "5"), ("chr1", "5"), ("chr11", "5"), ("hr1", "5"), ("r1", "5")] * 50
    text = make_vcf(rows)
    for c in ["1", "11", "10", "21", LONG, LONG + "x", "chr1", "hr1", "r1", "chr", "2", LONG[:-1]]:
        both(text, c)
        both(text, c, 5, 5)
        both(text, c, 6, 60)
    both(text, None, 6, 60)
    both(text)
    # the pattern "\n1\t" also occurs inside other fields only if a field could hold '\n' -- it cannot; but
    # "1\t" at the START of a later field must not count
    both(make_vcf([("2", "5")], extra_cols="\t1\t1\t1\t1\tPASS\t1"), "1")
    # a literal that no CHROM can equal
    for lit in ["", "1\t2", "1\n", "x" * 300]:
        assert gpu_count(gpu_ctx, [text], lit, strict=strict) == 0


def test_long_lines(gpu_ctx):
    """Rows far longer than a tile (the reference's bigger-index fixture has 3202 samples per row)."""
    rng = np.random.default_rng(7)
    samples = "\tGT" + "".join("\t0|1" for _ in range(3000))
    rows = [(("chr1", "1", "chr22")[int(rng.integers(0, 3))], str(int(rng.integers(1, 3_000_000)))) for _ in range(600)]
    text = make_vcf(rows, extra_cols="\t.\tA\tC\t50\tPASS\t." + samples)
    assert len(text) > 7_000_000
    for strict in (False, True):
        for variant in (0, 5):
            with gpu_ctx.open_vcf(strict=strict, kernel_variant=variant) as s:
                s.feed(text)
                for q in [("chr1", 1_000_000, 2_000_000), ("1", None, None), ("chr22", 1, 1_500_000), (None, None, None),
                          (None, 1_000_000, 2_000_000)]:
                    assert s.filter_count(make_region(*q)) == oracle.filter_count(text, *q)[0], q
    with gpu_ctx.open_vcf() as s:
        s.feed(text)
        pos = np.concatenate([b.column("pos") for b in s.batches()])
    assert np.array_equal(pos, np.concatenate([b["pos"] for b in oracle.read_batches(text)]))


@pytest.mark.parametrize("strict", [False, True])
def test_malformed_records_raise(gpu_ctx, strict):
    """Error behaviour of the reference: a record whose POS cannot be parsed fails the query (ArrowError)."""
    bad = [make_vcf([("1", "0")]), make_vcf([("1", "12x")]), make_vcf([("1", "")]), make_vcf([("1", "-5")]),
           make_vcf([("1", str(2**63))]), make_vcf([("1", "99999999999999999999999")]), b"1\t5\n"]
    for text in bad:
        with pytest.raises(ExonGpuError) as e:
            gpu_count(gpu_ctx, [text], "1", 1, 10, strict=strict)
        assert e.value.code == _abi.ERR_PARSE
        with pytest.raises(ValueError):
            oracle.filter_count(text, "1", 1, 10)
        with gpu_ctx.open_vcf() as s:
            s.feed(text)
            with pytest.raises(ExonGpuError):
                s.next_batch()


def test_strict_validates_unselected_rows(gpu_ctx):
    """strict = 1 validates CHROM/POS of every row, selected or not, like LazyVCFArrayBuilder::append does
    (lazy_array_builder.rs:159-168); the default fast path validates the rows whose CHROM matches."""
    for text in [make_vcf([("2", "12x"), ("1", "5")]), make_vcf([("2", "0"), ("1", "5")]), b"1\t5\t.\n2\n", b"2\n"]:
        with pytest.raises(ExonGpuError):
            gpu_count(gpu_ctx, [text], "1", 1, 10, strict=True)
        with pytest.raises(ValueError):
            oracle.filter_count(text, "1", 1, 10)
        with gpu_ctx.open_vcf() as s:
            s.feed(text)
            with pytest.raises(ExonGpuError):
                s.next_batch()


def test_state_errors(gpu_ctx):
    with gpu_ctx.open_vcf() as s:
        s.feed(make_vcf([("1", "5")]))
        list(s.batches())
        with pytest.raises(ExonGpuError) as e:
            s.feed(make_vcf([("1", "5")]))
        assert e.value.code == _abi.ERR_STATE
        s.reset()
        s.feed(make_vcf([("1", "5"), ("1", "6")]))
        assert s.filter_count(make_region("1")) == 2
    with pytest.raises(ExonGpuError) as e:
        gpu_ctx.open_vcf(projection=(9,))     # 0..8 are the VCF file-schema columns
    assert e.value.code == _abi.ERR_ARG


# ---- size-independent properties at a larger size ------------------------------------------------------------

def test_properties_large(gpu_ctx):
    """10M rows: counts equal the generator's integer truth; per-contig counts sum to COUNT(*); interval
    counts are additive over a partition of the position axis; strict == fast path."""
    from synth import vcf

    cols = vcf.columns(10_000_000, seed=99)
    files = vcf.shards(cols, 16)
    with gpu_ctx.open_vcf() as s, gpu_ctx.open_vcf(strict=True) as t:
        for f in files:
            s.feed(f, is_last=True)
            t.feed(f, is_last=True)
        total = s.filter_count(None)
        assert total == cols.n == t.filter_count(None)
        per = [s.filter_count(make_region(c)) for c, _ in vcf.CONTIGS]
        assert sum(per) == total and per == [cols.truth_count(c, None, None) for c, _ in vcf.CONTIGS]
        cuts = [1, 1_000_000, 2_000_001, 50_000_000, 150_000_000, 250_000_000]
        parts = [s.filter_count(make_region("1", a, b - 1)) for a, b in zip(cuts[:-1], cuts[1:])]
        assert sum(parts) == per[0]
        assert parts[1] == cols.truth_count("1", 1_000_000, 2_000_000) == t.filter_count(make_region("1", 1_000_000, 2_000_000))
        assert s.filter_count(make_region(None, 1_000_000, 2_000_000)) == cols.truth_count(None, 1_000_000, 2_000_000)
        pos = np.concatenate([b.column("pos") for b in s.batches()])
        assert np.array_equal(pos, cols.pos)


# ---- K2 specifics: ranks inside a tile, files inside one run, stream-level K3 -------------------------------------

def py_rows(text: bytes):
    """(chrom bytes, pos) of every record by plain Python splitting (for inputs the oracle rejects: < 8 fields)."""
    rows = []
    for ln in text.split(b"\n"):
        if not ln or ln.startswith(b"#"):
            continue
        f = ln.split(b"\t")
        rows.append((f[0], int(f[1])))
    return rows


def gpu_rows(ctx, feeds, batch_rows=8192, **kw):
    with ctx.open_vcf(batch_rows=batch_rows, **kw) as s:
        for f in feeds:
            s.feed(f, is_last=True)
        out, sizes = [], []
        for b in s.batches():
            sizes.append(b.num_rows)
            out += list(zip([c.encode() for c in b.chrom_strings()], [int(p) for p in b.column("pos")]))
            b.release()
        return out, sizes


def test_columns_short_lines_rank_exactly(gpu_ctx):
    """Lines shorter than 16 bytes put several line starts into one 16-byte chunk: rows must keep text order."""
    rng = np.random.default_rng(11)
    names = [b"1", b"22", b"X", b"chr1"]
    lines = [names[int(rng.integers(0, 4))] + b"\t" + str(int(rng.integers(1, 10 ** int(rng.integers(1, 8))))).encode() + b"\t."
             for _ in range(50_000)]
    text = b"#h\n" + b"\n".join(lines) + b"\n"
    got, sizes = gpu_rows(gpu_ctx, [text], batch_rows=1000)
    assert got == py_rows(text) and sizes == [1000] * 50
    # > 256 line starts inside one 512-byte warp row: 2-byte names, 1-digit positions ("1\t5\t\n" = 5 bytes)
    text2 = b"".join(b"%d\t%d\t\n" % (i % 10, 1 + i % 9) for i in range(20_000))
    got2, _ = gpu_rows(gpu_ctx, [text2])
    assert got2 == py_rows(text2)
    with gpu_ctx.open_vcf() as s:
        s.feed(text2)
        assert s.filter_count(make_region("3", 4, 4)) == sum(1 for c, p in got2 if c == b"3" and p == 4)


def test_columns_files_share_a_run(gpu_ctx, synth_small):
    """Host feeds append consecutive files to the same arena run; batches must still restart at every file,
    including empty and header-only files in between (FileStream opens one batch stream per file)."""
    cols, files = synth_small
    feeds = [files[0][:30_011], b"", make_vcf([]), make_vcf([("7", "77")], trailing_newline=False), files[1][:50_000 - 13],
             make_vcf([("chrUn_KI270742v1_decoy_long_name", "5")] * 3)]
    feeds = [bytes(f) for f in feeds]
    feeds = [f if (not f or f.endswith(b"\n") or f.count(b"\n") < 3) else f[: f.rfind(b"\n") + 1] for f in feeds]
    want_rows, want_sizes = [], []
    for f in feeds:
        for b in oracle.read_batches(f, batch_size=100):
            want_sizes.append(b["rows"])
            off, val = b["chrom_offsets"], b["chrom_values"].tobytes()
            want_rows += [(val[off[i]:off[i + 1]], int(b["pos"][i])) for i in range(b["rows"])]
    got, sizes = gpu_rows(gpu_ctx, feeds, batch_rows=100)
    assert sizes == want_sizes and got == want_rows


@pytest.mark.parametrize("on_device", [False, True])
def test_stream_filter_agg_matches_fused_scan(gpu_ctx, synth_small, on_device):
    """exon_gpu_vcf_filter_agg (K2 columns kept in HBM -> multi-batch K3, one launch) == K1 == oracle."""
    cols, files = synth_small
    with gpu_ctx.open_vcf(columns_on_device=on_device, batch_rows=1000) as s:
        for f in files:
            s.feed(f, is_last=True)
        for chrom, lo, hi in QUERIES:
            want = oracle.filter_count_files(files, chrom, lo, hi, target_partitions=4)[0]
            rg = make_region(chrom, lo, hi)
            cnt, _, _ = s.filter_agg(chrom_col=0, pos_col=1, region=rg)
            assert cnt == want == s.filter_count(rg), (chrom, lo, hi)
        # SUM / AVG state over pos for one contig
        sel = (cols.contig == 0) & (cols.pos >= 1_000_000) & (cols.pos <= 2_000_000)
        cnt, si, sf = s.filter_agg(chrom_col=0, pos_col=1, region=make_region("1", 1_000_000, 2_000_000), kind=_abi.AGG_SUM, value_col=1)
        assert cnt == int(sel.sum()) and si == int(cols.pos[sel].sum()) and sf == float(si)
        cnt, si, _ = s.filter_agg(kind=_abi.AGG_AVG, value_col=1)
        assert cnt == cols.n and si == int(cols.pos.sum())
        with pytest.raises(ExonGpuError):
            s.filter_agg(kind=_abi.AGG_SUM, value_col=0)
    with gpu_ctx.open_vcf(projection=(1,)) as s:  # the predicate may only read projected columns
        s.feed(files[0])
        assert s.filter_agg(pos_col=0, region=make_region(None, 1, 10**9))[0] == s.rows()
        with pytest.raises(ExonGpuError):
            s.filter_agg(chrom_col=0, region=make_region("1"))


def test_filter_agg_batches_one_launch(gpu_ctx, synth_small):
    """exon_gpu_filter_agg_batches over the device-resident batches of a stream: one kernel for all of them."""
    cols, files = synth_small
    with gpu_ctx.open_vcf(columns_on_device=True, batch_rows=4096) as s:
        for f in files:
            s.feed(f, is_last=True)
        batches = list(s.batches())
        l0 = gpu_ctx.launch_count()
        for chrom, lo, hi in QUERIES[:6]:
            cnt, _, _ = gpu_ctx.filter_agg_batches(batches, chrom_col=0, pos_col=1, region=make_region(chrom, lo, hi))
            assert cnt == cols.truth_count(chrom, lo, hi)
        assert gpu_ctx.launch_count() - l0 == 6
        cnt, si, _ = gpu_ctx.filter_agg_batches(batches, kind=_abi.AGG_SUM, value_col=1)
        assert cnt == cols.n and si == int(cols.pos.sum())
        for b in batches:
            b.release()
