"""GPU parity of the fused FASTQ scan -> mean-quality filter -> COUNT (BASELINE configs[1]) through the C ABI:
against the oracle, the reference fixture, the generator's integer truth, and on edge cases."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from exon_b200 import _abi
from exon_b200._abi import ExonGpuError

pytestmark = pytest.mark.gpu

THRESHOLDS = [None, 30, 20, 35, (61, 2), 0, 41]


def gpu_count(ctx, feeds, min_mean=None):
    with ctx.open_fastq() as s:
        for f in feeds:
            s.feed(f, is_last=True)
        return s.filter_count(min_mean), s.rows()


def test_reference_fixture(gpu_ctx):
    with open(os.path.join(GOLDEN, "test.fastq"), "rb") as f:
        text = f.read()
    assert gpu_count(gpu_ctx, [text]) == (2, 2)              # slt/fastq-scan-test.slt:51-54
    assert gpu_count(gpu_ctx, [text, text]) == (4, 4)        # the partition directory: :56-59
    for t in (0, 10, 17, 18, 30):
        assert gpu_count(gpu_ctx, [text], t) == oracle.fastq_filter_count(text, t)


@pytest.fixture(scope="module")
def synth_fq():
    from synth import fastq

    return fastq.shards(200_000, 6)


def test_synthetic_counts(gpu_ctx, synth_fq):
    sh = synth_fq
    with gpu_ctx.open_fastq() as s:
        for f in sh.files:
            s.feed(f, is_last=True)
        assert s.rows() == sh.n
        assert s.body_bytes() == sum(f.size for f in sh.files)
        for t in THRESHOLDS:
            want = oracle.fastq_filter_count_files(sh.files, t, target_partitions=3)[0]
            assert s.filter_count(t) == want, t
            if t is not None:
                num, den = t if isinstance(t, tuple) else (t, 1)
                assert want == sh.truth_count(num, den)


@pytest.mark.parametrize("chunk", [7, 4096, 100_000, 1 << 20])
def test_ragged_feeds_and_device_ranges(gpu_ctx, synth_fq, chunk):
    sh = synth_fq
    files = [f[: 316 * 300] for f in sh.files[:2]] if chunk < 4096 else sh.files[:3]
    want = oracle.fastq_filter_count_files(files, 30)
    with gpu_ctx.open_fastq() as s:
        for f in files:
            for o in range(0, f.size, chunk):
                s.feed(f[o:o + chunk], is_last=o + chunk >= f.size)
        assert (s.filter_count(30), s.rows()) == want
    if chunk == 4096:
        for shift in (0, 3, 15):
            bufs = []
            with gpu_ctx.open_fastq() as s:
                for f in files:
                    d = gpu_ctx.device_buffer(f.size + shift)
                    d.upload(np.ascontiguousarray(f), offset=shift)
                    bufs.append(d)
                    s.feed(None, device_ptr=d.ptr + shift, nbytes=f.size, is_last=True)
                assert (s.filter_count(30), s.rows()) == want
            for d in bufs:
                d.free()


def rand_fastq(rng, n, max_len, tiny=False):
    out = []
    for i in range(n):
        ln = int(rng.integers(0 if tiny else 1, max_len + 1))
        q = bytes(rng.integers(33, 75, ln).astype(np.uint8))
        desc = b" d@+ x" if i % 3 == 0 else b""
        out.append(b"@n%d" % i + desc + b"\n" + b"A" * ln + b"\n+" + (b"n%d" % i if i % 2 else b"") + b"\n" + q + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("max_len,n", [(3, 5000), (40, 3000), (600, 800), (20_000, 40)])
def test_variable_length_reads(gpu_ctx, max_len, n):
    """Tiny records (several quality lines per 16-byte chunk), reads longer than the staged halo, reads longer
    than a tile; '@' and '+' as first quality characters."""
    rng = np.random.default_rng(max_len)
    text = rand_fastq(rng, n, max_len, tiny=max_len <= 3)
    halves = [text[: len(text) // 2], text[len(text) // 2:]]  # one file fed in two ranges
    for t in (None, 10, 20, 30):
        want = oracle.fastq_filter_count(text, t)
        assert gpu_count(gpu_ctx, [text], t) == want, t
        with gpu_ctx.open_fastq() as s:
            s.feed(halves[0], is_last=False)
            s.feed(halves[1], is_last=True)
            assert (s.filter_count(t), s.rows()) == want


def test_edge_cases(gpu_ctx):
    assert gpu_count(gpu_ctx, [b""]) == (0, 0)
    assert gpu_count(gpu_ctx, [b"@a\nAC\n+\nII"], 30) == (1, 1) == oracle.fastq_filter_count(b"@a\nAC\n+\nII", 30)
    assert gpu_count(gpu_ctx, [b"@a\nAC\n+\n"], 0) == (0, 1) == oracle.fastq_filter_count(b"@a\nAC\n+\n", 0)
    t = b"@a\nAC\n+\n@@\n@b x y\n\n+b\n+I\n"
    assert gpu_count(gpu_ctx, [t], 30) == oracle.fastq_filter_count(t, 30) == (1, 2)
    assert gpu_count(gpu_ctx, [t, b"", t], 30) == (2, 4)
    for bad in [b"SEQ\nACGT\n+\n!!!!\n", b"@a\nACGT\n-\n!!!!\n", b"@a\nACGT\n", b"@a\n", b"@a\nAC\n+\n!!\n\n"]:
        with pytest.raises(ExonGpuError) as e:
            gpu_count(gpu_ctx, [bad], 30)
        assert e.value.code == _abi.ERR_PARSE
        with pytest.raises(ValueError):
            oracle.fastq_filter_count(bad, 30)
    # a truncated file is an error even when another complete file follows in the same partition
    with pytest.raises(ExonGpuError):
        gpu_count(gpu_ctx, [b"@a\nACGT\n", b"@a\nAC\n+\nII\n"], 30)


def test_properties_large(gpu_ctx):
    """2M reads: counts equal the generator's truth, are monotone in the threshold, and COUNT(*) = reads."""
    from synth import fastq

    sh = fastq.shards(2_000_000, 16, seed=5)
    with gpu_ctx.open_fastq() as s:
        for f in sh.files:
            s.feed(f, is_last=True)
        assert s.filter_count(None) == sh.n
        prev = sh.n
        for t in (5, 20, 25, 30, 35, 40):
            c = s.filter_count(t)
            assert c == sh.truth_count(t) and c <= prev
            prev = c


# ---- column batches (FASTQArrayBuilder) --------------------------------------------------------------------------

def gpu_records(ctx, feeds, batch_rows=8192, projection=(0, 1, 2, 3), gz=False):
    names = ["name", "description", "sequence", "quality_scores"]
    with ctx.open_fastq(batch_rows=batch_rows, projection=projection) as s:
        for f in feeds:
            (s.feed_gzip if gz else s.feed)(f)
        rows, sizes = [], []
        for b in s.batches():
            assert b.names == [names[k] for k in projection] and b.formats == ["u"] * len(projection)
            cols = [b.strings(n) for n in b.names]
            rows += list(zip(*cols))
            sizes.append(b.num_rows)
            b.release()
        return rows, sizes


def oracle_records(feeds, batch_rows=8192, projection=(0, 1, 2, 3)):
    keys = ["name", "description", "sequence", "quality"]
    rows, sizes = [], []
    for f in feeds:
        for b in oracle.fastq_read_batches(f, batch_size=batch_rows):
            rows += list(zip(*[b[keys[k]] for k in projection]))
            sizes.append(b["rows"])
    return rows, sizes


def test_columns_reference_fixture(gpu_ctx):
    with open(os.path.join(GOLDEN, "test.fastq"), "rb") as f:
        text = f.read()
    rows, sizes = gpu_records(gpu_ctx, [text])
    qual = b"!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65"
    seq = b"GATTTGGGGTExonAAGCAGTATCGAExonAATAGTAAATCCATTTGTExonACExonCAGTTT"
    # slt/fastq-scan-test.slt:6-10: name, description (NULL for the second record), quality_scores, sequence
    assert rows == [(b"SEQ_ID", b"This is a description", seq, qual), (b"SEQ_ID2", None, seq, qual)] and sizes == [2]
    assert gpu_records(gpu_ctx, [text, text], batch_rows=1)[1] == [1, 1, 1, 1]


def test_columns_match_oracle(gpu_ctx, synth_fq):
    files = [bytes(f[: 316 * 5000]) for f in synth_fq.files[:3]]
    for batch_rows, proj in [(8192, (0, 1, 2, 3)), (1000, (3, 0)), (777, (2,)), (64, (1,))]:
        assert gpu_records(gpu_ctx, files, batch_rows, proj) == oracle_records(files, batch_rows, proj), (batch_rows, proj)
    rng = np.random.default_rng(3)
    texts = [rand_fastq(rng, 3000, 40), rand_fastq(rng, 500, 3, tiny=True), b"", rand_fastq(rng, 40, 20_000), b"@a\nAC\n+\nII", b"@b desc only\nAC\n+\n"]
    assert gpu_records(gpu_ctx, texts, 100) == oracle_records(texts, 100)
    from bgzf_util import bgzf_compress

    assert gpu_records(gpu_ctx, [bgzf_compress(t) for t in texts[:2]], 100, gz=True) == oracle_records(texts[:2], 100)
    # ragged feeds of one file
    with gpu_ctx.open_fastq(projection=(0, 3), batch_rows=500) as s:
        t = texts[0]
        for o in range(0, len(t), 10_000):
            s.feed(t[o:o + 10_000], is_last=o + 10_000 >= len(t))
        got = [r for b in s.batches() for r in zip(b.strings("name"), b.strings("quality_scores"))]
    assert got == oracle_records([t], 500, (0, 3))[0]
    for bad in [b"SEQ\nACGT\n+\n!!!!\n", b"@a\nACGT\n-\n!!!!\n", b"@a\nACGT\n"]:
        with gpu_ctx.open_fastq(projection=(0,)) as s:
            s.feed(bad)
            with pytest.raises(ExonGpuError):
                s.next_batch()
