"""FASTA COUNT(*) (BASELINE configs[0]): oracle against slt/fasta-scan-tests.slt:72-85, GPU against the oracle."""
import gzip
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


def fixture():
    with open(os.path.join(GOLDEN, "test.fasta"), "rb") as f:
        return f.read()


def test_oracle_goldens():
    t = fixture()
    assert oracle.fasta_count(t) == 2                      # SELECT COUNT(*) FROM fasta_scan('.../test.fasta') -> 2
    assert oracle.fasta_count(t) + oracle.fasta_count(t) == 4  # the two-file partition directory -> 4
    assert oracle.fasta_count(b"") == 0 and oracle.fasta_count(b">only a definition") == 1
    assert oracle.fasta_count(b">a\nAC>GT\n>b\n\n>c") == 3  # '>' inside a sequence line is data
    with pytest.raises(ValueError):
        oracle.fasta_count(b"ACGT\n>a\nAC\n")


def synth(rng, n, width=60):
    out = []
    for i in range(n):
        ln = int(rng.integers(0, 500))
        seq = bytes(rng.choice(np.frombuffer(b"ACGT>N", dtype=np.uint8), ln))
        lines = [seq[o:o + width] for o in range(0, ln, width)]
        lines = [(b"A" + l[1:] if l[:1] == b">" else l) for l in lines]  # a sequence line may hold '>' but not start with it
        out.append(b">seq%d desc > x\n" % i + b"".join(l + b"\n" for l in lines))
    return b"".join(out)


@pytest.mark.gpu
def test_gpu_counts(gpu_ctx):
    from bgzf_util import bgzf_compress
    from exon_b200._abi import ExonGpuError

    t = fixture()
    with gpu_ctx.open_fasta() as s:
        s.feed(t)
        assert s.rows() == 2
        s.feed(t)
        assert s.rows() == 4
        s.feed_gzip(gzip.compress(t))
        s.feed_gzip(bgzf_compress(t))
        assert s.rows() == 8
    rng = np.random.default_rng(5)
    texts = [synth(rng, 20_000), synth(rng, 3, 10), b"", b">x", synth(rng, 5000, 7)]
    want = sum(oracle.fasta_count(x) for x in texts)
    with gpu_ctx.open_fasta() as s:
        for x in texts:
            for o in range(0, max(len(x), 1), 1 << 20):
                s.feed(x[o:o + (1 << 20)], is_last=o + (1 << 20) >= len(x))
        assert s.rows() == want
    for shift in (0, 9):
        x = np.frombuffer(texts[0], dtype=np.uint8)
        d = gpu_ctx.device_buffer(x.size + shift + 64)
        d.upload(np.ascontiguousarray(x), offset=shift)
        with gpu_ctx.open_fasta() as s:
            s.feed(None, device_ptr=d.ptr + shift, nbytes=x.size)
            assert s.rows() == oracle.fasta_count(texts[0])
        d.free()
    with gpu_ctx.open_fasta() as s:
        s.feed(b"ACGT\n>a\nAC\n")
        with pytest.raises(ExonGpuError):
            s.rows()


# ---- record batches {id, description, sequence} -------------------------------------------------------------------------

def test_oracle_record_goldens():
    # slt/fasta-scan-tests.slt:6-10: `a description ATCG`, `b description2 ATCG`
    assert oracle.fasta_records(fixture()) == [(b"a", b"description", b"ATCG"), (b"b", b"description2", b"ATCG")]
    assert oracle.fasta_records(b">x\nAC\nGT\r\n\nA\n>y  two  words \t\nN") == [(b"x", None, b"ACGTA"), (b"y", b"two  words", b"N")]
    assert oracle.fasta_records(b">z \nA\n") == [(b"z", b"", b"A")]          # whitespace but nothing after it: Some("")
    for bad in (b"ACGT\n>a\nAC\n", b"> desc\nAC\n", b">a\n>b\nAC\n", b">a\nAC\n>b"):
        with pytest.raises(ValueError):
            oracle.fasta_records(bad)


def synth_records(rng, n, width=60, max_len=500):
    out = []
    for i in range(n):
        ln = int(rng.integers(1, max_len))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), ln))
        eol = b"\r\n" if i % 7 == 3 else b"\n"
        desc = [b"", b" d%d" % i, b"  two words  ", b"\tx"][i % 4]
        out.append(b">s%d" % i + desc + eol + b"".join(seq[o:o + width] + eol for o in range(0, ln, width)))
    return b"".join(out)


def gpu_records(ctx, files, projection=(0, 1, 2), batch_rows=8192, gz=False):
    names = ["id", "description", "sequence"]
    cols = {names[p]: [] for p in projection}
    sizes = []
    with ctx.open_fasta(projection=projection, batch_rows=batch_rows) as s:
        for f in files:
            (s.feed_gzip if gz else s.feed)(f)
        for b in s.batches():
            rb = b.to_pyarrow()
            assert rb.schema.names == [names[p] for p in projection]
            sizes.append(rb.num_rows)
            for p in projection:
                cols[names[p]] += [None if x is None else x.encode() for x in rb.column(names[p]).to_pylist()]
    return cols, sizes


@pytest.mark.gpu
def test_gpu_record_batches(gpu_ctx):
    from bgzf_util import bgzf_compress
    from exon_b200 import _abi
    from exon_b200._abi import ExonGpuError

    t = fixture()
    got, sizes = gpu_records(gpu_ctx, [t])
    assert sizes == [2] and got == {"id": [b"a", b"b"], "description": [b"description", b"description2"], "sequence": [b"ATCG", b"ATCG"]}
    got, sizes = gpu_records(gpu_ctx, [gzip.compress(t), bgzf_compress(t)], gz=True)      # the gzip twin of the slt (:24-28)
    assert sizes == [2, 2] and got["sequence"] == [b"ATCG"] * 4
    rng = np.random.default_rng(11)
    files = [synth_records(rng, 3000), synth_records(rng, 5, 7, 40), synth_records(rng, 1200, 80, 3000)]
    want = [r for f in files for r in oracle.fasta_records(f)]
    for projection, batch_rows in (((0, 1, 2), 8192), ((2, 0), 1000), ((1,), 7)):
        got, sizes = gpu_records(gpu_ctx, files, projection, batch_rows)
        assert sum(sizes) == len(want) and max(sizes) <= batch_rows
        for p in projection:
            assert got[["id", "description", "sequence"][p]] == [r[p] for r in want]
    assert any(r[1] is None for r in want) and any(r[1] == b"" for r in want) is False
    # batches restart at every file
    _, sizes = gpu_records(gpu_ctx, [t, t, t], (0,), 8192)
    assert sizes == [2, 2, 2]
    for bad in (b"ACGT\n>a\nAC\n", b"> desc\nAC\n", b">a\n>b\nAC\n", b">a\nAC\n>b\n"):
        with pytest.raises(ExonGpuError) as e:
            gpu_records(gpu_ctx, [bad])
        assert e.value.code == _abi.ERR_PARSE
    assert gpu_records(gpu_ctx, [b""]) == ({"id": [], "description": [], "sequence": []}, [])
