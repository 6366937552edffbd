"""Indexed VCF scan: tabix chunk query and chunk-restricted device inflate, against the reference's
known answers -- the chunk of indexed_bgzf_file.rs:167-187 and the counts of slt/vcf-indexed-tests.slt:22-59."""
import gzip
import os

import numpy as np
import pytest

import oracle
from bgzf_util import bgzf_compress
from conftest import GOLDEN
from exon_b200._abi import make_region

pytestmark = pytest.mark.gpu


def raw(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def test_reference_chunk_golden(gpu_ctx):
    tbi = raw("bigger_index_test.vcf.gz.tbi")
    assert gpu_ctx.tabix_query(tbi, make_region("chr1", 1, 3388930)) == [(621346816, 3014113427456)]
    assert gpu_ctx.tabix_query(tbi, make_region("chr2", 1, 10)) == []            # contig absent: no chunks
    assert gpu_ctx.tabix_query(tbi, make_region("chr1")) == [(621346816, 3014113427456)]
    assert len(gpu_ctx.tabix_query(tbi, make_region("chr1", 1_000_000, 1_000_100))) == 1


def indexed_count(ctx, gz, tbi, region):
    chunks = ctx.tabix_query(tbi, region)
    with ctx.open_vcf() as s:
        for ch in chunks:
            lo = ch[0] >> 16                       # a ranged GET from the chunk's first member to the end of the file
            s.feed_bgzf_chunk(gz[lo:], ch, file_offset=lo)
        return s.filter_count(region), s.rows(), chunks


def test_indexed_goldens(gpu_ctx):
    gz, tbi = raw("index.vcf.gz"), raw("index.vcf.gz.tbi")
    text = gzip.decompress(gz)
    assert indexed_count(gpu_ctx, gz, tbi, make_region("1"))[0] == 191                     # slt/vcf-indexed-tests.slt:27-35 (per copy)
    assert indexed_count(gpu_ctx, gz, tbi, make_region("a"))[:2] == (0, 0)                  # :22-25
    for rg in [("1", 9999919, 10000000), ("2", None, None), ("10", 1, 5_000_000), ("1", 1, 100)]:
        cnt, rows, chunks = indexed_count(gpu_ctx, gz, tbi, make_region(*rg))
        assert cnt == oracle.filter_count(text, *rg)[0], rg
        assert rows <= 621
    # two copies in one partition (the vcf-partition table: 382)
    chunks = gpu_ctx.tabix_query(tbi, make_region("1"))
    with gpu_ctx.open_vcf() as s:
        for _ in range(2):
            for ch in chunks:
                s.feed_bgzf_chunk(gz, ch)
        assert s.filter_count(make_region("1")) == 382
    bgz, btbi = raw("biobear_vcf_file.vcf.gz"), raw("biobear_vcf_file.vcf.gz.tbi")
    assert indexed_count(gpu_ctx, bgz, btbi, make_region("1"))[0] == 11                    # :56-59
    assert indexed_count(gpu_ctx, bgz, btbi, make_region("1000"))[0] == 0                  # :51-54
    assert indexed_count(gpu_ctx, bgz, btbi, make_region("4"))[0] == 2


def test_chunks_skip_most_members(gpu_ctx):
    """A synthetic multi-member file with a hand-made chunk: only the covered members are inflated and scanned."""
    from synth import vcf

    cols = vcf.columns(200_000, seed=21)
    text = vcf.shards(cols, 1)[0].tobytes()
    gz = bgzf_compress(text, 6, block=30_000)
    # member table by walking BSIZE; a chunk from the start of member 5 (+ a record boundary) to the middle of member 9
    offs, p = [], 0
    while p < len(gz):
        offs.append(p)
        p += (gz[p + 16] | (gz[p + 17] << 8)) + 1
    ublocks = [text[i:i + 30_000] for i in range(0, len(text), 30_000)]
    u0 = ublocks[5].index(b"\n") + 1
    u1 = ublocks[9].rindex(b"\n") + 1
    chunk = ((offs[5] << 16) | u0, (offs[9] << 16) | u1)
    want_text = (ublocks[5][u0:] + b"".join(ublocks[6:9]) + ublocks[9][:u1])
    with gpu_ctx.open_vcf() as s:
        s.feed_bgzf_chunk(gz[offs[3]:], chunk, file_offset=offs[3])
        assert s.body_bytes() == len(want_text)
        assert s.rows() == want_text.count(b"\n")
        for q in [("1", None, None), (None, 1, 50_000_000), ("1", 1_000_000, 20_000_000)]:
            assert s.filter_count(make_region(*q)) == oracle.filter_count(want_text, *q)[0]
        pos = np.concatenate([b.column("pos") for b in s.batches()])
        assert np.array_equal(pos, np.concatenate([b["pos"] for b in oracle.read_batches(want_text)]))
