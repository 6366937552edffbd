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
