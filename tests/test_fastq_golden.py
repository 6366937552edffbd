"""The FASTQ oracle (oracle/fastq_oracle.c) against the reference's own known answers (later rows):
slt/fastq-scan-test.slt (2 records, the four column values, NULL description; bgzip twin -> 2) and
slt/quality-score-udfs.slt ('###' -> [2, 2, 2]); plus an independent pure-Python reader on synthetic input."""
import gzip
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN

QUAL = b"!''*((((***+))%%%++)(%%%%).1***-+*''))**55CCF>>>>>>CCCCCCC65"
SEQ = b"GATTTGGGGTExonAAGCAGTATCGAExonAATAGTAAATCCATTTGTExonACExonCAGTTT"


def fixture(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def py_records(text: bytes):
    """Independent restatement: 4 lines per record, name up to the first space."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    out = []
    for i in range(0, len(lines), 4):
        d = lines[i]
        assert d[:1] == b"@" and lines[i + 2][:1] == b"+"
        name, _, desc = d[1:].partition(b" ")
        out.append((name, desc or None, lines[i + 1], lines[i + 3] if i + 3 < len(lines) else b""))
    return out


def test_reference_fixture_rows():
    # slt/fastq-scan-test.slt:6-10
    text = fixture("test.fastq")
    (b,) = list(oracle.fastq_read_batches(text))
    assert b["rows"] == 2
    assert b["name"] == [b"SEQ_ID", b"SEQ_ID2"]
    assert b["description"] == [b"This is a description", None]
    assert b["quality"] == [QUAL, QUAL] and b["sequence"] == [SEQ, SEQ]
    assert oracle.fastq_filter_count(text) == (2, 2)                      # fastq-scan-test.slt:51-54
    assert oracle.fastq_filter_count(text + text)[0] == 4                 # the fastq-partition directory: :56-59
    assert oracle.fastq_filter_count(gzip.decompress(fixture("test_bgzip.fastq.gz"))) == (2, 2)  # :66-69
    mean = sum(c - 33 for c in QUAL) / len(QUAL)
    assert oracle.fastq_filter_count(text, 30) == (2 if mean > 30 else 0, 2)
    assert oracle.fastq_filter_count(text, int(mean)) == (2, 2) and oracle.fastq_filter_count(text, int(mean) + 1) == (0, 2)


def test_quality_scores_to_list():
    # slt/quality-score-udfs.slt:1-23
    assert oracle.quality_scores_to_list(b"###") == [2, 2, 2]
    assert oracle.quality_scores_to_list(b"!\"#$%&'()*+,-./0123456789:;<=>?@ABCDEFGHI") == list(range(41))


def test_batching_and_independent_reader():
    from synth import fastq

    sh = fastq.shards(20_000, 3)
    for f in sh.files:
        text = bytes(f)
        want = py_records(text)
        got = []
        sizes = []
        for b in oracle.fastq_read_batches(text, batch_size=1000):
            sizes.append(b["rows"])
            got += list(zip(b["name"], b["description"], b["sequence"], b["quality"]))
        assert got == want and all(s == 1000 for s in sizes[:-1])
    for t in (20, 30, (61, 2), 35):
        num, den = t if isinstance(t, tuple) else (t, 1)
        assert oracle.fastq_filter_count_files(sh.files, t, target_partitions=2) == (sh.truth_count(num, den), sh.n)


@pytest.mark.parametrize("text", [b"SEQ\nACGT\n+\n!!!!\n", b"@a\nACGT\n-\n!!!!\n", b"@a\nACGT\n", b"@a\n", b"@a\nAC\n+\n!!\n\n"])
def test_malformed(text):
    with pytest.raises(ValueError):
        oracle.fastq_filter_count(text)


def test_edge_semantics():
    assert oracle.fastq_filter_count(b"") == (0, 0)
    assert oracle.fastq_filter_count(b"@a\nAC\n+\nII", 30) == (1, 1)          # no trailing newline
    assert oracle.fastq_filter_count(b"@a\nAC\n+\n", 0) == (0, 1)             # missing quality line: empty string, no mean
    assert oracle.fastq_filter_count(b"@a\nAC\n+\n@@\n@b x y\n\n+b\n+I\n", 30) == (1, 2)  # '@' / '+' inside quality lines
    (b,) = list(oracle.fastq_read_batches(b"@b x y\n\n+b\n+I\n"))
    assert b["name"] == [b"b"] and b["description"] == [b"x y"] and b["sequence"] == [b""] and b["quality"] == [b"+I"]
