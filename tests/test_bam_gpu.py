"""GPU parity of the fused BAM scan -> flag / MAPQ filter -> per-reference COUNT (BASELINE configs[3]) through the C ABI."""
import os
import struct

import numpy as np
import pytest

import oracle
from bgzf_util import EOF_MARKER, bgzf_compress, bgzf_member
from conftest import GOLDEN
from exon_b200 import _abi
from exon_b200._abi import ExonGpuError

pytestmark = pytest.mark.gpu
PREDS = [dict(all_rows=True), dict(flag_exclude=0x904, min_mapq=30), dict(flag_exclude=0x904), dict(flag_require=0x10), dict(min_mapq=60),
         dict(flag_exclude=0xFFFF), dict(min_mapq=0)]


def fixture():
    with open(os.path.join(GOLDEN, "test.bam"), "rb") as f:
        return f.read()


def gpu(ctx, files, **kw):
    with ctx.open_bam() as s:
        for f in files:
            s.feed(f)
        return s.count_by_reference(**kw)


def test_reference_fixture(gpu_ctx):
    data = fixture()
    counts, rows = gpu(gpu_ctx, [data], all_rows=True)
    assert rows == 61 and counts["chr1"] == 61 and sum(counts.values()) == 61      # slt/bam-select-tests.slt:56-59
    assert gpu(gpu_ctx, [data, data], all_rows=True)[1] == 122                      # :61-64
    for kw in PREDS:
        assert gpu(gpu_ctx, [data], **kw) == oracle.bam_count_by_reference_files([data], **kw), kw
    assert len(counts) == 196 and None in counts
    # bam_region_filter goldens: slt/bam-indexed-select-tests.slt:11-14 (7), :22-25 (two files: 14)
    assert sum(gpu(gpu_ctx, [data], region=("chr1", 1, 12209145))[0].values()) == 7
    assert sum(gpu(gpu_ctx, [data, data], region=("chr1", 1, 12209145))[0].values()) == 14
    for rg in [("chr1", None, None), ("chr1", 12203704, 12203704), ("chr1", 12217174, None), ("chr2", 1, 10**9), ("nope", 1, 5)]:
        assert gpu(gpu_ctx, [data], region=rg) == oracle.bam_count_by_reference_files([data], region=rg), rg


def test_synthetic(gpu_ctx):
    from synth import bam

    sh = bam.shards(400_000, 5)
    with gpu_ctx.open_bam() as s:
        for f in sh.files:
            s.feed(f)
        for kw in PREDS:
            got, rows = s.count_by_reference(**kw)
            assert rows == sh.n
            assert got == sh.truth(**{k: v for k, v in kw.items() if k != "all_rows"}), kw
        assert got == oracle.bam_count_by_reference_files(sh.files, **PREDS[-1])[0]
        for kw in [dict(region=("7", 1_000_000, 2_000_000)), dict(region=("X", 500, 50_000_000), flag_exclude=0x904, min_mapq=30), dict(region=("MT", None, None))]:
            assert s.count_by_reference(**kw) == oracle.bam_count_by_reference_files(sh.files, **kw), kw
    # ranges of one file arrive in pieces
    with gpu_ctx.open_bam() as s:
        f = np.frombuffer(sh.files[0], dtype=np.uint8)
        for o in range(0, f.size, 1 << 20):
            s.feed(f[o:o + (1 << 20)], is_last=o + (1 << 20) >= f.size)
        assert s.count_by_reference(all_rows=True)[1] == oracle.Bam(sh.files[0]).count_by_reference(all_rows=True)[1]


def rebgzf(raw: bytes, block: int) -> bytes:
    return b"".join(bgzf_member(raw[o:o + block]) for o in range(0, len(raw), block)) + EOF_MARKER


@pytest.mark.parametrize("block", [0xFF00, 4096, 1000, 211, 64])
def test_records_straddling_members(gpu_ctx, block):
    """Members cut at arbitrary byte positions (records and even block_size fields straddle them): the speculative
    walks are corrected, or the serial fallback runs; the answer is the same."""
    import gzip
    from synth import bam

    sh = bam.shards(3000 if block >= 1000 else 300, 1)
    raw = gzip.decompress(sh.files[0])
    data = rebgzf(raw, block)
    want = oracle.bam_count_by_reference_files([sh.files[0]], flag_exclude=0x904, min_mapq=30)
    assert gpu(gpu_ctx, [data], flag_exclude=0x904, min_mapq=30) == want
    assert gpu(gpu_ctx, [data, sh.files[0]], all_rows=True)[1] == 2 * sh.n


def test_header_larger_than_the_probe(gpu_ctx):
    import gzip
    from synth import bam

    refs = [(f"contig_with_a_long_name_{i:06d}", 1000 + i) for i in range(4000)]   # ~150 KB of header
    sh = bam.shards(5000, 1, refs=refs)
    got, rows = gpu(gpu_ctx, sh.files, flag_exclude=0x904, min_mapq=30)
    assert rows == 5000 and got == sh.truth(flag_exclude=0x904, min_mapq=30) and len(got) == 4001


def test_malformed(gpu_ctx):
    import gzip
    from synth import bam

    sh = bam.shards(2000, 1)
    raw = bytearray(gzip.decompress(sh.files[0]))
    with pytest.raises(ExonGpuError):
        gpu(gpu_ctx, [bgzf_compress(bytes(raw[:-7]))], all_rows=True)            # truncated last record
    hdr_len = len(bam.header_bytes())
    bad = bytearray(raw)
    bad[hdr_len:hdr_len + 4] = struct.pack("<i", 5)                                # block_size < 32
    with pytest.raises(ExonGpuError) as e:
        gpu(gpu_ctx, [bgzf_compress(bytes(bad))], all_rows=True)
    assert e.value.code == _abi.ERR_PARSE
    with pytest.raises(ExonGpuError):
        gpu(gpu_ctx, [bgzf_compress(b"not a bam")], all_rows=True)
    for text in [bytes(raw[:-7]), bytes(bad)]:
        with pytest.raises(ValueError):
            oracle.Bam(bgzf_compress(text)).count_by_reference(all_rows=True)
