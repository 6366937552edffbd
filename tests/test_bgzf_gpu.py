"""Device BGZF / gzip inflate (exon_b200/csrc/bgzf.cu) against zlib, byte for byte, and through the VCF / FASTQ
streams against the reference's .gz goldens (slt/vcf-select-tests.slt:52-55 -> 621, slt/fastq-scan-test.slt:66-69 -> 2)."""
import gzip
import os
import zlib

import numpy as np
import pytest

import oracle
from bgzf_util import EOF_MARKER, bgzf_compress
from conftest import GOLDEN, make_vcf
from exon_b200 import _abi
from exon_b200._abi import ExonGpuError, make_region

pytestmark = pytest.mark.gpu


def raw(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def test_reference_fixtures_inflate_exactly(gpu_ctx):
    for name in ["index.vcf.gz", "index_plain.vcf.gz", "biobear_vcf_file.vcf.gz", "common_all_head.vcf.gz", "test_bgzip.fastq.gz"]:
        data = raw(name)
        assert gpu_ctx.gzip_inflate(data).tobytes() == gzip.decompress(data), name
    assert gpu_ctx.gzip_inflate(EOF_MARKER).size == 0 and gpu_ctx.gzip_inflate(b"").size == 0


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                                            (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
def test_synthetic_text_all_block_types(gpu_ctx, level, strategy):
    from synth import vcf

    cols = vcf.columns(150_000, seed=3)
    text = vcf.shards(cols, 1)[0].tobytes()
    comp = bgzf_compress(text, level, strategy)
    assert gpu_ctx.gzip_inflate(comp).tobytes() == text
    # plain single-member gzip of the same text (one warp walks the whole stream; distances reach back 32 KiB)
    if level == 6 and strategy == zlib.Z_DEFAULT_STRATEGY:
        assert gpu_ctx.gzip_inflate(gzip.compress(text[:3_000_000], 6)).tobytes() == text[:3_000_000]


def test_binary_and_degenerate_inputs(gpu_ctx):
    rng = np.random.default_rng(1)
    cases = [rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes(),           # incompressible
             b"\x00" * 300_000, b"ab" * 100_000, b"x",                            # runs: distance 1, distance 2
             bytes(rng.integers(0, 4, 500_000, dtype=np.uint8) + 65),             # 2-bit alphabet: long codes are rare, short frequent
             bytes(np.minimum(rng.geometric(0.02, 400_000), 255).astype(np.uint8))]  # skewed alphabet: code lengths up to 15
    for i, c in enumerate(cases):
        for level in (1, 9):
            assert gpu_ctx.gzip_inflate(bgzf_compress(c, level)).tobytes() == c, (i, level)
        assert gpu_ctx.gzip_inflate(bgzf_compress(c, 6, block=1000)).tobytes() == c  # many tiny members


def test_more_members_than_decoder_lanes(gpu_ctx):
    """One launch holds at most 148 x 12 x 32 = 56 832 members in flight: with more, lanes (and the copy kernel's warps) take
    several members each; with a handful, the members are spread over warps with few working lanes."""
    rng = np.random.default_rng(5)
    text = bytes(rng.integers(0, 6, 6_000_000, dtype=np.uint8) + 97)
    for block, n in ((80, 6_000_000), (80, 4000), (3000, 200_000)):
        comp = bgzf_compress(text[:n], 1, block=block)
        assert gpu_ctx.gzip_inflate(comp).tobytes() == text[:n], (block, n)


def test_corrupt_members_fail(gpu_ctx):
    text = make_vcf([("1", str(i + 1)) for i in range(5000)])
    comp = bytearray(bgzf_compress(text))
    bad = bytearray(comp)
    bad[40] ^= 0x55  # inside the first payload
    with pytest.raises(ExonGpuError):
        gpu_ctx.gzip_inflate(bytes(bad))
    with pytest.raises(ExonGpuError) as e:
        gpu_ctx.gzip_inflate(b"not a gzip file at all")
    assert e.value.code == _abi.ERR_PARSE
    with pytest.raises(ExonGpuError):
        gpu_ctx.gzip_inflate(bytes(comp[: len(comp) // 2]))  # truncated member


def test_streams_that_run_off_their_payload_fail_cleanly(gpu_ctx):
    """A member whose DEFLATE stream has lost its tail (no end-of-block inside the payload) and a plain gzip file cut short
    with a garbage trailer: the reader must never read past the payload (round-1 advice: it used to run up to 6 x ISIZE
    bytes ahead), the call fails with a parse error and the context stays usable."""
    import struct

    rng = np.random.default_rng(7)
    text = bytes(rng.integers(32, 127, 60_000, dtype=np.uint8))  # literal-heavy: a long stream per member
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    payload = co.compress(text) + co.flush()
    for keep in (len(payload) // 2, len(payload) - 3, 5):
        cut = payload[:keep]
        hdr = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cut) + 25)
        member = hdr + cut + struct.pack("<II", zlib.crc32(text) & 0xFFFFFFFF, len(text))  # trailer still claims the full ISIZE
        with pytest.raises(ExonGpuError) as e:
            gpu_ctx.gzip_inflate(member + EOF_MARKER)
        assert e.value.code == _abi.ERR_PARSE
    # plain gzip, truncated in the middle of the stream: the "trailer" is whatever bytes happen to be there
    gz = gzip.compress(text, 6)
    for keep in (len(gz) // 2, len(gz) - 9):
        with pytest.raises(ExonGpuError) as e:
            gpu_ctx.gzip_inflate(gz[:keep])
        assert e.value.code in (_abi.ERR_PARSE, _abi.ERR_ARG, _abi.ERR_OOM)
    # a stored block after which the stream is cut: the consumed-bit count must keep running across the stored block
    co = zlib.compressobj(0, zlib.DEFLATED, -15)
    stored = co.compress(text[:1000]) + co.flush(zlib.Z_FULL_FLUSH)
    co2 = zlib.compressobj(6, zlib.DEFLATED, -15)
    tail = co2.compress(text[1000:]) + co2.flush()
    cut = stored + tail[: len(tail) // 3]
    hdr = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cut) + 25)
    with pytest.raises(ExonGpuError):
        gpu_ctx.gzip_inflate(hdr + cut + struct.pack("<II", 0, len(text)) + EOF_MARKER)
    # and the context still works
    good = bgzf_compress(text)
    assert gpu_ctx.gzip_inflate(good).tobytes() == text


def test_vcf_gz_goldens_through_the_stream(gpu_ctx, index_vcf):
    data = raw("index.vcf.gz")  # the reference's BGZF fixture
    with gpu_ctx.open_vcf() as s:
        s.feed_gzip(data)
        assert s.filter_count(None) == 621                       # slt/vcf-select-tests.slt:52-55
        assert s.filter_count(make_region("1")) == 191
        assert s.filter_count(make_region("1", 9999919, 10000000)) == 82
        text = gzip.decompress(data)
        assert s.body_bytes() == len(text) - oracle.header_len(text)
        want = list(oracle.read_batches(text))
        got = list(s.batches())
        assert len(got) == len(want) == 1 and np.array_equal(got[0].column("pos"), want[0]["pos"])
    with gpu_ctx.open_vcf() as s:                                # vcf-partition: two copies, region '1' -> 382
        s.feed_gzip(data)
        s.feed_gzip(data[:5000], is_last=False)                  # ranges of one file are buffered until it is complete
        s.feed_gzip(data[5000:], is_last=True)
        s.feed(index_vcf)                                        # a plain-text file in the same partition
        assert s.filter_count(make_region("1")) == 3 * 191 and s.rows() == 3 * 621
    with gpu_ctx.open_vcf() as s:                                # biobear: region '1' -> 11 (slt/vcf-indexed-tests.slt:56-59)
        s.feed_gzip(raw("biobear_vcf_file.vcf.gz"))
        assert s.filter_count(make_region("1")) == 11 and s.filter_count(make_region("1000")) == 0


def test_vcf_gz_synthetic_equals_plain(gpu_ctx):
    from synth import vcf

    cols = vcf.columns(400_000, seed=8)
    files = vcf.shards(cols, 5)
    gz = [bgzf_compress(f.tobytes(), 6) for f in files]
    nonl = make_vcf([("1", "1500000")], trailing_newline=False)
    with gpu_ctx.open_vcf(pushdown=make_region("1", 1_000_000, 2_000_000)) as s, gpu_ctx.open_vcf() as t:
        for g in gz:
            s.feed_gzip(g)
            t.feed_gzip(g)
        s.feed_gzip(bgzf_compress(nonl))   # last record without '\n'
        t.feed_gzip(bgzf_compress(nonl))
        s.feed_gzip(EOF_MARKER)            # an empty file
        for q in [("1", 1_000_000, 2_000_000), ("X", None, None), (None, None, None), (None, 5, 40_000_000)]:
            want = cols.truth_count(*q) + (1 if q[0] in ("1", None) and (q[1] is None or q[1] <= 1_500_000 <= q[2]) else 0)
            assert s.filter_count(make_region(*q)) == want == t.filter_count(make_region(*q)), q
        pos = np.concatenate([b.column("pos") for b in t.batches()])
        assert np.array_equal(pos, np.concatenate([cols.pos, [1_500_000]]))


def test_fastq_gz(gpu_ctx):
    from synth import fastq

    with gpu_ctx.open_fastq() as s:
        s.feed_gzip(raw("test_bgzip.fastq.gz"))
        assert s.filter_count(None) == 2                        # slt/fastq-scan-test.slt:66-69
    sh = fastq.shards(100_000, 3)
    with gpu_ctx.open_fastq() as s:
        for f in sh.files:
            s.feed_gzip(bgzf_compress(f.tobytes(), 4))
        assert s.filter_count(30) == sh.truth_count(30) and s.rows() == sh.n
