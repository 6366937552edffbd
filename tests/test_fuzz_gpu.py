"""Seeded random-input parity sweeps: every fused path and column builder against its oracle on many small, irregular
inputs (ragged feeds, odd alignments, tiny and huge records, compressed and plain), complementing the fixed cases."""
import gzip
import struct
import zlib

import numpy as np
import pytest

import oracle
from bgzf_util import EOF_MARKER, bgzf_compress, bgzf_member
from exon_b200._abi import ExonGpuError, make_region

pytestmark = pytest.mark.gpu
CHROMS = ["1", "2", "10", "X", "chr1", "chr10", "chrUn_KI270742v1_decoy", "MT", "22"]


def rand_vcf(rng, n_rows, with_header=True):
    lines = []
    if with_header:
        lines += ["##fileformat=VCFv4.2"] + [f"##contig=<ID={c}>" for c in CHROMS[: int(rng.integers(0, len(CHROMS)))]]
        lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO")
    for _ in range(n_rows):
        c = CHROMS[int(rng.integers(0, len(CHROMS)))]
        p = int(rng.integers(1, 10 ** int(rng.integers(1, 12))))
        tail = "\t.\tA\tC\t50\tPASS\t" + ("." if rng.random() < 0.7 else "DP=" + "9" * int(rng.integers(1, 400)))
        lines.append(f"{c}\t{p}{tail}")
    text = "\n".join(lines)
    if n_rows or with_header:
        text += "\n" if rng.random() < 0.8 else ""
    return text.encode()


def feed_ragged(rng, s, data, gz=False):
    if gz:
        s.feed_gzip(bgzf_compress(data, int(rng.integers(0, 10)), block=int(rng.integers(200, 0xFF00))))
        return
    if rng.random() < 0.3 or len(data) < 2:
        s.feed(data)
        return
    cuts = sorted(set(int(x) for x in rng.integers(1, len(data), int(rng.integers(1, 6)))))
    prev = 0
    for c in cuts + [len(data)]:
        s.feed(data[prev:c], is_last=c == len(data))
        prev = c


@pytest.mark.parametrize("seed", range(12))
def test_vcf_counts_and_columns(gpu_ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    files = [rand_vcf(rng, int(rng.integers(0, 4000)), rng.random() < 0.9) for _ in range(int(rng.integers(1, 5)))]
    queries = [(CHROMS[int(rng.integers(0, len(CHROMS)))], None, None), (None, None, None)]
    lo = int(rng.integers(1, 10**6))
    queries += [(CHROMS[int(rng.integers(0, len(CHROMS)))], lo, lo * int(rng.integers(1, 1000))), (None, lo, lo * 50)]
    for strict in (False, True):
        with gpu_ctx.open_vcf(strict=strict, batch_rows=int(rng.integers(1, 3000))) as s:
            for f in files:
                feed_ragged(rng, s, f, gz=rng.random() < 0.4)
            for q in queries:
                want = sum(oracle.filter_count(f, *q)[0] for f in files)
                assert s.filter_count(make_region(*q)) == want, (seed, strict, q)
                cols = {"chrom_col": 0, "pos_col": 1}
                assert s.filter_agg(region=make_region(*q), **cols)[0] == want if any(x is not None for x in q) else True
    br = int(rng.integers(1, 3000))
    with gpu_ctx.open_vcf(batch_rows=br) as s:
        for f in files:
            feed_ragged(rng, s, f, gz=rng.random() < 0.4)
        got = [(b.num_rows, b.chrom_strings(), b.column("pos").tolist()) for b in s.batches()]
    want = []
    for f in files:
        for b in oracle.read_batches(f, batch_size=br):
            off, val = b["chrom_offsets"], b["chrom_values"].tobytes()
            want.append((b["rows"], [val[off[i]:off[i + 1]].decode() for i in range(b["rows"])], b["pos"].tolist()))
    assert got == want, seed


def rand_fastq(rng, n):
    out = []
    for i in range(n):
        ln = int(rng.integers(0, 3 if rng.random() < 0.2 else (400 if rng.random() < 0.95 else 9000)))
        q = bytes(rng.integers(33, 75, ln).astype(np.uint8))
        desc = (b" " + bytes(rng.integers(33, 126, int(rng.integers(0, 30))).astype(np.uint8))) if rng.random() < 0.5 else b""
        out.append(b"@r%d" % i + desc + b"\n" + b"ACGT"[: ln % 5] * (ln // max(ln % 5, 1) if ln % 5 else 0) + b"\n+\n" + q + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("seed", range(10))
def test_fastq_counts_and_columns(gpu_ctx, seed):
    rng = np.random.default_rng(2000 + seed)
    files = [rand_fastq(rng, int(rng.integers(0, 1500))) for _ in range(int(rng.integers(1, 4)))]
    if rng.random() < 0.5 and files[-1].endswith(b"\n"):
        files[-1] = files[-1][:-1]  # no trailing newline
    with gpu_ctx.open_fastq() as s:
        for f in files:
            feed_ragged(rng, s, f, gz=rng.random() < 0.4)
        for t in (None, 10, 20, (41, 2)):
            want = [oracle.fastq_filter_count(f, t) for f in files]
            assert s.filter_count(t) == sum(w[0] for w in want), (seed, t)
        assert s.rows() == sum(w[1] for w in want)
    br = int(rng.integers(1, 700))
    with gpu_ctx.open_fastq(projection=(0, 1, 2, 3), batch_rows=br) as s:
        for f in files:
            feed_ragged(rng, s, f, gz=rng.random() < 0.4)
        got = [r for b in s.batches() for r in zip(b.strings("name"), b.strings("description"), b.strings("sequence"), b.strings("quality_scores"))]
    want = [r for f in files for b in oracle.fastq_read_batches(f, batch_size=br)
            for r in zip(b["name"], b["description"], b["sequence"], b["quality"])]
    assert got == want, seed


@pytest.mark.parametrize("seed", range(8))
def test_inflate_random_streams(gpu_ctx, seed):
    rng = np.random.default_rng(3000 + seed)
    kind = seed % 4
    n = int(rng.integers(1, 400_000))
    if kind == 0:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    elif kind == 1:
        data = bytes(rng.choice(np.frombuffer(b"ACGT\n\t0123", dtype=np.uint8), n))
    elif kind == 2:
        data = (b"the quick brown fox " * (n // 20 + 1))[:n]
    else:
        data = bytes(np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), 50)[:n])
    level = int(rng.integers(0, 10))
    strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED][int(rng.integers(0, 5))]
    comp = bgzf_compress(data, level, strategy, block=int(rng.integers(1, 0xFF00)))
    assert gpu_ctx.gzip_inflate(comp).tobytes() == data
    assert gpu_ctx.gzip_inflate(gzip.compress(data, max(level, 1))).tobytes() == data


def rand_bam(rng, n, refs):
    hdr_text = "@HD\tVN:1.6\n" + "".join(f"@SQ\tSN:{r}\tLN:1000000\n" for r in refs)
    raw = [b"BAM\x01", struct.pack("<i", len(hdr_text)), hdr_text.encode(), struct.pack("<i", len(refs))]
    for r in refs:
        raw += [struct.pack("<i", len(r) + 1), r.encode() + b"\0", struct.pack("<i", 1000000)]
    for i in range(n):
        name = b"q%d" % i + b"x" * int(rng.integers(0, 40)) + b"\0"
        l_seq = int(rng.integers(0, 300))
        n_cig = int(rng.integers(0, 4))
        body = struct.pack("<iiBBHHHIiii", int(rng.integers(-1, len(refs))), int(rng.integers(0, 999000)), len(name), int(rng.choice([0, 10, 29, 30, 60, 255])),
                           4680, n_cig, int(rng.choice([0, 4, 16, 83, 99, 256, 1024, 2048, 2064])), l_seq, -1, -1, 0)
        body += name + b"".join(struct.pack("<I", (int(rng.integers(1, 100)) << 4) | int(rng.integers(0, 9))) for _ in range(n_cig))
        body += bytes(rng.integers(0, 256, (l_seq + 1) // 2, dtype=np.uint8)) + bytes(rng.integers(0, 60, l_seq, dtype=np.uint8))
        raw.append(struct.pack("<i", len(body)) + body)
    return b"".join(raw)


@pytest.mark.parametrize("seed", range(8))
def test_bam_counts(gpu_ctx, seed):
    rng = np.random.default_rng(4000 + seed)
    files = []
    for _ in range(int(rng.integers(1, 4))):
        refs = [f"c{j}" for j in range(int(rng.integers(1, 12)))]
        raw = rand_bam(rng, int(rng.integers(0, 3000)), refs)
        block = int(rng.integers(40, 0xFF00))  # members cut anywhere: records straddle them
        files.append(b"".join(bgzf_member(raw[o:o + block], int(rng.integers(0, 7))) for o in range(0, len(raw), block)) + EOF_MARKER)
    with gpu_ctx.open_bam() as s:
        for f in files:
            s.feed(f)
        for kw in [dict(all_rows=True), dict(flag_exclude=0x904, min_mapq=30), dict(flag_require=0x10), dict(min_mapq=0),
                   dict(region=("c0", int(rng.integers(1, 500000)), int(rng.integers(500000, 1000000)))), dict(region=("c3", None, None), flag_exclude=4)]:
            assert s.count_by_reference(**kw) == oracle.bam_count_by_reference_files(files, **kw), (seed, kw)
