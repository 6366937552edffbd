"""BAM column batches built on the device (exon_gpu_bam_next_batch) against the reference's own goldens
(/root/reference/exon/exon-core/tests/sqllogictests/slt/bam-select-tests.slt:9-31: first row
`READ_ID 83 chr1 12203704 12217173 NULL 55M13394N21M chr1`, the 76-base poly-A sequence, quality_score[1] of the first five
rows 23 20 37 34 31, list length 76) and against the oracle's restatement of BAMArrayBuilder::append
(/root/reference/exon/exon-bam/src/array_builder.rs:102-218) on the fixture and on synthetic htslib-style files.
Batches are imported through the Arrow C Data Interface into pyarrow (validate(full=True))."""
import os

import pytest

import oracle
from conftest import GOLDEN
from exon_b200 import _abi
from exon_b200.runtime import ExonGpuError

pytestmark = pytest.mark.gpu
NAMES = ["name", "flag", "reference", "start", "end", "mapping_quality", "cigar", "mate_reference", "sequence", "quality_score"]


@pytest.fixture(scope="module")
def test_bam() -> bytes:
    with open(os.path.join(GOLDEN, "test.bam"), "rb") as f:
        return f.read()


def gpu_table(ctx, files, projection, batch_rows=8192, on_device=False):
    cols = {NAMES[p]: [] for p in projection}
    sizes = []
    with ctx.open_bam(projection=projection, batch_rows=batch_rows, columns_on_device=on_device) as s:
        for f in files:
            s.feed(f)
        for b in s.batches():
            if on_device:
                sizes.append(b.num_rows)
                b.release()
                continue
            rb = b.to_pyarrow()
            assert rb.schema.names == [NAMES[p] for p in projection]
            sizes.append(rb.num_rows)
            for p in projection:
                cols[NAMES[p]] += rb.column(NAMES[p]).to_pylist()
    return cols, sizes


def oracle_table(files, n_rows_each=None):
    want = {n: [] for n in NAMES}
    for f in files:
        b = oracle.Bam(f)
        try:
            n = b.count_by_reference(all_rows=True)[1]
            for i in range(n):
                r = b.row(i)
                seq, qual = b.seq_qual(i)
                for k in ("name", "flag", "reference", "start", "end", "mapping_quality", "cigar", "mate_reference"):
                    want[k].append(r[k])
                want["sequence"].append(seq)
                want["quality_score"].append(qual)
        finally:
            b.close()
    return want


def test_reference_goldens(gpu_ctx, test_bam):
    got, sizes = gpu_table(gpu_ctx, [test_bam], tuple(range(10)))
    assert sizes == [61]
    first = [got[k][0] for k in NAMES[:8]]
    assert first == ["READ_ID", 83, "chr1", 12203704, 12217173, None, "55M13394N21M", "chr1"]  # bam-select-tests.slt:9-12
    assert got["sequence"][0] == "A" * 76                                                          # :14-17
    assert [q[0] for q in got["quality_score"][:5]] == [23, 20, 37, 34, 31]                       # :19-26
    assert [len(q) for q in got["quality_score"][:5]] == [76] * 5                                  # :28-35


@pytest.mark.parametrize("projection", [tuple(range(10)), (1, 5), (9,), (8, 0, 6), (3, 4, 2, 7)])
def test_fixture_against_oracle(gpu_ctx, test_bam, projection):
    got, _ = gpu_table(gpu_ctx, [test_bam], projection)
    want = oracle_table([test_bam])
    for p in projection:
        assert got[NAMES[p]] == want[NAMES[p]], NAMES[p]


def test_batches_restart_per_file(gpu_ctx, test_bam):
    got, sizes = gpu_table(gpu_ctx, [test_bam, test_bam], (0, 1, 3, 9), batch_rows=25)
    assert sizes == [25, 25, 11, 25, 25, 11]  # bam-partition: 122 rows (bam-select-tests.slt:61-64), batches never span files
    want = oracle_table([test_bam])
    for k in ("name", "flag", "start", "quality_score"):
        assert got[k] == want[k] * 2


def test_synthetic_files(gpu_ctx):
    from synth import bam

    sh = bam.shards(6000, 3)
    got, sizes = gpu_table(gpu_ctx, sh.files, tuple(range(10)), batch_rows=1000)
    assert sum(sizes) == sh.n and all(x <= 1000 for x in sizes)
    want = oracle_table(sh.files)
    for k in NAMES:
        assert got[k] == want[k], k
    assert any(v is None for v in got["mapping_quality"]) and any(v is None for v in got["reference"])


def test_device_resident_and_errors(gpu_ctx, test_bam):
    _, sizes = gpu_table(gpu_ctx, [test_bam], (1, 8, 9), batch_rows=16, on_device=True)
    assert sizes == [16, 16, 16, 13]
    with pytest.raises(ExonGpuError) as e:
        gpu_ctx.open_bam(projection=(10,))
    assert e.value.code == _abi.ERR_UNSUPPORTED
    with pytest.raises(ExonGpuError):
        gpu_ctx.open_bam(projection=(1, 1))
    # a stream that produced batches still answers the fused query
    with gpu_ctx.open_bam(projection=(1,)) as s:
        s.feed(test_bam)
        n = sum(b.to_pyarrow().num_rows for b in s.batches())
        assert n == 61 and s.count_by_reference(all_rows=True)[1] == 61
