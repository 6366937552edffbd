"""Arrow C stream export (seam B4: the reference's FFI hands out an FFI_ArrowArrayStream, exon/exon-core/src/ffi/mod.rs:25-73):
exon_gpu_stream_export imported with pyarrow.RecordBatchReader, against the per-batch calls and the reference's golden rows."""
import gzip
import os

import pytest

import oracle
from conftest import GOLDEN
from exon_b200 import _abi
from exon_b200._abi import ExonGpuError
from exon_b200.runtime import export_reader

pytestmark = pytest.mark.gpu


def test_vcf_stream_reader(gpu_ctx, index_vcf):
    hdr = bytes(index_vcf[: oracle.header_len(index_vcf)])
    with gpu_ctx.open_vcf(projection=(0, 1, 3, 5, 7, 8), batch_rows=100) as s:
        s.set_header(hdr)
        s.feed(index_vcf, is_last=True)
        rd = export_reader(s)
        assert rd.schema.names == ["chrom", "pos", "ref", "qual", "info", "formats"]
        tbl = rd.read_all()
    assert tbl.num_rows == 621 and len(tbl.to_batches()) == 7                       # slt/vcf-select-tests.slt:47-50
    assert tbl.column("info")[0].as_py() == "DP=1;I16=1,0,0,0,26,676,0,0,60,3600,0,0,0,0,0,0;QS=1,0;MQ0F=0"   # :6-10
    assert tbl.column("formats")[0].as_py() == "GT:PL:PG\t0/0:0,3,26:0"             # :12-15
    want = [b for b in oracle.read_batches(index_vcf)]
    import numpy as np

    assert np.array_equal(tbl.column("pos").to_numpy(), np.concatenate([b["pos"] for b in want]))


def test_stream_owns_and_reports_errors(gpu_ctx):
    s = gpu_ctx.open_vcf(projection=(0, 1))
    s.feed(b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n1\tx\t.\tA\tC\t.\t.\t.\n", is_last=True)
    rd = export_reader(s, take_ownership=True)   # releasing the reader closes the stream
    assert rd.schema.names == ["chrom", "pos"]
    with pytest.raises(Exception) as e:
        rd.read_all()
    assert "POS" in str(e.value)                 # get_last_error carries the library's message
    del rd
    with gpu_ctx.open_vcf(projection=(0, 1), columns_on_device=True) as d:
        with pytest.raises(ExonGpuError) as e:
            export_reader(d)
        assert e.value.code == _abi.ERR_UNSUPPORTED


def test_other_formats_through_the_stream(gpu_ctx):
    with gpu_ctx.open_bam(projection=(0, 1, 2, 3)) as s:
        s.feed(open(os.path.join(GOLDEN, "test.bam"), "rb").read())
        tbl = export_reader(s).read_all()
        assert tbl.num_rows == 61 and tbl.column("name")[0].as_py() == "READ_ID"   # slt/bam-select-tests.slt:9-12
    with gpu_ctx.open_fasta(projection=(0, 1, 2)) as s:
        s.feed(open(os.path.join(GOLDEN, "test.fasta"), "rb").read())
        assert export_reader(s).read_all().to_pylist()[0] == {"id": "a", "description": "description", "sequence": "ATCG"}
    with gpu_ctx.open_gff(projection=(0, 1, 3, 4)) as s:
        s.feed(gzip.open(os.path.join(GOLDEN, "test.gff.gz")).read())
        assert export_reader(s).read_all().num_rows == 5000                          # slt/gff-scan-tests.slt:80-83
    with gpu_ctx.open_fastq(projection=(0, 2)) as s:
        s.feed(open(os.path.join(GOLDEN, "test.fastq"), "rb").read())
        assert export_reader(s).read_all().num_rows == 2                             # slt/fastq-scan-test.slt:51-54
