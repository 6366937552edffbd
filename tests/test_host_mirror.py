"""The C++ mirror of the reference's host-side operators (exon_b200/host/): planning logic on the CPU, and the
reference's own VCF sqllogictests (restated in tests/slt/vcf.slt) executed on the GPU through ExonSession::sql."""
import ctypes as C
import os

import pytest

from conftest import ROOT, has_gpu
from host_util import build_datasources, load_host, parse_slt


def test_host_library_exports():
    L = load_host()
    for name in ["exon_host_session_new", "exon_host_session_free", "exon_host_sql", "exon_host_last_error", "exon_host_pushdown"]:
        assert hasattr(L, name)


def test_supports_filters_pushdown():
    """vcf/table_provider.rs:299-320: Exact only for vcf_region_filter (2 or 3 args) and hive-partition equality."""
    L = load_host()
    buf = C.create_string_buffer(64)

    def push(where, part_cols=""):
        assert L.exon_host_pushdown(None, part_cols.encode(), where.encode(), buf, 64) == 0, L.exon_host_last_error()
        return buf.value.decode()

    assert push("chrom = '1' AND pos BETWEEN 1000000 AND 2000000") == "UU"
    assert push("pos >= 5 AND pos <= 7 AND chrom = 'X'") == "UUU"
    assert push("vcf_region_filter('1', chrom) = true") == "E"
    assert push("vcf_region_filter('1:5-6', chrom, pos) = true AND chrom = '1'") == "EU"
    assert push("vcf_region_filter('1') = true") == "U"             # wrong arity is not pushed down
    assert push("sample = '1' AND chrom = '1'", "sample") == "EU"   # filter_matches_partition_cols
    assert push("sample = '1'") == "U"
    assert push("region_match(chrom, pos, '1:1-1') = true") == "U"  # evaluated UDFs stay in FilterExec


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_session_fails_loudly_without_gpu():
    L = load_host()
    assert L.exon_host_session_new(0) is None
    assert b"no CPU fallback" in L.exon_host_last_error()


@pytest.mark.gpu
def test_reference_slt_on_gpu(tmp_path):
    L = load_host()
    data = build_datasources(str(tmp_path / "datasources"))
    text = open(os.path.join(ROOT, "tests", "slt", "vcf.slt")).read().replace("$DATA", data)
    s = L.exon_host_session_new(0)
    assert s, L.exon_host_last_error()
    try:
        n_queries = 0
        for kind, sql, want in parse_slt(text):
            out = C.c_char_p()
            rc = L.exon_host_sql(s, sql.encode(), C.byref(out))
            if kind == "error":
                assert rc != 0, f"expected an error: {sql}"
            else:
                assert rc == 0, f"{sql}: {L.exon_host_last_error().decode()}"
                if kind == "query":
                    assert out.value.decode().splitlines() == want, sql
                    n_queries += 1
        assert n_queries >= 19 and L.exon_host_gpu_launches(s) > 0
    finally:
        L.exon_host_session_free(s)


@pytest.mark.gpu
def test_error_kinds_and_malformed_input(tmp_path):
    """Plan / NotImplemented / External errors as the reference raises them."""
    L = load_host()
    data = build_datasources(str(tmp_path / "datasources"))
    bad = tmp_path / "bad.vcf"
    bad.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n2\t12x\t.\tA\tC\t.\t.\t.\n1\t5\t.\tA\tC\t.\t.\t.\n")
    s = L.exon_host_session_new(0)
    try:
        def run(sql):
            out = C.c_char_p()
            rc = L.exon_host_sql(s, sql.encode(), C.byref(out))
            return rc, (out.value.decode() if rc == 0 else L.exon_host_last_error().decode())

        rc, msg = run(f"CREATE EXTERNAL TABLE t STORED AS INDEXED_VCF LOCATION '{data}/vcf/index.vcf.gz' OPTIONS (compression gzip)")
        assert rc == 0
        rc, msg = run("SELECT COUNT(*) FROM t")
        assert rc == 1 and "requires a region filter" in msg                       # DataFusionError::Plan
        rc, msg = run("SELECT COUNT(*) FROM t WHERE vcf_region_filter('1', chrom) = true AND vcf_region_filter('2', chrom) = true")
        assert rc == 3 and "Multiple regions" in msg                               # NotImplemented
        rc, msg = run(f"SELECT COUNT(*) FROM vcf_scan('{bad}') WHERE chrom = '1'")
        assert rc == 4 and "malformed VCF record" in msg                           # External (strict: row 0 is bad)
        rc, msg = run("SET exon.gpu_strict = false")
        rc, msg = run(f"SELECT COUNT(*) FROM vcf_scan('{bad}') WHERE chrom = '1'")
        assert (rc, msg.strip()) == (0, "1")                                       # lazy: only matching rows are parsed
        rc, msg = run(f"SELECT COUNT(*) FROM vcf_scan('{tmp_path}/missing.vcf')")
        assert rc != 0 and "not found" in msg
    finally:
        L.exon_host_session_free(s)
