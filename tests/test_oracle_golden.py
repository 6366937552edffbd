"""Pins the CPU oracle (oracle/vcf_oracle.c) to the reference's own known answers."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, REF_DATA, golden_text, make_vcf


def test_count_star_index_vcf(index_vcf, index_vcf_gz_twin, goldens):
    g = goldens["index.vcf"]["reference_pinned"]
    assert oracle.filter_count(index_vcf) == (g["count_star"]["value"], 621)
    assert oracle.filter_count(index_vcf_gz_twin)[0] == 621  # vcf-select-tests.slt:52-55 (gzip twin)


def test_region_counts_index_vcf(index_vcf, goldens):
    g = goldens["index.vcf"]
    assert oracle.filter_count(index_vcf, "1")[0] == g["reference_pinned"]["chrom_1"]["value"] == 191
    assert oracle.filter_count(index_vcf, "a")[0] == g["reference_pinned"]["chrom_a"]["value"] == 0
    # vcf-partition = two copies of the same data -> 382 (slt/vcf-indexed-tests.slt:32-43)
    c, rows, parts = oracle.filter_count_files([index_vcf, index_vcf], "1", target_partitions=8)
    assert (c, rows, parts) == (g["reference_pinned"]["chrom_1_two_copies"]["value"], 1242, 2)
    d = g["derived"]
    assert oracle.filter_count(index_vcf, "2")[0] == d["chrom_2"] == 219
    assert oracle.filter_count(index_vcf, "10")[0] == d["chrom_10"] == 211
    assert oracle.filter_count(index_vcf, "1", 9999919, 10000000)[0] == d["chrom_1_pos_9999919_10000000"] == 82
    assert oracle.filter_count(index_vcf, "1", 1000000, 2000000)[0] == d["chrom_1_pos_1000000_2000000"] == 0
    assert oracle.filter_count(index_vcf, None, 10000000, None)[0] == d["pos_ge_10000000"]


def test_biobear(biobear_vcf, goldens):
    g = goldens["biobear_vcf_file.vcf"]
    assert oracle.filter_count(biobear_vcf, "1")[0] == g["reference_pinned"]["chrom_1"]["value"] == 11
    assert oracle.filter_count(biobear_vcf, "1000")[0] == g["reference_pinned"]["chrom_1000"]["value"] == 0
    assert oracle.filter_count(biobear_vcf)[0] == g["derived"]["count_star"] == 15
    (b,) = list(oracle.read_batches(biobear_vcf))
    vals = b["chrom_values"].tobytes()
    rows = [[vals[b["chrom_offsets"][i]:b["chrom_offsets"][i + 1]].decode(), int(b["pos"][i])] for i in range(b["rows"])]
    assert rows == g["derived"]["rows"]  # duplicate POS and multi-allelic rows included


def test_columns_index_vcf(index_vcf, common_all_vcf, goldens):
    d = goldens["index.vcf"]["derived"]
    batches = list(oracle.read_batches(index_vcf, batch_size=8192))
    assert len(batches) == 1 and batches[0]["rows"] == 621
    b = batches[0]
    assert int(b["pos"].sum()) == d["pos_sum"] and int(b["chrom_offsets"][-1]) == d["chrom_bytes"]
    assert b["chrom_offsets"][0] == 0
    # smaller batches: offsets restart at 0 in every batch
    small = list(oracle.read_batches(index_vcf, batch_size=100))
    assert [x["rows"] for x in small] == [100] * 6 + [21]
    assert all(x["chrom_offsets"][0] == 0 for x in small)
    assert np.concatenate([x["pos"] for x in small]).tolist() == b["pos"].tolist()
    g = goldens["common_all_head.vcf"]["derived"]
    (c,) = list(oracle.read_batches(common_all_vcf))
    assert c["rows"] == g["count_star"] and c["pos"].tolist() == [r[1] for r in g["rows"]]


def test_udf_truth_tables(goldens):
    t = goldens["udf_truth_tables"]
    chroms = [r[0] for r in t["rows"]]
    pos = np.array([r[1] for r in t["rows"]], dtype=np.int64)
    off = np.zeros(len(chroms) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(c) for c in chroms])
    val = np.frombuffer("".join(chroms).encode(), dtype=np.uint8)
    assert oracle.region_match(off, val, pos, "1:1-1").tolist() == t["region_match(chrom,pos,'1:1-1')"]
    assert oracle.interval_match(pos, "1-1").tolist() == t["interval_match(pos,'1-1')"]
    assert oracle.chrom_match(off, val, "1").tolist() == t["chrom_match(chrom,'1')"]


def test_physical_expr_vectors(goldens):
    v = goldens["physical_expr_vectors"]["region chr1:1-1"]
    chroms = [r[0] for r in v["rows"]]
    pos = np.array([r[1] for r in v["rows"]], dtype=np.int64)
    off = np.zeros(len(chroms) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(c) for c in chroms])
    val = np.frombuffer("".join(chroms).encode(), dtype=np.uint8)
    assert oracle.region_match(off, val, pos, "chr1:1-1").tolist() == v["expect"]
    p = goldens["physical_expr_vectors"]["pos = 1"]
    assert oracle.interval_match(np.array(p["pos"], dtype=np.int64), "1-1").tolist() == p["expect"]


def test_region_parsing():
    r = oracle.parse_region("1")
    assert (r.name[:r.name_len], r.has_interval) == (b"1", 0)
    r = oracle.parse_region("1:9999921")
    assert (r.name[:r.name_len], r.has_interval, r.lo, r.hi) == (b"1", 1, 9999921, oracle.INT64_MAX)
    r = oracle.parse_region("chr1:1-3388930")  # indexed_bgzf_file.rs:167-187
    assert (r.name[:r.name_len], r.lo, r.hi) == (b"chr1", 1, 3388930)
    assert oracle.parse_interval("1-1") == (1, 1)


def test_regroup_files_by_size():
    # exon_file_scan_config.rs:79-110: ascending by size, round-robin
    parts, g = oracle.regroup_files_by_size([30, 10, 20, 40], 2)
    assert parts == 2 and g == [0, 0, 1, 1]  # sorted 10,20,30,40 -> 0,1,0,1
    parts, g = oracle.regroup_files_by_size([5, 5, 5], 8)
    assert parts == 3 and sorted(g) == [0, 1, 2]


def test_malformed_records():
    with pytest.raises(ValueError):
        oracle.filter_count(make_vcf([("1", "0")]), "1", 1, 10)  # POS 0 -> None in a non-nullable column
    with pytest.raises(ValueError):
        oracle.filter_count(make_vcf([("1", "12x")]), "1", 1, 10)
    with pytest.raises(ValueError):
        oracle.filter_count(b"1\t100\t.\tA\n", "1", 1, 10)  # fewer than 8 fields
    assert oracle.filter_count(make_vcf([("1", "+7")]), "1", 1, 10)[0] == 1  # Rust usize::from_str takes '+'
    assert oracle.filter_count(make_vcf([("1", "7")], trailing_newline=False), "1", 1, 10)[0] == 1
    assert oracle.filter_count(b"", "1")[0] == 0
    assert oracle.filter_count(make_vcf([]), "1")[0] == 0


def test_bigger_index_columns_golden():
    """Very long lines (3202 samples): the oracle's columns equal the pure-Python derivation (committed npz).
    Needs the 44 MB reference fixture, so it only runs where /root/reference exists."""
    path = os.path.join(REF_DATA, "bigger-index", "test.vcf.gz")
    if not os.path.exists(path):
        pytest.skip("reference fixture not present on this machine")
    import gzip

    text = gzip.open(path).read()
    want = np.load(os.path.join(GOLDEN, "bigger_index_cols.npz"))["pos"]
    got = np.concatenate([b["pos"] for b in oracle.read_batches(text)])
    assert got.tolist() == want.tolist()
    assert oracle.filter_count(text, "chr1", 1000000, 2000000) == (37232, 99904)


def test_synthetic_truth_matches_oracle():
    from synth import vcf

    cols = vcf.columns(200_000)
    shards = vcf.shards(cols, 8)
    for chrom, lo, hi in [("1", 1000000, 2000000), ("X", None, None), ("22", 5, 30000000), (None, 1, 1000000), (None, None, None)]:
        want = cols.truth_count(chrom, lo, hi)
        got, rows, _ = oracle.filter_count_files(shards, chrom, lo, hi, target_partitions=4)
        assert (got, rows) == (want, cols.n), (chrom, lo, hi)
