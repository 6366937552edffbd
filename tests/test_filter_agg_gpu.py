"""K3 (columnar filter + partial aggregate) and the evaluated region UDFs, through the C ABI, against numpy /
pyarrow.compute restatements of DataFusion's FilterExec + AggregateExec semantics and the reference's slt truth
tables (tests/golden/vcf_goldens.json <- slt/vcf-udfs.slt, region_physical_expr.rs, pos_interval_physical_expr.rs)."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from exon_b200 import _abi
from exon_b200._abi import ExonGpuError, make_region

pytestmark = pytest.mark.gpu


def export(struct_array: pa.StructArray):
    arr, sch = _abi.ArrowArray(), _abi.ArrowSchema()
    struct_array._export_to_c(C.addressof(arr), C.addressof(sch))
    return arr, sch


def make_batch(n, seed, null_frac=0.1):
    rng = np.random.default_rng(seed)
    names = np.array(["1", "2", "10", "11", "X", "chr1", "chrUn_KI270742v1"])
    chrom = names[rng.integers(0, len(names), n)]
    pos = rng.integers(1, 3_000_000, n).astype(np.int64)
    f64 = rng.lognormal(8, 2, n)
    f32 = rng.normal(30, 5, n).astype(np.float32)
    i32 = rng.integers(-1000, 1000, n).astype(np.int32)

    def nulls():
        return rng.random(n) < null_frac if null_frac else None

    cols = [pa.array(chrom, type=pa.utf8(), mask=nulls()), pa.array(pos, mask=nulls()), pa.array(f64, mask=nulls()),
            pa.array(f32, mask=nulls()), pa.array(i32, mask=nulls()), pa.array(pos * 3, mask=nulls())]
    return pa.StructArray.from_arrays(cols, names=["chrom", "pos", "f64", "f32", "i32", "i64"])


def expected(batch, chrom, lo, hi, kind, value_col):
    sel = pa.array(np.ones(len(batch), bool))
    if chrom is not None:
        sel = pc.and_kleene(sel, pc.equal(batch.field("chrom"), chrom))
    if lo is not None:
        sel = pc.and_kleene(sel, pc.and_kleene(pc.greater_equal(batch.field("pos"), lo), pc.less_equal(batch.field("pos"), hi)))
    kept = batch.filter(sel)  # FilterExec drops NULL and false
    if kind == _abi.AGG_COUNT_STAR:
        return len(kept), 0.0
    v = kept.field(value_col)
    cnt = len(v) - v.null_count
    s = pc.sum(pc.cast(v, pa.float64())).as_py() or 0.0
    return cnt, s


@pytest.mark.parametrize("null_frac", [0.0, 0.15])
@pytest.mark.parametrize("n", [0, 1, 31, 8192, 100_003])
def test_filter_agg_matches_arrow_semantics(gpu_ctx, n, null_frac):
    batch = make_batch(n, seed=n + 7, null_frac=null_frac)
    arr, sch = export(batch)
    for chrom, lo, hi in [("1", 1_000_000, 2_000_000), ("chrUn_KI270742v1", None, None), (None, 5, 1_500_000), (None, None, None), ("nope", 1, 10)]:
        rg = make_region(chrom, lo, hi)
        for kind, col in [(_abi.AGG_COUNT_STAR, None), (_abi.AGG_COUNT, "f64"), (_abi.AGG_SUM, "f64"), (_abi.AGG_AVG, "f32"),
                          (_abi.AGG_SUM, "i32"), (_abi.AGG_SUM, "i64")]:
            vi = batch.type.get_field_index(col) if col else -1
            cnt, si, sf = gpu_ctx.filter_agg(arr, sch, on_device=False, chrom_col=0, pos_col=1, region=rg, kind=kind, value_col=vi)
            want_cnt, want_sum = expected(batch, chrom, lo, hi, kind, col)
            assert cnt == want_cnt, (chrom, lo, hi, kind, col)  # bit-exact
            if kind in (_abi.AGG_SUM, _abi.AGG_AVG):
                if col in ("i32", "i64"):
                    assert si == int(want_sum)                   # integer sums are exact
                else:
                    assert sf == pytest.approx(want_sum, rel=1e-6, abs=1e-9)  # north_star tolerance for float aggregates


def test_sliced_batches(gpu_ctx):
    batch = make_batch(5000, seed=3)
    for off, ln in [(1, 100), (7, 4000), (4999, 1), (13, 0)]:
        sl = batch.slice(off, ln)
        arr, sch = export(sl)
        cnt, _, sf = gpu_ctx.filter_agg(arr, sch, on_device=False, chrom_col=0, pos_col=1, region=make_region("2", 1, 2_000_000),
                                        kind=_abi.AGG_SUM, value_col=2)
        want_cnt, want_sum = expected(sl, "2", 1, 2_000_000, _abi.AGG_SUM, "f64")
        assert cnt == want_cnt and sf == pytest.approx(want_sum, rel=1e-6, abs=1e-9)


def test_unfused_path_equals_fused(gpu_ctx):
    """K2 batches left on the device + K3 accumulate == K1 fused count == generator truth."""
    from synth import vcf

    cols = vcf.columns(400_000, seed=5)
    files = vcf.shards(cols, 5)
    acc = gpu_ctx.device_buffer(64)
    for q in [("1", 1_000_000, 2_000_000), ("X", None, None), (None, 1, 50_000_000), (None, None, None)]:
        rg = make_region(*q)
        with gpu_ctx.open_vcf(columns_on_device=True) as s:
            for f in files:
                s.feed(f, is_last=True)
            fused = s.filter_count(rg)
            gpu_ctx.memset(acc.ptr, 0, 64)
            rows = 0
            for b in s.batches():
                gpu_ctx.filter_agg_accumulate(b.c_array, b.c_schema, acc.ptr, chrom_col=0, pos_col=1,
                                              region=rg if rg is not None else None)
                rows += b.num_rows
                b.release()
            cnt, _, _ = gpu_ctx.partial_read(acc.ptr)
        assert rows == cols.n and cnt == fused == cols.truth_count(*q), q
    acc.free()


def _table(rows):
    return pa.StructArray.from_arrays([pa.array([r[0] for r in rows], type=pa.utf8()), pa.array([r[1] for r in rows], type=pa.int64())],
                                      names=["chrom", "positions"])


def test_region_udf_truth_tables(gpu_ctx, goldens):
    t = goldens["udf_truth_tables"]
    arr, sch = export(_table(t["rows"]))
    lib = _abi.load()

    def region(s):
        buf = C.create_string_buffer(256)
        r = _abi.Region()
        _abi.check(lib.exon_gpu_region_parse(s.encode(), buf, 256, C.byref(r)))
        r._keep = buf
        return r

    def interval(s):
        r = _abi.Region()
        _abi.check(lib.exon_gpu_interval_parse(s.encode(), C.byref(r)))
        return r

    v, valid = gpu_ctx.region_udf(_abi.UDF_REGION_MATCH, arr, sch, on_device=False, chrom_col=0, pos_col=1, region=region("1:1-1"))
    assert v.tolist() == t["region_match(chrom,pos,'1:1-1')"] and valid.all()
    v, _ = gpu_ctx.region_udf(_abi.UDF_INTERVAL_MATCH, arr, sch, on_device=False, pos_col=1, region=interval("1-1"))
    assert v.tolist() == t["interval_match(pos,'1-1')"]
    v, _ = gpu_ctx.region_udf(_abi.UDF_CHROM_MATCH, arr, sch, on_device=False, chrom_col=0, region=region("1"))
    assert v.tolist() == t["chrom_match(chrom,'1')"]
    p = goldens["physical_expr_vectors"]["region chr1:1-1"]
    arr2, sch2 = export(_table(p["rows"]))
    v, _ = gpu_ctx.region_udf(_abi.UDF_REGION_MATCH, arr2, sch2, on_device=False, chrom_col=0, pos_col=1, region=region("chr1:1-1"))
    assert v.tolist() == p["expect"]
    q = goldens["physical_expr_vectors"]["pos = 1"]
    arr3, sch3 = export(_table([["x", x] for x in q["pos"]]))
    v, _ = gpu_ctx.region_udf(_abi.UDF_INTERVAL_MATCH, arr3, sch3, on_device=False, pos_col=1, region=interval("1-1"))
    assert v.tolist() == q["expect"]
    # whole-contig region, open-ended region
    v, _ = gpu_ctx.region_udf(_abi.UDF_REGION_MATCH, arr, sch, on_device=False, chrom_col=0, pos_col=1, region=region("2"))
    assert v.tolist() == [False, False, False, True, True]
    v, _ = gpu_ctx.region_udf(_abi.UDF_REGION_MATCH, arr, sch, on_device=False, chrom_col=0, pos_col=1, region=region("2:3"))
    assert v.tolist() == [False, False, False, False, True]


def test_region_udf_null_and_zero_behaviour(gpu_ctx):
    """udfs/vcf/mod.rs: region_match errors on NULL operands and pos 0; chrom_match propagates NULL;
    interval_match maps NULL to false and errors on pos 0."""
    chrom = pa.array(["1", None, "1"], type=pa.utf8())
    pos = pa.array([5, 5, None], type=pa.int64())
    arr, sch = export(pa.StructArray.from_arrays([chrom, pos], names=["chrom", "pos"]))
    rg = make_region("1", 1, 10)
    with pytest.raises(ExonGpuError):
        gpu_ctx.region_udf(_abi.UDF_REGION_MATCH, arr, sch, on_device=False, chrom_col=0, pos_col=1, region=rg)
    v, valid = gpu_ctx.region_udf(_abi.UDF_CHROM_MATCH, arr, sch, on_device=False, chrom_col=0, region=rg)
    assert valid.tolist() == [True, False, True] and v[0] and v[2]
    v, valid = gpu_ctx.region_udf(_abi.UDF_INTERVAL_MATCH, arr, sch, on_device=False, pos_col=1, region=rg)
    assert v.tolist() == [True, True, False] and valid.all()
    arr0, sch0 = export(pa.StructArray.from_arrays([pa.array(["1"]), pa.array([0], type=pa.int64())], names=["chrom", "pos"]))
    for kind in (_abi.UDF_REGION_MATCH, _abi.UDF_INTERVAL_MATCH):
        with pytest.raises(ExonGpuError):
            gpu_ctx.region_udf(kind, arr0, sch0, on_device=False, chrom_col=0, pos_col=1, region=rg)
