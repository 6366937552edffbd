"""The C-ABI library loads without a GPU and exports every symbol include/exon_gpu.h declares; host-only
entry points (region parsing, file regrouping) behave like the oracle; compute entry points fail loudly."""
import ctypes as C
import os
import re

import pytest

import oracle
from conftest import ROOT, has_gpu
from exon_b200 import _abi


def header_functions():
    text = open(os.path.join(ROOT, "include", "exon_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(exon_gpu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _abi.load()
    declared = header_functions()
    assert declared, "no functions parsed from include/exon_gpu.h"
    for name in declared:
        assert hasattr(lib, name), f"libexon_gpu.so does not export {name}"
    assert sorted(_abi.SYMBOLS) == declared
    assert lib.exon_gpu_version().decode().endswith("sm_100a")


def test_struct_layouts_match_header():
    # sizes a bindgen-generated Rust struct would have (x86-64 SysV)
    assert C.sizeof(_abi.Region) == 40
    assert C.sizeof(_abi.VcfOpts) == 40
    assert C.sizeof(_abi.Partial) == 24
    assert C.sizeof(_abi.ArrowArray) == 80 and C.sizeof(_abi.ArrowSchema) == 72


@pytest.mark.parametrize("s", ["1", "1:9999921", "chr1:1-3388930", "1:1-1", "HLA-A*01:01", "chrUn:5-", "x:-7", "a:b", "1:0-5", "chr1:", "a:b:", ":100-200", "a::5"])
def test_region_parse_matches_oracle(s):
    lib = _abi.load()
    buf = C.create_string_buffer(256)
    r = _abi.Region()
    assert lib.exon_gpu_region_parse(s.encode(), buf, 256, C.byref(r)) == 0
    o = oracle.parse_region(s)
    assert (buf.value, r.has_interval, r.lo, r.hi) == (o.name[: o.name_len], o.has_interval, o.lo, o.hi)
    assert r.has_chrom == 1 and r.chrom_len == o.name_len


def test_region_parse_empty_suffix_and_name():
    # noodles-core 0.15 Region::from_str: rsplit_once(':') + Interval::from_str("") == unbounded
    lib = _abi.load()
    buf = C.create_string_buffer(256)
    r = _abi.Region()
    assert lib.exon_gpu_region_parse(b"chr1:", buf, 256, C.byref(r)) == 0
    assert (buf.value, r.has_interval, r.lo, r.hi) == (b"chr1", 1, 1, _abi.INT64_MAX)
    assert lib.exon_gpu_region_parse(b":100-200", buf, 256, C.byref(r)) == 0
    assert (buf.value, r.chrom_len, r.has_interval, r.lo, r.hi) == (b"", 0, 1, 100, 200)
    assert lib.exon_gpu_region_parse(b"a:b:", buf, 256, C.byref(r)) == 0 and buf.value == b"a:b" and r.has_interval == 1


def test_interval_parse_and_errors():
    lib = _abi.load()
    r = _abi.Region()
    assert lib.exon_gpu_interval_parse(b"1-1", C.byref(r)) == 0 and (r.has_chrom, r.lo, r.hi) == (0, 1, 1)
    assert lib.exon_gpu_interval_parse(b"5-", C.byref(r)) == 0 and (r.lo, r.hi) == (5, _abi.INT64_MAX)
    assert lib.exon_gpu_interval_parse(b"x", C.byref(r)) == _abi.ERR_ARG
    assert b"interval" in lib.exon_gpu_last_error()
    assert lib.exon_gpu_region_parse(b"", C.create_string_buffer(8), 8, C.byref(r)) == _abi.ERR_ARG


@pytest.mark.parametrize("sizes,target", [([30, 10, 20, 40], 2), ([5, 5, 5], 8), ([7], 4), ([9, 1, 8, 2, 7, 3, 6, 4, 5], 4), ([], 3)])
def test_regroup_files_matches_oracle(sizes, target):
    lib = _abi.load()
    n = len(sizes)
    out = (C.c_int32 * max(n, 1))()
    parts = C.c_int32()
    assert lib.exon_gpu_regroup_files_by_size((C.c_int64 * max(n, 1))(*sizes), n, target, out, C.byref(parts)) == 0
    want_parts, want = oracle.regroup_files_by_size(sizes, target) if n else (0, [])
    assert parts.value == want_parts and list(out)[:n] == want


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    lib = _abi.load()
    h = C.c_void_p()
    assert lib.exon_gpu_ctx_create(0, None, C.byref(h)) == _abi.ERR_CUDA
    assert b"no CPU fallback" in lib.exon_gpu_last_error()
    with pytest.raises(_abi.ExonGpuError):
        from exon_b200.runtime import Context

        Context(0)
