"""Columns 2..6 (id, ref, alt, qual, filter): the oracle restatement of LazyVCFArrayBuilder::append
(/root/reference/exon/exon-vcf/src/array_builder/lazy_array_builder.rs:169-216) against an independent pure-Python
split of the reference fixtures, and the QUAL parser (Rust f32::from_str semantics; exon_gpu_parse_f32 is the host
instance of the routine the column kernel runs) against exact rational arithmetic.  No reference test prints these
columns, so this is the only pin they have."""
import ctypes as C
import random
import struct
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np
import pytest

import oracle
from exon_b200 import _abi


def python_rows(text: bytes):
    """Independent statement: str.split on the fixture, field semantics straight from lazy_array_builder.rs:169-216."""
    out = {"id": [], "ref": [], "alt": [], "qual": [], "filter": []}
    for line in text.split(b"\n"):
        if not line or line.startswith(b"#"):
            continue
        f = line.split(b"\t")
        out["id"].append(None if f[2] in (b".", b"") else f[2].split(b";"))
        out["ref"].append(f[3])
        out["alt"].append(None if f[4] in (b".", b"") else [])
        out["qual"].append(None if f[5] == b"." else exact_f32_bits(f[5].decode()))
        out["filter"].append([] if f[6] in (b".", b"") else f[6].split(b";"))
    return out


def _bits_value(b: int) -> Fraction:
    if b < 0x00800000:
        return Fraction(b, 2 ** 149)
    return Fraction((b & 0x7FFFFF) | 0x800000) * Fraction(2) ** ((b >> 23) - 150)


def exact_f32_bits(s: str) -> int:
    """Round-to-nearest-even f32 bit pattern of a decimal literal, by exact rational arithmetic."""
    t = s.lower()
    neg = t.startswith("-")
    t = t.lstrip("+-")
    mant, ex = (t.split("e") + ["0"])[:2]
    ip, fp = (mant.split(".") + [""])[:2]
    v = Fraction(int((ip + fp) or "0")) * Fraction(10) ** (int(ex) - len(fp))

    def val(b):
        return Fraction(2) ** 128 if b == 0x7F800000 else _bits_value(b)

    lo, hi = 0, 0x7F800000
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if val(mid) <= v:
            lo = mid
        else:
            hi = mid
    if v >= val(0x7F800000):
        b = 0x7F800000
    else:
        a, c = val(lo), val(lo + 1)
        b = lo if v - a < c - v else lo + 1 if v - a > c - v else (lo if lo % 2 == 0 else lo + 1)
    return b | (0x80000000 if neg else 0)


def lib_f32_bits(s: bytes):
    out = C.c_float()
    rc = _abi.load().exon_gpu_parse_f32(s, len(s), C.byref(out))
    return ("err", rc) if rc else struct.unpack("<I", struct.pack("<f", out.value))[0]


@pytest.mark.parametrize("name", ["index_vcf", "biobear_vcf", "common_all_vcf"])
def test_oracle_wide_columns_match_python_split(name, request):
    text = request.getfixturevalue(name)
    got, want = oracle.vcf_wide_rows(text), python_rows(text)
    assert got == want
    assert len(got["ref"]) > 0


def test_oracle_wide_known_rows(biobear_vcf, index_vcf):
    w = oracle.vcf_wide_rows(biobear_vcf)
    assert w["id"][2] == [b"id3D"] and w["id"][0] is None
    assert w["alt"][9] is None and w["alt"][0] == []  # ALT "." -> NULL, anything else -> []
    assert w["qual"][13] is None and w["filter"][2] == [b"q10"]
    w = oracle.vcf_wide_rows(index_vcf)
    assert w["filter"][0] == [] and w["qual"][0] == 0 and w["ref"][:3] == [b"G", b"T", b"A"]  # FILTER "." -> valid empty list


def test_parse_f32_exact():
    rng = random.Random(20241017)
    cases = ["0", "1", "100", "99", "0.5", "1e10", "1e-10", "16777217", "16777216", "33554433", "3.4028235e38",
             "3.4028235677973366e38", "3.4028235677973367e38", "1e39", "1e-45", "7e-46", "7.006492321624085e-46",
             "7.006492321624086e-46", "1.17549435e-38", "0.1", "0.30000001192092896", "123456789012345678901234567890",
             "1.", "+.5", "-0", "-1.5e3", "1E5", "00012.500", "000", "1e-100", "1e100", "9007199254740993", "8388608.5",
             "8388609.5", "8388608.50000000000000000001", "1.00000017881393432617187500", "1.00000017881393432617187501",
             "1.0000001788139343261718749999", "29.9999", "62.8"]
    for _ in range(400):
        b = rng.randrange(0, 0x7F7FFFFF)
        mid = (_bits_value(b) + _bits_value(b + 1)) / 2
        for prec in (36, rng.choice([5, 9, 12, 17, 20, 30])):
            getcontext().prec = prec
            cases.append(format(Decimal(mid.numerator) / Decimal(mid.denominator), "e"))
    for _ in range(400):
        cases.append("%d.%0*d" % (rng.randrange(0, 10 ** rng.randrange(1, 9)), rng.randrange(1, 8), rng.randrange(0, 10 ** 6)))
        cases.append("%de%d" % (rng.randrange(1, 10 ** rng.randrange(1, 20)), rng.randrange(-50, 40)))
    for s in cases:
        assert lib_f32_bits(s.encode()) == exact_f32_bits(s), s
    # glibc strtof (what the oracle uses after checking the grammar) agrees on the same literals
    text = ("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n" + "".join(f"1\t{i + 1}\t.\tA\tC\t{s}\t.\t.\n" for i, s in enumerate(cases))).encode()
    assert oracle.vcf_wide_rows(text)["qual"] == [exact_f32_bits(s) for s in cases]


def test_parse_f32_grammar():
    for s in [b"", b".", b"+", b"e5", b"1e", b"1e+", b"1..2", b"1_0", b" 1", b"1 ", b"0x10", b"infinit", b"nan(1)", b"1f"]:
        got = lib_f32_bits(s)
        assert got == ("err", _abi.ERR_PARSE), s
    assert lib_f32_bits(b"inf") == 0x7F800000 and lib_f32_bits(b"-Infinity") == 0xFF800000 and lib_f32_bits(b"+INF") == 0x7F800000
    assert np.isnan(np.array([lib_f32_bits(b"NaN")], np.uint32).view(np.float32)[0])
    assert lib_f32_bits(b"1" * 37) == ("err", _abi.ERR_UNSUPPORTED)  # 37 significant digits: refused, not approximated
    assert lib_f32_bits(b"1" + b"0" * 40) == 0x7F800000  # trailing zeros are not significant digits


def test_info_oracle_goldens(index_vcf, biobear_vcf):
    # slt/vcf-select-tests.slt:6-10: SELECT info FROM vcf_table LIMIT 2
    got = oracle.vcf_info_strings(index_vcf)
    assert got[:2] == [b"DP=1;I16=1,0,0,0,26,676,0,0,60,3600,0,0,0,0,0,0;QS=1,0;MQ0F=0", b"DP=1;I16=1,0,0,0,34,1156,0,0,60,3600,0,0,1,1,0,0;QS=1,0;MQ0F=0"]
    assert len(got) == 621
    assert oracle.vcf_info_strings(biobear_vcf)[2] == b"DP4=1,2,3,4;AN=4;AC=2;INDEL=true;STR=test"   # a flag is printed as key=true
    hdr = b"##fileformat=VCFv4.2\n##INFO=<ID=AF,Number=A,Type=Float,Description=\"x, y\">\n##INFO=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    rows = b"1\t5\t.\tA\tC\t.\t.\tAF=0.50,1e-3,.;DP=007\n1\t6\t.\tA\tC\t.\t.\t.\n"
    assert oracle.vcf_info_strings(hdr + rows) == [b"AF=0.5,0.001,.;DP=7", b""]   # numbers go through Display; '.' info is ""
