"""Regenerates tests/golden/ from the reference's own test fixtures (run in the build container only:
/root/reference does not exist on the GPU box).

  python tests/golden/make_golden.py

What is committed
  index.vcf.gz, biobear_vcf_file.vcf.gz   byte-for-byte copies of the reference's small BGZF DATA fixtures
  index_plain.vcf.gz, common_all_head.vcf.gz   the reference's uncompressed fixtures vcf/index.vcf and
                      vcf-broad/00-common_all.head.vcf, gzip'ed (test inputs, not source)
  vcf_goldens.json    known answers: the values the reference's slt / unit tests assert, plus
                      answers derived here by an independent pure-Python line splitter (marked "derived")
  bigger_index_cols.npz  (chrom, pos) columns of bigger-index/test.vcf.gz (44 MB, too large to commit) parsed
                      by the same pure-Python splitter -- pins the oracle on very long lines (3202 samples)
"""
import gzip
import json
import os
import shutil

import numpy as np

REF = "/root/reference/exon/exon-core/test-data/datasources"
OUT = os.path.dirname(os.path.abspath(__file__))


def py_columns(text: bytes):
    """Independent restatement used only to derive goldens: str.split, int()."""
    chrom, pos = [], []
    for line in text.split(b"\n"):
        if not line or line.startswith(b"#"):
            continue
        f = line.split(b"\t", 2)
        chrom.append(f[0].decode())
        pos.append(int(f[1]))
    return chrom, np.array(pos, dtype=np.int64)


def count(chrom, pos, c=None, lo=None, hi=None):
    m = np.ones(len(pos), dtype=bool)
    if c is not None:
        m &= np.array([x == c for x in chrom], dtype=bool)
    if lo is not None:
        m &= pos >= lo
    if hi is not None:
        m &= pos <= hi
    return int(m.sum())


def main():
    copies = {
        "index.vcf.gz": f"{REF}/vcf/index.vcf.gz",
        "biobear_vcf_file.vcf.gz": f"{REF}/biobear-vcf/vcf_file.vcf.gz",
    }
    for dst, src in copies.items():
        shutil.copyfile(src, os.path.join(OUT, dst))
    for dst, src in {"common_all_head.vcf.gz": f"{REF}/vcf-broad/00-common_all.head.vcf",
                     "index_plain.vcf.gz": f"{REF}/vcf/index.vcf"}.items():
        # uncompressed reference fixtures, gzip'ed here only to keep the repository small
        with open(src, "rb") as f, open(os.path.join(OUT, dst), "wb") as raw:
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as g:
                g.write(f.read())

    idx = open(f"{REF}/vcf/index.vcf", "rb").read()
    ic, ip = py_columns(idx)
    # the reference's own .gz twin has a slightly different header but the same 621 records
    zc, zp = py_columns(gzip.open(f"{REF}/vcf/index.vcf.gz").read())
    assert zc == ic and (zp == ip).all()
    bio = gzip.open(f"{REF}/biobear-vcf/vcf_file.vcf.gz").read()
    bc, bp = py_columns(bio)
    broad = open(f"{REF}/vcf-broad/00-common_all.head.vcf", "rb").read()
    rc, rp = py_columns(broad)
    big = gzip.open(f"{REF}/bigger-index/test.vcf.gz").read()
    gc, gp = py_columns(big)
    assert set(gc) == {"chr1"}
    np.savez_compressed(os.path.join(OUT, "bigger_index_cols.npz"), pos=gp)

    goldens = {
        "index.vcf": {
            "reference_pinned": {
                "count_star": {"value": 621, "source": "exon/exon-core/tests/sqllogictests/slt/vcf-select-tests.slt:47-55"},
                "chrom_1": {"value": 191, "source": "exon/exon-core/tests/sqllogictests/slt/vcf-indexed-tests.slt:27-30; exon_context_ext.rs:1054-1090"},
                "chrom_1_two_copies": {"value": 382, "source": "slt/vcf-indexed-tests.slt:32-43"},
                "chrom_a": {"value": 0, "source": "slt/vcf-indexed-tests.slt:22-25"},
            },
            "derived": {
                "chrom_2": count(ic, ip, "2"), "chrom_10": count(ic, ip, "10"),
                "chrom_1_pos_9999919_10000000": count(ic, ip, "1", 9999919, 10000000),
                "chrom_1_pos_1000000_2000000": count(ic, ip, "1", 1000000, 2000000),
                "pos_ge_10000000": count(ic, ip, None, 10000000, None),
                "first_rows": [[ic[i], int(ip[i])] for i in range(5)],
                "pos_sum": int(ip.sum()), "chrom_bytes": sum(len(x) for x in ic),
            },
        },
        "biobear_vcf_file.vcf": {
            "reference_pinned": {
                "chrom_1": {"value": 11, "source": "slt/vcf-indexed-tests.slt:56-59"},
                "chrom_1000": {"value": 0, "source": "slt/vcf-indexed-tests.slt:51-54"},
            },
            "derived": {"count_star": len(bp), "rows": [[bc[i], int(bp[i])] for i in range(len(bp))]},
        },
        "common_all_head.vcf": {"derived": {"count_star": len(rp), "rows": [[rc[i], int(rp[i])] for i in range(len(rp))]}},
        "bigger-index/test.vcf": {
            "derived": {"count_star": len(gp), "chrom": "chr1", "text_bytes": len(big),
                        "chr1_pos_1000000_2000000": count(gc, gp, "chr1", 1000000, 2000000),
                        "pos_min": int(gp.min()), "pos_max": int(gp.max())},
        },
        "udf_truth_tables": {
            "source": "exon/exon-core/tests/sqllogictests/slt/vcf-udfs.slt:1-41",
            "rows": [["1", 1], ["1", 1], ["1", 2], ["2", 2], ["2", 3]],
            "region_match(chrom,pos,'1:1-1')": [True, True, False, False, False],
            "interval_match(pos,'1-1')": [True, True, False, False, False],
            "chrom_match(chrom,'1')": [True, True, True, False, False],
        },
        "physical_expr_vectors": {
            "region chr1:1-1": {"rows": [["chr1", 1], ["chr1", 2], ["chr2", 3]], "expect": [True, False, False],
                                "source": "exon/exon-core/src/physical_plan/region_physical_expr.rs:306-345"},
            "pos = 1": {"pos": [1, 2, 3], "expect": [True, False, False],
                        "source": "exon/exon-core/src/physical_plan/pos_interval_physical_expr.rs:277-316"},
        },
    }
    with open(os.path.join(OUT, "vcf_goldens.json"), "w") as f:
        json.dump(goldens, f, indent=1)
    print(json.dumps({k: v.get("derived", {}).get("count_star") for k, v in goldens.items() if isinstance(v, dict)}))


if __name__ == "__main__":
    main()
