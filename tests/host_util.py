"""ctypes access to exon_b200/host/libexon_host.so (the C++ mirror of the reference's host-side operators)."""
import ctypes as C
import gzip
import os
import shutil

from conftest import GOLDEN, ROOT

HOST_LIB = os.path.join(ROOT, "exon_b200", "host", "libexon_host.so")


def load_host():
    L = C.CDLL(HOST_LIB)
    L.exon_host_session_new.restype = C.c_void_p
    L.exon_host_session_new.argtypes = [C.c_int]
    L.exon_host_session_free.argtypes = [C.c_void_p]
    L.exon_host_last_error.restype = C.c_char_p
    L.exon_host_sql.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]
    L.exon_host_pushdown.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    L.exon_host_gpu_launches.restype = C.c_int64
    L.exon_host_gpu_launches.argtypes = [C.c_void_p]
    return L


def build_datasources(root):
    """Lay the committed fixtures out like exon/exon-core/test-data/datasources/."""
    os.makedirs(os.path.join(root, "vcf"))
    with gzip.open(os.path.join(GOLDEN, "index_plain.vcf.gz")) as f, open(os.path.join(root, "vcf", "index.vcf"), "wb") as o:
        o.write(f.read())
    shutil.copy(os.path.join(GOLDEN, "index.vcf.gz"), os.path.join(root, "vcf", "index.vcf.gz"))
    for s in ("1", "2"):
        d = os.path.join(root, "vcf-partition", f"sample={s}")
        os.makedirs(d)
        shutil.copy(os.path.join(GOLDEN, "index.vcf.gz"), os.path.join(d, "index.vcf.gz"))
    os.makedirs(os.path.join(root, "biobear-vcf"))
    shutil.copy(os.path.join(GOLDEN, "biobear_vcf_file.vcf.gz"), os.path.join(root, "biobear-vcf", "vcf_file.vcf.gz"))
    # tabix indexes next to the BGZF files (the indexed scans ask them for chunks)
    shutil.copy(os.path.join(GOLDEN, "index.vcf.gz.tbi"), os.path.join(root, "vcf", "index.vcf.gz.tbi"))
    for s in ("1", "2"):
        shutil.copy(os.path.join(GOLDEN, "index.vcf.gz.tbi"), os.path.join(root, "vcf-partition", f"sample={s}", "index.vcf.gz.tbi"))
    shutil.copy(os.path.join(GOLDEN, "biobear_vcf_file.vcf.gz.tbi"), os.path.join(root, "biobear-vcf", "vcf_file.vcf.gz.tbi"))
    # the other formats' fixtures
    os.makedirs(os.path.join(root, "fastq"))
    shutil.copy(os.path.join(GOLDEN, "test.fastq"), os.path.join(root, "fastq", "test.fastq"))
    shutil.copy(os.path.join(GOLDEN, "test_bgzip.fastq.gz"), os.path.join(root, "fastq", "test_bgzip.fastq.gz"))
    with open(os.path.join(GOLDEN, "test.fastq"), "rb") as f, gzip.open(os.path.join(root, "fastq", "test.fastq.gz"), "wb") as o:
        o.write(f.read())
    for s in ("1", "2"):
        d = os.path.join(root, "fastq-partition", f"sample={s}")
        os.makedirs(d)
        shutil.copy(os.path.join(GOLDEN, "test.fastq"), os.path.join(d, "test.fastq"))
        d = os.path.join(root, "bam-partition", f"sample={s}")
        os.makedirs(d)
        shutil.copy(os.path.join(GOLDEN, "test.bam"), os.path.join(d, "test.bam"))
    os.makedirs(os.path.join(root, "bam"))
    shutil.copy(os.path.join(GOLDEN, "test.bam"), os.path.join(root, "bam", "test.bam"))
    os.makedirs(os.path.join(root, "mzml"))
    shutil.copy(os.path.join(GOLDEN, "test.mzML"), os.path.join(root, "mzml", "test.mzML"))
    with open(os.path.join(GOLDEN, "test.mzML"), "rb") as f, gzip.open(os.path.join(root, "mzml", "test.mzML.gz"), "wb") as o:
        o.write(f.read())
    os.makedirs(os.path.join(root, "fasta"))
    shutil.copy(os.path.join(GOLDEN, "test.fasta"), os.path.join(root, "fasta", "test.fasta"))
    with open(os.path.join(GOLDEN, "test.fasta"), "rb") as f, gzip.open(os.path.join(root, "fasta", "test.fasta.gz"), "wb") as o:
        o.write(f.read())
    for s in ("1", "2"):
        d = os.path.join(root, "fasta-partition", f"sample={s}")
        os.makedirs(d)
        shutil.copy(os.path.join(GOLDEN, "test.fasta"), os.path.join(d, "test.fasta"))
    os.makedirs(os.path.join(root, "gff"))
    with gzip.open(os.path.join(GOLDEN, "test.gff.gz")) as f, open(os.path.join(root, "gff", "test.gff"), "wb") as o:
        o.write(f.read())
    shutil.copy(os.path.join(GOLDEN, "test.gff.gz"), os.path.join(root, "gff", "test.gff.gz"))
    os.makedirs(os.path.join(root, "two-vcf"))
    for n in ("a.vcf", "b.vcf"):
        shutil.copy(os.path.join(root, "vcf", "index.vcf"), os.path.join(root, "two-vcf", n))
    return root


def parse_slt(text):
    """[(kind, sql, expected_rows)] with kind in {'ok', 'error', 'query'}."""
    out, lines, i = [], text.splitlines(), 0
    while i < len(lines):
        ln = lines[i].strip()
        if not ln or ln.startswith("#"):
            i += 1
            continue
        if ln.startswith("statement"):
            kind = "ok" if ln.split()[1] == "ok" else "error"
            i += 1
            sql = []
            while i < len(lines) and lines[i].strip():
                sql.append(lines[i])
                i += 1
            out.append((kind, " ".join(sql), None))
        elif ln.startswith("query"):
            i += 1
            sql = []
            while lines[i].strip() != "----":
                sql.append(lines[i])
                i += 1
            i += 1
            rows = []
            while i < len(lines) and lines[i].strip():
                rows.append(lines[i].rstrip())
                i += 1
            out.append(("query", " ".join(sql), rows))
        else:
            raise ValueError(f"slt: cannot parse line {i + 1}: {ln}")
    return out
