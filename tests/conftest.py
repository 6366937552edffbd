import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DATA = "/root/reference/exon/exon-core/test-data/datasources"  # build container only


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(GOLDEN, "vcf_goldens.json")) as f:
        return json.load(f)


def golden_text(name: str) -> bytes:
    with gzip.open(os.path.join(GOLDEN, name)) as f:
        return f.read()


@pytest.fixture(scope="session")
def index_vcf() -> bytes:
    return golden_text("index_plain.vcf.gz")


@pytest.fixture(scope="session")
def index_vcf_gz_twin() -> bytes:
    return golden_text("index.vcf.gz")


@pytest.fixture(scope="session")
def biobear_vcf() -> bytes:
    return golden_text("biobear_vcf_file.vcf.gz")


@pytest.fixture(scope="session")
def common_all_vcf() -> bytes:
    return golden_text("common_all_head.vcf.gz")


@pytest.fixture(scope="session")
def gpu_ctx():
    from exon_b200.runtime import Context

    ctx = Context(0)
    yield ctx
    ctx.close()


def make_vcf(rows, header=True, trailing_newline=True, extra_cols="\t.\tA\tC\t50\tPASS\t."):
    """Small hand-made VCF text: rows = [(chrom, pos_text), ...]."""
    out = []
    if header:
        out.append("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
    body = "\n".join(f"{c}\t{p}{extra_cols}" for c, p in rows)
    out.append(body)
    if rows and trailing_newline:
        out.append("\n")
    return "".join(out).encode()
