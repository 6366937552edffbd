"""Builds exon_b200/libexon_gpu.so (CUDA kernels + C ABI) for sm_100a with nvcc.  In-tree, so the library
travels with a repo snapshot; `python -m exon_b200.build` or `__graft_entry__.build()`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libexon_gpu.so")
SOURCES = ["abi.cu", "vcf_stream.cu", "vcf_scan.cu", "vcf_columns.cu", "vcf_wide.cu", "filter_agg.cu", "fastq_scan.cu", "bgzf.cu", "inflate.cu", "bam.cu", "mzml.cu", "mzml_columns.cu", "fasta_scan.cu", "fasta_columns.cu", "gff_scan.cu", "gff_columns.cu", "nccl.cu", "host_logic.cpp", "tabix.cpp", "arrow_stream.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr", "-cudart", "static"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "exon_gpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        build_host(verbose=verbose)
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out, file=sys.stderr)
    link = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", LIB, *objs, "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    build_host(force=True, verbose=verbose)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_LIB = os.path.join(HOST_DIR, "libexon_host.so")


def build_host(force: bool = False, verbose: bool = False) -> str:
    """exon_b200/host/libexon_host.so: the C++ mirror of the reference's host-side operators (g++, links libexon_gpu)."""
    srcs = [os.path.join(HOST_DIR, "exon_host.cpp")]
    deps = srcs + [os.path.join(HOST_DIR, "exon_host.hpp"), os.path.join(HERE, "..", "include", "exon_gpu.h"), LIB]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_LIB) for d in deps):
        return HOST_LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", HOST_LIB, *srcs, "-L" + HERE, "-lexon_gpu",
           "-lpthread", "-Wl,-rpath,$ORIGIN/.."]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed on exon_host.cpp:\n{r.stdout}{r.stderr}")
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
