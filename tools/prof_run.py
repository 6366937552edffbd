#!/usr/bin/env python
"""Short driver for ncu captures: uploads the bench workload once and launches the fused scan a few times.

    ncu --set full --clock-control none --import-source on -k regex:vcf_scan -c 6 -o gpurun_out/prof \
        python tools/prof_run.py [--rows N] [--variant V] [--modes lazy,lazy,strict,count_star,interval]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from exon_b200 import _abi  # noqa: E402
from exon_b200.runtime import Context  # noqa: E402
from synth import vcf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--mzml-spectra", type=int, default=0, help="profile the mzML path")
ap.add_argument("--bam-alignments", type=int, default=0, help="profile the BAM path")
ap.add_argument("--gz-rows", type=int, default=0, help="profile the BGZF inflate on this many VCF rows")
ap.add_argument("--fastq-reads", type=int, default=0, help="profile the FASTQ fused scan instead of VCF")
ap.add_argument("--shards", type=int, default=64)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--modes", default="lazy,lazy,strict,count_star,interval")
args = ap.parse_args()

if args.mzml_spectra:
    from synth import mzml

    sh = mzml.shards(args.mzml_spectra, 16)
    with Context(0) as ctx:
        s = ctx.open_mzml()
        keep = []
        for f in sh.files:
            d = ctx.device_buffer(f.size + 64)
            d.upload(np.ascontiguousarray(f))
            keep.append(d)
            s.feed(None, device_ptr=d.ptr, nbytes=f.size)
        for _ in range(3):
            print("mzml", s.filter_sum(sh.lo, sh.hi), sh.truth_count, f"{ctx.last_kernel_ms():.3f} ms", flush=True)
        s.close()
    raise SystemExit(0)
if args.bam_alignments:
    from synth import bam

    sh = bam.shards(args.bam_alignments, 16)
    with Context(0) as ctx:
        s = ctx.open_bam()
        for f in sh.files:
            s.feed(f)
        for _ in range(3):
            print("bam", s.count_by_reference(flag_exclude=0x904, min_mapq=30)[1], f"{ctx.last_kernel_ms():.3f} ms", flush=True)
        s.close()
    raise SystemExit(0)
if args.gz_rows:
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from bgzf_util import bgzf_compress

    cols = vcf.columns(args.gz_rows)
    files = vcf.shards(cols, 16)
    with ThreadPoolExecutor(16) as ex:
        gz = b"".join(ex.map(lambda f: bgzf_compress(f.tobytes(), 6), files))
    raw = sum(f.size for f in files)
    with Context(0) as ctx:
        d = ctx.device_buffer(raw + 64)
        a = np.frombuffer(gz, dtype=np.uint8)
        for _ in range(3):
            n = C.c_size_t()
            _abi.check(ctx.lib.exon_gpu_gzip_inflate(ctx.handle, C.c_void_p(a.ctypes.data), a.size, C.c_void_p(d.ptr), d.nbytes, 1, C.byref(n)))
            print("inflate", n.value, raw, f"{ctx.last_kernel_ms():.3f} ms", flush=True)
    raise SystemExit(0)
if args.fastq_reads:
    from synth import fastq

    sh = fastq.shards(args.fastq_reads, 32)
    with Context(0) as ctx:
        s = ctx.open_fastq()
        keep = []
        for f in sh.files:
            d = ctx.device_buffer(f.size)
            d.upload(np.ascontiguousarray(f))
            keep.append(d)
            s.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
        for _ in range(3):
            print("fastq", s.filter_count(30), sh.truth_count(30), f"{ctx.last_kernel_ms():.3f} ms", flush=True)
        s.close()
    raise SystemExit(0)
cols = vcf.columns(args.rows)
files = vcf.shards(cols, args.shards)
region = _abi.make_region("1", 1_000_000, 2_000_000)
with Context(0) as ctx:
    dbufs = []
    lazy = ctx.open_vcf(kernel_variant=args.variant)
    strict = ctx.open_vcf(kernel_variant=args.variant, strict=True)
    for f in files:
        d = ctx.device_buffer(f.size)
        d.upload(np.ascontiguousarray(f))
        dbufs.append(d)
        lazy.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
        strict.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
    for m in args.modes.split(","):
        if m == "lazy":
            c = lazy.filter_count(region)
        elif m == "strict":
            c = strict.filter_count(region)
        elif m == "count_star":
            c = lazy.filter_count(None)
        elif m == "interval":
            c = lazy.filter_count(_abi.make_region(None, 1_000_000, 2_000_000))
        elif m in ("columns", "k3"):
            with ctx.open_vcf(projection=(0, 1), columns_on_device=True) as st:
                for d, f in zip(dbufs, files):
                    st.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
                b = st.next_batch()  # K2: measure, scan, emit, offsets
                b.release()
                c = st.filter_agg(chrom_col=0, pos_col=1, region=region)[0] if m == "k3" else st.rows()
        elif m == "wide":
            with ctx.open_vcf(projection=(2, 3, 4, 5, 6), columns_on_device=True) as st:
                for d, f in zip(dbufs, files):
                    st.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
                b = st.next_batch()  # line index, measure, scans, emit (vcf_wide.cu)
                b.release()
                c = st.rows()
        else:
            raise SystemExit(m)
        print(m, c, f"{ctx.last_kernel_ms():.3f} ms", flush=True)
    lazy.close()
    strict.close()
