#!/usr/bin/env python
"""Secondary bench lines for the BASELINE configs that are parity-test cases rather than the headline
(configs[1] FASTQ mean-quality filter + COUNT, ...).  Same measurement rules as bench.py: inputs resident in HBM
for `value`, CUDA events on the library's stream, W >= 3 warm-ups, input >> L2; `e2e` feeds pinned host buffers
through the C ABI inside the timed region; the CPU arm is the oracle on all host cores.

    python tools/bench_formats.py fastq [--reads 10000000] [--steps 20]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from exon_b200.runtime import Context  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return (float(json.load(open(p))["hbm_gbs"]), "measured") if os.path.exists(p) else (6650.0, "fallback")


def timed(ctx, tstream, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    kms = []
    l0 = ctx.launch_count()
    e0.record(tstream)
    for _ in range(steps):
        out = fn()
        kms.append(ctx.last_kernel_ms())
    e1.record(tstream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(np.mean(kms)), out, (ctx.launch_count() - l0) / steps


def bench_fastq(args):
    import oracle
    from synth import fastq

    tstream = torch.cuda.Stream()
    ctx = Context(0, cuda_stream=tstream.cuda_stream)
    pins = []

    def alloc(nb):
        p = ctx.pinned(nb)
        pins.append(p)
        return p.array

    t0 = time.perf_counter()
    sh = fastq.shards(args.reads, args.shards, alloc=alloc)
    gen_s = time.perf_counter() - t0
    truth = sh.truth_count(30)
    total = int(sum(f.size for f in sh.files))
    dbufs = []
    res = ctx.open_fastq()
    for f in sh.files:
        d = ctx.device_buffer(f.size)
        d.upload(f)
        dbufs.append(d)
        res.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
    ms, kms, cnt, launches = timed(ctx, tstream, lambda: res.filter_count(30), args.steps, 3)
    assert cnt == truth, (cnt, truth)

    e2e_s = ctx.open_fastq()

    def e2e():
        e2e_s.reset()
        for f in sh.files:
            e2e_s.feed(f, is_last=True)
        return e2e_s.filter_count(30)

    e_ms, _, ecnt, _ = timed(ctx, tstream, e2e, max(2, args.steps // 4), 2)
    assert ecnt == truth
    cores = os.cpu_count() or 1
    n_cpu = max(1, min(len(sh.files), cores))
    t0 = time.perf_counter()
    c_cnt, c_rows = oracle.fastq_filter_count_files(sh.files[:n_cpu], 30, target_partitions=cores)
    cpu_s = time.perf_counter() - t0
    peak, src = peak_gbs()
    line = {"metric": "fastq_mean_quality_filter_count_reads_per_sec", "value": sh.n / ms * 1e3, "unit": "reads/s", "n_gpus": 1,
            "steps": args.steps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"FASTQ scan + WHERE mean(quality) > 30 + COUNT(*), {sh.n} synthetic {sh.read_len}-base reads in "
                                   f"{len(sh.files)} files (BASELINE configs[1])", "l2": "input >> 126 MB L2, no flush"},
            "e2e": {"value": sh.n / e_ms * 1e3, "unit": "reads/s", "h2d_bytes_per_step": total, "d2h_bytes_per_step": 24, "ms_per_step": e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "fq_lines_kernel + fq_filter_kernel (two passes)", "achieved": total / kms / 1e6,
                         "peak": peak, "peak_source": src, "unit": "GB/s", "frac": total / kms / 1e6 / peak,
                         "algorithmic_bytes_per_step": total, "kernel_ms": kms, "traffic": None},
            "cpu_baseline": {"value": c_rows / cpu_s, "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} of {len(sh.files)} files ({c_rows} reads), one worker per file"},
            "count": cnt, "count_matches_truth": True, "gen_seconds": gen_s}
    if args.gz:
        # the same reads as BGZF files (.fastq.gz is how FASTQ is stored; the reference's tests read gzip and bgzip twins,
        # slt/fastq-scan-test.slt:60-69): compressed bytes in pinned host memory -> H2D -> device inflate -> fused scan, next
        # to the CPU arm that has to inflate too (zlib + oracle, one worker per file)
        from concurrent.futures import ThreadPoolExecutor

        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from bgzf_util import bgzf_compress

        with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
            gz = list(ex.map(lambda f: bgzf_compress(f.tobytes(), args.level), sh.files))
        gpins = []
        for g in gz:
            p = ctx.pinned(len(g))
            p.array[:] = np.frombuffer(g, dtype=np.uint8)
            gpins.append(p)
        comp = int(sum(len(g) for g in gz))
        gz_s = ctx.open_fastq()

        def e2e_gz():
            gz_s.reset()
            for p in gpins:
                gz_s.feed_gzip(p.array)
            return gz_s.filter_count(30)

        g_ms, _, gcnt, _ = timed(ctx, tstream, e2e_gz, max(3, args.steps // 4), 2)
        assert gcnt == truth

        def cpu_one(g):
            return oracle.fastq_filter_count(oracle.gunzip_all(g), 30)

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            rs = list(ex.map(cpu_one, gz[:n_cpu]))
        gcpu_s = time.perf_counter() - t0
        line["bgzf_files"] = {"e2e": {"value": sh.n / g_ms * 1e3, "unit": "reads/s", "ms_per_step": g_ms, "h2d_bytes_per_step": comp, "d2h_bytes_per_step": 24},
                              "compressed_bytes": comp, "zlib_level": args.level,
                              "cpu_baseline": {"value": sum(r[1] for r in rs) / gcpu_s, "unit": "reads/s", "cores": cores, "kind": "port",
                                               "sample": f"{n_cpu} of {len(gz)} files: zlib inflate + oracle, one worker per file"}}
        gz_s.close()
        for p in gpins:
            p.free()
    res.close()
    e2e_s.close()
    for d in dbufs:
        d.free()
    for p in pins:
        p.free()
    ctx.close()
    print(json.dumps(line), flush=True)


def bench_vcfgz(args):
    """BGZF-compressed VCF shards: device inflate throughput, and end to end (compressed bytes in pinned host memory ->
    H2D -> device inflate -> fused scan) next to the CPU arm (zlib inflate + oracle, one worker per file)."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from exon_b200 import _abi
    from synth import vcf
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from bgzf_util import bgzf_compress

    tstream = torch.cuda.Stream()
    ctx = Context(0, cuda_stream=tstream.cuda_stream)
    cols = vcf.columns(args.rows)
    truth = cols.truth_count("1", 1_000_000, 2_000_000)
    files = vcf.shards(cols, args.shards)
    raw_bytes = int(sum(f.size for f in files))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        gz = list(ex.map(lambda f: bgzf_compress(f.tobytes(), args.level), files))
    comp_s = time.perf_counter() - t0
    pins = []
    for g in gz:
        p = ctx.pinned(len(g))
        p.array[:] = np.frombuffer(g, dtype=np.uint8)
        pins.append(p)
    comp_bytes = int(sum(len(g) for g in gz))
    region = _abi.make_region("1", 1_000_000, 2_000_000)
    st = ctx.open_vcf(pushdown=region)

    def e2e():
        st.reset()
        for p in pins:
            st.feed_gzip(p.array)
        return st.filter_count(region)

    e_ms, _, cnt, launches = timed(ctx, tstream, e2e, max(3, args.steps // 4), 2)
    assert cnt == truth, (cnt, truth)
    # inflate alone: every member of every shard in ONE launch (the concatenation of BGZF files is a BGZF file),
    # into a device buffer; kernel time from the library's events
    import ctypes as C
    allgz = ctx.pinned(comp_bytes)
    o = 0
    for g in gz:
        allgz.array[o:o + len(g)] = np.frombuffer(g, dtype=np.uint8)
        o += len(g)
    dbuf = ctx.device_buffer(raw_bytes + 64)
    kms = 0.0
    for rep in range(3):
        n = C.c_size_t()
        _abi.check(ctx.lib.exon_gpu_gzip_inflate(ctx.handle, C.c_void_p(allgz.ptr), allgz.nbytes, C.c_void_p(dbuf.ptr), dbuf.nbytes, 1, C.byref(n)))
        assert n.value == raw_bytes
        kms = ctx.last_kernel_ms()
    allgz.free()
    cores = os.cpu_count() or 1
    n_cpu = min(len(gz), cores)
    t0 = time.perf_counter()
    c_cnt, c_rows = oracle.filter_count_gz_files(gz[:n_cpu], "1", 1_000_000, 2_000_000, target_partitions=cores)
    cpu_s = time.perf_counter() - t0
    peak, src = peak_gbs()
    line = {"metric": "vcf_gz_region_filter_count_rows_per_sec", "value": cols.n / e_ms * 1e3, "unit": "rows/s", "n_gpus": 1,
            "ms_per_step": e_ms, "higher_is_better": True, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2] as {len(gz)} BGZF files (zlib level {args.level}): {cols.n} variants, "
                                   f"{raw_bytes} B text, {comp_bytes} B compressed; value == e2e (compressed bytes start in pinned host memory)"},
            "e2e": {"value": cols.n / e_ms * 1e3, "unit": "rows/s", "h2d_bytes_per_step": comp_bytes, "d2h_bytes_per_step": 64 * len(gz),
                    "ms_per_step": e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "latency (serial Huffman decode per member; see DESIGN.md)", "kernel": "inflate_decode_kernel + inflate_copy_kernel",
                         "achieved": (raw_bytes + comp_bytes) / kms / 1e6, "unit": "GB/s", "peak": peak, "peak_source": src,
                         "frac": (raw_bytes + comp_bytes) / kms / 1e6 / peak, "kernel_ms_total": kms,
                         "inflate_output_gbs": raw_bytes / kms / 1e6, "algorithmic_bytes": raw_bytes + comp_bytes},
            "cpu_baseline": {"value": c_rows / cpu_s, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} of {len(gz)} files ({c_rows} rows): zlib inflate + oracle, one worker per file"},
            "count": cnt, "count_matches_truth": True, "compress_seconds": comp_s}
    st.close()
    dbuf.free()
    for p in pins:
        p.free()
    ctx.close()
    print(json.dumps(line), flush=True)


def bench_bam(args):
    """BASELINE configs[3]: BAM flag + MAPQ filter + per-reference COUNT.  `value`: records resident (inflated) in HBM,
    one step = speculative walks + verification; `e2e`: BGZF bytes in pinned host memory -> H2D -> device inflate -> walks."""
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from synth import bam

    tstream = torch.cuda.Stream()
    ctx = Context(0, cuda_stream=tstream.cuda_stream)
    t0 = time.perf_counter()
    sh = bam.shards(args.alignments, args.shards, level=1)
    gen_s = time.perf_counter() - t0
    kw = dict(flag_exclude=0x904, min_mapq=30)
    truth = sh.truth(**kw)
    pins = []
    for f in sh.files:
        p = ctx.pinned(len(f))
        p.array[:] = np.frombuffer(f, dtype=np.uint8)
        pins.append(p)
    comp_bytes = int(sum(len(f) for f in sh.files))
    s = ctx.open_bam()
    for p in pins:
        s.feed(p.array)
    ms, kms, (got, rows), launches = timed(ctx, tstream, lambda: s.count_by_reference(**kw), args.steps, 3)
    assert got == truth and rows == sh.n

    def e2e():
        s.reset()
        for p in pins:
            s.feed(p.array)
        return s.count_by_reference(**kw)

    e_ms, _, (egot, _), _ = timed(ctx, tstream, e2e, max(3, args.steps // 4), 2)
    assert egot == truth
    cores = os.cpu_count() or 1
    n_cpu = min(len(sh.files), cores)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        parts = list(ex.map(lambda f: oracle.bam_count_by_reference_files([f], **kw), sh.files[:n_cpu]))
    cpu_s = time.perf_counter() - t0
    c_rows = sum(p[1] for p in parts)
    peak, src = peak_gbs()
    rec_bytes = sh.raw_bytes / sh.n
    line = {"metric": "bam_flag_mapq_filter_count_by_reference_alignments_per_sec", "value": sh.n / ms * 1e3, "unit": "alignments/s", "n_gpus": 1,
            "steps": args.steps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "dtype": "int64", "data": "synthetic",
            "config": {"workload": f"BAM (flag & 0x904) = 0 AND mapq >= 30 GROUP BY reference, {sh.n} synthetic alignments "
                                   f"(l_seq 100, 26 references) in {len(sh.files)} BGZF files = one GPU's share of BASELINE configs[3] "
                                   f"(200M over 8 GPUs); {sh.raw_bytes} B of records, {comp_bytes} B compressed",
                       "l2": "record stream >> 126 MB L2, no flush"},
            "e2e": {"value": sh.n / e_ms * 1e3, "unit": "alignments/s", "h2d_bytes_per_step": comp_bytes, "d2h_bytes_per_step": 8 * len(truth) + 128,
                    "ms_per_step": e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "bam_walk_kernel", "achieved": 36 * sh.n / kms / 1e6, "unit": "GB/s", "peak": peak,
                         "peak_source": src, "frac": 36 * sh.n / kms / 1e6 / peak, "kernel_ms": kms, "algorithmic_bytes_per_record": 36,
                         "full_record_gbs": sh.raw_bytes / kms / 1e6, "bytes_per_record": rec_bytes, "traffic": None},
            "cpu_baseline": {"value": c_rows / cpu_s, "unit": "alignments/s", "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} of {len(sh.files)} files ({c_rows} alignments): zlib inflate + record walk, one worker per file"},
            "count_matches_truth": True, "gen_seconds": gen_s}
    s.close()
    # column batches (exon_gpu_bam_next_batch), device resident: records already inflated in HBM -> Arrow columns; the
    # first next_batch() builds every batch of the partition
    cols = {}
    for name, proj in (("flag_reference_start_end_mapq", (1, 2, 3, 4, 5)), ("all_0_9", tuple(range(10)))):
        for rep in range(2):
            with ctx.open_bam(projection=proj, columns_on_device=True) as cs:
                for p in pins:
                    cs.feed(p.array)
                cs.count_by_reference(all_rows=True)   # inflate + verified walks: not part of the column build
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(tstream)
                if os.environ.get("EXON_BENCH_MEM"):
                    print("mem before next_batch", name, rep, [x >> 20 for x in torch.cuda.mem_get_info()], file=sys.stderr, flush=True)
                b = cs.next_batch()
                e1.record(tstream)
                torch.cuda.synchronize()
                b.release()
                if rep == 1:
                    cols[name] = {"ms": e0.elapsed_time(e1), "alignments_per_s": sh.n / e0.elapsed_time(e1) * 1e3}
    line["column_batches"] = cols
    for p in pins:
        p.free()
    ctx.close()
    print(json.dumps(line), flush=True)


def bench_mzml(args):
    """BASELINE configs[4]: mzML scan + m/z range filter + SUM(intensity)."""
    import math
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from synth import mzml

    tstream = torch.cuda.Stream()
    ctx = Context(0, cuda_stream=tstream.cuda_stream)
    pins = []

    def alloc(nb):
        p = ctx.pinned(nb)
        pins.append(p)
        return p.array

    t0 = time.perf_counter()
    sh = mzml.shards(args.spectra, args.shards, peaks=args.peaks, alloc=alloc)
    gen_s = time.perf_counter() - t0
    total = int(sum(f.size for f in sh.files))
    dbufs = []
    res = ctx.open_mzml()
    for f in sh.files:
        d = ctx.device_buffer(f.size + 64)
        d.upload(f)
        dbufs.append(d)
        res.feed(None, device_ptr=d.ptr, nbytes=f.size)
    ms, kms, (ssum, n_sel, n_sp), launches = timed(ctx, tstream, lambda: res.filter_sum(sh.lo, sh.hi), args.steps, 3)
    assert n_sp == sh.n and n_sel == sh.truth_count and math.isclose(ssum, sh.truth_sum, rel_tol=1e-6)
    e2e_s = ctx.open_mzml()

    def e2e():
        e2e_s.reset()
        for f in sh.files:
            e2e_s.feed(f)
        return e2e_s.filter_sum(sh.lo, sh.hi)

    e_ms, _, (esum, en, _), _ = timed(ctx, tstream, e2e, max(3, args.steps // 4), 2)
    assert en == n_sel
    cores = os.cpu_count() or 1
    n_cpu = min(len(sh.files), cores)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        parts = list(ex.map(lambda f: oracle.mzml_scan(f, sh.lo, sh.hi).n_spectra, sh.files[:n_cpu]))
    cpu_s = time.perf_counter() - t0
    peak, src = peak_gbs()
    col_bytes = 16 * sh.peaks * sh.n
    line = {"metric": "mzml_mz_range_filter_sum_intensity_spectra_per_sec", "value": sh.n / ms * 1e3, "unit": "spectra/s", "n_gpus": 1,
            "steps": args.steps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"mzML m/z BETWEEN {sh.lo} AND {sh.hi} + SUM(intensity), {sh.n} synthetic spectra x {sh.peaks} peaks "
                                   f"(f64, base64, uncompressed) in {len(sh.files)} files (BASELINE configs[4]); {total} B of text",
                       "l2": "text >> 126 MB L2, no flush", "tolerance": "sum within 1e-6 relative of the oracle / generator truth; selected-peak count exact"},
            "e2e": {"value": sh.n / e_ms * 1e3, "unit": "spectra/s", "h2d_bytes_per_step": total, "d2h_bytes_per_step": 192, "ms_per_step": e_ms},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "mzml_events_kernel + sort + mzml_spectra_kernel + mzml_sum_kernel", "achieved": total / kms / 1e6,
                         "unit": "GB/s", "peak": peak, "peak_source": src, "frac": total / kms / 1e6 / peak, "kernel_ms": kms,
                         "algorithmic_bytes_per_step": total, "columnar_bytes_per_step": col_bytes, "traffic": None},
            "cpu_baseline": {"value": sum(parts) / cpu_s, "unit": "spectra/s", "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} of {len(sh.files)} files ({sum(parts)} spectra): tag scan + base64 + f64 sum, one worker per file"},
            "sum": ssum, "selected_peaks": n_sel, "sum_matches_truth_1e-6": True, "gen_seconds": gen_s}
    res.close()
    e2e_s.close()
    for d in dbufs:
        d.free()
    for p in pins:
        p.free()
    ctx.close()
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("fmt", choices=["fastq", "vcfgz", "bam", "mzml"])
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--alignments", type=int, default=25_000_000)
    ap.add_argument("--spectra", type=int, default=1_000_000)
    ap.add_argument("--peaks", type=int, default=200)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--shards", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--gz", action="store_true", help="fastq: also the BGZF-compressed variant of the workload")
    a = ap.parse_args()
    {"fastq": bench_fastq, "vcfgz": bench_vcfgz, "bam": bench_bam, "mzml": bench_mzml}[a.fmt](a)
