#!/usr/bin/env python
"""bench.py -- VCF region filter (chrom='1' AND pos BETWEEN 1000000 AND 2000000) + COUNT over 100M synthetic
variants per GPU (BASELINE.json configs[2]); one rank per GPU, weak scaling, one all-reduce of the int64 partial per
step (exchanged over peer memory / NVLink by the library, NCCL as its fallback).

    python bench.py --gpus 1 --steps 50 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU restatement of the reference path on the host cores

A step = one pass of the hot path over the whole resident workload: fused scan->filter->COUNT kernel over every
shard body in HBM, (N>1: all-reduce,) count read back to the host.  `value` = rows all ranks processed / time.
`e2e` = the same through exon_gpu_vcf_feed with HOST (pinned) buffers: H2D of every shard inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

QUERY = ("1", 1_000_000, 2_000_000)
METRIC = "vcf_region_filter_count_rows_per_sec"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="variants per GPU (BASELINE config: 100M)")
    ap.add_argument("--shards", type=int, default=64)
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (exon_gpu_vcf_opts.kernel_variant)")
    ap.add_argument("--strict", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time strict mode, COUNT(*) and the column build")
    return ap.parse_args()


# ---- clocks ---------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock and clock-event reasons of one GPU through NVML while a timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.004):
        self.samples, self.reason_bits, self.period = [], 0, period_s
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            # CUDA_VISIBLE_DEVICES remaps CUDA ordinals; NVML does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report it, never fake a clock
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        s = sorted(self.samples)
        bits = self.reason_bits & ~0x1  # gpu_idle between regions is not a throttle
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": [n for b, n in self.REASONS.items() if bits & b]}


# ---- CPU arm (oracle port of the reference path) -------------------------------------------------------------

def cpu_rows_per_sec(files, rows_per_file, cores, target_s, steps=1, warmup=0):
    """Times the oracle (restated reference CPU path: one worker per file partition, 8192-row batches, column
    build, predicate, count) on a bounded sample of the shard files.  Returns (rows/s, sample description,
    seconds per step)."""
    import oracle

    t0 = time.perf_counter()
    oracle.filter_count_files(files[:1], *QUERY, target_partitions=1)
    t1 = max(time.perf_counter() - t0, 1e-4)  # one file, one core
    n = int(max(cores, min(len(files), round(target_s * cores / t1))))
    n = min(len(files), (n // cores) * cores if n >= cores else n)
    sample = files[:n]
    for _ in range(warmup):
        oracle.filter_count_files(sample, *QUERY, target_partitions=cores)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle.filter_count_files(sample, *QUERY, target_partitions=cores)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    rows = sum(rows_per_file[:n])
    return rows / dt, f"{n} of {len(files)} shard files ({rows} rows), {cores} worker threads, page-cache resident", dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from synth import vcf

    cores = os.cpu_count() or 1
    cols = vcf.columns(args.rows)
    bounds = vcf.shard_bounds(cols.n, args.shards)
    # only the sampled prefix of the shard list is ever touched: materialise a bounded number of files
    n_make = min(args.shards, max(cores, 16))
    sub = vcf.VcfColumns(cols.contig[: bounds[n_make - 1][1]], cols.pos[: bounds[n_make - 1][1]],
                         cols.ref[: bounds[n_make - 1][1]], cols.alt[: bounds[n_make - 1][1]],
                         cols.qual[: bounds[n_make - 1][1]], cols.contigs)
    files = []
    hdr = vcf.header_text(cols.contigs)
    import numpy as np

    for lo, hi in bounds[:n_make]:
        buf = np.empty(len(hdr) + (hi - lo) * vcf.MAX_LINE, dtype=np.uint8)
        buf[: len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
        w = vcf.format_rows(sub, lo, hi, buf[len(hdr):])
        files.append(buf[: len(hdr) + w])
    rows_per_file = [hi - lo for lo, hi in bounds[:n_make]]
    v, sample, dt = cpu_rows_per_sec(files, rows_per_file, cores, target_s=1.5, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is Rust (no cargo in this image): timed arm is oracle/vcf_oracle.c, the C restatement "
                    "of exon's VCFScan + LazyVCFArrayBuilder + FilterExec + COUNT, one worker per file partition"}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    return {"workload": f"VCF region filter chrom='1' AND pos BETWEEN 1000000 AND 2000000 + COUNT(*), "
                        f"{args.rows} synthetic variants per GPU in {args.shards} shard files (BASELINE configs[2])",
            "rows_per_gpu": args.rows, "shards_per_gpu": args.shards, "batch_rows": 8192,
            "parallelism": f"file-shard x{args.gpus}, one int64 all-reduce per step (peer-memory exchange over NVLink, NCCL fallback)" if args.gpus > 1 else "1 GPU",
            "l2": "input (~2.75 GB per GPU) is >20x the 126 MB L2; no flush needed between steps"}


# ---- GPU arm ------------------------------------------------------------------------------------------------

def run_b200(args):
    import numpy as np
    import torch

    from exon_b200 import _abi
    from exon_b200.runtime import Context
    from synth import vcf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: exon_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tstream = torch.cuda.Stream()
    ctx = Context(local, cuda_stream=tstream.cuda_stream)
    if world > 1:
        from exon_b200 import sharding

        sharding.init_final_aggregate(ctx, dist, rank, world)

    # ---- synthetic workload: this rank's file group (seed differs per rank), pinned on the host ----
    t_gen = time.perf_counter()
    cols = vcf.columns(args.rows, seed=vcf.SEED + rank)
    pins = []

    def alloc(nb):
        p = ctx.pinned(nb)
        pins.append(p)
        return p.array

    files = vcf.shards(cols, args.shards, alloc=alloc)
    truth = cols.truth_count(*QUERY)
    rows_per_file = [hi - lo for lo, hi in vcf.shard_bounds(cols.n, args.shards)]
    n_rows = cols.n
    del cols
    t_gen = time.perf_counter() - t_gen
    region = _abi.make_region(*QUERY)
    total_file_bytes = int(sum(f.size for f in files))

    # ---- resident copy in HBM, fed zero-copy (one run per shard file) ----
    dbufs = []
    resident = ctx.open_vcf(projection=(0, 1), kernel_variant=args.variant, strict=bool(args.strict))
    for f in files:
        d = ctx.device_buffer(f.size)
        d.upload(f)
        dbufs.append(d)
        resident.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
    body_bytes = resident.body_bytes()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if world > 1:
            return resident.filter_count_global(region)
        c = resident.filter_count(region)
        return c, c

    def timed(fn, steps, warmup, sampler=None):
        barrier()  # ranks finish generating / uploading their shards at different times
        for _ in range(warmup):
            fn()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        kms, out = [], None
        l0 = ctx.launch_count()
        if sampler:
            sampler.__enter__()
        ev0.record(tstream)
        for _ in range(steps):
            out = fn()
            kms.append(ctx.last_kernel_ms())
        ev1.record(tstream)
        barrier()
        if sampler:
            sampler.__exit__()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, kms, out, ctx.launch_count() - l0

    sampler = ClockSampler(local)
    ms_step, kms, (loc, glob), launches = timed(step, args.steps, max(args.warmup, 3), sampler)
    assert loc == truth, f"rank {rank}: GPU count {loc} != generator truth {truth}"
    truths = [truth]
    if dist is not None:
        t = torch.tensor([truth], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        truths = [int(t.item())]
    assert glob == truths[0], f"global count {glob} != sum of per-rank truths {truths[0]}"
    kernel_ms = float(np.mean(kms))
    value = world * n_rows / (ms_step * 1e-3)

    # ---- end to end through the C ABI with host buffers: H2D of every shard inside the timed region ----
    e2e_stream = ctx.open_vcf(projection=(0, 1), kernel_variant=args.variant, strict=bool(args.strict), pushdown=region)

    def e2e_step():
        e2e_stream.reset()
        for f in files:  # views of the pinned allocations, trimmed to each file's length
            e2e_stream.feed(f, is_last=True)
        if world > 1:
            return e2e_stream.filter_count_global(region)
        c = e2e_stream.filter_count(region)
        return c, c

    e2e_ms, _, (eloc, eglob), _ = timed(e2e_step, max(1, min(args.e2e_steps, args.steps)), 2)
    assert eloc == truth and eglob == glob
    e2e = {"value": world * n_rows / (e2e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": total_file_bytes,
           "d2h_bytes_per_step": 64, "ms_per_step": e2e_ms, "steps": max(1, min(args.e2e_steps, args.steps)),
           "api": "exon_gpu_vcf_feed(host pinned, per shard file) + exon_gpu_vcf_filter_count (pushdown declared)"}

    # ---- roofline of the dominant (only) kernel in the step ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    achieved = body_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("rows") == n_rows and tj.get("variant") == args.variant:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "vcf_scan_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": body_bytes, "bytes_per_row": body_bytes / n_rows,
                "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_step,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    line = {"metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": workload_config(args),
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": sampler.summary(),
            "count": glob, "count_matches_truth": True, "gen_seconds": t_gen,
            "kernel_variant": args.variant, "strict": int(args.strict)}

    if args.extra:
        extra = {}
        with ctx.open_vcf(projection=(0, 1), kernel_variant=args.variant, strict=True) as st:
            for d, f in zip(dbufs, files):
                st.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
            for name, rg in [("strict", region), ("count_star", None), ("interval_only", _abi.make_region(None, *QUERY[1:]))]:
                src = st if name == "strict" else resident
                ms, k, _, _ = timed(lambda: (src.filter_count(rg),) * 2, 10, 3)
                extra[name] = {"ms_per_step": ms, "kernel_ms": float(np.mean(k)),
                               "gbs": body_bytes / (float(np.mean(k)) * 1e-3) / 1e9}
        # K2: text -> Arrow {chrom, pos} batches left in device memory (second build = steady state: scratch areas
        # warm); K3: FilterExec + COUNT over all of those batches in one launch (exon_gpu_vcf_filter_agg)
        col_bytes = None
        for rep in range(2):
            with ctx.open_vcf(projection=(0, 1), columns_on_device=True) as st:
                for d, f in zip(dbufs, files):
                    st.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
                e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                torch.cuda.synchronize()
                l0 = ctx.launch_count()
                e0.record(tstream)
                first = st.next_batch()          # builds every column of the resident partition
                e1.record(tstream)
                torch.cuda.synchronize()
                k2_ms, k2_launches = e0.elapsed_time(e1), ctx.launch_count() - l0
                first.release()
                if rep == 0:
                    continue
                batches = -(-n_rows // 8192)
                col_bytes = 4 * (n_rows + batches) + 8 * n_rows + int(1.3 * n_rows)  # offsets + pos + ~chrom bytes
                extra["k2_columns"] = {"ms": k2_ms, "rows_per_s": n_rows / k2_ms * 1e3, "launches": k2_launches,
                                       "algorithmic_bytes": body_bytes + col_bytes,
                                       "algorithmic_gbs": (body_bytes + col_bytes) / k2_ms / 1e6,
                                       "frac_of_measured_peak": (body_bytes + col_bytes) / k2_ms / 1e6 / peak}
                def k3():
                    c, _, _ = st.filter_agg(chrom_col=0, pos_col=1, region=region)
                    return c, c
                ms, k, (cnt, _), nl = timed(k3, 20, 3)
                assert cnt == truth
                kk = float(np.mean(k))
                extra["k3_filter_count_columns"] = {"ms_per_step": ms, "kernel_ms": kk, "launches_per_step": nl / 20,
                                                    "rows_per_s": n_rows / ms * 1e3, "algorithmic_bytes": col_bytes,
                                                    "algorithmic_gbs": col_bytes / kk / 1e6,
                                                    "frac_of_measured_peak": col_bytes / kk / 1e6 / peak}
        # wide columns: text -> Arrow {id, ref, alt, qual, filter} (vcf_wide.cu), device resident, second build timed
        for rep in range(2):
            with ctx.open_vcf(projection=(2, 3, 4, 5, 6), columns_on_device=True) as st:
                for d, f in zip(dbufs, files):
                    st.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
                e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                torch.cuda.synchronize()
                l0 = ctx.launch_count()
                e0.record(tstream)
                first = st.next_batch()
                e1.record(tstream)
                torch.cuda.synchronize()
                w_ms, w_launches = e0.elapsed_time(e1), ctx.launch_count() - l0
                first.release()
                if rep == 1:
                    batches = -(-n_rows // 8192)
                    # id: validity + list offsets (all NULL here); ref: offsets + 1 B; alt: validity; qual: validity + f32;
                    # filter: list offsets + child offsets + "PASS"
                    out_bytes = 3 * n_rows // 8 + 4 * 3 * (n_rows + batches) + n_rows + 4 * n_rows + 4 * (n_rows + batches) + 4 * n_rows
                    extra["wide_columns_2_6"] = {"ms": w_ms, "rows_per_s": n_rows / w_ms * 1e3, "launches": w_launches,
                                                 "algorithmic_bytes": body_bytes + out_bytes,
                                                 "algorithmic_gbs": (body_bytes + out_bytes) / w_ms / 1e6,
                                                 "frac_of_measured_peak": (body_bytes + out_bytes) / w_ms / 1e6 / peak}
        line["extra"] = extra

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample, _ = cpu_rows_per_sec(files, rows_per_file, cores, target_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample}
    elif rank == 0:
        line["cpu_baseline"] = None

    resident.close()
    e2e_stream.close()
    for d in dbufs:
        d.free()
    for p in pins:
        p.free()
    ctx.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
