#!/usr/bin/env python
"""bench.py -- VCF region filter (chrom='1' AND pos BETWEEN 1000000 AND 2000000) + COUNT over ONE set of 100M
synthetic variants in 64 shard files (BASELINE.json configs[2]); one rank per GPU.  Default = strong scaling: the files
are assigned to the ranks by the reference's own file -> partition rule (regroup_files_by_size) and the int64 partials
are exchanged in the scan kernel's tail over peer memory / NVLink (NCCL as fallback); the weak series (every rank scans
the whole set) is reported next to it under `other_series`.

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
    ap.add_argument("--rows", type=int, default=100_000_000, help="variants in the file set (BASELINE config: 100M)")
    ap.add_argument("--shards", type=int, default=64)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): ONE --rows file set, files assigned to ranks by the reference's regroup_files_by_size "
                         "rule; weak: every rank scans the whole set. The other series is reported under its own key.")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (exon_gpu_vcf_opts.kernel_variant)")
    ap.add_argument("--strict", type=int, default=1,
                    help="1 (default, what INTEGRATION.md opens the stream with): every row's CHROM/POS is validated like the "
                         "reference's builder does; 0: only rows whose CHROM matches are validated (same counts on valid input)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time COUNT(*), interval-only and the column builds")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the FASTQ / BAM / mzML sub-lines (tools/bench_formats.py in subprocesses, N = 1 only, about a minute)")
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

    if target_s is None:
        n = len(files)
    else:
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


def make_files(args, n_files=None, alloc=None):
    """The synthetic file set of the workload (same on every rank): (columns, files, rows per file)."""
    from synth import vcf

    cols = vcf.columns(args.rows)
    bounds = vcf.shard_bounds(cols.n, args.shards)
    files = vcf.shards(cols, args.shards, alloc=alloc)
    return cols, files, bounds


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    cols, files, bounds = make_files(args)
    rows_per_file = [hi - lo for lo, hi in bounds]
    # the whole file set, every step: one worker per file partition on all host cores (the reference's own parallelism)
    v, sample, dt = cpu_rows_per_sec(files, rows_per_file, cores, target_s=None, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is Rust (no cargo in this image): timed arm is oracle/vcf_oracle.c, the C restatement "
                    "of exon's VCFScan + LazyVCFArrayBuilder + FilterExec + COUNT, one worker per file partition"}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args):
    n = args.gpus
    return {"workload": f"VCF region filter chrom='1' AND pos BETWEEN 1000000 AND 2000000 + COUNT(*), "
                        f"{args.rows} synthetic variants in {args.shards} shard files (BASELINE configs[2])",
            "rows": args.rows, "shards": args.shards, "batch_rows": 8192,
            "parallelism": (f"{args.scaling} scaling: the {args.shards} files are assigned to {n} ranks by regroup_files_by_size "
                            f"(exon_file_scan_config.rs:79-110); the int64 partials are exchanged over peer memory (NVLink) in "
                            f"the scan kernel's tail, NCCL as fallback") if n > 1 else "1 GPU",
            "l2": "input per GPU (2.75 GB / N ranks) stays > 2.7x the 126 MB L2 at N = 8; no flush needed between steps"}


# ---- GPU arm ------------------------------------------------------------------------------------------------

def run_b200(args):
    import numpy as np
    import torch

    from exon_b200 import _abi, sharding
    from exon_b200.runtime import Context

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
        sharding.init_final_aggregate(ctx, dist, rank, world)

    # ---- the synthetic file set (identical on every rank) and this rank's share of it ----
    t_gen = time.perf_counter()
    cols, files, bounds = make_files(args)
    rows_per_file = [hi - lo for lo, hi in bounds]
    sizes = [int(f.size) for f in files]
    mine = sharding.files_of_rank(sizes, rank, world)  # regroup_files_by_size: the reference's file -> partition rule
    chrom_idx = [c for c, _ in cols.contigs].index(QUERY[0])
    hit = (cols.contig == chrom_idx) & (cols.pos >= QUERY[1]) & (cols.pos <= QUERY[2])
    truth_file = [int(hit[lo:hi].sum()) for lo, hi in bounds]
    truth_all, truth_mine = int(sum(truth_file)), int(sum(truth_file[i] for i in mine))
    n_rows_all, n_rows_mine = cols.n, int(sum(rows_per_file[i] for i in mine))
    del cols, hit
    t_gen = time.perf_counter() - t_gen
    region = _abi.make_region(*QUERY)

    # ---- resident copies in HBM, fed zero-copy (one run per shard file): the whole set, and this rank's share ----
    dbufs = []
    for f in files:
        d = ctx.device_buffer(f.size)
        d.upload(f)
        dbufs.append(d)

    def open_resident(idx, strict):
        st = ctx.open_vcf(projection=(0, 1), kernel_variant=args.variant, strict=bool(strict))
        for i in idx:
            st.feed(None, device_ptr=dbufs[i].ptr, nbytes=files[i].size, is_last=True)
        return st

    everything = list(range(len(files)))
    strong = args.scaling == "strong"
    head_idx = mine if strong else everything
    head = open_resident(head_idx, args.strict)
    head_bytes = head.body_bytes()
    head_rows_total = n_rows_all if strong else world * n_rows_all  # rows ALL ranks process per step
    head_truth_local = truth_mine if strong else truth_all
    head_truth_global = truth_all if strong else world * truth_all

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def stepper(stream):
        if world > 1:
            return lambda: stream.filter_count_global(region)

        def one():
            c = stream.filter_count(region)
            return c, c
        return one

    def timed(fn, steps, warmup, sampler=None):
        barrier()  # ranks finish generating / uploading their shards at different times
        for _ in range(warmup):
            fn()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        out = None
        l0 = ctx.launch_count()
        if sampler:
            sampler.__enter__()
        ev0.record(tstream)
        for _ in range(steps):
            out = fn()
        ev1.record(tstream)
        barrier()
        if sampler:
            sampler.__exit__()
        ms = ev0.elapsed_time(ev1)
        # device time of each launch of the region: CUDA events the library records around the kernel on its stream
        kms = ctx.kernel_ms_history(min(steps, 64))
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, kms, out, ctx.launch_count() - l0

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    traffic_table = {}
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic_table = json.load(open(tp))

    def roofline_of(kms, body_bytes, ms_step, mode):
        kernel_ms = float(np.mean(kms))
        achieved = body_bytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tj = traffic_table.get(mode) if isinstance(traffic_table.get(mode), dict) else None
        if tj and tj.get("rows") == args.rows and world == 1:
            traffic = tj.get("dram_bytes_per_launch")
        return {"bound": "hbm", "kernel": "vcf_scan_kernel", "mode": mode, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": body_bytes, "bytes_per_row": body_bytes / max(1, (n_rows_mine if strong else n_rows_all)),
                "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_step,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    mode_name = {0: "lazy", 1: "strict"}
    sampler = ClockSampler(local)
    ms_step, kms, (loc, glob), launches = timed(stepper(head), args.steps, max(args.warmup, 3), sampler)
    assert loc == head_truth_local, f"rank {rank}: GPU count {loc} != generator truth {head_truth_local}"
    assert glob == head_truth_global, f"global count {glob} != generator truth {head_truth_global}"
    value = head_rows_total / (ms_step * 1e-3)
    roofline = roofline_of(kms, head_bytes, ms_step, mode_name[int(bool(args.strict))])

    # ---- the other validation mode on the same resident bytes (same counts on valid input) ----
    other = open_resident(head_idx, not args.strict)
    o_ms, o_kms, (o_loc, o_glob), _ = timed(stepper(other), args.steps, 3)
    assert o_loc == head_truth_local and o_glob == head_truth_global
    other_mode = mode_name[int(not args.strict)]
    modes = {roofline["mode"]: {"ms_per_step": ms_step, "value": value, "roofline": roofline},
             other_mode: {"ms_per_step": o_ms, "value": head_rows_total / (o_ms * 1e-3),
                          "roofline": roofline_of(o_kms, head_bytes, o_ms, other_mode)}}
    other.close()

    # ---- the other scaling series (N > 1): weak = every rank scans the whole set, strong = its share ----
    series = None
    if world > 1:
        alt_idx = everything if strong else mine
        alt = open_resident(alt_idx, args.strict)
        a_ms, a_kms, (a_loc, a_glob), _ = timed(stepper(alt), args.steps, 3)
        a_rows = world * n_rows_all if strong else n_rows_all
        assert a_loc == (truth_all if strong else truth_mine) and a_glob == (world * truth_all if strong else truth_all)
        series = {"scaling": "weak" if strong else "strong", "value": a_rows / (a_ms * 1e-3), "ms_per_step": a_ms,
                  "kernel_ms": float(np.mean(a_kms)), "rows_per_step_all_ranks": a_rows}
        alt.close()

    # ---- end to end through the C ABI with host buffers: H2D of this rank's shard files inside the timed region ----
    pins = []
    e2e_files = []
    for i in head_idx:
        p = ctx.pinned(files[i].size)
        p.array[:] = files[i]
        pins.append(p)
        e2e_files.append(p.array)
    e2e_bytes = int(sum(f.size for f in e2e_files))
    e2e_stream = ctx.open_vcf(projection=(0, 1), kernel_variant=args.variant, strict=bool(args.strict), pushdown=region)

    def e2e_step():
        e2e_stream.reset()
        for f in e2e_files:
            e2e_stream.feed(f, is_last=True)
        if world > 1:
            return e2e_stream.filter_count_global(region)
        c = e2e_stream.filter_count(region)
        return c, c

    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    e2e_ms, _, (eloc, eglob), _ = timed(e2e_step, e2e_steps, 2)
    assert eloc == head_truth_local and eglob == head_truth_global

    # the bare host->device copy of the same pinned bytes on every rank at once: the ceiling e2e can reach on this box
    def h2d_only():
        for f, i in zip(e2e_files, head_idx):
            dbufs[i].upload_async(f)
        ctx.synchronize()
        return 0, 0

    h2d_ms, _, _, _ = timed(h2d_only, 3, 1)
    e2e = {"value": head_rows_total / (e2e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": e2e_bytes,
           "d2h_bytes_per_step": 64, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_gbs_per_gpu": e2e_bytes / e2e_ms / 1e6, "h2d_ceiling_gbs_per_gpu": e2e_bytes / h2d_ms / 1e6,
           "h2d_ceiling_note": "bare cudaMemcpyAsync of the same pinned bytes, all ranks at once (max over ranks)",
           "api": "exon_gpu_vcf_feed(host pinned, per shard file) + exon_gpu_vcf_filter_count" + ("_global" if world > 1 else "") + " (pushdown declared)"}

    line = {"metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": workload_config(args),
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "modes": modes, "clocks": sampler.summary(),
            "count": glob, "count_matches_truth": True, "gen_seconds": t_gen,
            "kernel_variant": args.variant, "strict": int(args.strict),
            "files_of_rank0": len(mine), "rows_of_rank0": n_rows_mine}
    if series:
        line["other_series"] = series

    if args.extra:
        line["extra"] = extra_measurements(args, ctx, tstream, timed, dbufs, files, region, n_rows_all, truth_all, peak)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample, _ = cpu_rows_per_sec(files, rows_per_file, cores, target_s=None, steps=3, warmup=2)
        line["cpu_baseline"] = {"value": v, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample + "; 2 warm-up + 3 timed passes"}
    elif rank == 0:
        line["cpu_baseline"] = None

    head.close()
    e2e_stream.close()
    for d in dbufs:
        d.free()
    for p in pins:
        p.free()
    ctx.close()
    if rank == 0 and world == 1 and not args.no_extra_configs and args.rows == 100_000_000:
        line["extra_configs"] = extra_configs()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def extra_configs():
    """The other BASELINE configs (FASTQ mean-quality filter, BAM flag / MAPQ per-reference count, mzML m/z filter + SUM), each
    measured by tools/bench_formats.py in its own process after this one has released the GPU: the compact form of the lines
    profiles/r2_{fastq_config2,bam_config4,mzml_config5}.json hold in full.  A failure is recorded, never raised."""
    import subprocess

    out = {}
    for fmt in ("fastq", "bam", "mzml"):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_formats.py"), fmt, "--steps", "10"], capture_output=True, text=True,
                               timeout=240, cwd=ROOT)
            rows = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode != 0 or not rows:
                out[fmt] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            d = json.loads(rows[-1])
            out[fmt] = {"metric": d["metric"], "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                        "workload": d["config"]["workload"], "e2e": d.get("e2e"), "roofline": d.get("roofline"),
                        "cpu_baseline": d.get("cpu_baseline"),
                        "matches_truth": bool(d.get("count_matches_truth", d.get("sum_matches_truth_1e-6", False)))}
        except Exception as e:  # noqa: BLE001
            out[fmt] = {"error": repr(e)[:300]}
    return out


def extra_measurements(args, ctx, tstream, timed, dbufs, files, region, n_rows, truth, peak):
    """COUNT(*), interval-only, K2 / K3 / wide column builds over the whole resident set (1 GPU)."""
    import numpy as np
    import torch

    from exon_b200 import _abi

    extra = {}
    with ctx.open_vcf(projection=(0, 1), kernel_variant=args.variant, strict=False) as st:
        for d, f in zip(dbufs, files):
            st.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
        body_bytes = st.body_bytes()
        for name, rg in [("count_star", None), ("interval_only", _abi.make_region(None, *QUERY[1:]))]:
            ms, k, _, _ = timed(lambda: (st.filter_count(rg),) * 2, 10, 3)
            extra[name] = {"ms_per_step": ms, "kernel_ms": float(np.mean(k)),
                           "gbs": body_bytes / (float(np.mean(k)) * 1e-3) / 1e9}
    # K2: text -> Arrow {chrom, pos} batches left in device memory (second build = steady state: scratch areas
    # warm); K3: FilterExec + COUNT over all of those batches in one launch (exon_gpu_vcf_filter_agg)
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
    return extra


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
