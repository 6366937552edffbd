#!/usr/bin/env python
"""Multi-GPU lines for the BASELINE configs that name a sharding other than the headline's:

  bam   configs[3]: BAM flag + MAPQ filter + per-contig COUNT, 200M synthetic alignments, CONTIG-sharded across the ranks
        (every rank holds the alignments of its own group of references; groups are made largest-first so that bytes
        balance); the per-reference count vectors are merged with exon_gpu_allreduce_counts (one
        ncclAllReduce(sum, int64, n_groups)) -- the analogue of AggregateExec(Final) of the GROUP BY.
  mzml  configs[4]: mzML m/z range filter + SUM(intensity), 1M synthetic spectra, files sharded with the reference's
        regroup_files_by_size rule; {count, spectra, f64 sum} merged with exon_gpu_allreduce_partial (1e-6 relative).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_multi.py {bam|mzml} --gpus N [--steps K]

One JSON line on rank 0; timing = CUDA events on the library's stream, max over ranks, barrier on both sides.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def setup(args):
    import torch.distributed as dist

    from exon_b200 import sharding
    from exon_b200.runtime import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tstream = torch.cuda.Stream()
    ctx = Context(local, cuda_stream=tstream.cuda_stream)
    if world > 1:
        sharding.init_final_aggregate(ctx, dist, rank, world)
    return dist if world > 1 else None, rank, world, tstream, ctx


def timed(dist, tstream, fn, steps, warmup):
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(tstream)
    for _ in range(steps):
        out = fn()
    e1.record(tstream)
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, out


def contig_groups(refs, world):
    """Largest-first greedy assignment of references to ranks (balances reference length, i.e. alignments and bytes)."""
    load = [0] * world
    group = [[] for _ in range(world)]
    for i in sorted(range(len(refs)), key=lambda i: -refs[i][1]):
        r = min(range(world), key=lambda r: load[r])
        group[r].append(i)
        load[r] += refs[i][1]
    return [sorted(g) for g in group]


def bench_bam(args):
    from synth import bam

    dist, rank, world, tstream, ctx = setup(args)
    refs = bam.REFS
    groups = contig_groups(refs, world)
    share = [sum(refs[i][1] for i in g) for g in groups]
    per_rank = [int(round(args.alignments * s / sum(share))) for s in share]
    per_rank[-1] = args.alignments - sum(per_rank[:-1])
    mine = [refs[i] if i in groups[rank] else (refs[i][0], 0) for i in range(len(refs))]  # other contigs: probability 0
    t0 = time.perf_counter()
    sh = bam.shards(per_rank[rank], args.shards, seed=bam.SEED + rank, level=1, refs=_with_probabilities(refs, mine))
    gen_s = time.perf_counter() - t0
    kw = dict(flag_exclude=0x904, min_mapq=30)
    truth_local = sh.truth(**kw)
    s = ctx.open_bam()
    for f in sh.files:
        s.feed(np.frombuffer(f, dtype=np.uint8))
    names = None

    def step():
        nonlocal names
        got, rows = s.count_by_reference(**kw)
        names = list(got.keys())  # header order + NULL: the same on every rank
        vec = ctx.allreduce_counts([got[k] for k in names] + [rows]) if world > 1 else [got[k] for k in names] + [rows]
        return got, vec

    ms, (got, vec) = timed(dist, tstream, step, args.steps, 3)
    assert got == truth_local, "local per-reference counts differ from the generator's truth"
    # global truth: every rank's local truth, summed on the host side of the harness (gloo-free: one more all-reduce of the truth)
    tv = [truth_local[k] for k in names] + [sh.n]
    if world > 1:
        t = torch.tensor(tv, device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        tv = [int(x) for x in t.tolist()]
    assert vec == tv, "merged per-reference counts differ from the sum of the ranks' truths"
    assert vec[-1] == args.alignments
    if rank == 0:
        line = {"metric": "bam_flag_mapq_filter_count_by_reference_alignments_per_sec", "value": args.alignments / ms * 1e3,
                "unit": "alignments/s", "n_gpus": world, "steps": args.steps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "dtype": "int64", "data": "synthetic",
                "config": {"workload": f"BAM (flag & 0x904) = 0 AND mapq >= 30 GROUP BY reference, {args.alignments} synthetic alignments "
                                       f"(l_seq 100, {len(refs)} references), contig-sharded over {world} ranks (BASELINE configs[3])",
                           "contigs_per_rank": [[refs[i][0] for i in g] for g in groups], "alignments_per_rank": per_rank,
                           "final_aggregate": "exon_gpu_allreduce_counts: ncclAllReduce(sum, int64, n_groups + 1)",
                           "l2": "record stream per rank >> 126 MB L2, no flush"},
                "groups": len(names), "selected_total": int(sum(vec[:-1])), "counts_match_truth": True, "gen_seconds": gen_s}
        print(json.dumps(line), flush=True)
    s.close()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _with_probabilities(refs, mine):
    """synth.bam draws a record's reference with probability proportional to the length column: a rank's generator sees the
    real names (the header must be the same everywhere) and length 0 for the contigs of other ranks."""
    return [(n, l) for (n, _), (_, l) in zip(refs, mine)]


def bench_mzml(args):
    from exon_b200 import sharding
    from synth import mzml

    dist, rank, world, tstream, ctx = setup(args)
    t0 = time.perf_counter()
    sh = mzml.shards(args.spectra, args.shards, peaks=args.peaks)
    gen_s = time.perf_counter() - t0
    sizes = [int(f.size) for f in sh.files]
    mine = sharding.files_of_rank(sizes, rank, world)
    dbufs = []
    s = ctx.open_mzml()
    for i in mine:
        f = sh.files[i]
        d = ctx.device_buffer(f.size + 64)
        d.upload(f)
        dbufs.append(d)
        s.feed(None, device_ptr=d.ptr, nbytes=f.size)

    def step():
        ssum, n_sel, n_sp = s.filter_sum(sh.lo, sh.hi)
        if world > 1:
            n_sel, n_sp, ssum = ctx.allreduce_partial(count=n_sel, sum_i64=n_sp, sum_f64=ssum)
        return ssum, n_sel, n_sp

    ms, (ssum, n_sel, n_sp) = timed(dist, tstream, step, args.steps, 3)
    assert n_sp == sh.n and n_sel == sh.truth_count and math.isclose(ssum, sh.truth_sum, rel_tol=1e-6), (ssum, sh.truth_sum)
    if rank == 0:
        total = int(sum(sizes))
        line = {"metric": "mzml_mz_range_filter_sum_intensity_spectra_per_sec", "value": sh.n / ms * 1e3, "unit": "spectra/s",
                "n_gpus": world, "steps": args.steps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"mzML m/z BETWEEN {sh.lo} AND {sh.hi} + SUM(intensity), {sh.n} synthetic spectra x {sh.peaks} peaks in "
                                       f"{len(sh.files)} files ({total} B of text), files assigned to {world} ranks by regroup_files_by_size "
                                       f"(BASELINE configs[4])",
                           "final_aggregate": "exon_gpu_allreduce_partial: ncclAllReduce(sum) of {int64 count, int64 spectra} + {f64 sum}",
                           "tolerance": "sum within 1e-6 relative of the generator truth; counts exact"},
                "sum": ssum, "truth_sum": sh.truth_sum, "rel_err": abs(ssum - sh.truth_sum) / abs(sh.truth_sum),
                "selected_peaks": n_sel, "sum_matches_truth_1e-6": True, "gen_seconds": gen_s}
        print(json.dumps(line), flush=True)
    s.close()
    for d in dbufs:
        d.free()
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("fmt", choices=["bam", "mzml"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--alignments", type=int, default=200_000_000)
    ap.add_argument("--spectra", type=int, default=1_000_000)
    ap.add_argument("--peaks", type=int, default=200)
    ap.add_argument("--shards", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    {"bam": bench_bam, "mzml": bench_mzml}[a.fmt](a)
