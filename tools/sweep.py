#!/usr/bin/env python
"""Kernel-variant sweep for the fused VCF scan (tuning aid; prints one JSON line per configuration).

    python tools/sweep.py [--rows 100000000] [--reps 10]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from exon_b200 import _abi  # noqa: E402
from exon_b200.runtime import Context  # noqa: E402
from synth import vcf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--shards", type=int, default=64)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--variants", default="0,1,2,3,4,5")
args = ap.parse_args()

cols = vcf.columns(args.rows)
files = vcf.shards(cols, args.shards)
queries = {"region": ("1", 1_000_000, 2_000_000), "chrom22": ("22", None, None), "interval": (None, 1_000_000, 2_000_000),
           "count_star": (None, None, None)}
truth = {k: cols.truth_count(*q) for k, q in queries.items()}
with Context(0) as ctx:
    dbufs = []
    for f in files:
        d = ctx.device_buffer(f.size)
        d.upload(np.ascontiguousarray(f))
        dbufs.append(d)
    for variant in [int(v) for v in args.variants.split(",")]:
        for strict in (0, 1):
            with ctx.open_vcf(kernel_variant=variant, strict=bool(strict)) as s:
                for d, f in zip(dbufs, files):
                    s.feed(None, device_ptr=d.ptr, nbytes=f.size, is_last=True)
                body = s.body_bytes()
                for name, q in queries.items():
                    if strict and name in ("interval", "count_star"):
                        continue
                    rg = _abi.make_region(*q)
                    ms = []
                    for i in range(args.reps + 3):
                        c = s.filter_count(rg)
                        assert c == truth[name], (name, c, truth[name])
                        if i >= 3:
                            ms.append(ctx.last_kernel_ms())
                    ms.sort()
                    print(json.dumps({"variant": variant, "strict": strict, "query": name, "kernel_ms_med": ms[len(ms) // 2],
                                      "kernel_ms_min": ms[0], "GBps_med": body / ms[len(ms) // 2] / 1e6,
                                      "rows_per_s": args.rows / ms[len(ms) // 2] * 1e3}), flush=True)
