#!/usr/bin/env python
"""K1 kernel time against resident bytes on ONE GPU: the shapes a strong-scaling run gives each rank (64, 32, 16, 8 of
the 64 shard files of the 100M-row set), both validation modes.  Separates the kernel's own non-linearity (ramp-up,
tail, boundary tiles) from anything the multi-GPU exchange adds.  Prints one JSON line.

    python tools/k1_size_sweep.py [--rows 100000000] [--variant 0]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--shards", type=int, default=64)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import numpy as np

    from exon_b200 import _abi, sharding
    from exon_b200.runtime import Context
    from synth import vcf

    cols = vcf.columns(args.rows)
    files = vcf.shards(cols, args.shards)
    sizes = [int(f.size) for f in files]
    region = _abi.make_region("1", 1_000_000, 2_000_000)
    out = {"rows": args.rows, "variant": args.variant, "points": []}
    with Context(0) as ctx:
        dbufs = []
        for f in files:
            d = ctx.device_buffer(f.size)
            d.upload(f)
            dbufs.append(d)
        for world in (1, 2, 4, 8):
            idx = sharding.files_of_rank(sizes, 0, world)
            for strict in (1, 0):
                with ctx.open_vcf(projection=(0, 1), strict=bool(strict), kernel_variant=args.variant) as s:
                    for i in idx:
                        s.feed(None, device_ptr=dbufs[i].ptr, nbytes=files[i].size, is_last=True)
                    body = s.body_bytes()
                    for _ in range(5):
                        s.filter_count(region)
                    import time

                    ctx.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(args.steps):
                        s.filter_count(region)
                    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
                    k = ctx.kernel_ms_history(args.steps)
                    out["points"].append({"files": len(idx), "strict": strict, "body_bytes": body, "kernel_ms": float(np.mean(k)),
                                          "kernel_ms_min": float(np.min(k)), "wall_ms_per_step": wall_ms,
                                          "gbs": body / float(np.mean(k)) / 1e6})
        for d in dbufs:
            d.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
