"""File-group sharding across GPUs and the rendezvous for the final aggregate.

The reference partitions a scan by whole files (`ExonFileScanConfig::regroup_files_by_size`,
exon/exon-core/src/datasources/exon_file_scan_config.rs:79-110; `VCFScan::repartitioned`,
exon/exon-core/src/datasources/vcf/scanner.rs:103-124): one partition stream per file group, partial aggregates
merged by CoalescePartitionsExec + AggregateExec(Final).  Here a partition is a GPU: the same rule (through the C
ABI) assigns files to ranks, and the merge is one ncclAllReduce inside the library.  torch.distributed is used only
to hand the NCCL unique id from rank 0 to the other ranks.
"""
from __future__ import annotations

import ctypes as C

from . import _abi


def assign_files(sizes, world_size: int) -> list[int]:
    """Rank of every file: stable sort by size ascending, file i of that order -> rank i % min(world, n_files)."""
    lib = _abi.load()
    n = len(sizes)
    out = (C.c_int32 * max(n, 1))()
    parts = C.c_int32()
    _abi.check(lib.exon_gpu_regroup_files_by_size((C.c_int64 * max(n, 1))(*[int(s) for s in sizes]), n, int(world_size),
                                                  out, C.byref(parts)))
    return list(out)[:n]


def files_of_rank(sizes, rank: int, world_size: int) -> list[int]:
    """Indices of the files rank `rank` scans, in input (file) order -- the order FileStream would open them."""
    return [i for i, r in enumerate(assign_files(sizes, world_size)) if r == rank]


def init_final_aggregate(ctx, dist, rank: int, world_size: int) -> None:
    """Create the library's NCCL communicator on every rank (id made on rank 0, broadcast by the host)."""
    ids = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.nccl_init(ids[0], world_size, rank)
