"""world_size-2 host-side test (gloo, CPU): the file -> rank assignment is the reference's regroup rule, every file
is scanned exactly once, and partial counts merge to the global answer the way AggregateExec(Final) would."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from exon_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sizes, per_file_counts, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.files_of_rank(sizes, rank, world)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    partial = torch.tensor([sum(per_file_counts[i] for i in mine)], dtype=torch.int64)
    dist.all_reduce(partial)  # the stand-in for the library's ncclAllReduce(sum, int64, 1)
    if rank == 0:
        out.put((everyone, int(partial.item())))
    dist.destroy_process_group()


def test_two_rank_sharding_and_final_aggregate():
    from synth import vcf

    cols = vcf.columns(40_000)
    files = vcf.shards(cols, 7)
    files[3] = files[3][: files[3].size // 2 + int(np.argmax(files[3][files[3].size // 2:] == 10)) + 1]  # ragged sizes
    sizes = [int(f.size) for f in files]
    per_file = [oracle.filter_count(f, "1", 1_000_000, 50_000_000)[0] for f in files]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, per_file, q)) for r in range(2)]
    for p in procs:
        p.start()
    everyone, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(i for m in everyone for i in m) == list(range(7))       # every file exactly once
    assert total == sum(per_file) == oracle.filter_count_files(files, "1", 1_000_000, 50_000_000, target_partitions=2)[0]
    # the assignment equals the reference rule as restated by the oracle
    parts, groups = oracle.regroup_files_by_size(sizes, 2)
    assert parts == 2 and sharding.assign_files(sizes, 2) == groups


def test_assignment_edge_cases():
    assert sharding.assign_files([], 4) == []
    assert sharding.assign_files([5], 8) == [0]
    assert sharding.assign_files([3, 1, 2], 8) == [2, 0, 1]          # partitions = min(target, files)
    assert sharding.assign_files([10, 10, 10, 10], 2) == [0, 1, 0, 1]  # stable on ties
