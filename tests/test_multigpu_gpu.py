"""Two-rank GPU tests of the final aggregate: the peer-memory exchange fused into the scan kernel's tail
(exon_gpu_vcf_filter_count_global), its NCCL fallback, exon_gpu_allreduce_counts and exon_gpu_allreduce_partial.
Needs two GPUs on the box (`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`); skipped otherwise.
Every expected value comes from the oracle (local counts) and plain integer / float sums (the merge)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _two_gpus() -> bool:
    try:
        return torch.cuda.is_available() and torch.cuda.device_count() >= 2
    except Exception:
        return False


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, peer_xchg, steps, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), EXON_GPU_PEER_XCHG=str(peer_xchg))
        import time

        import torch.distributed as dist

        import oracle
        from exon_b200 import _abi, sharding
        from exon_b200.runtime import Context, ExonGpuError
        from synth import vcf

        torch.cuda.set_device(rank)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ctx = Context(rank)
        sharding.init_final_aggregate(ctx, dist, rank, world)

        # ---- vectors and aggregate state over NCCL ----
        counts = ctx.allreduce_counts([rank + 1, 10 * (rank + 1), 0, 7])
        assert counts == [sum(r + 1 for r in range(world)), sum(10 * (r + 1) for r in range(world)), 0, 7 * world], counts
        assert ctx.allreduce_counts([]) == []
        c, si, sf = ctx.allreduce_partial(count=3 + rank, sum_i64=-5 * (rank + 1), sum_f64=0.25 * (rank + 1))
        assert c == sum(3 + r for r in range(world)) and si == sum(-5 * (r + 1) for r in range(world))
        assert abs(sf - sum(0.25 * (r + 1) for r in range(world))) <= 1e-12

        # ---- the file set, sharded with the reference's rule; every rank knows every file's oracle count ----
        cols = vcf.columns(120_000)
        files = vcf.shards(cols, 6)
        sizes = [int(f.size) for f in files]
        mine = sharding.files_of_rank(sizes, rank, world)
        queries = [("1", 1_000_000, 60_000_000), ("X", None, None), (None, 5, 40_000_000), (None, None, None), ("nope", 1, 2)]
        per_file = {qi: [oracle.filter_count(f, *qq)[0] for f in files] for qi, qq in enumerate(queries)}
        for strict in (False, True):
            with ctx.open_vcf(projection=(0, 1), strict=strict) as s:
                for i in mine:
                    s.feed(files[i], is_last=True)
                for qi, qq in enumerate(queries):
                    loc, glob = s.filter_count_global(_abi.make_region(*qq))
                    assert loc == sum(per_file[qi][i] for i in mine), (qq, loc)
                    assert glob == sum(per_file[qi]), (qq, glob)
                # many back-to-back exchanges; one rank is late now and then (the parity protocol must never mix steps)
                rg = [_abi.make_region(*queries[0]), _abi.make_region(*queries[2])]
                want = [(sum(per_file[qi][i] for i in mine), sum(per_file[qi])) for qi in (0, 2)]
                for k in range(steps):
                    if rank == (k // 97) % world and k % 97 == 0:
                        time.sleep(0.01)
                    assert s.filter_count_global(rg[k & 1]) == want[k & 1], k
                # an empty partition (nothing fed) still takes part
            with ctx.open_vcf(projection=(0, 1), strict=strict) as s:
                if rank == 0:
                    s.feed(files[0], is_last=True)
                loc, glob = s.filter_count_global(_abi.make_region(*queries[0]))
                assert loc == (per_file[0][0] if rank == 0 else 0) and glob == per_file[0][0]
            # pushdown (eager) streams publish through the tail-only launch
            rgp = _abi.make_region(*queries[0])
            with ctx.open_vcf(projection=(0, 1), strict=strict, pushdown=rgp) as s:
                for i in mine:
                    s.feed(files[i][: files[i].size // 2], is_last=False)
                    s.feed(files[i][files[i].size // 2:], is_last=True)
                for _ in range(3):
                    assert s.filter_count_global(rgp) == (sum(per_file[0][i] for i in mine), sum(per_file[0]))

        # ---- a rank whose local scan fails still delivers a partial; the next query is in step again ----
        bad = b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n1\t12x\t.\tA\tC\t5\tPASS\t.\n1\t13\t.\tA\tC\t5\tPASS\t.\n"
        good = b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n1\t12\t.\tA\tC\t5\tPASS\t.\n1\t13\t.\tA\tC\t5\tPASS\t.\n"
        with ctx.open_vcf(projection=(0, 1), strict=True) as s:
            s.feed(bad if rank == 1 else good, is_last=True)
            rg = _abi.make_region("1", 1, 100)
            if rank == 1:
                with pytest.raises(ExonGpuError) as ei:
                    s.filter_count_global(rg)
                assert ei.value.code == _abi.ERR_PARSE
            else:
                loc, glob = s.filter_count_global(rg)
                assert loc == 2 and glob in (2, 3)  # the failing rank's partial is whatever it counted before failing
            # the exchange stays in step: the next query delivers again (the bad file still fails its own rank)
            if rank == 1:
                with pytest.raises(ExonGpuError):
                    s.filter_count_global(_abi.make_region("2", None, None))
            else:
                assert s.filter_count_global(_abi.make_region("2", None, None)) == (0, 0)
        with ctx.open_vcf(projection=(0, 1), strict=False) as s:
            s.feed(good, is_last=True)
            assert s.filter_count_global(_abi.make_region("1", 13, 13)) == (1, world)
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001 - report to the parent, whatever it was
        import traceback

        q.put((rank, "".join(traceback.format_exception(type(e), e, e.__traceback__))))


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs on the box")
@pytest.mark.parametrize("peer_xchg", [1, 0])
def test_two_rank_final_aggregate(peer_xchg):
    world = 2
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, peer_xchg, 1000 if peer_xchg else 100, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, msg in got:
        assert msg == "ok", f"rank {rank}:\n{msg}"
    assert all(p.exitcode == 0 for p in procs)
