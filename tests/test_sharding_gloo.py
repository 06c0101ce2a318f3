"""Multi-GPU path on CPU: two gloo ranks shard a heterogeneous batch by file, each runs the host prepass on its
shard, and the join (the only communication the path has) reproduces the single-process totals."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, seeds, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import audio_formats_b200 as af
    from audio_formats_b200 import shard, synth
    streams = [synth.generate(synth.config4_params(s, 1.0)) for s in seeds]
    costs = [s.granules * s.params.nch for s in streams]
    mine = shard.shard_lpt(costs, world)[rank]
    scans = [af.Scan(streams[i].data) for i in mine]
    local = torch.tensor([sum(s.granules * s.channels for s in scans), sum(s.delivered_samples for s in scans),
                          len(mine)], dtype=torch.int64)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)     # the timing join: max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put(([g.tolist() for g in gathered], float(t.item()), sum(costs)))
    dist.destroy_process_group()


def test_two_rank_file_sharding(built):
    from audio_formats_b200 import shard
    seeds = list(range(200, 224))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seeds, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax, total_cost = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert tmax == 2.0
    assert gathered[0][2] + gathered[1][2] == len(seeds)
    assert gathered[0][0] + gathered[1][0] == total_cost
    loads = [g[0] for g in gathered]
    assert max(loads) <= 1.25 * (sum(loads) / 2)     # LPT keeps the two GPUs balanced


def test_shard_helpers():
    from audio_formats_b200 import shard
    assert [list(shard.shard_contiguous(10, r, 4)) for r in range(4)] == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    parts = shard.shard_lpt([5, 1, 1, 1, 4, 3, 3], 2)
    assert sorted(sum(parts, [])) == list(range(7))
    assert abs(sum([5, 1, 1, 1, 4, 3, 3][i] for i in parts[0]) - 9) <= 1
