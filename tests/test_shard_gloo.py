"""World-size-2 gloo test of the N>1 host logic used by bench.py: utterance sharding without a
data-path collective + the max-over-ranks timing / sum-over-ranks throughput reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from summertts_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    lengths = rng.integers(5, 200, size=37).tolist()
    parts = shard.balanced(lengths, world)
    mine = parts[rank]
    # stand-in for the per-rank engine run: samples = 256 * 5 frames per id
    samples = torch.tensor([float(sum(lengths[i] for i in mine) * 5 * 256)], dtype=torch.float64)
    ms = torch.tensor([10.0 + 3.0 * rank], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(samples, op=dist.ReduceOp.SUM)
    owned = [None] * world
    dist.all_gather_object(owned, mine)
    if rank == 0:
        q.put((ms.item(), samples.item(), owned, lengths))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, samples, owned, lengths = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 13.0  # max over ranks
    assert sorted(i for part in owned for i in part) == list(range(len(lengths)))  # exact cover, no overlap
    assert samples == sum(lengths) * 5 * 256
    loads = [sum(lengths[i] for i in part) for part in owned]
    assert max(loads) - min(loads) <= max(lengths)


def test_partitions():
    assert shard.round_robin(10, 4, 1) == [1, 5, 9]
    assert shard.weak_batch(8, 4, 2) == list(range(16, 24))
    parts = shard.balanced([5, 9, 1, 7, 7, 3], 3)
    assert sorted(i for p in parts for i in p) == list(range(6))
