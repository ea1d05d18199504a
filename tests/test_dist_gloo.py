"""world_size-2 gloo test of the multi-GPU host logic (clip sharding + the single metric all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iip_uavsal_saliency_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    D.init_process_group("gloo")
    g = torch.Generator().manual_seed(0)
    vals = torch.rand(37, 4, generator=g)                      # every rank builds the same global table
    vals[5, 2] = float("nan")
    mine = vals[D.shard_indices(37, rank, world)]
    means = D.allreduce_metric_means(D.metric_partial(mine))
    t = D.max_over_ranks(1.0 + rank)
    q.put((rank, means.tolist(), t))
    dist.destroy_process_group()


def test_sharded_metric_means_equal_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(30)
    g = torch.Generator().manual_seed(0)
    vals = torch.rand(37, 4, generator=g)
    vals[5, 2] = float("nan")
    ok = ~torch.isnan(vals).any(1)
    expect = vals[ok].double().mean(0)
    for rank, means, t in res:
        assert torch.allclose(torch.tensor(means, dtype=torch.float64), expect, atol=1e-12)
        assert t == 2.0                                        # max over ranks
    assert sorted(D.shard_indices(10, 0, 4) + D.shard_indices(10, 1, 4) + D.shard_indices(10, 2, 4) + D.shard_indices(10, 3, 4)) == list(range(10))
