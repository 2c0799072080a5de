"""CPU, world_size 2 over gloo: the N>1 path = shard independent sequences, no data-path collective, one
metric all-reduce at the end."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bde2vid_b200.dist import finalize_means, reduce_metric_sums, shard_units


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [100, 50, 50, 25, 25, 25, 25]
    mine = shard_units(costs, world)[rank]
    # stand-in per-sequence metric: depends only on the sequence id, so the reduced result is checkable
    sums = {"mse": sum(0.01 * (i + 1) * costs[i] for i in mine), "n": float(sum(costs[i] for i in mine))}
    total = reduce_metric_sums(sums)
    ret[rank] = (mine, total)
    dist.destroy_process_group()


def test_two_rank_shard_and_metric_reduce():
    world, port = 2, 29541
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    costs = [100, 50, 50, 25, 25, 25, 25]
    a, b = ret[0], ret[1]
    assert sorted(a[0] + b[0]) == list(range(len(costs)))
    assert a[1] == b[1]                                    # every rank holds the same reduced sums
    assert abs(a[1]["n"] - sum(costs)) < 1e-9
    assert abs(a[1]["mse"] - sum(0.01 * (i + 1) * c for i, c in enumerate(costs))) < 1e-9
    assert abs(finalize_means(a[1])["mse"] - a[1]["mse"] / a[1]["n"]) < 1e-12
