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


def _eval_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bde2vid_b200.eval_seq import eval_files
    files = ["HQF/a.npz", "HQF/b.npz", "ECD/c.npz", "ECD/d.npz", "ECD/e.npz"]
    seen = []

    def fake_eval(f):           # per-file "metrics" that depend only on the file name: the gathered table is checkable
        seen.append(f)
        n = 2 + files.index(f)
        d = {"mse": [0.1 * files.index(f)] * n, "ssim": [0.5] * n}
        return {k: sum(v) / len(v) for k, v in d.items()}, d

    results, details, overall = eval_files(files, fake_eval, costs=[5, 4, 3, 2, 1])
    ret[rank] = (seen, results, overall)
    dist.destroy_process_group()


def test_two_rank_headless_driver_sharding_and_gather():
    """eval_seq.eval_files: files sharded over 2 ranks, every rank ends with the full {dataset: {file: row}} table
    (reference schema, eval_models_seq.py:122-135) and the frame-weighted overall means from ONE all-reduce."""
    world, port = 2, 29543
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_eval_worker, args=(world, port, ret), nprocs=world, join=True)
    (s0, r0, o0), (s1, r1, o1) = ret[0], ret[1]
    assert sorted(s0 + s1) == sorted(["HQF/a.npz", "HQF/b.npz", "ECD/c.npz", "ECD/d.npz", "ECD/e.npz"]) and s0 and s1
    assert r0 == r1 and set(r0) == {"HQF", "ECD"} and set(r0["ECD"]) == {"c", "d", "e"}
    assert abs(r0["ECD"]["d"]["mse"] - 0.3) < 1e-12
    n = [2, 3, 4, 5, 6]
    want = sum(0.1 * i * n[i] for i in range(5)) / sum(n)
    assert abs(o0["mse"] - want) < 1e-12 and abs(o1["mse"] - want) < 1e-12 and abs(o0["ssim"] - 0.5) < 1e-12


def _pair_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bde2vid_b200.engine import PairSplit
    pair = PairSplit(rank)
    T, Tc = 19, 8
    # stand-ins for the per-level hidden-state sequences: every rank computes only its own direction
    hf = torch.full((T, 4), float("nan"))
    hb = torch.full((T, 4), float("nan"))
    if pair.owns_direction(False):
        hf = torch.arange(T * 4, dtype=torch.float32).view(T, 4)
    if pair.owns_direction(True):
        hb = -torch.arange(T * 4, dtype=torch.float32).view(T, 4) * 0.5
    pair.broadcast(hf, 0)
    pair.broadcast(hb, 1)
    # decoder chunks alternate; foreign chunks are zero, the frames meet in one all-reduce(sum)
    img = torch.full((T, 3), float("nan"))
    owned = []
    for c in range((T + Tc - 1) // Tc):
        sl = slice(c * Tc, min((c + 1) * Tc, T))
        if pair.owns_chunk(c):
            img[sl] = (hf + hb)[sl, :3] + 100.0
            owned.append(c)
        else:
            img[sl] = 0.0
    pair.sum_frames(img)
    ret[rank] = (owned, hf.clone(), hb.clone(), img.clone())
    dist.destroy_process_group()


def test_pair_split_schedule_and_exchange():
    """engine.PairSplit: rank 0 = forward chains, rank 1 = backward chains, decoder chunks alternate,
    hidden states by broadcast, frames by one all-reduce: both ranks end with identical, complete data."""
    world, port = 2, 29545
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pair_worker, args=(world, port, ret), nprocs=world, join=True)
    (o0, hf0, hb0, i0), (o1, hf1, hb1, i1) = ret[0], ret[1]
    assert sorted(o0 + o1) == [0, 1, 2] and not set(o0) & set(o1) and o0 and o1
    T = 19
    hf = torch.arange(T * 4, dtype=torch.float32).view(T, 4)
    hb = -hf * 0.5
    assert torch.equal(hf0, hf) and torch.equal(hf1, hf) and torch.equal(hb0, hb) and torch.equal(hb1, hb)
    want = (hf + hb)[:, :3] + 100.0
    assert torch.equal(i0, want) and torch.equal(i1, want)
