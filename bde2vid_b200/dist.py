"""Multi-GPU plumbing: independent sequences are sharded across ranks (one process per GPU); the only
collective on the path is the final metric reduction.

The reference has no distributed code at all; its per-file metric averaging is eval_models_seq.py:278-282.
"""
import torch
import torch.distributed as dist


def shard_units(costs, world_size):
    """Greedy longest-first assignment of independent units (sequences / sub-sequence chunks) to ranks.
    ``costs[i]`` ~ T * Hp * Wp of unit i.  Returns a list of unit-index lists, one per rank; deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    for r in range(world_size):
        out[r].sort()
    return out


def reduce_metric_sums(sums, device=None):
    """All-reduce (sum) a dict of float metric sums + frame counts across ranks (float64).
    With no process group initialised it returns the input unchanged."""
    keys = sorted(sums)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(sums)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(sums[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def finalize_means(total):
    """{'mse': sum, 'ssim': sum, ..., 'n': frames} -> per-frame means (eval_models_seq.py:278-282)."""
    n = max(total.get("n", 0.0), 1.0)
    return {k: v / n for k, v in total.items() if k != "n"}
