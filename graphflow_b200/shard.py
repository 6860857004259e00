"""Sharding of a batch of independent contraction instances (one per (graph, vertex, level)) over the ranks of one
node, and the one collective the path has: the sum of the parameter gradients.

Mirrors the reference's only parallel scheme, SMP_beta::Threaded_BatchLearn (SMP_beta.h:697-739): a replica per
worker, examples dealt out to the replicas, `add_gradient` (a host loop, :677-687) summing the replicas' parameter
gradients.  Here a replica is a rank (one process per GPU), instances never interact inside forward/backward, so the
data path has no collective at all; `allreduce_gradients` is `add_gradient`.
"""
import numpy as np


def contiguous_shard(batch, world, rank):
    """[start, stop) of `rank`'s contiguous share of `batch` units; shares differ by at most one unit."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    base, extra = divmod(int(batch), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def instance_cost(n, C):
    """Bytes one instance streams forward + backward: 8 (n^3 C + 18 n^2 C + n^2)."""
    n = np.asarray(n, np.int64)
    return 8 * (n ** 3 * C + 18 * n * n * C + n * n)


def balanced_shards(sizes, C, world):
    """Ragged batches: deal instances to ranks so the streamed bytes are balanced (longest-processing-time greedy
    over instance_cost).  Returns a list of `world` int64 index arrays, each sorted by instance size so that one
    rank's launch sees instances of similar n next to each other."""
    sizes = np.asarray(sizes, np.int64)
    cost = instance_cost(sizes, C)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world, np.int64)
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += cost[i]
    out = []
    for b in bins:
        b = np.asarray(b, np.int64)
        out.append(b[np.argsort(sizes[b], kind="stable")] if len(b) else b)
    return out


def allreduce_gradients(tensors, group=None):
    """Sum the parameter gradients of all ranks in place (NCCL on GPUs, gloo on CPU): one flat all-reduce, the
    payload being ~1 MB it is latency-bound, so one launch beats one per tensor."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        k = t.numel()
        t.copy_(flat[off:off + k].view_as(t))
        off += k
    return tensors
