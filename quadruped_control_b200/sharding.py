"""Batch sharding for the multi-GPU path: every QP is independent, so a rank owns a
contiguous index range and there is no data-path collective.  The only exchange is the report:
max over ranks of the elapsed time and sums of the counters -- and, for a single consumer that wants the
results of every shard (optional), one all-gather of the result records."""


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of rank ``rank``; remainders go to the lowest ranks."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_report(elapsed_ms, counters, dist=None, device=None):
    """(max elapsed over ranks, element-wise sum of counters).  ``dist`` is torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(elapsed_ms), [float(c) for c in counters]
    import torch

    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(v) for v in counters], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(v) for v in c.tolist()]


def gather_outputs(local, dist=None, out=None):
    """All-gather of the ranks' result records: ``local`` is this rank's shard (a contiguous torch tensor, the same number
    of bytes on every rank -- equal shards); returns a tensor of world x local.numel() elements holding the shards in rank
    order, i.e. the batch in its original order (shards are contiguous index ranges).  NCCL over NVLink / NVSwitch for
    device tensors, gloo for host tensors; ``dist`` None or a world of one: the shard itself.  ``out``: reuse a buffer."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world = dist.get_world_size()
    flat = local.reshape(-1)
    if out is None:
        out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    assert out.numel() == world * flat.numel() and out.dtype == flat.dtype
    dist.all_gather_into_tensor(out, flat)
    return out
