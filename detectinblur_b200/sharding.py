"""Data-parallel sharding of the blur path: one process per GPU, images split by rank, no collective on the hot path.

The reference shards with ``torch.utils.data.DistributedSampler`` (train.py:187-189: every rank takes a strided
slice of the index list padded to a common length) and only communicates for gradients, logging and COCO results.
The blur kernels need nothing from other ranks; the one optional message is an all-gather of a 64-bit checksum per
rank so that a run can be compared shard by shard (the pattern of ``utils.all_gather``, utils.py:536-576).
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world_size, drop_last=False):
    """Indices rank ``rank`` owns: DistributedSampler's strided split without shuffling (train.py:187)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    idx = list(range(n_items))
    if drop_last:
        total = (n_items // world_size) * world_size
        idx = idx[:total]
    else:
        total = -(-n_items // world_size) * world_size
        if n_items and total > n_items:
            idx += (idx * (-(-(total - n_items) // n_items)))[:total - n_items]   # pad by wrapping, like the sampler
    return idx[rank:total:world_size]


def gather_checksums(local_checksum, device=None):
    """All-gather one 64-bit value per rank; returns the list ordered by rank (a single-element list without a group)."""
    if isinstance(local_checksum, torch.Tensor):
        t = local_checksum.reshape(1).to(torch.int64)
    else:
        v = int(local_checksum) & 0xFFFFFFFFFFFFFFFF
        if v >= 1 << 63:
            v -= 1 << 64
        t = torch.tensor([v], dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(t.item()) & 0xFFFFFFFFFFFFFFFF]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(o.item()) & 0xFFFFFFFFFFFFFFFF for o in out]


def max_over_ranks(value, device=None):
    """Max of a python float over the ranks (timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
