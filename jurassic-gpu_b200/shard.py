"""Ray sharding across ranks: packages are independent, every rank owns a contiguous slice,
tables are broadcast once, results are gathered on rank 0.  The helpers work with any torch.distributed backend
(NCCL on the GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous balanced slice [first, first+count) of `total` packages for `rank` of `world`."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def broadcast_blob(dist, blob, src, device):
    """Broadcast a uint8 tensor whose size only `src` knows; returns the tensor on every rank."""
    import torch
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if dist.get_rank() == src:
        n[0] = blob.numel()
    dist.broadcast(n, src)
    if dist.get_rank() != src:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(blob, src)
    return blob


def gather_rows(dist, mine, counts, dst, device):
    """Gather per-rank row blocks (rank r contributes counts[r] rows of equal width) on `dst`, in rank order."""
    import torch
    world = dist.get_world_size()
    width, cap = mine.shape[1], max(counts)
    send = mine.contiguous()
    if send.shape[0] < cap:  # dist.gather wants equal shapes: pad to the largest block
        pad = torch.zeros((cap - send.shape[0], width), dtype=mine.dtype, device=device)
        send = torch.cat([send, pad], dim=0)
    if dist.get_rank() == dst:
        bufs = [torch.empty((cap, width), dtype=mine.dtype, device=device) for _ in range(world)]
        dist.gather(send, bufs, dst=dst)
        return torch.cat([bufs[r][: counts[r]] for r in range(world)], dim=0)
    dist.gather(send, None, dst=dst)
    return None
