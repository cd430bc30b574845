"""Multi-GPU plumbing: independent IQ channels, one per rank.

The receive chain has no cross-channel dependency, so there is no collective on the data path. NCCL (or
gloo in the CPU tests) is used only to (1) fan the input buffers out from the ingest rank, (2) agree on the
slowest rank's time and (3) gather the tiny per-rank results.
"""
from __future__ import annotations


def channels_of(rank: int, world: int, n_channels: int):
    """Round-robin channel ownership when there are more channels than ranks."""
    return [c for c in range(n_channels) if c % world == rank]


def fan_out(make_channel, rank: int, world: int, device, numel: int):
    """Rank 0 builds channel c with make_channel(c) -> (float32 tensor[numel], meta) and sends it to rank c.
    Returns this rank's (tensor, meta)."""
    if world == 1:
        return make_channel(0)
    import torch
    import torch.distributed as dist
    if rank == 0:
        metas = [None] * world
        mine = None
        for ch in range(world - 1, -1, -1):
            iq, meta = make_channel(ch)
            metas[ch] = meta
            if ch != 0:
                dist.send(iq, dst=ch)
                del iq
            else:
                mine = iq
    else:
        mine = torch.empty(numel, dtype=torch.float32, device=device)
        dist.recv(mine, src=0)
        metas = None
    out = [None]
    dist.scatter_object_list(out, metas, src=0)
    return mine, out[0]


def max_over_ranks(value: float, world: int, device) -> float:
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(value: int, world: int, device):
    """All ranks' integer results (e.g. frames decoded) on every rank."""
    if world == 1:
        return [int(value)]
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
