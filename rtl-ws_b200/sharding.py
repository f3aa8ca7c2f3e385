"""Stream sharding across ranks.

Every piece of DSP state is per stream (rf_decimator.c:23,28; audio_main.c:77-79), so the
path shards by stream with no collective on the data path: stream s lives on rank s mod G.
The only exchange is at the end, when one consumer wants every stream's averaged u8 dB
spectrum (the 1 KiB the reference ships to its UI client per update): a gather to rank 0
over the process group (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations


def streams_for_rank(n_streams: int, world: int, rank: int) -> list[int]:
    """Global stream ids owned by `rank`: s mod world == rank, ascending."""
    return list(range(rank, n_streams, world))


def local_count(n_streams: int, world: int, rank: int) -> int:
    return len(range(rank, n_streams, world))


def gather_spectra(local, n_streams: int, world: int, rank: int, dist=None, dst: int = 0):
    """Gather per-stream rows to `dst` and put them back in global stream order.

    local: [n_local, ...] tensor of this rank's streams (ascending global id).  Ranks may own
    different counts when world does not divide n_streams; rows are padded to the maximum for
    the collective.  Returns the [n_streams, ...] tensor on `dst`, None elsewhere."""
    import torch
    if world == 1:
        return local
    n_max = local_count(n_streams, world, 0)
    if local.shape[0] < n_max:
        pad = torch.zeros((n_max - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, parts, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_streams,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = streams_for_rank(n_streams, world, r)
        out[ids] = parts[r][:len(ids)]
    return out


def max_over_ranks(value: float, world: int, dist=None, device="cpu") -> float:
    """Timing convention of bench.py: a step takes as long as its slowest rank."""
    if world == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
