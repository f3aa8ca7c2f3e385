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


class SpectraGatherer:
    """Gather of per-stream rows (e.g. [n_local, 1024] u8 spectra) to rank `dst`, with every buffer
    allocated once: a step costs exactly one collective and no allocation or reordering kernel.

    The receive buffer is rank-major, [world, n_max, ...]; `ordered()` puts it in global stream
    order on demand (stream s sits at [s mod world, s div world])."""

    def __init__(self, n_streams: int, world: int, rank: int, row_shape, dtype, device, dist=None, dst: int = 0):
        import torch
        self.n_streams, self.world, self.rank, self.dist, self.dst = n_streams, world, rank, dist, dst
        self.n_local = local_count(n_streams, world, rank)
        self.n_max = local_count(n_streams, world, 0)
        row_shape = tuple(row_shape)
        # the send buffer kernels write into directly (padded to n_max rows for the collective)
        self.send = torch.zeros((self.n_max,) + row_shape, dtype=dtype, device=device)
        self.local = self.send[:self.n_local]
        self.recv = None
        self.parts = None
        if world > 1 and rank == dst:
            self.recv = torch.zeros((world, self.n_max) + row_shape, dtype=dtype, device=device)
            self.parts = list(self.recv.unbind(0))

    def gather(self):
        if self.world > 1:
            self.dist.gather(self.send, self.parts, dst=self.dst)

    def ordered(self):
        """[n_streams, ...] in global stream order (on dst; None elsewhere)."""
        import torch
        if self.world == 1:
            return self.local
        if self.rank != self.dst:
            return None
        out = torch.empty((self.n_streams,) + tuple(self.recv.shape[2:]), dtype=self.recv.dtype, device=self.recv.device)
        for r in range(self.world):
            ids = streams_for_rank(self.n_streams, self.world, r)
            out[ids] = self.recv[r, :len(ids)]
        return out


def gather_spectra(local, n_streams: int, world: int, rank: int, dist=None, dst: int = 0):
    """One-shot convenience over SpectraGatherer: [n_local, ...] -> [n_streams, ...] on dst."""
    g = SpectraGatherer(n_streams, world, rank, local.shape[1:], local.dtype, local.device, dist=dist, dst=dst)
    g.local.copy_(local)
    g.gather()
    return g.ordered()


def max_over_ranks(value: float, world: int, dist=None, device="cpu") -> float:
    """Timing convention of bench.py: a step takes as long as its slowest rank."""
    if world == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
