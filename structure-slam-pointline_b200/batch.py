"""Batched-frames mode across GPUs (north_star: frames are independent units; matching needs pairs (f, f+1)).

One process per GPU.  The global batch of `total` frames is cut into `world` contiguous blocks; rank r extracts its
block plus ONE halo frame (the first frame of block r+1, cyclically) so that every pair (f, f+1 mod total) is local to
exactly one rank.  The only collective is one all-gather of the fixed-size match tables (NCCL over NVLink on GPUs,
gloo in the CPU tests) — SURVEY.md 8(e).  No CPU compute happens here: the tables are produced by the CUDA path.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class FrameShard:
    total: int          # frames in the global batch
    world: int
    rank: int

    def __post_init__(self):
        if self.total % self.world != 0:
            raise ValueError("the batch must divide evenly over the ranks (fixed-size all-gather)")

    @property
    def per_rank(self):
        return self.total // self.world

    @property
    def start(self):
        return self.rank * self.per_rank

    def frame_ids(self):
        """Global ids of the frames this rank extracts: its block + the halo frame."""
        ids = [self.start + i for i in range(self.per_rank)]
        ids.append((self.start + self.per_rank) % self.total)
        return ids

    def pair_ids(self):
        """Global ids p of the pairs (p, p+1 mod total) matched on this rank (local pair i uses local frames i, i+1)."""
        return [self.start + i for i in range(self.per_rank)]


def all_gather_tables(local, group=None):
    """local: [per_rank, cap] int32 tensor (device for NCCL, CPU for gloo) -> [total, cap] on every rank,
    rows ordered by global pair id."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class PackedGather:
    """ONE all-gather per step for both match tables (north_star: "a single NCCL all-gather of match tables"): the point
    table [per_rank, cap] of this step and a line table [per_rank, nl] travel in one [per_rank, cap + nl] buffer.
    Buffers are allocated once; all copies are issued on the caller's current stream."""

    def __init__(self, per_rank, cap, nl, world, device, dtype=None):
        import torch
        dtype = dtype or torch.int32
        self.cap, self.nl, self.world, self.per_rank = cap, nl, world, per_rank
        self.local = torch.full((per_rank, cap + nl), -1, dtype=dtype, device=device)
        self.full = torch.empty((world * per_rank, cap + nl), dtype=dtype, device=device)

    def stage_lines(self, line_table):
        self.local[:, self.cap:].copy_(line_table, non_blocking=True)

    def gather(self, point_table, group=None):
        """-> (points [total, cap], lines [total, nl]) views of the gathered buffer, rows ordered by global pair id."""
        import torch.distributed as dist
        self.local[:, :self.cap].copy_(point_table, non_blocking=True)
        dist.all_gather_into_tensor(self.full, self.local, group=group)
        return self.full[:, :self.cap], self.full[:, self.cap:]
