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

    The collective runs on its OWN stream with two buffer sets: the producing stream only records an event and moves on, so
    ranks are no longer lock-stepped by the all-gather (the gathered tables of step i are complete one step later; `wait()`
    joins them into the caller's stream).  Buffers are allocated once."""

    def __init__(self, per_rank, cap, nl, world, device, dtype=None, own_stream=True, nbuf=2):
        import torch
        dtype = dtype or torch.int32
        self.cap, self.nl, self.world, self.per_rank = cap, nl, world, per_rank
        self.nbuf = max(2, int(nbuf)) if own_stream else 1     # buffer sets: a rank may run nbuf - 1 steps ahead of the slowest all-gather
        self.locals = [torch.full((per_rank, cap + nl), -1, dtype=dtype, device=device) for _ in range(self.nbuf)]
        self.fulls = [torch.empty((world * per_rank, cap + nl), dtype=dtype, device=device) for _ in range(self.nbuf)]
        self.cuda = own_stream and device is not None and getattr(device, "type", str(device)) == "cuda"
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = [torch.cuda.Event() for _ in range(self.nbuf)] if self.cuda else None      # all-gather of buffer b finished
        self.b = 0
        self.last = None

    # kept for callers that look at the buffers directly
    @property
    def local(self):
        return self.locals[self.b]

    @property
    def full(self):
        return self.fulls[self.last if self.last is not None else 0]

    def _reuse(self, b):
        """The caller's stream may write locals[b] only after the collective that last read it has finished."""
        if self.cuda and not getattr(self, "_fresh", None) == b:
            import torch
            torch.cuda.current_stream().wait_event(self.done[b])
            self._fresh = b

    def stage_lines(self, line_table):
        """Line table of a finished ring slot: rides along with the next gather() (copied now, on the caller's stream)."""
        self._reuse(self.b)
        self.locals[self.b][:, self.cap:].copy_(line_table, non_blocking=True)

    def gather(self, point_table, group=None):
        """Start the all-gather of (point_table, staged line table).  Returns views of the gathered buffer; with the own stream
        they are valid after wait()."""
        import torch
        import torch.distributed as dist
        b = self.b
        self._reuse(b)
        self.locals[b][:, :self.cap].copy_(point_table, non_blocking=True)
        if self.cuda:
            cur = torch.cuda.current_stream()
            ready = torch.cuda.Event(); ready.record(cur)
            self.comm.wait_event(ready)
            with torch.cuda.stream(self.comm):
                dist.all_gather_into_tensor(self.fulls[b], self.locals[b], group=group)
                self.done[b].record(self.comm)
        else:
            dist.all_gather_into_tensor(self.fulls[b], self.locals[b], group=group)
        self._fresh = None
        self.last = b
        self.b = (b + 1) % self.nbuf
        return self.fulls[b][:, :self.cap], self.fulls[b][:, self.cap:]

    def wait(self):
        """Make the caller's current stream wait for every all-gather issued so far."""
        if self.cuda:
            import torch
            cur = torch.cuda.current_stream()
            for e in self.done:
                cur.wait_event(e)
