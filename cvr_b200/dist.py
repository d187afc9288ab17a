"""Multi-GPU host logic: one process per GPU, row shards by nnz balance, one collective.

The matrix is cut into contiguous row ranges with (nearly) equal nnz (cvr_b200.shard); rank g
converts its re-based shard to its own CVR, keeps a full replicated x and writes only its y
range.  A single SpMV needs no communication.  Iterated SpMV (x <- y, square A) needs exactly
one exchange per iteration: an all-gather of the y shards into every rank's x.  The reference
has no counterpart (single process, shared memory) -- this is the north star's
multi-GPU extension.

torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo on CPU for the tests); the
local SpMV is whatever callable the caller passes (CvrMatrix.spmv_device on GPUs).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class RowShardExchange:
    """y shards -> replicated x.  `cuts` are the 1-based row cut points of shard.partition_rows_by_nnz
    (cuts[g] .. cuts[g+1]-1 belong to rank g); x has n_rows+1 entries, x[0] is the phantom."""

    def __init__(self, cuts, rank: int, world: int, device, group=None):
        self.cuts = [int(c) for c in cuts]
        self.rank, self.world, self.group = rank, world, group
        self.counts = [self.cuts[g + 1] - self.cuts[g] for g in range(world)]
        self.n_local = self.counts[rank]
        self.n_rows = self.cuts[-1] - 1
        self.equal = len(set(self.counts)) == 1
        self.max_count = max(self.counts)
        if not self.equal:
            # all-gather needs equal contributions: pad every shard to the longest one and compact
            # afterwards (the all-gather-v idiom of SURVEY.md 8e); still one collective per iteration
            self.send = torch.zeros(self.max_count, dtype=torch.float64, device=device)
            self.stage = torch.zeros(world * self.max_count, dtype=torch.float64, device=device)

    def bytes_received_per_rank(self) -> int:
        return 8 * (self.n_rows - self.n_local)

    def __call__(self, y_local: torch.Tensor, x: torch.Tensor) -> None:
        """y_local: this rank's y (n_local+1 entries, [0] phantom); x: replicated vector to rebuild."""
        if self.world == 1:
            x[1:1 + self.n_rows].copy_(y_local[1:1 + self.n_local])
            return
        if self.equal:
            dist.all_gather_into_tensor(x[1:1 + self.n_rows], y_local[1:1 + self.n_local], group=self.group)
            return
        self.send[:self.n_local].copy_(y_local[1:1 + self.n_local])
        dist.all_gather_into_tensor(self.stage, self.send, group=self.group)
        for g in range(self.world):
            x[self.cuts[g]:self.cuts[g + 1]].copy_(self.stage[g * self.max_count:g * self.max_count + self.counts[g]])


def iterate(local_spmv, exchange: RowShardExchange, x: torch.Tensor, y_local: torch.Tensor, iters: int) -> None:
    """`iters` iterations of x <- A x on row shards: local SpMV into y_local, then the exchange."""
    for _ in range(iters):
        local_spmv(x, y_local)
        exchange(y_local, x)
