"""Row partition of the propagation across the GPUs of one box (BASELINE.json north_star).

Rank r owns the equal block of rows [r*rows, (r+1)*rows) of L and of every per-layer tensor (E_k, S_k, gradients);
N is padded to ``world * rows`` with empty rows so one ``all_gather_into_tensor`` assembles a full, contiguous
``[N_pad, d]`` tensor whose row index is the global node id.  The reference has no multi-device code at all; this
is the design the contract prescribes: per layer one all-gather of the E shards (forward) and of the gS shards
(backward), an all-reduce of the W/b gradients and an all-gather of the table-gradient shards (the two embedding
tables stay replicated so ``state_dict`` is unchanged).  The helpers in this file are pure host logic (CPU-testable).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class RowShards:
    """Equal contiguous row blocks.  ``rows`` per rank, ``N_pad = world * rows`` >= N."""

    def __init__(self, N: int, world: int, rank: int):
        if not (0 <= rank < world):
            raise ValueError("rank out of range")
        self.N, self.world, self.rank = int(N), int(world), int(rank)
        self.rows = -(-self.N // self.world)
        self.N_pad = self.rows * self.world
        self.r0 = self.rank * self.rows
        self.valid = max(0, min(self.N, self.r0 + self.rows) - self.r0)      # rows of this rank that exist in L

    def bounds(self, rank: int):
        r0 = rank * self.rows
        return r0, max(r0, min(self.N, r0 + self.rows))


def shard_coo(L: torch.Tensor, sh: RowShards):
    """Entries of the row shard of L and of L^T as (row_local, col_global, coo_position) triples.

    forward  side: rows [r0, r1) of L        -> (i - r0, j) for entries with i in the block
    backward side: rows [r0, r1) of L^T      -> (j - r0, i) for entries with j in the block
    Both are ``rows x N_pad`` matrices; for a symmetric L they are the same matrix."""
    idx = L._indices()
    row, col = idx[0], idx[1]
    r0, r1 = sh.r0, sh.r0 + sh.rows
    pos = torch.arange(row.numel(), device=row.device)
    mf = (row >= r0) & (row < r1)
    mb = (col >= r0) & (col < r1)
    fwd = (row[mf] - r0, col[mf], pos[mf])
    bwd = (col[mb] - r0, row[mb], pos[mb])
    return fwd, bwd


def all_gather_rows(out: torch.Tensor, shard: torch.Tensor, group=None):
    """out[rank*rows : (rank+1)*rows] = shard of every rank (out: [world*rows, d], shard: [rows, d])."""
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, shard, group=group)
    else:                                   # gloo (CPU tests): list form
        world = dist.get_world_size(group)
        rows = shard.shape[0]
        dist.all_gather([out[r * rows:(r + 1) * rows] for r in range(world)], shard, group=group)
    return out
