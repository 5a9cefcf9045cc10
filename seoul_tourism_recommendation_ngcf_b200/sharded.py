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

    # node id <-> position in the rank-contiguous [N_pad] layout (identity here)
    def position(self, nodes: torch.Tensor) -> torch.Tensor:
        return nodes

    def node(self, positions: torch.Tensor) -> torch.Tensor:
        """Node id at each position, -1 for padding."""
        return torch.where(positions < self.N, positions, torch.full_like(positions, -1))


class DealtShards(RowShards):
    """Entry-balanced blocks: nodes are dealt round-robin,
    node i lives at position ``p(i) = (i mod W) * rows + i div W``, so rank r owns nodes r, r+W, r+2W, ... as local rows
    0, 1, 2, ... and one ``all_gather_into_tensor`` still assembles the whole ``[N_pad, d]`` layer in position order.
    Entries per rank, max / mean, 8 ranks at Gowalla shape: 1.21 for equal CONTIGUOUS blocks on the bench's shuffled
    synthetic ids and > 3 when ids are sorted by popularity (heavy rows first); 1.05 for dealt blocks either way.  Host logic only so far: the CUDA path
    (NGCF.shard) still uses RowShards — relabelling the batch row ids, table rows and RNG keys is the part to wire."""

    def __init__(self, N: int, world: int, rank: int):
        super().__init__(N, world, rank)
        self.valid = max(0, -(-(self.N - self.rank) // self.world))          # nodes r, r+W, ... below N

    def bounds(self, rank: int):
        r0 = rank * self.rows
        return r0, r0 + max(0, -(-(self.N - rank) // self.world))

    def position(self, nodes: torch.Tensor) -> torch.Tensor:
        return (nodes % self.world) * self.rows + nodes // self.world

    def node(self, positions: torch.Tensor) -> torch.Tensor:
        i = (positions % self.rows) * self.world + positions // self.rows
        return torch.where(i < self.N, i, torch.full_like(i, -1))


def shard_coo(L: torch.Tensor, sh: RowShards):
    """Entries of the row shard of L and of L^T as (row_local, col_global, coo_position) triples.

    forward  side: rows [r0, r1) of L        -> (i - r0, j) for entries with i in the block
    backward side: rows [r0, r1) of L^T      -> (j - r0, i) for entries with j in the block
    Both are ``rows x N_pad`` matrices; for a symmetric L they are the same matrix."""
    idx = L._indices()
    row, col = sh.position(idx[0]), sh.position(idx[1])          # identity unless the nodes are dealt (DealtShards)
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


def parity_vs_unsharded(emb: int, layers: list, L: torch.Tensor, num_dict: dict, batch: dict, batch_size: int, device,
                        node_p: float = 0.3, mess_p: float = 0.1, weight_decay: float = 0.025, group=None) -> dict:
    """Runs ONE training step (forward + BPR + backward, node and message dropout ON, device RNG) twice on this rank's
    GPU — unsharded, and row-sharded over the ranks of ``group`` — and returns the relative differences
    ``{out, loss, all_E, worst_grad}`` (max|a-b| / max|b|).  The dropout keys are global coordinates, so both runs draw
    the same masks and the sharded step must reproduce the 1-GPU step (reference semantics: NGCF.py:102-156 computed
    once over the whole graph).  Collective: every rank of the group must call it with the same arguments."""
    from .NGCF import NGCF
    from .bprloss import BPR
    res = []
    b = {k: (v if k == "year" else v.to(device)) for k, v in batch.items()}
    for sharded in (False, True):
        torch.manual_seed(0)
        m = NGCF(emb, list(layers), node_p, [mess_p] * len(layers), 1.0, [L, L], num_dict, batch_size, device).to(device)
        if sharded:
            m.shard(group)
        m.train()
        torch.manual_seed(7)
        uu, pp, nn_ = m(b["year"], b["u_id"], b["age"], b["sex"], b["month"], b["day"], b["dow"], b["pos_item"],
                        b["neg_item"], True)
        loss = BPR(weight_decay, batch_size)(uu, pp, nn_)
        loss.backward()
        torch.cuda.synchronize(device)
        res.append((uu.detach(), float(loss), {k: p.grad for k, p in m.named_parameters() if p.grad is not None},
                    m.all_users_emb.clone(), m.all_items_emb.clone()))
        del m

    def rel(a, c):
        return float((a - c).abs().max() / c.abs().max().clamp_min(1e-30))
    (u0, l0, g0, a0, i0), (u1, l1, g1, a1, i1) = res
    return {"out": rel(u1, u0), "loss": abs(l1 - l0) / max(abs(l0), 1e-30),
            "all_E": max(rel(a1, a0), rel(i1, i0)), "worst_grad": max(rel(g1[k], g0[k]) for k in g0),
            "loss_value": l0, "world": dist.get_world_size(group) if dist.is_initialized() else 1}
