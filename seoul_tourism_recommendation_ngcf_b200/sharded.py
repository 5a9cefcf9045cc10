"""Row partition of the propagation across the GPUs of one box (BASELINE.json north_star).

Rank r owns a contiguous block of rows of L and of every per-layer tensor (E_k, S_k, gradients): equal blocks
(``RowShards``; N is padded to ``world * rows`` with empty rows) or blocks cut by work (``BalancedShards``).  The
reference has no multi-device code at all; this is the design the contract prescribes.  Per layer and direction every
rank needs every rank's rows of one ``[N_pad, d]`` matrix; the table gradient has to be complete everywhere (the two
embedding tables stay replicated so ``state_dict`` is unchanged) and the W/b gradients are sums over the blocks.
``PeerExchange`` moves all of that with peer-memory stores over NVLink (``ngcf_push_rows`` on symmetric memory, no
library collective); ``all_gather_rows`` / ``dist.all_reduce`` are the NCCL fallback for equal blocks.  The shard
descriptors are pure host logic (CPU-testable).
"""
from __future__ import annotations

import os

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
        self.equal = True

    def bounds(self, rank: int):
        r0 = rank * self.rows
        return r0, max(r0, min(self.N, r0 + self.rows))

    # node id <-> position in the rank-contiguous [N_pad] layout (identity here)
    def position(self, nodes: torch.Tensor) -> torch.Tensor:
        return nodes

    def node(self, positions: torch.Tensor) -> torch.Tensor:
        """Node id at each position, -1 for padding."""
        return torch.where(positions < self.N, positions, torch.full_like(positions, -1))


class BalancedShards(RowShards):
    """Contiguous row blocks of UNEQUAL length, cut so that every rank gets the same share of the work
    ``entries + row_weight * rows`` (the SpMM cost follows the entries, the dense kernels the rows; item rows of a
    bipartite interaction graph are several times heavier than user rows, so equal blocks leave the item ranks with
    most of the entries).  No padding: N_pad = N.  Needs the peer-memory exchange (sharded.PeerExchange): NCCL's
    all_gather_into_tensor wants equal blocks."""

    def __init__(self, N: int, world: int, rank: int, starts):
        if not (0 <= rank < world):
            raise ValueError("rank out of range")
        starts = [int(x) for x in starts]
        if len(starts) != world + 1 or starts[0] != 0 or starts[-1] != N or any(b < a for a, b in zip(starts, starts[1:])):
            raise ValueError("starts must be world + 1 non-decreasing row indices from 0 to N")
        self.N, self.world, self.rank = int(N), int(world), int(rank)
        self.starts = starts
        self.r0 = starts[rank]
        self.rows = starts[rank + 1] - starts[rank]
        self.valid = self.rows
        self.N_pad = self.N
        self.equal = False

    def bounds(self, rank: int):
        return self.starts[rank], self.starts[rank + 1]

    @staticmethod
    def cut(row_work, world: int):
        """Block boundaries from a per-row work array (numpy / torch, length N): prefix sums cut into equal shares."""
        import numpy as np
        w = np.asarray(row_work.cpu() if hasattr(row_work, "cpu") else row_work, dtype=np.float64)
        c = np.concatenate([[0.0], np.cumsum(w)])
        targets = c[-1] * np.arange(1, world) / world
        inner = np.searchsorted(c, targets, side="left")
        return [0] + [int(x) for x in np.clip(inner, 0, w.size)] + [int(w.size)]

    @staticmethod
    def cut_bipartite(n_user: int, n_item: int, n_edges: int, world: int, row_weight: float = 32.0):
        """The same for a bipartite graph whose users (items) all have the class's mean degree - exact in expectation
        for the synthetic power-law graphs, whose Zipf ranks are scattered uniformly over each id range."""
        wu, wi = n_edges / n_user + row_weight, n_edges / n_item + row_weight
        total = wu * n_user + wi * n_item
        starts = [0]
        for r in range(1, world):
            t = total * r / world
            row = t / wu if t <= wu * n_user else n_user + (t - wu * n_user) / wi
            starts.append(int(round(row)))
        return starts + [n_user + n_item]


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


class PeerExchange:
    """All-gathers of row blocks by peer-memory stores (``ngcf_push_rows``) instead of NCCL collectives.

    Every exchanged matrix lives in symmetric memory (``torch.distributed._symmetric_memory``: the same allocation
    mapped into every rank of the box over NVLink); the producing kernel writes this rank's rows straight into the
    local copy, one launch then stores them into every peer's copy and returns once the peers' rows have landed here.
    Buffers are persistent (one per key), so a captured CUDA graph replays against fixed addresses; reuse across steps is
    safe because a peer only stores into a copy after its owner has reached the same exchange (see exchange.cu)."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._symm = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = torch.device(device)
        lib = _lib.load()
        enable = getattr(symm_mem, "enable_symm_mem_for_group", None)
        if enable is not None:
            try:
                enable(self.group.group_name)
            except Exception:
                pass
        self.flags = symm_mem.empty(lib.ngcf_exchange_flag_words(), dtype=torch.int32, device=self.device)
        self.flags.zero_()
        self._flag_hdl = symm_mem.rendezvous(self.flags, self.group)
        self._flag_ptrs = _lib.ptr_array_int([int(p) for p in self._flag_hdl.buffer_ptrs])
        self.local_state = torch.zeros(2, dtype=torch.int32, device=self.device)
        self._mats = {}
        # NVLS multicast (one multimem.st per element, replicated by the switch): the sender's egress shrinks (W-1)-fold but
        # the switch also returns the sender its own rows - measured slower at 2 ranks (426 vs 231 us for 128 MB), faster
        # inside the 8-rank step (75 vs 85 ms at pl-1b).  NGCF_B200_MULTICAST=0/1 overrides.
        mc_env = os.environ.get("NGCF_B200_MULTICAST", "auto")
        self.use_multicast = mc_env == "1" or (mc_env == "auto" and self.world >= 4)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                    # every rank's flag block is zero before anyone signals

    def multicast(self) -> bool:
        return any(m[3] for m in self._mats.values())

    def matrix(self, key, n_rows: int, d: int) -> torch.Tensor:
        """The persistent symmetric [n_rows, d] fp32 matrix of ``key`` (allocated collectively on first use)."""
        ent = self._mats.get(key)
        if ent is None:
            t = self._symm.empty(n_rows * d, dtype=torch.float32, device=self.device)
            t.zero_()
            hdl = self._symm.rendezvous(t, self.group)
            from . import _lib
            ptrs = _lib.ptr_array_int([int(p) for p in hdl.buffer_ptrs])
            mc = 0
            if self.use_multicast:
                try:
                    mc = int(hdl.multicast_ptr) if hdl.has_multicast_support(self.device.type, self.device.index) else 0
                except Exception:
                    mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)                # zero-filled everywhere before the first stores arrive
            ent = self._mats[key] = (t.view(n_rows, d), hdl, ptrs, mc)
        if ent[0].shape != (n_rows, d):
            raise RuntimeError(f"exchange matrix {key!r} was created as {tuple(ent[0].shape)}, asked for {(n_rows, d)}")
        return ent[0]

    def push(self, key, row0: int, n_rows: int):
        """This rank's rows [row0, row0 + n_rows) of matrix ``key`` -> every peer; returns (in stream order) with every
        peer's rows present in the local copy."""
        from . import _lib
        mat, _, ptrs, mc = self._mats[key]
        lib = _lib.load()
        _lib.check(lib.ngcf_push_rows(ptrs, self._flag_ptrs, self.local_state.data_ptr(), self.world, self.rank,
                                      int(row0), int(n_rows), int(mat.shape[1]), mc or None, _lib.current_stream()),
                   "push_rows")

    def push_selected(self, key, row0: int, n_rows: int, row_lists, offsets):
        """Only the rows ``row_lists[q] + offsets[q]`` (int64 CUDA tensors of node ids) that this rank owns -> every peer."""
        from . import _lib
        mat, _, ptrs, mc = self._mats[key]
        lib = _lib.load()
        _lib.check(lib.ngcf_push_selected_rows(ptrs, self._flag_ptrs, self.local_state.data_ptr(), self.world, self.rank,
                                               int(row0), int(n_rows), int(mat.shape[1]), _lib.ptr_array(row_lists),
                                               _lib.i64_array(offsets), _lib.i64_array([t.numel() for t in row_lists]),
                                               len(row_lists), mc or None, _lib.current_stream()), "push_selected_rows")


def parity_vs_unsharded(emb: int, layers: list, L, num_dict: dict, batch: dict, batch_size: int, device,
                        node_p: float = 0.3, mess_p: float = 0.1, weight_decay: float = 0.025, group=None,
                        L_shard=None, shards=None) -> dict:
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
        Lm = L_shard if (sharded and L_shard is not None) else L    # (a device-built CSR row shard, plgraph.py)
        m = NGCF(emb, list(layers), node_p, [mess_p] * len(layers), 1.0, [Lm, Lm], num_dict, batch_size, device).to(device)
        if sharded:
            m.shard(group, shards)
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
