"""Laplacians built on the device, straight into CSR (BASELINE.json config 5).

The reference's ``Matrix.create_matrix`` (matrix.py:41-83) goes through dense N x N arrays; round 1's sparse builder still
emitted the reference's 20-byte-per-entry int64 COO on the host.  Neither can hold the 1 B-edge graph (40 GB of COO).
Here a rank generates (``ngcf_plgraph_entries``) or receives the adjacency entries of ITS rows only, as 64-bit keys
``row_local << 32 | col``; one device sort + dedup turns them into the shard's CSR, degrees are exchanged once, and the
values ``L[r, c] = deg(r)^-1/2 * a * deg(c)^-1/2`` (count degrees, matrix.py:55-62; a = 1 for these implicit graphs) are
evaluated in float64 and stored as float32 like the reference does.  No COO, no host pass, no Python loop over years or
tiles."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


class CsrLaplacian:
    """One ``lap_list`` element that already is a CSR row shard on the device (what ``NGCF._plan`` accepts beside the
    reference's ``torch.sparse_coo``): rows [row0, row0 + n_rows) of the symmetric N x N Laplacian, global column ids."""
    is_sparse = False

    def __init__(self, rowptr, colidx, vals, N: int, row0: int = 0):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.shape = (int(N), int(N))
        self.row0 = int(row0)
        self.n_rows = int(rowptr.numel()) - 1
        self.nnz = int(colidx.numel())

    def _nnz(self):
        return self.nnz


def powerlaw_entries(n_user: int, n_item: int, n_edges: int, alpha: float, seed: int, row0: int, n_rows: int, device):
    """Sorted, deduplicated int64 keys ``row_local << 32 | col`` of the rows [row0, row0 + n_rows)."""
    lib = _lib.load()
    total = torch.zeros(1, dtype=torch.int64, device=device)
    st = _lib.current_stream()
    _lib.check(lib.ngcf_plgraph_entries(n_user, n_item, n_edges, float(alpha), int(seed), int(row0), int(n_rows),
                                        total.data_ptr(), None, 0, st), "plgraph_entries(count)")
    n = int(total)
    keys = torch.empty(n, dtype=torch.int64, device=device)
    total.zero_()
    _lib.check(lib.ngcf_plgraph_entries(n_user, n_item, n_edges, float(alpha), int(seed), int(row0), int(n_rows),
                                        total.data_ptr(), keys.data_ptr(), n, st), "plgraph_entries(fill)")
    keys = torch.sort(keys).values
    return torch.unique_consecutive(keys)


def laplacian_from_keys(keys: torch.Tensor, N: int, row0: int, n_rows: int, group=None, world: int = 1,
                        shard=None) -> CsrLaplacian:
    """CSR row shard of D^-1/2 A D^-1/2 from the shard's sorted unique adjacency keys (a_ij = 1)."""
    dev = keys.device
    row = keys >> 32
    col = (keys & 0xFFFFFFFF).to(torch.int32)
    del keys
    deg_local = torch.bincount(row, minlength=n_rows)[:n_rows]
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(deg_local, 0)
    if world > 1:                                        # degrees of every node: one all-gather of the row blocks
        sizes = [shard.bounds(r)[1] - shard.bounds(r)[0] for r in range(world)] if shard is not None else [n_rows] * world
        pad = max(max(sizes), n_rows)
        mine = torch.zeros(pad, dtype=deg_local.dtype, device=dev)
        mine[:n_rows] = deg_local
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        deg = torch.cat([p_[:n] for p_, n in zip(parts, sizes)])
        if deg.numel() < N + 1:                          # (equal blocks pad the last one)
            deg = torch.cat([deg, torch.zeros(N + 1 - deg.numel(), dtype=deg.dtype, device=dev)])
    else:
        deg = deg_local
    d_sqrt = torch.pow(deg.to(torch.float32), -0.5)      # float32 d^-1/2 with inf -> 0 (matrix.py:56-57)
    d_sqrt[torch.isinf(d_sqrt)] = 0.0
    # d_i * (a_ij * d_j) in float64, cast to float32 (matrix.py:58-62,82), in chunks to bound the float64 temporaries
    vals = torch.empty(col.numel(), dtype=torch.float32, device=dev)
    step = 1 << 27
    for s in range(0, col.numel(), step):
        e = min(s + step, col.numel())
        r = row[s:e] + row0
        vals[s:e] = (d_sqrt[r].double() * d_sqrt[col[s:e].long()].double()).float()
    return CsrLaplacian(rowptr.to(torch.int32), col, vals, N, row0)


def powerlaw_laplacian(n_user: int, n_item: int, n_edges: int, device, alpha: float = 0.8, seed: int = 0, shard=None,
                       group=None) -> CsrLaplacian:
    """The Laplacian (row shard, if ``shard`` is a sharded.RowShards) of the synthetic power-law graph of that shape."""
    N = n_user + n_item
    if shard is None:
        row0, n_rows, world = 0, N, 1
    else:
        row0, n_rows, world = shard.r0, shard.rows, shard.world
    keys = powerlaw_entries(n_user, n_item, n_edges, alpha, seed, row0, n_rows, device)
    return laplacian_from_keys(keys, N, row0, n_rows, group, world, shard)


def laplacian_csr_device(users, items, ratings, n_user: int, n_item: int, device) -> CsrLaplacian:
    """One year's Laplacian (matrix.py:48-67) from distinct (user, item, rating) triples, built on the device straight
    into CSR: both directions of every non-zero rating as keys ``row << 32 | col``, one sort, count degrees
    (matrix.py:55), values ``d_i * (a_ij * d_j)`` evaluated in float64 and stored as float32 (matrix.py:56-62,82).
    Same structure as ``laplacian.laplacian_coo`` and values within 5e-7 (numpy's float32 power is not correctly
    rounded)."""
    import numpy as np
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("laplacian_csr_device builds on a CUDA device only")
    u = torch.as_tensor(np.asarray(users, dtype=np.int64)).to(device)
    i = torch.as_tensor(np.asarray(items, dtype=np.int64)).to(device) + n_user
    r = torch.as_tensor(np.asarray(ratings, dtype=np.float32)).to(device)
    nz = r != 0                                                  # zero ratings are not edges (dok_matrix drops them)
    u, i, r = u[nz], i[nz], r[nz]
    N = n_user + n_item
    keys = torch.cat([(u << 32) | i, (i << 32) | u])
    keys, order = torch.sort(keys)
    a = torch.cat([r, r])[order]
    row, col = keys >> 32, (keys & 0xFFFFFFFF)
    deg = torch.bincount(row, minlength=N)
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(deg, 0)
    d_sqrt = torch.pow(deg.to(torch.float32), -0.5)
    d_sqrt[torch.isinf(d_sqrt)] = 0.0
    vals = (d_sqrt[row].double() * (a.double() * d_sqrt[col].double())).float()
    keep = vals != 0
    if not bool(keep.all()):                                     # (a product that underflows to 0 is no entry either)
        row, col, vals = row[keep], col[keep], vals[keep]
        rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=N), 0)
    return CsrLaplacian(rowptr.to(torch.int32), col.to(torch.int32), vals, N, 0)
