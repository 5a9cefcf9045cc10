"""Sparse construction of ``lap_list`` in the reference's in-memory / pickle format.

The reference builds L = D^-1/2 A D^-1/2 through dense N x N arrays (matrix.py:55-62: O(N^2) memory, O(N^3)
time), which cannot produce any of the named benchmark shapes.  This is the same arithmetic restricted to the
non-zeros: count-degree D (matrix.py:55), float32 d^-1/2 with inf -> 0 (matrix.py:56-57), the product
d_i * (a_ij * d_j) evaluated in float64 (matrix.py:58-62) and cast to float32 (matrix.py:82), emitted as the
uncoalesced int64/fp32 ``torch.sparse_coo`` with row-major sorted indices that ``NGCF.forward`` consumes
(matrix.py:79-83, NGCF.py:117-118).  Host-side, one-off, not on the timed path.
"""
from __future__ import annotations

import numpy as np
import torch


def laplacian_coo(users, items, ratings, n_user: int, n_item: int) -> torch.Tensor:
    """One year's Laplacian from distinct (user, item, rating) triples; zero ratings are not edges."""
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    ratings = np.asarray(ratings, dtype=np.float32)
    nz = ratings != 0
    users, items, ratings = users[nz], items[nz], ratings[nz]
    N = n_user + n_item
    row = np.concatenate([users, items + n_user])
    col = np.concatenate([items + n_user, users])
    a = np.concatenate([ratings, ratings])
    order = np.lexsort((col, row))
    row, col, a = row[order], col[order], a[order]
    deg = np.bincount(row, minlength=N)
    with np.errstate(divide="ignore"):
        d_sqrt = np.power(deg, -0.5, dtype=np.float32)
    d_sqrt[np.isinf(d_sqrt)] = 0.0
    v = d_sqrt[row].astype(np.float64) * (a.astype(np.float64) * d_sqrt[col].astype(np.float64))
    keep = v != 0.0
    idx = torch.from_numpy(np.stack([row[keep], col[keep]]))
    val = torch.from_numpy(v[keep].astype(np.float32))
    return torch.sparse_coo_tensor(idx, val, (N, N), is_coalesced=False, check_invariants=False)


def laplacian_coo_device(users, items, ratings, n_user: int, n_item: int, device) -> torch.Tensor:
    """Same result as ``laplacian_coo`` built on the GPU (``ngcf_laplacian_entries``: degree count + normalisation of
    the 2·nnz entries in two launches; the row-major order of matrix.py:79-83 by one device sort).  Values agree with the
    host builder's to 5e-7 relative (numpy's SIMD float32 power is not correctly rounded; the device value is).  CUDA only."""
    from . import _lib
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("laplacian_coo_device builds on a CUDA device only; use laplacian_coo on the host")
    lib = _lib.load()
    u = torch.as_tensor(np.asarray(users, dtype=np.int64)).to(device)
    i = torch.as_tensor(np.asarray(items, dtype=np.int64)).to(device)
    r = torch.as_tensor(np.asarray(ratings, dtype=np.float32)).to(device)
    nz = r != 0                                                  # zero ratings are not edges (dok_matrix drops them)
    u, i, r = u[nz].contiguous(), i[nz].contiguous(), r[nz].contiguous()
    n, N = u.numel(), n_user + n_item
    deg = torch.empty(N, dtype=torch.int32, device=device)
    row = torch.empty(2 * n, dtype=torch.int64, device=device)
    col = torch.empty(2 * n, dtype=torch.int64, device=device)
    val = torch.empty(2 * n, dtype=torch.float32, device=device)
    _lib.check(lib.ngcf_laplacian_entries(u.data_ptr(), i.data_ptr(), r.data_ptr(), n, n_user, n_item, deg.data_ptr(),
                                          row.data_ptr(), col.data_ptr(), val.data_ptr(), _lib.current_stream()),
               "laplacian_entries")
    keep = val != 0
    row, col, val = row[keep], col[keep], val[keep]
    order = torch.argsort(row * N + col)
    idx = torch.stack([row[order], col[order]])
    return torch.sparse_coo_tensor(idx, val[order], (N, N), is_coalesced=False, check_invariants=False)


def accumulate_years(years, users, items, ratings, n_item: int):
    """The year loop of Matrix.create_matrix (matrix.py:41-47) without a per-edge Python loop: R is never reset, so
    after year y it holds every (user, item) seen so far with the LAST rating written (rows in frame order, years in
    first-seen order).  Yields (year, keys = user * n_item + item, ratings) of the accumulated R after each year."""
    years = np.asarray(years)
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    ratings = np.asarray(ratings, dtype=np.float32)
    keys_all = users * n_item + items
    acc_k = np.empty(0, dtype=np.int64)
    acc_v = np.empty(0, dtype=np.float32)
    for y in dict.fromkeys(years.tolist()):
        m = years == y
        k = np.concatenate([acc_k, keys_all[m]])               # later positions win: the year's rows come after R
        v = np.concatenate([acc_v, ratings[m]])
        _, last = np.unique(k[::-1], return_index=True)        # first occurrence in the reversed array = last write
        pos = np.sort(k.size - 1 - last)                       # keep dict insertion order irrelevant: sorted by position
        acc_k, acc_v = k[pos], v[pos]
        yield y, acc_k, acc_v


def build_lap_list(years, users, items, ratings, n_user: int, n_item: int, device=None, fmt: str = "coo") -> list:
    """Year loop of Matrix.create_matrix (matrix.py:41-67): R is never reset, so each year's graph is the
    previous one overwritten by that year's ratings; the slot is ``year % 18``.  fmt="coo": the reference's
    ``torch.sparse_coo`` elements (host builder, or the device builder when ``device`` is CUDA); fmt="csr": device-built
    ``plgraph.CsrLaplacian`` elements - the kernels' layout directly, no COO in between (needs a CUDA ``device``)."""
    uniq = list(dict.fromkeys(np.asarray(years).tolist()))
    lap_list = [[] for _ in uniq]
    for y, keys, vals in accumulate_years(years, users, items, ratings, n_item):
        if fmt == "csr":
            from .plgraph import laplacian_csr_device
            lap_list[y % 18] = laplacian_csr_device(keys // n_item, keys % n_item, vals, n_user, n_item, device)
        elif device is not None and torch.device(device).type == "cuda":
            lap_list[y % 18] = laplacian_coo_device(keys // n_item, keys % n_item, vals, n_user, n_item, device)
        else:
            lap_list[y % 18] = laplacian_coo(keys // n_item, keys % n_item, vals, n_user, n_item)
    return lap_list
