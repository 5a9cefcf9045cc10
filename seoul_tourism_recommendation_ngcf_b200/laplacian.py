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


def build_lap_list(years, users, items, ratings, n_user: int, n_item: int) -> list:
    """Year loop of Matrix.create_matrix (matrix.py:41-67): R is never reset, so each year's graph is the
    previous one overwritten by that year's ratings; the slot is ``year % 18``."""
    years = np.asarray(years)
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    ratings = np.asarray(ratings, dtype=np.float32)
    uniq = list(dict.fromkeys(years.tolist()))
    lap_list = [[] for _ in uniq]
    acc = {}                                                    # (user*n_item+item) -> rating, later rows win
    keys_all = users * n_item + items
    for y in uniq:
        m = years == y
        for k, r in zip(keys_all[m].tolist(), ratings[m].tolist()):
            acc[k] = r
        keys = np.fromiter(acc.keys(), dtype=np.int64, count=len(acc))
        vals = np.fromiter(acc.values(), dtype=np.float32, count=len(acc))
        lap_list[y % 18] = laplacian_coo(keys // n_item, keys % n_item, vals, n_user, n_item)
    return lap_list
