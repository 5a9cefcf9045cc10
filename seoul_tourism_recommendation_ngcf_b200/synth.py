"""Seeded synthetic inputs of the shapes BASELINE.json names.

Host-side numpy only; nothing here is on the timed path.  The bipartite user-item graph has
Zipf-distributed user activity and item popularity (P(rank r) ~ r^-alpha), sampled without
replacement until exactly ``n_edges`` distinct (user, item) pairs exist.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: (n_user, n_item, n_interactions, emb, layers)
    "seoul": (5840, 100, 394236, 65, 2),          # ~ nnz(L) 0.79 M
    "gowalla": (29858, 40981, 1027370, 64, 3),
    "yelp2018": (31668, 38048, 1561406, 64, 3),
    "amazon-book": (52643, 91599, 2984108, 128, 4),
    # generated on the device per row shard (plgraph.py); the smaller ones are the same family at 1/10 and 1/100 scale
    "pl-1b": (10_000_000, 5_000_000, 1_000_000_000, 64, 3),
    "pl-100m": (1_000_000, 500_000, 100_000_000, 64, 3),
    "pl-10m": (100_000, 50_000, 10_000_000, 64, 3),
}

# the reference dataset's feature cardinalities (saved_model_data/num_dict.pkl)
FEATURE_CARD = {"sex": 2, "age": 76, "month": 13, "day": 32, "dayofweek": 7}


def num_dict_for(n_user: int, n_item: int) -> dict:
    d = {"user": int(n_user), "item": int(n_item)}
    d.update(FEATURE_CARD)
    return d


def zipf_probs(n: int, alpha: float, rng: np.random.Generator) -> np.ndarray:
    p = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    p /= p.sum()
    return p[rng.permutation(n)]          # hubs scattered over the id range, like real id maps


def powerlaw_bipartite(n_user: int, n_item: int, n_edges: int, alpha: float = 0.8, seed: int = 0,
                       weighted: bool = False):
    """Returns (users int64[E], items int64[E], ratings float32[E]) with distinct pairs, sorted
    by (user, item).  ratings are 1.0 (implicit) unless ``weighted`` (U(0,3], Seoul shape)."""
    if n_edges > n_user * n_item:
        raise ValueError("more edges than user-item pairs")
    rng = np.random.default_rng(seed)
    pu, pi = zipf_probs(n_user, alpha, rng), zipf_probs(n_item, alpha, rng)
    cu, ci = np.cumsum(pu), np.cumsum(pi)
    keys = np.empty(0, dtype=np.int64)
    while keys.size < n_edges:
        need = n_edges - keys.size
        m = int(need * 1.3) + 1024
        u = np.minimum(np.searchsorted(cu, rng.random(m)), n_user - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(ci, rng.random(m)), n_item - 1).astype(np.int64)
        keys = np.unique(np.concatenate([keys, u * n_item + i]))
    if keys.size > n_edges:
        keys = np.sort(rng.choice(keys, size=n_edges, replace=False))
    users, items = keys // n_item, keys % n_item
    if weighted:
        ratings = (3.0 * (1.0 - rng.random(n_edges))).astype(np.float32)   # (0, 3]
    else:
        ratings = np.ones(n_edges, dtype=np.float32)
    return users, items, ratings


def random_batch(n_user: int, n_item: int, batch: int, seed: int = 1, year: int = 18) -> dict:
    """Uniform random training triples + feature ids in their num_dict ranges (int64 numpy)."""
    rng = np.random.default_rng(seed)
    return {
        "year": np.full(batch, year, dtype=np.int64),
        "u_id": rng.integers(0, n_user, batch, dtype=np.int64),
        "age": rng.integers(0, FEATURE_CARD["age"], batch, dtype=np.int64),
        "sex": rng.integers(0, FEATURE_CARD["sex"], batch, dtype=np.int64),
        "month": rng.integers(0, FEATURE_CARD["month"], batch, dtype=np.int64),
        "day": rng.integers(0, FEATURE_CARD["day"], batch, dtype=np.int64),
        "dow": rng.integers(0, FEATURE_CARD["dayofweek"], batch, dtype=np.int64),
        "pos_item": rng.integers(0, n_item, batch, dtype=np.int64),
        "neg_item": rng.integers(0, n_item, batch, dtype=np.int64),
    }
