"""Host-side logic of the row-sharded multi-GPU path (sharded.py) on CPU: the partition itself, and — with two gloo
ranks — that per-layer all-gather + local shard products reproduce the unsharded L·X and L^T·gS."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seoul_tourism_recommendation_ngcf_b200 import laplacian, synth
from seoul_tourism_recommendation_ngcf_b200.sharded import RowShards, all_gather_rows, shard_coo


def test_row_shards_cover_all_rows():
    for N, world in ((10, 1), (10, 2), (10, 3), (70839, 8), (7, 8), (16, 4)):
        shards = [RowShards(N, world, r) for r in range(world)]
        assert all(s.rows == shards[0].rows and s.N_pad == world * s.rows >= N for s in shards)
        assert sum(s.valid for s in shards) == N
        covered = np.concatenate([np.arange(s.r0, s.r0 + s.valid) for s in shards])
        assert np.array_equal(covered, np.arange(N))
        assert all(s.bounds(r) == (shards[r].r0, shards[r].r0 + shards[r].valid) for s in shards for r in range(world))
    with pytest.raises(ValueError):
        RowShards(10, 2, 2)


def test_shard_coo_partitions_entries_and_is_symmetric_for_symmetric_L():
    u, i, r = synth.powerlaw_bipartite(50, 37, 600, seed=1)
    L = laplacian.laplacian_coo(u, i, r, 50, 37)
    world = 3
    seen_f, seen_b = [], []
    for rank in range(world):
        sh = RowShards(87, world, rank)
        (rf, cf, pf), (rb, cb, pb) = shard_coo(L, sh)
        assert rf.min() >= 0 and rf.max() < sh.rows and cf.max() < 87
        seen_f.append(pf); seen_b.append(pb)
        vals = L._values()
        A = torch.sparse_coo_tensor(torch.stack([rf, cf]), vals[pf], (sh.rows, sh.N_pad)).to_dense()
        B = torch.sparse_coo_tensor(torch.stack([rb, cb]), vals[pb], (sh.rows, sh.N_pad)).to_dense()
        assert torch.equal(A, B)                                   # L symmetric -> the L^T shard is the L shard
        assert torch.equal(A[:sh.valid, :87], L.to_dense()[sh.r0:sh.r0 + sh.valid])
    for seen in (seen_f, seen_b):
        assert np.array_equal(np.sort(torch.cat(seen).numpy()), np.arange(L._nnz()))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_user, n_item, d = 61, 42, 8
        N = n_user + n_item
        g = np.random.default_rng(0)
        # a NON-symmetric matrix exercises the separate L^T shard
        nnz = 900
        row, col = g.integers(0, N, nnz), g.integers(0, N, nnz)
        val = torch.from_numpy(g.standard_normal(nnz).astype(np.float32))
        L = torch.sparse_coo_tensor(torch.from_numpy(np.stack([row, col])), val, (N, N))
        sh = RowShards(N, world, rank)
        (rf, cf, pf), (rb, cb, pb) = shard_coo(L, sh)
        Lf = torch.sparse_coo_tensor(torch.stack([rf, cf]), val[pf], (sh.rows, sh.N_pad)).coalesce()
        Lb = torch.sparse_coo_tensor(torch.stack([rb, cb]), val[pb], (sh.rows, sh.N_pad)).coalesce()
        torch.manual_seed(1)                                       # same on every rank: replicated tables
        X = torch.randn(N, d)
        Xpad = torch.zeros(sh.N_pad, d)
        Xpad[:N] = X
        # forward layer: local rows of S = L X, then the all-gather every rank needs for the next layer
        S_loc = torch.sparse.mm(Lf, Xpad)
        S_all = all_gather_rows(torch.empty(sh.N_pad, d), S_loc)
        want = torch.sparse.mm(L.coalesce(), X)
        ok = torch.allclose(S_all[:N], want, atol=1e-5) and float(S_all[N:].abs().max() if sh.N_pad > N else 0) == 0
        # backward: gE rows = (L^T gS)[rows] from the all-gathered gS and the L^T shard
        gS_loc = torch.randn(sh.rows, d, generator=torch.Generator().manual_seed(10 + rank))
        gS_all = all_gather_rows(torch.empty(sh.N_pad, d), gS_loc)
        gE_loc = torch.sparse.mm(Lb, gS_all)
        gE_all = all_gather_rows(torch.empty(sh.N_pad, d), gE_loc)
        want_b = torch.sparse.mm(L.coalesce().t().coalesce(), gS_all[:N])
        ok = ok and torch.allclose(gE_all[:N], want_b, atol=1e-5)
        # W/b gradients: sum of per-rank partial sums over the row blocks
        part = (S_loc[:sh.valid] ** 2).sum(0)
        dist.all_reduce(part)
        ok = ok and torch.allclose(part, (want ** 2).sum(0), rtol=1e-5)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_products_match_unsharded_over_gloo(world):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) is True for r in range(world)), dict(ret)
