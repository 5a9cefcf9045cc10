"""GPU tests of the host-side contracts tightened in round 2: the pinned staging ring of GraphedStep, Adam's single
step counter with several parameter groups, the live-table freshness check (and snapshot=True), the device guard, and
the row-sharded step on 2 real ranks (NCCL, spawned here when the box has >= 2 GPUs).

Tolerances: 1e-6 relative where both sides run the same kernels in the same order (sharded vs unsharded outputs,
graph replays), 2e-6 against torch.optim.Adam, 1e-5 on gradients whose atomics may reorder."""
import os
import sys

import numpy as np
import pytest
import torch

import seoul_tourism_recommendation_ngcf_b200 as pkg
from seoul_tourism_recommendation_ngcf_b200 import laplacian, synth
from tests._golden import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _graph(n_user=900, n_item=700, edges=40000, seed=5):
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, edges, seed=seed)
    return laplacian.laplacian_coo(u, i, r, n_user, n_item), synth.num_dict_for(n_user, n_item)


def _call(m, b, node_flag):
    d = {k: v.to(m.user_embedding.weight.device) for k, v in b.items()}
    return m(year=d["year"], u_id=d["u_id"], age=d["age"], sex=d["sex"], month=d["month"], day=d["day"], dow=d["dow"],
             pos_item=d["pos_item"], neg_item=d["neg_item"], node_flag=node_flag)


def test_graphed_step_host_batches_survive_a_host_running_ahead():
    """with ONE pinned staging buffer a host that runs ahead of the GPU overwrote a batch before its
    asynchronous H2D copy had run.  40 distinct host batches issued back to back without any sync must give the same
    losses as the same batches issued with a sync after each."""
    n_user, n_item, B = 900, 700, 256
    L, nd = _graph(n_user, n_item)
    batches = [{k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=100 + j).items()}
               for j in range(40)]
    losses = []
    for sync in (True, False):
        torch.manual_seed(0)
        m = pkg.NGCF(64, [64, 64], 0.0, [0.0, 0.0], 0.0, [L, L], nd, B, torch.device(DEV)).to(DEV).eval()
        step = pkg.GraphedStep(m, pkg.BPR(0.025, B), B, node_flag=False)
        step(batches[0])
        torch.cuda.synchronize()
        out = torch.zeros(len(batches), device=DEV)
        # a long-running kernel ahead of the loop makes the host run far ahead of the device in the unsynced pass
        big = torch.randn(8192, 8192, device=DEV)
        for _ in range(4):
            big = big @ big * 1e-4
        for j, b in enumerate(batches):
            out[j] = step(b).detach()
            if sync:
                torch.cuda.synchronize()
        losses.append(out.cpu().numpy())
    assert len(set(np.round(losses[0], 6))) > 30                     # the batches really differ
    # (the loss is summed with float atomics: equal to rounding; a batch mix-up moves it in the third digit)
    assert np.allclose(losses[0], losses[1], rtol=2e-6, atol=0)
    assert np.abs(np.diff(losses[0])).min() > 1e-5


def test_adam_two_param_groups_match_torch_adam():
    """the shared step counter must advance once per step(), not once per parameter group."""
    gen = torch.Generator().manual_seed(5)
    shapes = [(70, 64), (64, 64), (64,), (33, 7)]
    ref = [torch.nn.Parameter(torch.randn(*sh, generator=gen)) for sh in shapes]
    ours = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref]

    def groups(ps):
        return [dict(params=ps[:1], lr=1e-2), dict(params=ps[1:3], lr=3e-3, weight_decay=0.01), dict(params=ps[3:], lr=1e-3)]
    o_ref, o_our = torch.optim.Adam(groups(ref)), pkg.Adam(groups(ours))
    for step in range(4):
        for a, b in zip(ref, ours):
            g = torch.randn(a.shape, generator=gen)
            a.grad, b.grad = g.clone(), g.clone().to(DEV)
        o_ref.step()
        o_our.step()
        for i, (a, b) in enumerate(zip(ref, ours)):
            assert rel_err(b.detach().cpu().numpy().reshape(-1), a.detach().numpy().reshape(-1)) <= 2e-6, (step, i)
    sd = o_our.state_dict()
    assert all(float(sd["state"][i]["step"]) == 4.0 for i in range(4))
    # a parameter whose first gradient arrives late has its own t in torch: refused, not silently mis-corrected
    late = torch.nn.Parameter(torch.zeros(5, device=DEV))
    o = pkg.Adam([ours[0], late], lr=1e-3)
    ours[0].grad = torch.ones_like(ours[0])
    o.step()
    late.grad = torch.ones_like(late)
    with pytest.raises(RuntimeError, match="one step count"):
        o.step()


def test_live_table_freshness_is_enforced_and_snapshot_lifts_it():
    """E_0 of a forward is the live parameter table (the reference's torch.cat copies it, NGCF.py:120): a second forward
    or an optimizer step before backward / all_users_emb must raise, and snapshot=True must give the reference's
    semantics — the first forward's gradients, unaffected by the second forward's feature mix."""
    n_user, n_item, B = 900, 700, 128
    L, nd = _graph(n_user, n_item)
    b1 = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=1).items()}
    b2 = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=2).items()}
    crit = pkg.BPR(0.025, B)

    def make(**kw):
        torch.manual_seed(0)
        return pkg.NGCF(64, [64, 64], 0.0, [0.0, 0.0], 0.5, [L, L], nd, B, torch.device(DEV), **kw).to(DEV).eval()
    m = make()
    loss = crit(*_call(m, b1, False))
    _call(m, b2, False)
    with pytest.raises(RuntimeError, match="snapshot=True"):
        loss.backward()
    m = make()
    opt = pkg.Adam(m.parameters(), lr=1e-3)
    crit(*_call(m, b1, False)).backward()
    m.all_users_emb                                                  # fresh: fine
    m._all_E = None
    opt.step()
    with pytest.raises(RuntimeError, match="optimizer step"):
        m.all_users_emb
    m = make()
    step = pkg.GraphedStep(m, crit, B, node_flag=False, optimizer=pkg.Adam(m.parameters(), lr=1e-3))
    step(b1)
    with pytest.raises(RuntimeError, match="also ran the optimizer"):
        m.all_users_emb
    # reference semantics with snapshot=True: backward of forward #1 after forward #2 == backward right after forward #1
    ma, mb = make(), make(snapshot=True)
    la = crit(*_call(ma, b1, False))
    la.backward()
    lb = crit(*_call(mb, b1, False))
    all_b = mb.all_users_emb.clone()
    with torch.no_grad():
        _call(mb, b2, False)
    mb._last = None                                                  # (the attributes now belong to forward #2)
    lb.backward()
    assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
    for (k, p), (_, q) in zip(ma.named_parameters(), mb.named_parameters()):
        if p.grad is not None:
            assert rel_err(q.grad.cpu().numpy(), p.grad.cpu().numpy()) <= 1e-5, k
    assert rel_err(all_b.cpu().numpy(), ma.all_users_emb.cpu().numpy()) <= 1e-6


def test_reference_rng_mode_keeps_the_host_stream_with_message_dropout():
    """in rng='reference' the device-RNG key of message dropout must not consume torch's CPU generator,
    whose stream belongs to the node masks (NGCF.py:94): the masks of step 2 equal the reference stream's."""
    n_user, n_item, B = 300, 200, 64
    L, nd = _graph(n_user, n_item, 6000)
    b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=1).items()}
    torch.manual_seed(0)
    m = pkg.NGCF(64, [64, 64], 0.3, [0.1, 0.1], 1.0, [L, L], nd, B, torch.device(DEV), rng="reference").to(DEV).train()
    nnz = int(L._nnz())
    torch.manual_seed(42)
    want = []
    for _ in range(2):                                               # two steps of the reference's host draws
        alive = nnz
        for _k in range(2):
            keep = torch.nn.Dropout(0.3)(torch.tensor(np.ones(alive))).type(torch.bool)
            alive = int(keep.sum())
        want.append(alive)
    torch.manual_seed(42)
    got = []
    for _ in range(2):
        masks = []
        orig = m._reference_node_masks

        def spy(n, dev, _o=orig):
            out = _o(n, dev)
            masks.extend(out)
            return out
        m._reference_node_masks = spy
        _call(m, b, True)
        m._reference_node_masks = orig
        got.append(int(masks[-1].sum()))
    assert got == want


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_module_on_a_non_current_device():
    """kernels must launch on the tensors' device, not on torch's current device."""
    n_user, n_item, B = 600, 400, 128
    L, nd = _graph(n_user, n_item, 20000)
    b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=1).items()}
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)
        torch.manual_seed(0)
        m = pkg.NGCF(64, [64, 64], 0.0, [0.0, 0.0], 1.0, [L, L], nd, B, torch.device(dev)).to(dev).eval()
        u, p, n = _call(m, b, False)
        loss = pkg.BPR(0.025, B)(u, p, n)
        loss.backward()
        val, idx = pkg.score_topk(u.detach(), m.all_items_emb, 10)
        torch.cuda.synchronize(dev)
        outs.append((u.detach().cpu(), float(loss), m.user_embedding.weight.grad.cpu(), idx.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and abs(outs[0][1] - outs[1][1]) <= 1e-6 * abs(outs[0][1])
    assert rel_err(outs[1][2].numpy(), outs[0][2].numpy()) <= 1e-5 and torch.equal(outs[0][3], outs[1][3])


def _rank_main(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from seoul_tourism_recommendation_ngcf_b200.sharded import parity_vs_unsharded
        n_user, n_item, B = 3001, 2000, 512                          # N not divisible by the world size: padded shards
        L, nd = _graph(n_user, n_item, 120000, seed=0)
        b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=1).items()}
        res = parity_vs_unsharded(64, [64, 64, 64], L, nd, b, B, torch.device("cuda", rank))
        ret[rank] = res
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (the driver's multi-GPU tier / gpurun --gpus 2)")
def test_row_sharded_step_on_two_ranks_equals_the_one_gpu_step():
    """the row-sharded step (dropout ON, device RNG keyed on global coordinates) on 2 NCCL ranks against
    the unsharded step each rank runs on its own GPU: outputs, loss, all_E and every gradient."""
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_rank_main, args=(world, 29571, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            res = ret[r]
            assert res["out"] <= 1e-6 and res["loss"] <= 1e-6 and res["all_E"] <= 1e-6 and res["worst_grad"] <= 1e-5, res


def test_device_built_power_law_laplacian_matches_the_host_builder():
    """SURVEY.md 8(f) #2 / BASELINE config 5: plgraph builds the CSR Laplacian of a synthetic power-law graph on the
    device.  Its structure and values must equal laplacian.laplacian_coo (the bit-exact sparse restatement of the
    reference's Matrix.create_matrix, matrix.py:41-83) fed with the same edges; the generator must be symmetric,
    heavy-tailed and nearly duplicate-free; device-built SpMM tiles must equal the host's greedy tiles; and a plan built
    from the CSR must multiply like torch's sparse mm."""
    from seoul_tourism_recommendation_ngcf_b200 import plgraph
    from seoul_tourism_recommendation_ngcf_b200.plan import LaplacianPlan, device_tiles, greedy_tiles, spmm
    from seoul_tourism_recommendation_ngcf_b200 import _lib
    n_user, n_item, n_edges = 3000, 1700, 90000
    N = n_user + n_item
    csr = plgraph.powerlaw_laplacian(n_user, n_item, n_edges, torch.device(DEV), alpha=0.8, seed=3)
    rp, col, val = csr.rowptr.cpu().numpy().astype(np.int64), csr.colidx.cpu().numpy().astype(np.int64), csr.vals.cpu().numpy()
    row = np.repeat(np.arange(N), np.diff(rp))
    assert rp[-1] == col.size and 1.5 * n_edges < col.size <= 2 * n_edges            # both directions; repeated pairs collapse
    up = row < n_user
    assert (col[up] >= n_user).all() and (col[~up] < n_user).all()                    # bipartite
    a = set(zip(row[up].tolist(), col[up].tolist()))
    assert a == set(zip(col[~up].tolist(), row[~up].tolist()))                        # symmetric
    deg = np.diff(rp)
    assert deg.max() > 8 * deg.mean()                                                 # heavy tail
    L = laplacian.laplacian_coo(row[up], col[up] - n_user, np.ones(up.sum(), np.float32), n_user, n_item).coalesce()
    idx, v = L.indices().numpy(), L.values().numpy()
    assert np.array_equal(idx[0], row) and np.array_equal(idx[1], col)
    assert np.abs(val - v).max() <= 5e-7 * np.abs(v).max()      # numpy's float32 pow is not correctly rounded
    # tiles
    lib = _lib.load()
    tr, te = lib.ngcf_spmm_tile_rows(), lib.ngcf_spmm_tile_entries()
    t_dev = device_tiles(csr.rowptr, tr, te).cpu().numpy()
    assert np.array_equal(t_dev, greedy_tiles(rp, tr, te))
    # a plan straight from the CSR
    plan = LaplacianPlan(csr, torch.device(DEV))
    X = torch.randn(N, 64, device=DEV)
    want = torch.sparse.mm(L.to(DEV), X)
    got = spmm(plan.fwd, None, X, 64)
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) <= 2e-6
    assert plan.fwd.n_hub > 0


def _rank_pl(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from seoul_tourism_recommendation_ngcf_b200 import plgraph
        from seoul_tourism_recommendation_ngcf_b200.sharded import RowShards, parity_vs_unsharded
        n_user, n_item, n_edges, B = 30001, 20000, 1500000, 512
        dev = torch.device("cuda", rank)
        sh = RowShards(n_user + n_item, world, rank)
        full = plgraph.powerlaw_laplacian(n_user, n_item, n_edges, dev, seed=1)
        part = plgraph.powerlaw_laplacian(n_user, n_item, n_edges, dev, seed=1, shard=sh)
        # the shard IS the row block of the unsharded CSR
        r0, r1 = sh.r0, min(sh.r0 + sh.rows, n_user + n_item)
        e0, e1 = int(full.rowptr[r0]), int(full.rowptr[r1])
        same = bool(torch.equal(part.colidx, full.colidx[e0:e1]) and torch.equal(part.vals, full.vals[e0:e1]) and
                    torch.equal(part.rowptr[:r1 - r0 + 1].long(), full.rowptr[r0:r1 + 1].long() - e0))
        b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=1).items()}
        res = parity_vs_unsharded(64, [64, 64, 64], full, synth.num_dict_for(n_user, n_item), b, B, dev, L_shard=part)
        res["shard_is_row_block"] = same
        ret[rank] = res
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_device_built_row_shards_train_like_the_unsharded_graph():
    """BASELINE config 5 at test scale: every rank generates ONLY its own rows on the device; the shard must be the row
    block of the unsharded CSR bit for bit, and the sharded training step must equal the 1-GPU step."""
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_rank_pl, args=(world, 29573, ret), nprocs=world, join=True)
        for r in range(world):
            res = ret[r]
            assert res["shard_is_row_block"], res
            assert res["out"] <= 1e-6 and res["loss"] <= 1e-6 and res["all_E"] <= 1e-6 and res["worst_grad"] <= 1e-5, res


def test_matrix_csr_builder_matches_the_reference_lap_list_and_trains_the_same():
    """``Matrix(builder="csr")`` builds every year's Laplacian on the device straight into CSR (no
    COO, no per-edge Python loop).  Against the reference's own ``Matrix.create_matrix`` output (golden fixture): same
    structure, values within 5e-7; and a training step on the CSR ``lap_list`` equals the step on the reference-format
    ``lap_list`` bit for bit (same plan)."""
    import pandas as pd
    from seoul_tourism_recommendation_ngcf_b200.matrix import Matrix
    from tests._golden import Golden
    g = Golden("seoul_small")
    f = g.group("frame")
    df = pd.DataFrame({k: f[k] for k in ("year", "userid", "itemid", "visitor")})
    nd = {"user": g.cfg["n_user"], "item": g.cfg["n_item"]}
    m = Matrix(df, ["year", "userid", "itemid", "visitor"], "visitor", nd, ".", False, torch.device(DEV), builder="csr")
    laps = m.create_matrix()
    ref = g.lap_list()
    assert len(laps) == len(ref)
    for Lc, Lr in zip(laps, ref):
        Lr = Lr.coalesce()
        idx, v = Lr.indices().numpy(), Lr.values().numpy()
        rp = Lc.rowptr.cpu().numpy().astype(np.int64)
        row = np.repeat(np.arange(Lc.n_rows), np.diff(rp))
        assert np.array_equal(row, idx[0]) and np.array_equal(Lc.colidx.cpu().numpy(), idx[1])
        assert np.abs(Lc.vals.cpu().numpy() - v).max() <= 5e-7 * np.abs(v).max()
    # one step on either lap_list format
    cfg = g.cfg
    full_nd = synth.num_dict_for(cfg["n_user"], cfg["n_item"])
    outs = []
    for lap in (laps, [L.to(DEV) for L in ref]):
        torch.manual_seed(0)
        mod = pkg.NGCF(cfg["emb"], cfg["layers"], 0.0, [0.0] * len(cfg["layers"]), 1.0, lap, full_nd, cfg.get("B", 16),
                       torch.device(DEV))
        mod.load_state_dict(g.params())
        mod = mod.to(DEV).eval()
        u, p, n = _call(mod, g.batch(), False)
        loss = pkg.BPR(cfg["wd"], cfg["B_ctor"])(u, p, n)
        loss.backward()
        outs.append((u.detach().cpu().numpy(), float(loss), mod.item_embedding.weight.grad.cpu().numpy()))
    assert rel_err(outs[0][0], outs[1][0]) <= 1e-6 and abs(outs[0][1] - outs[1][1]) <= 1e-6 * abs(outs[1][1])
    assert rel_err(outs[0][2], outs[1][2]) <= 1e-5
    assert rel_err(outs[1][0], g.out("u")) <= 1e-4                                    # ... and both equal the reference
