"""GPU parity: the CUDA path (through the C ABI, behind the drop-in NGCF / BPR modules) against
(1) the golden vectors produced by the reference itself and (2) the CPU oracle on seeded inputs.

Tolerance (BASELINE.json north_star): max|a-b| / max|b| <= 1e-4 on layer embeddings, loss and gradients;
identical top-k lists apart from score ties.
"""
import numpy as np
import pytest
import torch

import seoul_tourism_recommendation_ngcf_b200 as pkg
from oracle import ngcf_oracle as O
from seoul_tourism_recommendation_ngcf_b200 import laplacian, synth
from tests._golden import STEP_CASES, Golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda"


def _build(g, rng="device"):
    cfg = g.cfg
    nd = synth.num_dict_for(cfg["n_user"], cfg["n_item"])
    m = pkg.NGCF(cfg["emb"], cfg["layers"], cfg.get("node_p", 0.3), cfg.get("mess_p", [0.1] * len(cfg["layers"])),
                 cfg.get("emb_ratio", 1.0), g.lap_list(), nd, cfg.get("B", 512), torch.device(DEV), rng=rng)
    m.load_state_dict(g.params())
    return m.to(DEV)


def _call(m, b, node_flag, neg=True):
    d = {k: v.to(DEV) for k, v in b.items()}
    return m(year=d["year"], u_id=d["u_id"], age=d["age"], sex=d["sex"], month=d["month"], day=d["day"], dow=d["dow"],
             pos_item=d["pos_item"], neg_item=d["neg_item"] if neg else torch.empty(0), node_flag=node_flag)


def _check_step(m, g, u, p, n):
    cfg = g.cfg
    assert rel_err(u.detach().cpu().numpy(), g.out("u")) <= TOL
    assert rel_err(p.detach().cpu().numpy(), g.out("pos")) <= TOL
    assert rel_err(n.detach().cpu().numpy(), g.out("neg")) <= TOL
    assert rel_err(m.all_users_emb.cpu().numpy(), g.out("all_users_emb")) <= TOL
    assert rel_err(m.all_items_emb.cpu().numpy(), g.out("all_items_emb")) <= TOL
    assert rel_err(m.user_embedding.weight.detach().cpu().numpy(), g.out("user_after")) <= 1e-6
    loss = pkg.BPR(cfg["wd"], cfg["B_ctor"])(u, p, n)
    assert loss.dim() == 0
    assert abs(float(loss) - float(g.out("loss"))) <= TOL * abs(float(g.out("loss")))
    m.zero_grad()
    loss.backward()
    for k, gr in g.grads().items():
        got = dict(m.named_parameters())[k].grad
        assert got is not None, k
        assert rel_err(got.cpu().numpy(), gr) <= TOL, k
    for k in g.nograd_keys():
        assert dict(m.named_parameters())[k].grad is None, k       # feature tables: no gradient (NGCF.py:115)


@pytest.mark.parametrize("name", STEP_CASES)
def test_training_step_matches_reference_golden(name):
    """Forward + BPR + backward against the reference's own outputs; dropout decisions injected as masks
    reproduced from the reference's host RNG stream."""
    g = Golden(name)
    cfg = g.cfg
    m = _build(g)
    m.train(cfg["train_mode"])
    L = g.lap_list()[O.select_year(g.batch()["year"])]
    if cfg["rng_seed"] >= 0:
        torch.manual_seed(cfg["rng_seed"])
    keep, mult = O.reference_dropout_draws(L._nnz(), L.shape[0], cfg["layers"], cfg["node_p"], cfg["mess_p"],
                                           cfg["node_flag"], cfg["train_mode"])
    inj = {}
    if keep is not None:
        inj["edge_keep"] = [torch.from_numpy(k.astype(np.uint8)) for k in keep]
    if mult is not None:
        inj["mess_mult"] = mult
    m._inject = inj or None
    u, p, n = _call(m, g.batch(), cfg["node_flag"])
    _check_step(m, g, u, p, n)


def test_reference_rng_mode_reproduces_host_node_dropout():
    """rng='reference': the node mask comes from the same torch CPU calls as NGCF.py:94 — no injection."""
    g = Golden("node_dropout")
    m = _build(g, rng="reference")
    m.eval()
    torch.manual_seed(g.cfg["rng_seed"])
    u, p, n = _call(m, g.batch(), True)
    _check_step(m, g, u, p, n)


def test_no_negatives_and_eval_style_bpr():
    """experiment.py:82-100: neg_item=torch.empty(0) -> third output is an empty CPU tensor; BPR is then called
    with a single positive row broadcast against the batch."""
    g = Golden("no_negatives")
    m = _build(g)
    m.eval()
    with torch.no_grad():
        u, p, n = _call(m, g.batch(), False, neg=False)
    assert n.numel() == 0 and n.device.type == "cpu"
    assert rel_err(u.cpu().numpy(), g.out("u")) <= TOL and rel_err(p.cpu().numpy(), g.out("pos")) <= TOL
    neg = torch.cat((p[1:], p[:1]))
    got = pkg.BPR(0.025, 25)(u, p[:1], neg)
    want = O.bpr_loss(u.cpu(), p[:1].cpu(), neg.cpu(), 0.025, 25)
    assert abs(float(got) - float(want)) <= TOL * abs(float(want))


def test_demo_checkpoint_topk():
    """demo.py:220-235 with the real checkpoint: CPU index tensors for year / pos_item, full ranking of 100 items."""
    g = Golden("ckpt_demo")
    m = _build(g)
    m.eval()
    info = g.batch()
    with torch.no_grad():
        u, _, _ = m(year=torch.LongTensor([0]), u_id=info["u_id"].to(DEV), age=info["age"].to(DEV),
                    sex=info["sex"].to(DEV), month=info["month"].to(DEV), day=info["day"].to(DEV),
                    dow=info["dow"].to(DEV), pos_item=torch.LongTensor([0]), neg_item=torch.empty(0), node_flag=False)
    assert rel_err(u.cpu().numpy(), g.out("u")) <= TOL
    assert rel_err(m.all_items_emb.cpu().numpy(), g.out("all_items_emb")) <= TOL
    val, idx = pkg.score_topk(u, m.all_items_emb, 100)
    _assert_topk(val.cpu().numpy(), idx.cpu().numpy(), g.out("scores"), 100)
    val20, idx20 = pkg.score_topk(u, m.all_items_emb, 20)
    _assert_topk(val20.cpu().numpy(), idx20.cpu().numpy(), g.out("scores"), 20)
    assert np.array_equal(idx20.cpu().numpy(), idx.cpu().numpy()[:, :20])


def _assert_topk(val, idx, ref_scores, k, tie=2e-5):
    """Same list as sorting the reference scores, except where reference scores tie within `tie` (relative)."""
    scale = np.abs(ref_scores).max()
    order = np.argsort(-ref_scores, axis=1, kind="stable")[:, :k]
    for r in range(ref_scores.shape[0]):
        assert len(set(idx[r].tolist())) == k
        got_ref_scores = ref_scores[r, idx[r]]
        want = ref_scores[r, order[r]]
        assert np.all(np.abs(got_ref_scores - want) <= tie * scale), (r, idx[r], order[r])
        assert np.all(np.abs(val[r] - want) <= 1e-4 * scale)
        assert np.all(np.diff(val[r]) <= 0)


# ---- kernel-level checks against torch on seeded inputs ------------------------------------------------
def _random_coo(N, nnz, hubs, seed, dup=True):
    rng = np.random.default_rng(seed)
    row = rng.integers(0, N, nnz)
    col = rng.integers(0, N, nnz)
    for h, cnt in hubs:                       # hub rows (longer than the split threshold) and hub columns
        row = np.concatenate([row, np.full(cnt, h)]); col = np.concatenate([col, rng.integers(0, N, cnt)])
        col = np.concatenate([col, np.full(cnt, h)]); row = np.concatenate([row, rng.integers(0, N, cnt)])
    if dup:
        row = np.concatenate([row, row[:50]]); col = np.concatenate([col, col[:50]])
    perm = rng.permutation(row.size)          # unsorted, with duplicates: what coalesce() would have to fix
    row, col = row[perm], col[perm]
    val = rng.standard_normal(row.size).astype(np.float32)
    return torch.sparse_coo_tensor(torch.from_numpy(np.stack([row, col])), torch.from_numpy(val), (N, N))


@pytest.mark.parametrize("d", [4, 20, 64, 65, 96, 128])
def test_spmm_forward_and_transpose_vs_torch(d):
    from seoul_tourism_recommendation_ngcf_b200.plan import LaplacianPlan, spmm
    N = 3000
    L = _random_coo(N, 40000, hubs=[(7, 5000), (1500, 700), (2999, 300)], seed=d)
    plan = LaplacianPlan(L, DEV)
    assert plan.fwd.n_hub >= 3 and not plan.symmetric
    X = torch.randn(N, d, generator=torch.Generator().manual_seed(1))
    Ld = L.to(torch.float64).coalesce()
    want = torch.sparse.mm(Ld, X.double())
    got = spmm(plan.fwd, None, X.to(DEV), d)
    assert rel_err(got.cpu().numpy(), want.numpy()) <= 1e-5
    want_t = torch.sparse.mm(Ld.t().coalesce(), X.double())
    add = torch.randn(N, d, generator=torch.Generator().manual_seed(2))
    got_t = spmm(plan.side(True, False), None, X.to(DEV), d, addend=add.to(DEV))
    assert rel_err(got_t.cpu().numpy(), (want_t + add.double()).numpy()) <= 1e-5
    # layout invariants: the permutation is a bijection onto the COO entries; ordinary rows first (row CSR with
    # the hub rows emptied), then the hub rows' entries in row order, cut into chunks of <= split entries
    side = plan.fwd
    perm = side.perm.cpu().numpy()
    assert np.array_equal(np.sort(perm), np.arange(plan.nnz))
    rp = side.rowptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == side.nnz_short and np.all(np.diff(rp) >= 0)
    coo_rows = L._indices()[0].numpy()
    rows_of = np.repeat(np.arange(N), np.diff(rp))
    assert np.array_equal(rows_of, coo_rows[perm[:side.nnz_short]])
    cp, crow = side.chunk_ptr.cpu().numpy(), side.chunk_row.cpu().numpy()
    assert cp[0] == 0 and cp[-1] == side.nnz_hub and np.all(np.diff(cp) > 0) and np.diff(cp).max() <= 128
    assert np.array_equal(np.repeat(crow, np.diff(cp)), coo_rows[perm[side.nnz_short:]])
    hub_of = side.hub_of_row.cpu().numpy()
    assert set(np.nonzero(hub_of >= 0)[0]) == set(crow) and np.all(np.diff(rp)[hub_of >= 0] == 0)
    for tiles, ptr in ((side.tiles, rp), (side.chunk_tiles, cp)):
        t = tiles.cpu().numpy()
        assert t[0, 0] == 0 and t[-1, 1] == ptr.size - 1 and np.array_equal(t[1:, 0], t[:-1, 1])
        assert np.array_equal(t[:, 2], ptr[t[:, 0]]) and np.array_equal(t[:, 3], ptr[t[:, 1]])


def test_symmetric_laplacian_shares_csr_and_empty_rows():
    from seoul_tourism_recommendation_ngcf_b200.plan import LaplacianPlan, spmm
    u, i, r = synth.powerlaw_bipartite(400, 300, 3000, seed=3)
    L = laplacian.laplacian_coo(u, i, r, 400, 320)            # 20 items without any edge -> empty rows
    ref = O.laplacian_from_R(__import__("scipy.sparse", fromlist=["x"]).csr_matrix((r, (u, i)), shape=(400, 320)), 400, 320)
    assert torch.equal(L._indices(), ref._indices()) and torch.equal(L._values(), ref._values())
    plan = LaplacianPlan(L, DEV)
    assert plan.symmetric and plan.side(True, False) is plan.fwd and plan.side(True, True) is plan.bwd
    X = torch.randn(720, 64)
    got = spmm(plan.fwd, None, X.to(DEV), 64)
    want = torch.mm(L, X)
    assert rel_err(got.cpu().numpy(), want.numpy()) <= 1e-5
    assert float(got[700:].abs().max()) == 0.0


def test_feature_mix_last_duplicate_wins():
    g = Golden("emb64_k3")
    m = _build(g)
    b = g.batch()
    assert b["u_id"][0] == b["u_id"][32]
    with torch.no_grad():
        _call(m, b, False)
    assert np.array_equal(m.user_embedding.weight.detach().cpu().numpy(), g.out("user_after"))


@pytest.mark.parametrize("U,I,D,k", [(3, 100, 257, 100), (37, 5000, 256, 20), (1, 129, 65, 128), (200, 300, 640, 7),
                                     (1000, 20011, 256, 20), (129, 1300, 96, 32), (300, 4000, 256, 33), (5, 3000, 64, 20)])
def test_score_topk_vs_torch(U, I, D, k):
    gen = torch.Generator().manual_seed(U * 7 + k)
    u = torch.randn(U, D, generator=gen)
    it = torch.randn(I, D, generator=gen)
    it[I // 2] = it[I // 3]                                      # an exact tie
    val, idx = pkg.score_topk(u.to(DEV), it.to(DEV), k)
    scores = (u.double() @ it.double().T).numpy()
    _assert_topk(val.cpu().numpy(), idx.cpu().numpy(), scores, k)


def test_bpr_gradients_vs_autograd():
    gen = torch.Generator().manual_seed(5)
    u, p, n = (torch.randn(300, 195, generator=gen) * 0.3 for _ in range(3))
    p[3] = 0                                                      # sign(0) = 0 branch of |x|
    ref_in = [t.clone().requires_grad_(True) for t in (u, p, n)]
    want = O.bpr_loss(*ref_in, 0.025, 1024)
    want.backward()
    got_in = [t.clone().to(DEV).requires_grad_(True) for t in (u, p, n)]
    got = pkg.BPR(0.025, 1024)(*got_in)
    (got * 3.0).backward()
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    for a, b in zip(got_in, ref_in):
        assert rel_err(a.grad.cpu().numpy(), 3.0 * b.grad.numpy()) <= 1e-5


def test_device_rng_dropout_statistics_and_consistency():
    """Device-RNG mode: keep fractions follow the reference's semantics (node: cumulative (1-p)^k, unscaled;
    message: Bernoulli(1-p) scaled by 1/(1-p)) and forward/backward see the same decisions."""
    from seoul_tourism_recommendation_ngcf_b200.plan import LaplacianPlan, spmm
    # (a) the effective masked matrix, read back through the SpMM itself with X = identity (N = 128 = max width)
    us, is_, rs = synth.powerlaw_bipartite(70, 58, 2500, seed=9, alpha=0.2)
    Ls = laplacian.laplacian_coo(us, is_, rs, 70, 58)
    ps = LaplacianPlan(Ls, DEV)
    eye = torch.eye(128, device=DEV)
    dense = Ls.to_dense().to(DEV)
    m0 = spmm(ps.fwd, None, eye, 128, drop_p=0.3, seed=1234, layer=0)
    m2 = spmm(ps.fwd, None, eye, 128, drop_p=0.3, seed=1234, layer=2)
    nz = dense != 0
    f0, f2 = float((m0 != 0)[nz].float().mean()), float((m2 != 0)[nz].float().mean())
    assert abs(f0 - 0.7) < 0.04 and abs(f2 - 0.343) < 0.04
    assert bool(((m2 != 0) <= (m0 != 0)).all())                  # cumulative: survivors of layer 2 survived layer 0
    assert torch.equal(m0[m0 != 0], dense[m0 != 0])              # unscaled
    assert not torch.equal(m0, m0.T)                             # (i,j) and (j,i) are dropped independently
    for side in (ps.bwd, ps.fwd):                                # L^T from its own CSR, and from L's (symmetric L)
        mt = spmm(side, None, eye, 128, drop_p=0.3, seed=1234, layer=0, transposed=True)
        assert torch.equal(mt, m0.T)
    # the per-step precomputed decision bytes (ngcf_node_dropout_bits) are the same decisions
    from seoul_tourism_recommendation_ngcf_b200.plan import node_dropout_bits
    bl, bt = node_dropout_bits(ps.fwd, 0.3, 1234, None, 3, as_L=True, as_Lt=True)
    assert torch.equal(spmm(ps.fwd, None, eye, 128, layer=0, keep_bits=bl), m0)
    assert torch.equal(spmm(ps.fwd, None, eye, 128, layer=2, keep_bits=bl), m2)
    assert torch.equal(spmm(ps.fwd, None, eye, 128, layer=0, transposed=True, keep_bits=bt), m0.T)
    _, bt2 = node_dropout_bits(ps.bwd, 0.3, 1234, None, 3, as_L=False, as_Lt=True)
    assert torch.equal(spmm(ps.bwd, None, eye, 128, layer=0, transposed=True, keep_bits=bt2), m0.T)
    # ... and so are the per-step compacted survivors (ngcf_node_dropout_compact): only they are staged and gathered
    from seoul_tourism_recommendation_ngcf_b200.plan import node_dropout_compact
    cl, ct = node_dropout_compact(ps.fwd, 0.3, 1234, None, 3, as_L=True, as_Lt=True)
    assert torch.equal(spmm(ps.fwd, None, eye, 128, compact=cl[0]), m0)
    assert torch.equal(spmm(ps.fwd, None, eye, 128, compact=cl[2]), m2)
    assert torch.equal(spmm(ps.fwd, None, eye, 128, transposed=True, compact=ct[0]), m0.T)
    _, ct2 = node_dropout_compact(ps.bwd, 0.3, 1234, None, 3, as_L=False, as_Lt=True)
    assert torch.equal(spmm(ps.bwd, None, eye, 128, transposed=True, compact=ct2[2]), m2.T)
    from seoul_tourism_recommendation_ngcf_b200 import _lib
    # static per-entry keys (the plan's default) and keys derived from the coordinates inside the pass: same survivors
    cl_nk, ct_nk = node_dropout_compact(ps.fwd, 0.3, 1234, None, 3, as_L=True, as_Lt=True, static_keys=False)
    for k in range(3):
        for x, y in ((cl[k], cl_nk[k]), (ct[k], ct_nk[k])):
            assert torch.equal(spmm(ps.fwd, None, eye, 128, compact=x), spmm(ps.fwd, None, eye, 128, compact=y))
    tiles = ps.fwd.tiles.cpu().numpy()
    for k, frac in ((0, 0.7), (2, 0.343)):                       # survivors per tile = what the bits say, in order
        cnt = cl[k][1].cpu().numpy()
        ent_c, ent_o = cl[k][0].cpu().numpy(), ps.fwd.ent.cpu().numpy()
        keep = ((bl.cpu().numpy() >> k) & 1).astype(bool)
        kept = 0
        for t, (r0, r1, e0, e1) in enumerate(tiles):
            n = cnt[t]
            assert n == keep[e0:e1].sum()
            assert (ent_c[e0:e0 + n] == ent_o[e0:e1][keep[e0:e1]]).all()
            kept += n
        assert abs(kept / max(1, tiles[-1][3]) - frac) < 0.05
    m0b = spmm(ps.fwd, None, eye, 128, drop_p=0.3, seed=99, layer=0)
    assert not torch.equal(m0b, m0)
    sd = torch.tensor([1234 - 99], dtype=torch.int64, device=DEV)   # device-side seed offset (graph replay path)
    assert torch.equal(spmm(ps.fwd, None, eye, 128, drop_p=0.3, seed=99, seed_dev=sd, layer=0), m0)

    # whole module in training mode with device RNG: read back the decisions the FORWARD drew (node masks via the
    # identity trick above with the step's own seed, message multipliers from the zeros of the stored layer outputs),
    # hand them to the CPU oracle as explicit masks, and require loss and every gradient to agree.  Fails if the
    # forward and the hand-written backward (L^T product, dense backward) disagree on a single mask bit.
    n_user, n_item, B = 700, 500, 256
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, 30000, seed=9)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    N = n_user + n_item
    nd = synth.num_dict_for(n_user, n_item)
    torch.manual_seed(0)
    m = pkg.NGCF(64, [64, 64], 0.3, [0.2, 0.2], 1.0, [L, L], nd, B, torch.device("cpu"))
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    m.train()
    b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=4).items()}
    torch.manual_seed(77)
    uu, pp, nn_ = _call(m, b, True)
    loss = pkg.BPR(0.025, B)(uu, pp, nn_)
    loss.backward()
    st = m._last
    assert st.drop_p == pytest.approx(0.3) and st.seed != 0
    plan = st.plan
    assert plan.fwd.n_hub > 0                                    # hub chunks take part
    eyeN = torch.eye(N, device=DEV)
    pad = torch.zeros(N, 128 * ((N + 127) // 128) - N, device=DEV)
    coo = L._indices().numpy()
    keep = []
    for k in range(2):
        cols = []
        for c0 in range(0, N, 128):                              # L with layer-k dropout, 128 columns at a time
            blk = torch.cat([eyeN, pad], 1)[:, c0:c0 + 128].contiguous()
            cols.append(spmm(plan.fwd, None, blk, 128, drop_p=0.3, seed=st.seed, layer=k))
        Lk = torch.cat(cols, 1)[:, :N].cpu().numpy()
        keep.append(Lk[coo[0], coo[1]] != 0)
    assert 0.6 < keep[0].mean() < 0.8 and np.all(keep[1] <= keep[0])
    mult = [(st.E[k + 1] != 0).float().cpu() / 0.8 for k in range(2)]
    assert abs(float((mult[0] == 0).float().mean()) - 0.2) < 0.02     # message dropout zeroes ~p of the entries
    loss_ref, grads_ref, mid = O.train_step(params, L, b, emb_ratio=1.0, weight_decay=0.025, batch_size_ctor=B,
                                            edge_keep=keep, mess_mult=mult)
    assert rel_err(uu.detach().cpu().numpy(), mid["u"].numpy()) <= TOL
    assert abs(float(loss) - float(loss_ref)) <= TOL * abs(float(loss_ref))
    for k, gr in grads_ref.items():
        if gr is not None:
            assert rel_err(dict(m.named_parameters())[k].grad.cpu().numpy(), gr.numpy()) <= TOL, k


@pytest.mark.parametrize("shape", ["seoul", "gowalla", "yelp2018", "amazon-book"])
def test_full_size_step_vs_oracle(shape):
    """BASELINE.json configs 1-2 at full size against the CPU oracle (the reference's torch.sparse path):
    Seoul shape = width 65 (scalar kernels, item rows of ~4000 entries -> hub split), Gowalla shape = width 64."""
    n_user, n_item, n_edges, emb, K = synth.SHAPES[shape]
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, n_edges, seed=0, weighted=(shape == "seoul"),
                                       alpha=0.3 if shape == "seoul" else 0.8)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    nd = synth.num_dict_for(n_user, n_item)
    torch.manual_seed(0)
    m = pkg.NGCF(emb, [emb] * K, 0.3, [0.1] * K, 1.0, [L, L], nd, 1024, torch.device("cpu"))
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(DEV).eval()
    b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, 1024, seed=1).items()}
    loss_ref, grads_ref, mid = O.train_step(params, L, b, emb_ratio=1.0, weight_decay=0.025, batch_size_ctor=1024)
    uu, pp, nn_ = _call(m, b, False)
    assert rel_err(uu.detach().cpu().numpy(), mid["u"].numpy()) <= TOL
    assert rel_err(pp.detach().cpu().numpy(), mid["pos"].numpy()) <= TOL
    assert rel_err(nn_.detach().cpu().numpy(), mid["neg"].numpy()) <= TOL
    all_E = mid["out"]["all_E"].detach().numpy()
    assert rel_err(m.all_users_emb.cpu().numpy(), all_E[:n_user]) <= TOL
    assert rel_err(m.all_items_emb.cpu().numpy(), all_E[n_user:]) <= TOL
    loss = pkg.BPR(0.025, 1024)(uu, pp, nn_)
    assert abs(float(loss) - float(loss_ref)) <= TOL * abs(float(loss_ref))
    loss.backward()
    # Gradients.  LeakyReLU'(M) jumps at M = 0: a pre-activation that is zero to rounding (|M| ~ 1e-8 in a row of
    # magnitude 1) may take the other branch here than in the fp32 reference, and one such element in a
    # large-gradient row moves a (heavily cancelling) bias gradient by more than 1e-4.  So: the branches must agree
    # with the reference's except on numerically-zero pre-activations, and every gradient must match the float64
    # restatement evaluated on the branches THIS forward took.
    act_pos, n_flip = [], 0
    for k in range(K):
        ours, ref = m._last.E[k + 1].cpu().numpy(), mid["out"]["E"][k + 1].detach().numpy()
        flip = (ours > 0) != (ref > 0)
        n_flip += int(flip.sum())
        if flip.any():
            scale = np.abs(ref).max(1, keepdims=True)
            assert (np.abs(ref)[flip] <= 1e-5 * np.broadcast_to(scale, ref.shape)[flip]).all()
        act_pos.append(ours > 0)
    assert n_flip <= 1e-5 * K * all_E.shape[0] * emb
    _, grads64, _ = O.train_step_f64(params, L, b, emb_ratio=1.0, weight_decay=0.025, batch_size_ctor=1024,
                                     act_pos=act_pos)
    for k, gr in grads_ref.items():
        got = dict(m.named_parameters())[k].grad
        if gr is None:
            assert got is None, k
        else:
            assert rel_err(got.cpu().numpy(), grads64[k]) <= TOL, k
            if n_flip == 0:
                assert rel_err(got.cpu().numpy(), gr.numpy()) <= TOL, k
    # top-20 over every item for 64 batch users (BASELINE north_star: identical top-20 apart from ties)
    val, idx = pkg.score_topk(uu[:64].detach(), m.all_items_emb, 20)
    scores = (mid["u"][:64].double() @ torch.from_numpy(all_E[n_user:]).double().T).numpy()
    _assert_topk(val.cpu().numpy(), idx.cpu().numpy(), scores, 20, tie=5e-5)


def _dense_fwd(S, E, W1, b1, W2, b2, mess_mult=None):
    """ngcf_pack_weights + ngcf_dense_fwd through the C ABI on device tensors."""
    from seoul_tourism_recommendation_ngcf_b200 import _lib
    lib = _lib.load()
    N, d_in = S.shape
    d_out = W1.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    wcat = torch.empty(2 * d_in * d_out, device=DEV)
    bias = torch.empty(d_out, device=DEV)
    _lib.check(lib.ngcf_pack_weights(W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), d_in, d_out,
                                     wcat.data_ptr(), bias.data_ptr(), st))
    out = torch.empty(N, d_out, device=DEV)
    _lib.check(lib.ngcf_dense_fwd(S.data_ptr(), E.data_ptr(), N, d_in, d_out, wcat.data_ptr(), bias.data_ptr(), 0.2,
                                  _lib.ptr(mess_mult), None, 0.0, 0, None, 0, 0, out.data_ptr(), st), "dense_fwd")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("N,d_in,d_out", [(1000, 64, 64), (128, 64, 64), (70839, 64, 64), (333, 32, 48), (777, 64, 16),
                                          (500, 65, 64), (300, 128, 128), (5000, 128, 64), (1300, 64, 128), (257, 20, 36)])
def test_dense_forward_vs_float64(N, d_in, d_out):
    """Per-layer epilogue (NGCF.py:131-142) against a float64 restatement.  Widths 32/64 take the tcgen05 3xTF32
    kernel, everything else the FFMA kernel; both must be at fp32 accuracy (1e-5 here, far inside the 1e-4 bar:
    a plain TF32 product would sit near 5e-4)."""
    g = torch.Generator().manual_seed(N + d_in)
    S, E = torch.randn(N, d_in, generator=g), torch.randn(N, d_in, generator=g)
    W1, W2 = torch.randn(d_out, d_in, generator=g) * 0.2, torch.randn(d_out, d_in, generator=g) * 0.2
    b1, b2 = torch.randn(d_out, generator=g), torch.randn(d_out, generator=g)
    mult = (torch.rand(N, d_out, generator=g) > 0.3).float() * 1.25
    for mm in (None, mult):
        got = _dense_fwd(S.to(DEV), E.to(DEV), W1.to(DEV), b1.to(DEV), W2.to(DEV), b2.to(DEV),
                         None if mm is None else mm.to(DEV))
        Sd, Ed = S.double(), E.double()
        M = (Sd + Ed) @ W1.double().T + (Sd * Ed) @ W2.double().T + 2 * b1.double() + b2.double()
        want = torch.where(M > 0, M, 0.2 * M)
        if mm is not None:
            want = want * mm.double()
        assert rel_err(got.cpu().numpy(), want.numpy()) <= 1e-5


@pytest.mark.parametrize("N,d_in,d_out", [(1000, 64, 64), (128, 64, 64), (70839, 64, 64), (4321, 64, 32), (500, 65, 64),
                                          (300, 128, 128), (5000, 128, 64), (1300, 64, 128), (257, 20, 36)])
def test_dense_backward_vs_float64(N, d_in, d_out):
    """Row-local backward of one layer against a float64 restatement; d_in = 64 takes the
    tcgen05 kernel (both GEMMs 3xTF32, weight gradients accumulated in TMEM), other widths the FFMA kernel."""
    from seoul_tourism_recommendation_ngcf_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(N + d_out)
    S, E = torch.randn(N, d_in, generator=g), torch.randn(N, d_in, generator=g)
    W1, W2 = torch.randn(d_out, d_in, generator=g) * 0.2, torch.randn(d_out, d_in, generator=g) * 0.2
    E_out = torch.randn(N, d_out, generator=g)
    gE_next = torch.randn(N, d_out, generator=g)
    mult = (torch.rand(N, d_out, generator=g) > 0.3).float() * 1.25
    n_slot, D, col_off = min(N // 3, 700), d_out + 24, 8
    rows = torch.randperm(N, generator=g)[:n_slot]
    slot = torch.full((N,), -1, dtype=torch.int32)
    slot[rows] = torch.arange(n_slot, dtype=torch.int32)
    gsum = torch.randn(n_slot, D, generator=g)
    dev = lambda t: t.to(DEV).contiguous()
    for use_next, use_mult, pre in ((True, True, 0), (False, False, 0), (True, False, 1)):
        d = dict(S=dev(S), E=dev(E), W1=dev(W1), W2=dev(W2), E_out=dev(E_out), gE_next=dev(gE_next), mult=dev(mult),
                 slot=dev(slot), gsum=dev(gsum))
        if pre:          # F.normalize's backward applied to gsum up front (ngcf_rowgrad_normalize), as the module does
            rows_d = dev(rows.to(torch.int64))
            blocks = [torch.zeros(N, col_off, device=DEV), d["E_out"]] + \
                     ([torch.zeros(N, D - col_off - d_out, device=DEV)] if D - col_off - d_out > 0 else [])
            before = d["gsum"].clone()
            _lib.check(lib.ngcf_rowgrad_normalize(_lib.ptr_array([rows_d]), _lib.i64_array([0]), _lib.i64_array([n_slot]),
                                                  1, _lib.ptr_array(blocks), _lib.int_array([b.shape[1] for b in blocks]),
                                                  len(blocks), d["slot"].data_ptr(), d["gsum"].data_ptr(), D,
                                                  torch.cuda.current_stream().cuda_stream), "rowgrad_normalize")
            assert torch.equal(d["gsum"][:, :col_off], before[:, :col_off])          # block 0 is not normalised
        gS, gEl = torch.empty(N, d_in, device=DEV), torch.empty(N, d_in, device=DEV)
        gW1, gW2 = torch.zeros(d_out, d_in, device=DEV), torch.zeros(d_out, d_in, device=DEV)
        gb1, gb2 = torch.zeros(d_out, device=DEV), torch.zeros(d_out, device=DEV)
        scratch = torch.empty(N, d_out, device=DEV)
        _lib.check(lib.ngcf_dense_bwd(d["gE_next"].data_ptr() if use_next else None, d["slot"].data_ptr(),
                                      d["gsum"].data_ptr(), D, col_off, d["E_out"].data_ptr(), d["S"].data_ptr(),
                                      d["E"].data_ptr(), N, d_in, d_out, d["W1"].data_ptr(), d["W2"].data_ptr(), 0.2,
                                      d["mult"].data_ptr() if use_mult else None, None, 0.0, 0, None, 0, 0, pre, gS.data_ptr(),
                                      gEl.data_ptr(), gW1.data_ptr(), gb1.data_ptr(), gW2.data_ptr(), gb2.data_ptr(),
                                      scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "dense_bwd")
        torch.cuda.synchronize()
        Eo, Sd, Ed = E_out.double(), S.double(), E.double()
        n = Eo.norm(dim=1, keepdim=True).clamp_min(1e-12)
        H = Eo / n
        gH = torch.zeros(N, d_out, dtype=torch.float64)
        gH[rows] = gsum[:, col_off:col_off + d_out].double()
        gEp = (gE_next.double() if use_next else 0) + (gH - H * (H * gH).sum(1, keepdim=True)) / n
        gM = gEp * (mult.double() if use_mult else 1.0) * torch.where(Eo > 0, 1.0, 0.2)
        T1, T2 = gM @ W1.double(), gM @ W2.double()
        want = dict(gS=T1 + T2 * Ed, gEl=T1 + T2 * Sd, gW1=gM.T @ (Sd + Ed), gW2=gM.T @ (Sd * Ed), gb1=2 * gM.sum(0),
                    gb2=gM.sum(0))
        got = dict(gS=gS, gEl=gEl, gW1=gW1, gW2=gW2, gb1=gb1, gb2=gb2)
        for k in want:
            assert rel_err(got[k].cpu().numpy(), want[k].numpy()) <= 2e-5, (k, use_next)


def test_row_sharded_path_single_rank_equals_unsharded():
    """The row-sharded code path (sharded.py; all-gathers, shard plans, global RNG keys) with a one-rank NCCL group
    must reproduce the unsharded step exactly, dropout included.  (tools/mgpu_check.py runs the same comparison
    under torchrun on 2+ GPUs.)"""
    import torch.distributed as dist
    own = not dist.is_initialized()
    if own:
        import os
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        n_user, n_item, B = 900, 700, 256
        u, i, r = synth.powerlaw_bipartite(n_user, n_item, 40000, seed=5)
        L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
        nd = synth.num_dict_for(n_user, n_item)
        b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=6).items()}
        res = []
        for sharded in (False, True):
            torch.manual_seed(0)
            m = pkg.NGCF(64, [64, 64, 32], 0.3, [0.1] * 3, 1.0, [L, L], nd, B, torch.device(DEV)).to(DEV)
            if sharded:
                m.shard()
            m.train()
            torch.manual_seed(11)
            uu, pp, nn_ = _call(m, b, True)
            loss = pkg.BPR(0.025, B)(uu, pp, nn_)
            loss.backward()
            res.append((uu.detach().clone(), float(loss), {k: p.grad.clone() for k, p in m.named_parameters()
                                                           if p.grad is not None}, m.all_items_emb.clone()))
        (u0, l0, g0, a0), (u1, l1, g1, a1) = res
        assert torch.equal(u0, u1) and abs(l0 - l1) <= 1e-6 * abs(l0) and torch.equal(a0, a1)   # loss: atomic sum order
        assert g0.keys() == g1.keys()
        for k in g0:
            assert rel_err(g1[k].cpu().numpy(), g0[k].cpu().numpy()) <= 1e-6, k      # weight grads: atomics order
    finally:
        if own:
            dist.destroy_process_group()


def test_graphed_step_matches_eager_and_redraws_dropout():
    """graph.GraphedStep replays model(...) + criterion + backward as one CUDA graph: without dropout every replay
    must equal the eager step on the same batch; with dropout every replay draws fresh masks."""
    n_user, n_item, B = 1200, 900, 256
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, 50000, seed=8)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    nd = synth.num_dict_for(n_user, n_item)
    batches = [{k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=20 + j).items()} for j in range(3)]
    torch.manual_seed(0)
    m = pkg.NGCF(64, [64, 64], 0.3, [0.1, 0.1], 1.0, [L, L], nd, B, torch.device(DEV)).to(DEV)
    crit = pkg.BPR(0.025, B)
    m.eval()
    step = pkg.GraphedStep(m, crit, B, node_flag=False)
    for b in batches + batches[:1]:
        loss_g = float(step(b))                                      # host batch -> pinned copy -> replay
        grads_g = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        allE_g = m.all_users_emb.clone()
        m.zero_grad(set_to_none=True)
        uu, pp, nn_ = _call(m, b, False)
        loss_e = crit(uu, pp, nn_)
        loss_e.backward()
        assert abs(loss_g - float(loss_e)) <= 1e-6 * abs(float(loss_e))
        assert torch.equal(allE_g, m.all_users_emb)
        for k, p in m.named_parameters():
            if p.grad is not None:
                assert rel_err(grads_g[k].cpu().numpy(), p.grad.cpu().numpy()) <= 1e-5, k
    assert step.launches_per_step and step.launches_per_step >= 10
    m.train()
    step_t = pkg.GraphedStep(m, crit, B, node_flag=True)
    losses = [float(step_t(batches[0])) for _ in range(4)]
    assert all(np.isfinite(losses)) and len(set(losses)) == 4         # same batch, fresh dropout masks each replay
    assert max(losses) - min(losses) < 0.2 * abs(losses[0])


def test_precomputed_message_dropout_bits_equal_in_kernel_decisions():
    """ngcf_mess_dropout_bits draws the same message-dropout decisions as the in-kernel Philox path: a training
    step with the bits precomputed must reproduce the default step bit for bit (forward) / to rounding (grads)."""
    n_user, n_item, B = 800, 600, 128
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, 30000, seed=12)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    nd = synth.num_dict_for(n_user, n_item)
    b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=13).items()}
    res = []
    for use_bits in (False, True):
        torch.manual_seed(0)
        m = pkg.NGCF(64, [64, 32, 65], 0.2, [0.3, 0.2, 0.1], 1.0, [L, L], nd, B, torch.device(DEV)).to(DEV)
        m._mess_bits = use_bits
        m.train()
        torch.manual_seed(5)
        uu, pp, nn_ = _call(m, b, True)
        pkg.BPR(0.025, B)(uu, pp, nn_).backward()
        res.append((uu.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    assert torch.equal(res[0][0], res[1][0])
    for k in res[0][1]:
        assert rel_err(res[1][1][k].cpu().numpy(), res[0][1][k].cpu().numpy()) <= 1e-6, k


def test_node_dropout_modes_agree():
    """Device-RNG node dropout evaluated in-kernel, from per-step decision bytes, or as per-step compacted survivor
    lists (the default): one set of decisions, identical sums (hub rows included), forward and backward."""
    n_user, n_item, B = 900, 600, 256
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, 40000, seed=3)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    nd = synth.num_dict_for(n_user, n_item)
    b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=4).items()}
    res = {}
    for mode in ("inkernel", "bits", "compact"):
        torch.manual_seed(0)
        m = pkg.NGCF(64, [64, 64, 64], 0.3, [0.1] * 3, 1.0, [L, L], nd, B, torch.device("cpu")).to(DEV)
        m._node_mode = mode
        m.train()
        torch.manual_seed(5)
        uu, pp, nn_ = _call(m, b, True)
        assert m._last.plan.fwd.n_hub > 0
        loss = pkg.BPR(0.025, B)(uu, pp, nn_)
        loss.backward()
        res[mode] = [uu.detach().clone(), pp.detach().clone(), loss.detach().clone()] + \
                    [p.grad.clone() for p in m.parameters() if p.grad is not None]
    for i, (a, b_) in enumerate(zip(res["bits"], res["inkernel"])):
        if i < 2:
            assert torch.equal(a, b_)                             # output rows: the very same sums
        else:                                                     # loss, W/b gradients: summed with float atomics
            assert rel_err(a.cpu().numpy().reshape(-1), b_.cpu().numpy().reshape(-1)) <= 2e-6
    # compaction shifts an entry's position inside its row, hence its lane group in the row sum: same terms,
    # different fp32 summation tree
    for a, b_ in zip(res["compact"], res["inkernel"]):
        assert rel_err(a.cpu().numpy().reshape(-1), b_.cpu().numpy().reshape(-1)) <= 2e-6


def test_adam_matches_torch_adam():
    """ngcf_adam_step against torch.optim.Adam on the CPU (the reference's optimizer, main.py:74): five steps over
    tensors of awkward sizes, one without gradient (the feature tables never get one, NGCF.py:115)."""
    gen = torch.Generator().manual_seed(3)
    shapes = [(70, 64), (129, 64), (64, 64), (64,), (5, 13), (1,), (4097, 3)]
    ref = [torch.nn.Parameter(torch.randn(*sh, generator=gen)) for sh in shapes]
    ours = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref]
    for wd in (0.0, 0.01):
        o_ref = torch.optim.Adam(ref, lr=3e-3, weight_decay=wd)
        o_our = pkg.Adam(ours, lr=3e-3, weight_decay=wd)
        for step in range(5):
            for i, (a, b) in enumerate(zip(ref, ours)):
                if i == 4:
                    a.grad = b.grad = None
                    continue
                g = torch.randn(a.shape, generator=gen) * (10.0 ** (step - 2))
                a.grad, b.grad = g.clone(), g.clone().to(DEV)
            o_ref.step()
            o_our.step(zero_grads=(step == 4))
            for i, (a, b) in enumerate(zip(ref, ours)):
                assert rel_err(b.detach().cpu().numpy().reshape(-1), a.detach().numpy().reshape(-1)) <= 2e-6, (wd, step, i)
        assert all(float(b.grad.abs().max()) == 0.0 for i, b in enumerate(ours) if i != 4)      # zero_grads
        sd = o_our.state_dict()
        assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 5.0
        assert 4 not in sd["state"] or not sd["state"][4]
        assert rel_err(sd["state"][1]["exp_avg_sq"].cpu().numpy(), o_ref.state_dict()["state"][1]["exp_avg_sq"].numpy()) <= 2e-6


def test_graphed_training_iteration_with_adam_tracks_torch_adam():
    """GraphedStep(optimizer=Adam): forward + BPR + backward + Adam as one replayed graph, against the same three
    steps taken eagerly with torch.optim.Adam (dropout off so both see the same gradients)."""
    n_user, n_item, B = 600, 400, 256
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, 20000, seed=2)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    nd = synth.num_dict_for(n_user, n_item)
    batches = [{k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=10 + j).items()} for j in range(3)]
    models = []
    for _ in range(2):
        torch.manual_seed(0)
        m = pkg.NGCF(64, [64, 64], 0.0, [0.0, 0.0], 0.5, [L, L], nd, B, torch.device(DEV)).to(DEV)
        m.train()
        models.append(m)
    crit = pkg.BPR(0.025, B)
    ref_opt = torch.optim.Adam(models[0].parameters(), lr=1e-2)
    for b in batches:
        ref_opt.zero_grad()
        loss = crit(*_call(models[0], b, False))
        loss.backward()
        ref_opt.step()
    gstep = pkg.GraphedStep(models[1], crit, B, node_flag=False, optimizer=pkg.Adam(models[1].parameters(), lr=1e-2))
    for b in batches:
        loss_g = gstep({k: (v if k == "year" else v.to(DEV)) for k, v in b.items()})
    assert abs(float(loss_g) - float(loss)) <= 1e-4 * abs(float(loss))
    for (k, a), (_, c) in zip(models[0].named_parameters(), models[1].named_parameters()):
        assert rel_err(c.detach().cpu().numpy(), a.detach().cpu().numpy()) <= 1e-4, k


def test_edge_cases_empty_graph_single_row_batch_and_repeated_users():
    """Degenerate inputs the reference handles without special code: a Laplacian without any entry (L·E = 0), a
    one-row batch, a batch that repeats one user (the last duplicate's feature mix wins, rows gathered twice)."""
    n_user, n_item = 37, 23
    N = n_user + n_item
    nd = synth.num_dict_for(n_user, n_item)
    empty = torch.sparse_coo_tensor(torch.zeros(2, 0, dtype=torch.int64), torch.zeros(0), (N, N))
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, 300, seed=1)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    for lap, B, rep in ((empty, 4, False), (L, 1, False), (L, 6, True)):
        torch.manual_seed(0)
        m = pkg.NGCF(64, [64, 64], 0.0, [0.0, 0.0], 1.0, [lap, lap], nd, B, torch.device("cpu"))
        params = {k: v.detach().clone() for k, v in m.state_dict().items()}
        m = m.to(DEV).eval()
        b = {k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=2).items()}
        if rep:
            b["u_id"][:] = b["u_id"][0]
        want_loss, want_grads, mid = O.train_step(params, lap, b, emb_ratio=1.0, weight_decay=0.025, batch_size_ctor=B)
        uu, pp, nn_ = _call(m, b, False)
        loss = pkg.BPR(0.025, B)(uu, pp, nn_)
        loss.backward()
        assert rel_err(uu.detach().cpu().numpy(), mid["u"].numpy()) <= TOL
        assert abs(float(loss) - float(want_loss)) <= TOL * abs(float(want_loss))
        for k, g in want_grads.items():
            if g is not None:
                assert rel_err(dict(m.named_parameters())[k].grad.cpu().numpy(), g.numpy()) <= TOL, (k, B, rep)
