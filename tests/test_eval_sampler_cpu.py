"""CPU side of SURVEY.md section 8(f) #3/#4: the oracle's restatements of TourDataset._negative_sampling
(utils.py:213-275) and Experiment.eval (experiment.py:66-119) against what the reference itself produced
(tests/golden/eval_sampler.npz), and the sampler's host index against the reference's row layout."""
import numpy as np
import pytest
import torch

from oracle import ngcf_oracle as O
from seoul_tourism_recommendation_ngcf_b200 import sampler
from tests._golden import Golden, eval_frame_cols, eval_test_batches


@pytest.fixture(scope="module")
def g():
    return Golden("eval_sampler")


def test_oracle_sampler_reproduces_reference_with_numpy_seed(g):
    test, total = eval_frame_cols(g, "test"), eval_frame_cols(g, "total")
    np.random.seed(g.cfg["seed_train"])
    users, items = O.negative_sampling(test, total["itemid"], train=True)
    assert np.array_equal(users.numpy(), g.raw["sampler/train_users"])
    assert np.array_equal(items.numpy(), g.raw["sampler/train_items"])
    np.random.seed(g.cfg["seed_test"])
    users, items = O.negative_sampling(test, total["itemid"], train=False)
    assert np.array_equal(users.numpy(), g.raw["sampler/test_users"])
    assert np.array_equal(items.numpy(), g.raw["sampler/test_items"])


def test_oracle_eval_epoch_matches_reference_experiment(g):
    cfg = g.cfg
    kw = dict(emb_ratio=cfg["emb_ratio"], weight_decay=cfg["wd"], test_batch=cfg["test_batch"], ks=cfg["ks"])
    torch.set_num_threads(1)
    m1, per, after = O.eval_epoch(g.params(), g.lap_list(), eval_test_batches(g), **kw)
    assert np.allclose(m1, g.out("metrics"), rtol=1e-6, atol=0)
    assert np.array_equal(after["user_embedding.weight"].numpy(), g.out("user_after"))
    m2, _, _ = O.eval_epoch(after, g.lap_list(), eval_test_batches(g), **kw)
    assert np.allclose(m2, g.out("metrics_pass2"), rtol=1e-6, atol=0)
    assert len(per) == len(g.raw["sampler/test_users"]) // cfg["test_batch"]


def test_host_index_gives_reference_row_order_and_positive_lists(g):
    """sampler.index_frame: users in first-seen order, each user's positive rows in frame order, and per-user
    ascending unique positive lists — checked against the rows the reference's TourDataset emitted."""
    test, total = eval_frame_cols(g, "test"), eval_frame_cols(g, "total")
    ix = sampler.index_frame(test, total["itemid"], "rating")
    rows = ix["rows"]
    ref_u, ref_i = g.raw["sampler/train_users"], g.raw["sampler/train_items"]
    assert rows.size == len(ref_u)
    got_u = np.stack([test[c][rows] for c in sampler.CONTEXT_COLS], axis=1)
    assert np.array_equal(got_u, ref_u) and np.array_equal(test["itemid"][rows], ref_i[:, 0])
    assert np.array_equal(np.repeat(np.column_stack([got_u, test["rating"][rows].astype(np.int64)]), 25, axis=0),
                          g.raw["sampler/test_users"])
    assert np.array_equal(test["itemid"][rows], g.raw["sampler/test_items"][::25])
    cand, ptr, idx = ix["candidates"], ix["pos_ptr"], ix["pos_idx"]
    assert np.array_equal(cand, np.unique(total["itemid"]))
    for r in (0, len(rows) // 2, len(rows) - 1):
        u = ix["row_user"][r]
        mine = cand[idx[ptr[u]:ptr[u + 1]]]
        uid = test["userid"][rows[r]]
        want = np.unique(test["itemid"][(test["userid"] == uid) & (test["rating"] > 0)])
        assert np.array_equal(mine, want)
        assert np.all(np.diff(idx[ptr[u]:ptr[u + 1]]) > 0)
    # every reference negative lies outside its user's positive list (what the device sampler must also satisfy)
    for r in range(len(rows)):
        u = ix["row_user"][r]
        assert ref_i[r, 1] not in set(cand[idx[ptr[u]:ptr[u + 1]]].tolist())


def test_product_eval_and_sampler_refuse_cpu(g):
    import seoul_tourism_recommendation_ngcf_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.eval_groups(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, dtype=torch.int64), torch.zeros(4),
                        group=2, ks=1, weight_decay=0.025, batch_size=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.sample_negatives(torch.zeros(2, dtype=torch.int32), torch.zeros(0, dtype=torch.int32),
                             torch.zeros(1, dtype=torch.int64), torch.arange(4), 1, 0)
    test, total = eval_frame_cols(g, "test"), eval_frame_cols(g, "total")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.TourDataset(test, total, True, "rating", device="cpu")


@pytest.mark.parametrize("seed", range(6))
def test_host_index_row_order_matches_oracle_on_random_frames(seed):
    """Random frames (repeated (user, item) rows, zero ratings, users without positives, item ids with gaps): the
    sampler's host index emits the same rows in the same order as the restated reference loop, and every negative the
    reference draws lies in the index's free-candidate set of that row's user."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    ids = np.sort(rng.choice(1000, size=int(rng.integers(3, 40)), replace=False))        # item ids with gaps
    cols = {"year": rng.integers(18, 20, n), "userid": rng.integers(0, 25, n), "itemid": rng.choice(ids[:-2], n),
            "age": rng.integers(0, 70, n), "sex": rng.integers(0, 2, n), "month": rng.integers(1, 13, n),
            "day": rng.integers(1, 29, n), "dayofweek": rng.integers(0, 7, n),
            "rating": (rng.integers(0, 4, n) * (rng.random(n) < 0.7)).astype(np.int64)}
    ix = sampler.index_frame(cols, ids, "rating")
    np.random.seed(seed)
    users, items = O.negative_sampling(cols, ids, train=True)
    rows = ix["rows"]
    assert len(users) == rows.size
    if rows.size == 0:
        return
    assert np.array_equal(np.stack([cols[c][rows] for c in sampler.CONTEXT_COLS], axis=1), users.numpy())
    assert np.array_equal(cols["itemid"][rows], items.numpy()[:, 0])
    cand, ptr, idx = ix["candidates"], ix["pos_ptr"], ix["pos_idx"]
    assert np.array_equal(cand, ids)
    for r in range(rows.size):
        u = ix["row_user"][r]
        pos = cand[idx[ptr[u]:ptr[u + 1]]]
        want = np.unique(cols["itemid"][(cols["userid"] == cols["userid"][rows[r]]) & (cols["rating"] > 0)])
        assert np.array_equal(pos, want)
        assert items.numpy()[r, 1] in np.setdiff1d(cand, pos)


def test_oracle_group_metrics_against_independent_float64():
    """eval_group_metrics (the torch restatement of experiment.py:92-116) against a from-scratch float64 evaluation,
    including a ground-truth id that also appears further down the batch (``pred_items.index(gt)`` takes the
    best-ranked occurrence)."""
    rng = np.random.default_rng(3)
    for trial in range(20):
        n, D, ks = 25, 33, 10
        u = rng.standard_normal((n, D)).astype(np.float32) * 0.4
        p = rng.standard_normal((n, D)).astype(np.float32) * 0.4
        ids = rng.permutation(500)[:n]
        if trial % 3 == 0:
            ids[5] = ids[0]
        rating = rng.integers(0, 6, n)
        bpr, hit, ndcg, rmse, sc = O.eval_group_metrics(torch.from_numpy(u), torch.from_numpy(p), torch.from_numpy(ids),
                                                        torch.from_numpy(rating), 0.025, 25, ks)
        U, P = u.astype(np.float64), p.astype(np.float64)
        s = P @ U[0]
        neg = np.concatenate([P[1:], P[1:2]])
        x = np.abs(U @ P[0]) - np.abs((U * neg).sum(1))
        want_bpr = (-(np.minimum(x, 0) - np.log1p(np.exp(-np.abs(x)))).sum()
                    + 0.025 * ((U ** 2).sum() + (P[0] ** 2).sum() + (neg ** 2).sum())) / 25
        order = np.argsort(-s, kind="stable")
        ranked = ids[order].tolist()
        pos = ranked.index(ids[0])
        assert abs(float(bpr) - want_bpr) <= 1e-5 * abs(want_bpr)
        assert hit == int(pos < 3) and abs(ndcg - (1 / np.log2(pos + 2) if pos < ks else 0.0)) <= 1e-12
        assert abs(float(rmse) - abs(s[0] - rating[0])) <= 1e-5 and np.allclose(sc.numpy(), s, atol=1e-5)
