"""CPU oracle for the NGCF embedding-propagation hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, what the reference (haesungpyun/seoul_tourism_recommendation_NGCF)
computes on the path BASELINE.json's north_star names.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package (``seoul_tourism_recommendation_ngcf_b200``) never does and has no CPU fallback.

Where the arithmetic lives: the reference's arithmetic is torch's (un-pinned third-party dependency,
reference README.md:11 "torch >= 1.10.2"; this image: torch 2.11.0).  The fp32 functions below
therefore issue the *same torch CPU calls in the same order* as the reference lines they cite, so
on the same inputs they are bit-identical to the reference module executed on ``device='cpu'``
(``torch.mm(sparse_coo, dense)`` -> s_addmm_out_sparse_dense_cpu, ``F.linear``, ``F.leaky_relu``,
``F.normalize``, autograd for the backward).  The ``*_f64`` functions restate the same algebra in
explicit numpy float64 (forward and the hand-derived backward) and are the independent truth
the hand-written CUDA backward kernels are checked against.

Parity pin: the reference ships no tests or golden vectors for this path, so
the pin is "outputs of the reference itself run here": ``tests/golden/make_golden.py`` imports the
unmodified reference modules from /root/reference, runs them on seeded inputs and commits the
inputs+outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function in
this file against those vectors.

Reference citations are relative to /root/reference/model/.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

FEATURE_ORDER = ("age", "sex", "month", "day", "dow")  # concat order, NGCF.py:110


# --------------------------------------------------------------------------------------------
# Laplacian  (matrix.py:41-67, 79-83)
# --------------------------------------------------------------------------------------------
def laplacian_from_R(R_dense_or_sparse, n_user: int, n_item: int) -> torch.Tensor:
    """Sparse restatement of one year of ``Matrix.create_matrix`` (matrix.py:48-67).

    A = [[0, R], [R^T, 0]] (matrix.py:48-53); degree = COUNT of non-zeros per row (matrix.py:55),
    d^-1/2 in float32 with inf -> 0 (matrix.py:56-57); L = D^-1/2 A D^-1/2 evaluated by the
    reference in float64 dense (matrix.py:58-62, d_mat_inv is a float64 zeros array) and cast to
    float32 at matrix.py:82.  Output: uncoalesced fp32/int64 ``torch.sparse_coo`` [N, N] with
    row-major sorted indices (scipy dok->coo of a dense-built matrix, matrix.py:63,80-83).
    Entries whose product underflows to exactly 0 vanish in the reference (dok_matrix(dense) keeps
    non-zeros only); the same filter is applied here.
    """
    import scipy.sparse as sp

    R = sp.csr_matrix(R_dense_or_sparse, dtype=np.float32)
    R.eliminate_zeros()
    N = n_user + n_item
    A = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    A = sp.csr_matrix(A, shape=(N, N))
    deg = np.diff(A.indptr).astype(np.int64)                      # count_nonzero, matrix.py:55
    with np.errstate(divide="ignore"):
        d_sqrt = np.power(deg, -0.5, dtype=np.float32)            # matrix.py:56: float32 pow of the int count
    d_sqrt[np.isinf(d_sqrt)] = 0.0                                # matrix.py:57
    coo = A.tocoo()
    order = np.lexsort((coo.col, coo.row))
    row, col, a = coo.row[order], coo.col[order], coo.data[order]
    # multi_dot([D, A, D]) in float64 (matrix.py:58-62).  numpy's three-matrix rule evaluates
    # D(AD) when both orders cost the same (all N x N), so the entry is d_i * (a_ij * d_j).
    v = d_sqrt[row].astype(np.float64) * (a.astype(np.float64) * d_sqrt[col].astype(np.float64))
    keep = v != 0.0
    row, col, v = row[keep], col[keep], v[keep]
    idx = torch.from_numpy(np.stack([row, col]).astype(np.int64))
    val = torch.from_numpy(v.astype(np.float32))                  # matrix.py:82
    return torch.sparse_coo_tensor(idx, val, (N, N), is_coalesced=False)


def build_lap_list(years, users, items, ratings, n_user: int, n_item: int):
    """Restates the year loop of ``Matrix.create_matrix`` (matrix.py:41-67).

    R is a member that is never reset, so each year's R is the previous R overwritten with that
    year's ratings (matrix.py:33,45); the slot is ``year % 18`` (matrix.py:66-67).
    """
    years = np.asarray(years)
    users = np.asarray(users)
    items = np.asarray(items)
    ratings = np.asarray(ratings, dtype=np.float32)
    uniq = list(dict.fromkeys(years.tolist()))                    # df['year'].unique(): first-seen order
    lap_list = [[] for _ in uniq]                                 # matrix.py:38-39
    R = np.zeros((n_user, n_item), dtype=np.float32)
    for y in uniq:
        m = years == y
        R[users[m], items[m]] = ratings[m]                        # matrix.py:45 (later rows win)
        lap_list[y % 18] = laplacian_from_R(R, n_user, n_item)
    return lap_list


# --------------------------------------------------------------------------------------------
# node dropout masks (NGCF.py:93-100, 124-126)
# --------------------------------------------------------------------------------------------
def reference_node_masks(nnz: int, p: float, n_layer: int):
    """Reproduces the reference's cumulative edge masks from the CURRENT torch CPU RNG state.

    Per layer the reference draws ``nn.Dropout(p)(torch.tensor(np.ones(nnz_current))).bool()``
    (NGCF.py:94; float64, CPU generator, a fresh Dropout is always in training mode) over the
    SURVIVING entries only, does not rescale the values, and keeps the dropped matrix for the next
    layer (NGCF.py:124-126).  Returns ``n_layer`` bool arrays over the ORIGINAL nnz ordering, with
    mask[k] subset of mask[k-1].
    """
    alive = np.arange(nnz)
    out = []
    for _ in range(n_layer):
        m = torch.nn.Dropout(p)(torch.tensor(np.ones(alive.size))).type(torch.bool).numpy()
        alive = alive[m]
        full = np.zeros(nnz, dtype=bool)
        full[alive] = True
        out.append(full)
    return out


def reference_dropout_draws(nnz: int, N: int, out_dims, node_p, mess_p, node_flag: bool, training: bool):
    """Reproduces BOTH dropout streams of one reference forward from the current torch CPU RNG
    state, in the order the reference consumes it: per layer the node mask first (NGCF.py:124-126,
    only if node_flag), then the message-dropout multipliers (NGCF.py:142, only in training mode;
    CPU nn.Dropout = bernoulli_(1-p) noise / (1-p), independent of the input values).
    Returns (edge_keep or None, mess_mult or None)."""
    alive = np.arange(nnz)
    keep, mult = [], []
    for k, dk in enumerate(out_dims):
        if node_flag:
            m = torch.nn.Dropout(node_p)(torch.tensor(np.ones(alive.size))).type(torch.bool).numpy()
            alive = alive[m]
            full = np.zeros(nnz, dtype=bool)
            full[alive] = True
            keep.append(full)
        if training:
            mult.append(torch.nn.Dropout(mess_p[k])(torch.ones(N, dk)))
    return (keep if node_flag else None), (mult if training else None)


# --------------------------------------------------------------------------------------------
# forward  (NGCF.py:102-156)
# --------------------------------------------------------------------------------------------
def feature_mix_(user_w: torch.Tensor, tables: dict, idx: dict, u_id: torch.Tensor, emb_ratio: float):
    """In-place batch-user overwrite, NGCF.py:103-115 (outside autograd).

    Duplicate user ids in one batch: torch's CPU index_put_ is last-write-wins when it runs on one
    thread and element-wise racy when it is parallel (measured here, torch 2.11, 8 threads: a row can
    end up a mixture of two samples' features).  The pinned semantics are the deterministic ones
    -- the LAST occurrence in the batch wins -- so the assignment runs single-threaded here and the
    golden vectors are generated with torch.set_num_threads(1).  In the reference's real data a user
    id determines its features (utils.py:70-74), so duplicates write identical rows anyway."""
    feats = torch.cat([tables[k][idx[k]] for k in FEATURE_ORDER], dim=1)       # NGCF.py:103-110
    nt = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        user_w[u_id] = user_w[u_id] * (1 - emb_ratio) + feats.detach().clone() * emb_ratio  # NGCF.py:114-115
    finally:
        torch.set_num_threads(nt)
    return feats


def select_year(year: torch.Tensor) -> int:
    """NGCF.py:117 — smallest year in the batch, modulo 18."""
    return int(year.unique()[0] % 18)


def propagate(L, E0, W1, b1, W2, b2, edge_keep=None, mess_mult=None, slope: float = 0.2):
    """K-layer embedding propagation, NGCF.py:120-147, same torch ops in the same order.

    L: sparse_coo [N,N]; E0: [N,d0]; W1[k]/W2[k]: [d_{k+1}, d_k]; b1[k]/b2[k]: [d_{k+1}].
    edge_keep: optional list of K bool arrays over L's nnz (cumulative, see reference_node_masks);
    mess_mult: optional list of K [N,d_{k+1}] multipliers standing in for nn.Dropout (NGCF.py:142).
    Returns dict(all_E=[N,D], S=[...], E=[E0, E'_1..E'_K] un-normalised, H=[H_1..H_K]).
    """
    E = E0
    all_E = [E]
    S_list, E_list, H_list = [], [E0], []
    idx, val = L._indices(), L._values()
    for k in range(len(W1)):
        if edge_keep is not None:
            m = torch.as_tensor(edge_keep[k])
            Lk = torch.sparse_coo_tensor(idx[:, m], val[m], L.shape, is_coalesced=False)  # NGCF.py:95-99
        else:
            Lk = L
        L_E = torch.mm(Lk, E)                                     # NGCF.py:130
        L_E_W1 = F.linear(L_E, W1[k], b1[k])                      # NGCF.py:131
        E_W1 = F.linear(E, W1[k], b1[k])                          # NGCF.py:133 (bias b1 counted twice)
        L_E_E = L_E * E                                           # NGCF.py:135
        L_E_E_W2 = F.linear(L_E_E, W2[k], b2[k])                  # NGCF.py:136
        M = L_E_W1 + E_W1 + L_E_E_W2                              # NGCF.py:138
        E = F.leaky_relu(M, negative_slope=slope)                 # NGCF.py:140
        if mess_mult is not None:
            E = E * mess_mult[k]                                  # NGCF.py:142
        H = F.normalize(E, p=2, dim=1)                            # NGCF.py:144
        all_E.append(H)                                           # NGCF.py:146
        S_list.append(L_E); E_list.append(E); H_list.append(H)
    return dict(all_E=torch.cat(all_E, dim=1), S=S_list, E=E_list, H=H_list)  # NGCF.py:147


def gather_outputs(all_E, n_user, u_id, pos_item, neg_item):
    """NGCF.py:148-156."""
    users, items = all_E[:n_user, :], all_E[n_user:, :]
    u = users[u_id, :]
    p = items[pos_item, :]
    n = torch.empty(0)
    if len(neg_item) > 0:
        n = items[neg_item, :]
    return u, p, n


def bpr_loss(u, pos, neg, weight_decay: float, batch_size: int):
    """bprloss.py:15-22."""
    x_upos = torch.mul(u, pos).sum(dim=1)
    x_uneg = torch.mul(u, neg).sum(dim=1)
    x_upn = torch.abs(x_upos) - torch.abs(x_uneg)
    log_prob = F.logsigmoid(x_upn).sum()
    reg = weight_decay * (torch.linalg.norm(u, dim=1).pow(2).sum()
                          + torch.linalg.norm(pos, dim=1).pow(2).sum()
                          + torch.linalg.norm(neg, dim=1).pow(2).sum())
    return (-log_prob + reg) / batch_size


def score_topk(u, items, k: int):
    """demo.py:234-235 / experiment.py:93,104,109: dense scores then torch.topk."""
    scores = torch.mm(u, items.T)
    return torch.topk(scores, k)


def train_step(params: dict, L, batch: dict, *, emb_ratio, weight_decay, batch_size_ctor,
               edge_keep=None, mess_mult=None):
    """One reference training step's math: forward + BPR + autograd backward
    (experiment.py:45-57).  ``params`` holds fp32 leaf tensors named as in the reference
    state_dict; user_embedding.weight is mutated in place like the reference does.
    Returns (loss, grads dict, intermediates dict)."""
    K = sum(1 for k in params if k.startswith("w1_list.") and k.endswith(".weight"))
    n_user = params["user_embedding.weight"].shape[0]
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    with torch.no_grad():
        tables = {"age": leaves["age_emb.weight"], "sex": leaves["sex_emb.weight"],
                  "month": leaves["month_emb.weight"], "day": leaves["day_emb.weight"],
                  "dow": leaves["dow_emb.weight"]}
        idx = {k: batch[k] for k in FEATURE_ORDER}
        feature_mix_(leaves["user_embedding.weight"], tables, idx, batch["u_id"], emb_ratio)
    E0 = torch.cat((leaves["user_embedding.weight"], leaves["item_embedding.weight"]), dim=0)  # NGCF.py:120
    W1 = [leaves[f"w1_list.{k}.weight"] for k in range(K)]
    b1 = [leaves[f"w1_list.{k}.bias"] for k in range(K)]
    W2 = [leaves[f"w2_list.{k}.weight"] for k in range(K)]
    b2 = [leaves[f"w2_list.{k}.bias"] for k in range(K)]
    out = propagate(L, E0, W1, b1, W2, b2, edge_keep=edge_keep, mess_mult=mess_mult)
    u, p, n = gather_outputs(out["all_E"], n_user, batch["u_id"], batch["pos_item"], batch["neg_item"])
    loss = bpr_loss(u, p, n, weight_decay, batch_size_ctor)
    loss.backward()
    grads = {k: (v.grad.detach() if v.grad is not None else None) for k, v in leaves.items()}
    new_user = leaves["user_embedding.weight"].detach()
    return loss.detach(), grads, dict(out=out, u=u.detach(), pos=p.detach(), neg=n.detach(), user_after=new_user)


# --------------------------------------------------------------------------------------------
# evaluation loop and metrics  (experiment.py:66-130)
# --------------------------------------------------------------------------------------------
def forward_eval(params: dict, lap_list, batch: dict, *, emb_ratio):
    """One eval-style forward (experiment.py:82-91): node_flag=False, eval mode, neg_item=torch.empty(0).
    ``params['user_embedding.weight']`` is mutated in place by the feature mix, as in the reference."""
    K = sum(1 for k in params if k.startswith("w1_list.") and k.endswith(".weight"))
    n_user = params["user_embedding.weight"].shape[0]
    with torch.no_grad():
        tables = {"age": params["age_emb.weight"], "sex": params["sex_emb.weight"], "month": params["month_emb.weight"],
                  "day": params["day_emb.weight"], "dow": params["dow_emb.weight"]}
        feature_mix_(params["user_embedding.weight"], tables, {k: batch[k] for k in FEATURE_ORDER}, batch["u_id"],
                     emb_ratio)
        L = lap_list[select_year(batch["year"])]
        E0 = torch.cat((params["user_embedding.weight"], params["item_embedding.weight"]), dim=0)
        out = propagate(L, E0, [params[f"w1_list.{k}.weight"] for k in range(K)],
                        [params[f"w1_list.{k}.bias"] for k in range(K)],
                        [params[f"w2_list.{k}.weight"] for k in range(K)],
                        [params[f"w2_list.{k}.bias"] for k in range(K)])
        u, p, _ = gather_outputs(out["all_E"], n_user, batch["u_id"], batch["pos_item"], torch.empty(0))
    return u, p


def eval_group_metrics(u, pos_emb, pos_item, rating, weight_decay: float, test_batch: int, ks: int):
    """The metric block of one test batch, experiment.py:92-116.  Returns (bpr, hit, ndcg, rmse, scores)."""
    gt = int(pos_item[0])                                                      # :92
    pred = torch.mm(u, pos_emb.T)                                              # :93
    neg = pos_emb[1:]                                                          # :96
    neg = torch.cat((neg, neg[:1]))                                            # :97
    bpr = bpr_loss(u, pos_emb[:1], neg, weight_decay, test_batch)              # :98-100
    _, rank = torch.topk(pred[0], 3)                                           # :104
    rec = torch.take(pos_item, rank).tolist()                                  # :105
    hit = 1 if gt in rec else 0                                                # :106, :127-130
    _, rank = torch.topk(pred[0], ks)                                          # :109
    rec = torch.take(pos_item, rank).tolist()                                  # :110
    ndcg = float(np.reciprocal(np.log2(rec.index(gt) + 2))) if gt in rec else 0.0   # :111, :120-126
    rmse = torch.sqrt(F.mse_loss(pred[0, 0], rating[0].to(pred.dtype)))        # :114-116, :133-141
    return bpr, hit, ndcg, rmse, pred[0]


def eval_epoch(params: dict, lap_list, test_batches, *, emb_ratio, weight_decay, test_batch, ks):
    """Experiment.eval, experiment.py:66-119: one full-graph forward PER test batch (each one overwrites its
    users' table rows), metrics accumulated the reference's way.  Returns (BPR, HR, NDCG, RMSE) and per-batch lists."""
    params = {k: v.detach().clone() for k, v in params.items()}
    BPR, RMSE, HR, NDCG, per = 0, 0, [], [], []
    for b in test_batches:
        u, p = forward_eval(params, lap_list, b, emb_ratio=emb_ratio)
        bpr, hit, ndcg, rmse, sc = eval_group_metrics(u, p, b["pos_item"], b["rating"], weight_decay, test_batch, ks)
        BPR += bpr; RMSE += rmse; HR.append(hit); NDCG.append(ndcg)
        per.append(dict(bpr=float(bpr), hit=hit, ndcg=ndcg, rmse=float(rmse), scores=sc.numpy().copy()))
    n = len(test_batches)
    return (float(BPR / n), float(np.mean(HR)), float(np.mean(NDCG)), float(RMSE / n)), per, params


# --------------------------------------------------------------------------------------------
# triple sampler  (utils.py:213-275)
# --------------------------------------------------------------------------------------------
def negative_sampling(cols: dict, total_items, train: bool, rng=np.random):
    """TourDataset._negative_sampling restated over plain arrays (``cols``: the frame's columns year, userid, age,
    sex, month, day, dayofweek, rating, itemid as 1-D arrays in frame order).  Same loop order and the same
    ``rng.choice(neg, ng_ratio, replace=False)`` call per positive row, so with the same numpy seed it returns what
    the reference returns.  -> (users int64 [R, 7|8], items int64 [R, 2] | [R*25])."""
    all_dest = np.unique(np.asarray(total_items))                              # utils.py:224 (set semantics)
    ng_ratio = 1 if train else 24                                              # utils.py:227-230
    users_list, items_list = [], []
    uid = np.asarray(cols["userid"])
    _, first = np.unique(uid, return_index=True)
    for user in uid[np.sort(first)]:                                           # df['userid'].unique(): first-seen order
        rows = np.flatnonzero(uid == user)
        rows = rows[np.asarray(cols["rating"])[rows] > 0]                      # utils.py:238
        neg_items = np.setxor1d(all_dest, np.asarray(cols["itemid"])[rows])    # utils.py:240
        for r in rows:
            ctx = [cols[k][r] for k in ("year", "userid", "age", "sex", "month", "day", "dayofweek")]
            negs = rng.choice(neg_items.copy(), ng_ratio, replace=False)       # utils.py:258
            if train:
                items_list.append([cols["itemid"][r]] + negs.tolist())         # utils.py:253,261,268
                users_list.append(ctx)
            else:
                items_list.append(cols["itemid"][r])                           # utils.py:249-250
                users_list.append(ctx + [cols["rating"][r]])
                items_list += negs.tolist()                                    # utils.py:263-265
                users_list += [ctx + [cols["rating"][r]]] * ng_ratio
    return torch.LongTensor(np.asarray(users_list)), torch.LongTensor(np.asarray(items_list))


# --------------------------------------------------------------------------------------------
# explicit float64 restatement, forward and hand-derived backward
# --------------------------------------------------------------------------------------------
def _csr64(L):
    import scipy.sparse as sp
    idx = L._indices().numpy()
    val = L._values().numpy().astype(np.float64)
    return sp.csr_matrix((val, (idx[0], idx[1])), shape=tuple(L.shape))


def propagate_f64(L, E0, W1, b1, W2, b2, edge_keep=None, mess_mult=None, slope=0.2):
    """float64 numpy restatement of NGCF.py:120-147 using the merged form
    M = (S+E) W1^T + (S*E) W2^T + (2 b1 + b2)."""
    idx = L._indices().numpy()
    val = L._values().numpy().astype(np.float64)
    import scipy.sparse as sp
    E = np.asarray(E0, dtype=np.float64)
    Es, Ss, Ms, Hs, ns, Ls = [E], [], [], [], [], []
    for k in range(len(W1)):
        if edge_keep is not None:
            m = np.asarray(edge_keep[k])
            Lk = sp.csr_matrix((val[m], (idx[0][m], idx[1][m])), shape=tuple(L.shape))
        else:
            Lk = sp.csr_matrix((val, (idx[0], idx[1])), shape=tuple(L.shape))
        w1, w2 = np.asarray(W1[k], np.float64), np.asarray(W2[k], np.float64)
        S = Lk @ E
        M = (S + E) @ w1.T + (S * E) @ w2.T + (2.0 * np.asarray(b1[k], np.float64) + np.asarray(b2[k], np.float64))
        A = np.where(M > 0, M, slope * M)
        if mess_mult is not None:
            A = A * np.asarray(mess_mult[k], np.float64)
        n = np.maximum(np.sqrt((A * A).sum(1, keepdims=True)), 1e-12)
        Ss.append(S); Ms.append(M); Es.append(A); Hs.append(A / n); ns.append(n); Ls.append(Lk)
        E = A
    return dict(E=Es, S=Ss, M=Ms, H=Hs, n=ns, L=Ls, all_E=np.concatenate([Es[0]] + Hs, axis=1))


def bpr_f64(eu, ep, en, wd, batch_size):
    """float64 loss and row gradients, bprloss.py:15-22 differentiated by hand."""
    eu, ep, en = (np.asarray(x, np.float64) for x in (eu, ep, en))
    xp, xn = (eu * ep).sum(1), (eu * en).sum(1)
    x = np.abs(xp) - np.abs(xn)
    logsig = np.minimum(x, 0) - np.log1p(np.exp(-np.abs(x)))
    loss = (-logsig.sum() + wd * ((eu ** 2).sum() + (ep ** 2).sum() + (en ** 2).sum())) / batch_size
    c = -1.0 / (1.0 + np.exp(x))                                   # d(-logsig)/dx = -sigmoid(-x)
    sp_, sn_ = np.sign(xp)[:, None], np.sign(xn)[:, None]
    c = c[:, None]
    g_u = (c * sp_ * ep - c * sn_ * en + 2 * wd * eu) / batch_size
    g_p = (c * sp_ * eu + 2 * wd * ep) / batch_size
    g_n = (-c * sn_ * eu + 2 * wd * en) / batch_size
    return loss, g_u, g_p, g_n


def backward_f64(fw: dict, W1, W2, G, mess_mult=None, slope=0.2, act_pos=None):
    """Hand-derived backward of ``propagate_f64`` given G = dLoss/d all_E  [N, D_total]
    (SURVEY.md section 3.4).  Returns dict(gE0, gW1[k], gb1[k], gW2[k], gb2[k]).
    act_pos: optional list of K boolean [N, d_k] arrays used as the LeakyReLU branch (M > 0) instead of this
    forward's own signs: LeakyReLU'(M) jumps at M = 0, so two correct forwards that differ by rounding can
    disagree on the branch of a numerically-zero pre-activation; passing the other implementation's branches
    makes the comparison of the gradients well defined."""
    K = len(W1)
    dims = [fw["E"][0].shape[1]] + [fw["E"][k + 1].shape[1] for k in range(K)]
    offs = np.cumsum([0] + dims)
    gE_next = np.zeros_like(fw["E"][K])
    gW1, gb1, gW2, gb2 = [None] * K, [None] * K, [None] * K, [None] * K
    for k in range(K - 1, -1, -1):
        H, n, M, S, E = fw["H"][k], fw["n"][k], fw["M"][k], fw["S"][k], fw["E"][k]
        gH = G[:, offs[k + 1]:offs[k + 2]]
        gEp = gE_next + (gH - H * (H * gH).sum(1, keepdims=True)) / n
        gM = gEp * np.where(M > 0 if act_pos is None else np.asarray(act_pos[k]), 1.0, slope)
        if mess_mult is not None:
            gM = gM * np.asarray(mess_mult[k], np.float64)
        w1, w2 = np.asarray(W1[k], np.float64), np.asarray(W2[k], np.float64)
        gW1[k] = gM.T @ (S + E); gb1[k] = 2.0 * gM.sum(0)
        gW2[k] = gM.T @ (S * E); gb2[k] = gM.sum(0)
        T1, T2 = gM @ w1, gM @ w2
        gS = T1 + T2 * E
        gE_next = T1 + T2 * S + fw["L"][k].T @ gS
    gE0 = gE_next + G[:, :dims[0]]
    return dict(gE0=gE0, gW1=gW1, gb1=gb1, gW2=gW2, gb2=gb2)


def train_step_f64(params: dict, L, batch: dict, *, emb_ratio, weight_decay, batch_size_ctor, act_pos=None):
    """float64 restatement of one whole training step without dropout (feature mix in fp32 exactly as the reference
    does it, then propagate_f64 + bpr_f64 + backward_f64).  Returns (loss, grads dict keyed like the state_dict,
    forward dict).  See backward_f64 for act_pos."""
    K = sum(1 for k in params if k.startswith("w1_list.") and k.endswith(".weight"))
    n_user = params["user_embedding.weight"].shape[0]
    user = params["user_embedding.weight"].detach().clone()
    tables = {"age": params["age_emb.weight"], "sex": params["sex_emb.weight"], "month": params["month_emb.weight"],
              "day": params["day_emb.weight"], "dow": params["dow_emb.weight"]}
    feature_mix_(user, tables, {k: batch[k] for k in FEATURE_ORDER}, batch["u_id"], emb_ratio)
    E0 = torch.cat([user, params["item_embedding.weight"]]).numpy()
    W1 = [params[f"w1_list.{k}.weight"].numpy() for k in range(K)]
    b1 = [params[f"w1_list.{k}.bias"].numpy() for k in range(K)]
    W2 = [params[f"w2_list.{k}.weight"].numpy() for k in range(K)]
    b2 = [params[f"w2_list.{k}.bias"].numpy() for k in range(K)]
    fw = propagate_f64(L, E0, W1, b1, W2, b2)
    A = fw["all_E"]
    uid = batch["u_id"].numpy()
    pid, nid = batch["pos_item"].numpy() + n_user, batch["neg_item"].numpy() + n_user
    loss, gu, gp, gn = bpr_f64(A[uid], A[pid], A[nid], weight_decay, batch_size_ctor)
    G = np.zeros_like(A)
    np.add.at(G, uid, gu)
    np.add.at(G, pid, gp)
    np.add.at(G, nid, gn)
    bw = backward_f64(fw, W1, W2, G, act_pos=act_pos)
    grads = {"user_embedding.weight": bw["gE0"][:n_user], "item_embedding.weight": bw["gE0"][n_user:]}
    for k in range(K):
        grads[f"w1_list.{k}.weight"], grads[f"w1_list.{k}.bias"] = bw["gW1"][k], bw["gb1"][k]
        grads[f"w2_list.{k}.weight"], grads[f"w2_list.{k}.bias"] = bw["gW2"][k], bw["gb2"][k]
    return loss, grads, fw
