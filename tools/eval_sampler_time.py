"""Times the two section-8(f) kernels on a B200: ngcf_sample_negatives at Gowalla shape and ngcf_eval_groups against
the reference's per-batch metric block (experiment.py:92-116) issued with torch ops on the same GPU; then the drop-in
Experiment.eval in both modes.  `python tools/eval_sampler_time.py` (needs a GPU)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seoul_tourism_recommendation_ngcf_b200 as pkg  # noqa: E402
from seoul_tourism_recommendation_ngcf_b200 import laplacian, sampler, synth  # noqa: E402

dev = torch.device("cuda")


def ev_time(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


n_user, n_item, n_edges, emb, K = synth.SHAPES["gowalla"]
u, i, r = synth.powerlaw_bipartite(n_user, n_item, n_edges, alpha=0.8, seed=0)
cols = {c: np.zeros(n_edges, dtype=np.int64) for c in sampler.CONTEXT_COLS}
cols.update(userid=u.astype(np.int64), itemid=i.astype(np.int64), rating=np.ones(n_edges))
t0 = time.time()
ix = sampler.index_frame(cols, np.arange(n_item), "rating")
print(f"host index of {n_edges} rows: {time.time() - t0:.2f} s")
d = {k: torch.from_numpy(v).to(dev) for k, v in ix.items() if k != "rows"}
for ng in (1, 24):
    ms = ev_time(lambda: pkg.sample_negatives(d["pos_ptr"], d["pos_idx"], d["row_user"], d["candidates"], ng, 7), 5)
    print(f"sample_negatives: {n_edges} rows x {ng} negatives: {ms:.3f} ms ({n_edges * ng / ms / 1e6:.2f} G draws/s)")

G, grp, D, ks = 4000, 25, 256, 10
uu = torch.randn(G * grp, D, device=dev) * 0.2
pp = torch.randn(G * grp, D, device=dev) * 0.2
ids = torch.randint(0, n_item, (G * grp,), device=dev)
rt = torch.randint(0, 5, (G * grp,), device=dev).float()
ms = ev_time(lambda: pkg.eval_groups(uu, pp, ids, rt, group=grp, ks=ks, weight_decay=0.025, batch_size=grp))
print(f"eval_groups: {G} groups x {grp} x {D}: {ms:.4f} ms ({2 * G * grp * D * 4 / ms / 1e6:.1f} GB/s)")
crit = pkg.BPR(0.025, grp)


def torch_block(n):                      # the reference's per-batch block with torch ops + its host syncs
    hr = nd = 0
    for g in range(n):
        s = slice(g * grp, (g + 1) * grp)
        ue, pe, it = uu[s], pp[s], ids[s]
        gt = it[0].item()
        pred = torch.mm(ue, pe.T)
        neg = torch.cat((pe[1:], pe[1:2]))
        crit(ue, pe[:1], neg)
        rec = torch.take(it, torch.topk(pred[0], 3)[1]).cpu().numpy().tolist()
        hr += gt in rec
        rec = torch.take(it, torch.topk(pred[0], ks)[1]).cpu().numpy().tolist()
        nd += gt in rec
        torch.sqrt(torch.nn.functional.mse_loss(pred[0, 0], rt[s][0]))
    torch.cuda.synchronize()


t0 = time.time(); torch_block(400); dt = (time.time() - t0) / 400
print(f"torch per-batch metric block: {dt * 1e3:.3f} ms per group -> {dt * G * 1e3:.1f} ms for {G} groups")

# Experiment.eval at Gowalla shape: 200 test batches of 25 rows
L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
torch.manual_seed(0)
m = pkg.NGCF(emb, [emb] * K, 0.3, [0.1] * K, 1.0, [L, L], synth.num_dict_for(n_user, n_item), 1024, dev).to(dev)
rng = np.random.default_rng(1)
nb = 200
b = synth.random_batch(n_user, n_item, nb, seed=3)
users = np.stack([b["year"], b["u_id"], b["age"], b["sex"], b["month"], b["day"], b["dow"], rng.integers(1, 5, nb)], 1)
users = torch.from_numpy(np.repeat(users, grp, axis=0))
items = torch.from_numpy(rng.integers(0, n_item, nb * grp))


class DS(torch.utils.data.Dataset):
    def __len__(self):
        return len(users)

    def __getitem__(self, k):
        return tuple(users[k]) + (items[k],)


loader = torch.utils.data.DataLoader(DS(), batch_size=grp, shuffle=False, drop_last=True)
for mode in ("reference", "batched"):
    ex = pkg.Experiment(m, None, None, pkg.BPR(0.025, grp), None, loader, 1, ks, dev, eval_mode=mode, verbose=False)
    ex.eval(); torch.cuda.synchronize()
    t0 = time.time(); out = ex.eval(); torch.cuda.synchronize()
    print(f"Experiment.eval[{mode}] {nb} test batches at Gowalla shape: {(time.time() - t0) * 1e3:.1f} ms  -> {out}")
