"""ncu driver for the kernels outside the training step: fused score + top-k (1 024 users x 40 981 items x 256, k = 20),
the evaluation-metric launch (4 000 test batches of 25 x 256) and the negative sampler (1.03 M rows).
  ncu --set full --clock-control none -k regex:"score_topk|eval_groups|eval_reduce|sample_negatives" ... python tools/prof_topk_eval.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seoul_tourism_recommendation_ngcf_b200 as pkg  # noqa: E402
from seoul_tourism_recommendation_ngcf_b200 import sampler, synth  # noqa: E402

dev = torch.device("cuda:0")
I = torch.randn(40981, 256, device=dev)
U = torch.randn(1024, 256, device=dev)
pkg.score_topk(U, I, 20)
G, grp, D = 4000, 25, 256
uu, pp = torch.randn(G * grp, D, device=dev) * 0.2, torch.randn(G * grp, D, device=dev) * 0.2
ids = torch.randint(0, 40981, (G * grp,), device=dev)
rt = torch.randint(0, 5, (G * grp,), device=dev).float()
pkg.eval_groups(uu, pp, ids, rt, group=grp, ks=10, weight_decay=0.025, batch_size=grp)
n_user, n_item, n_edges, _, _ = synth.SHAPES["gowalla"]
u, i, _ = synth.powerlaw_bipartite(n_user, n_item, n_edges, alpha=0.8, seed=0)
ix = sampler.index_frame({"userid": u.astype(np.int64), "itemid": i.astype(np.int64), "rating": np.ones(n_edges)},
                         np.arange(n_item), "rating")
d = {k: torch.from_numpy(v).to(dev) for k, v in ix.items()}
pkg.sample_negatives(d["pos_ptr"], d["pos_idx"], d["row_user"], d["candidates"], 1, 7)
pkg.sample_negatives(d["pos_ptr"], d["pos_idx"], d["row_user"], d["candidates"], 24, 7)
torch.cuda.synchronize()
print("done")
