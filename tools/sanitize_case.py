"""Small, hub-heavy run of every kernel on the hot path, meant to be executed under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
A Seoul-like graph (few items, so every item row is a hub split into chunks and completed through the
fence/counter protocol of spmm.cu) at width 64 (tcgen05 kernels) and 65 (FFMA / scalar-gather kernels), device-RNG
node + message dropout, forward + BPR + backward + Adam, top-k, eval metrics, sampler."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import seoul_tourism_recommendation_ngcf_b200 as pkg
from seoul_tourism_recommendation_ngcf_b200 import laplacian, synth

dev = torch.device("cuda:0")
for emb, n_user, n_item, n_edges in ((64, 1500, 40, 30000), (65, 700, 30, 9000)):
    K, B = 2, 128
    u, i, r = synth.powerlaw_bipartite(n_user, n_item, n_edges, seed=0, weighted=True)
    L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
    torch.manual_seed(0)
    m = pkg.NGCF(emb, [emb] * K, 0.3, [0.1] * K, 1.0, [L, L], synth.num_dict_for(n_user, n_item), B, dev).to(dev).train()
    opt = pkg.Adam(m.parameters(), lr=1e-3)
    crit = pkg.BPR(0.025, B)
    for step in range(2):
        b = {k: torch.from_numpy(v).to(dev) for k, v in synth.random_batch(n_user, n_item, B, seed=1 + step).items()}
        opt.zero_grad()
        uu, pp, nn_ = m(b["year"].cpu(), b["u_id"], b["age"], b["sex"], b["month"], b["day"], b["dow"], b["pos_item"],
                        b["neg_item"], True)
        loss = crit(uu, pp, nn_)
        loss.backward()
        opt.step()
    with torch.no_grad():                                 # the optimizer stepped the table: score from a fresh forward
        uu, pp, _ = m(b["year"].cpu(), b["u_id"], b["age"], b["sex"], b["month"], b["day"], b["dow"], b["pos_item"],
                      torch.empty(0), False)
    val, idx = pkg.score_topk(uu, m.all_items_emb, 20)
    torch.cuda.synchronize()
    print(f"emb {emb}: loss {float(loss):.5f} hubs {m._last.plan.fwd.n_hub} chunks {m._last.plan.fwd.n_chunks} "
          f"top1 {int(idx[0, 0])}", flush=True)
print("sanitize_case done")
