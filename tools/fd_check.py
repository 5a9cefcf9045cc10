"""Debug helper: finite-difference vs analytic directional derivative under device-RNG dropout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seoul_tourism_recommendation_ngcf_b200 as pkg
from seoul_tourism_recommendation_ngcf_b200 import laplacian, synth

DEV = "cuda"
u, i, r = synth.powerlaw_bipartite(2000, 1500, 60000, seed=9)
L = laplacian.laplacian_coo(u, i, r, 2000, 1500)
nd = synth.num_dict_for(2000, 1500)
b = {k: torch.from_numpy(v).to(DEV) for k, v in synth.random_batch(2000, 1500, 256, seed=4).items()}
for node_p, mess_p in ((0.0, 0.0), (0.3, 0.0), (0.0, 0.2), (0.3, 0.2)):
    torch.manual_seed(0)
    m = pkg.NGCF(64, [64, 64], node_p, [mess_p, mess_p], 1.0, [L, L], nd, 256, torch.device(DEV)).to(DEV)
    m.train()
    crit = pkg.BPR(0.025, 256)

    def loss_at(seed):
        torch.manual_seed(seed)
        return crit(*m(b["year"], b["u_id"], b["age"], b["sex"], b["month"], b["day"], b["dow"], b["pos_item"],
                       b["neg_item"], True))
    for pname in ("w1_list.1.weight", "w2_list.0.weight", "item_embedding.weight"):
        w = dict(m.named_parameters())[pname]
        loss = loss_at(77)
        m.zero_grad(); loss.backward()
        torch.manual_seed(5)
        gdir = torch.randn_like(w)
        analytic = float((w.grad * gdir).sum())
        out = []
        for eps in (1e-2, 3e-3, 1e-3):
            with torch.no_grad():
                w.add_(eps * gdir); lp = float(loss_at(77)); w.sub_(2 * eps * gdir); lm = float(loss_at(77)); w.add_(eps * gdir)
            out.append((lp - lm) / (2 * eps))
        print(f"node_p {node_p} mess_p {mess_p} {pname:24s} analytic {analytic:+.6f} fd {out}")
