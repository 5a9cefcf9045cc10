"""Multi-GPU parity check, run under torchrun on one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/mgpu_check.py
Every rank runs the same training step (dropout on, device RNG) unsharded on its own GPU and row-sharded over all
ranks; outputs, loss and every gradient must agree (same Philox decisions: keys are global coordinates)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import seoul_tourism_recommendation_ngcf_b200 as pkg
from seoul_tourism_recommendation_ngcf_b200 import laplacian, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
shape = sys.argv[1] if len(sys.argv) > 1 else "small"
if shape == "small":
    n_user, n_item, n_edges, emb, K, B = 3001, 2000, 120000, 64, 3, 512
else:
    n_user, n_item, n_edges, emb, K = synth.SHAPES[shape]
    B = 1024
u, i, r = synth.powerlaw_bipartite(n_user, n_item, n_edges, seed=0)
L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
nd = synth.num_dict_for(n_user, n_item)
b = {k: torch.from_numpy(v).to(dev) for k, v in synth.random_batch(n_user, n_item, B, seed=1).items()}
res = []
for sharded in (False, True):
    torch.manual_seed(0)
    m = pkg.NGCF(emb, [emb] * K, 0.3, [0.1] * K, 1.0, [L, L], nd, B, dev).to(dev)
    if sharded:
        m.shard()
    m.train()
    torch.manual_seed(7)
    uu, pp, nn_ = m(b["year"], b["u_id"], b["age"], b["sex"], b["month"], b["day"], b["dow"], b["pos_item"], b["neg_item"], True)
    loss = pkg.BPR(0.025, B)(uu, pp, nn_)
    loss.backward()
    torch.cuda.synchronize()
    res.append((uu.detach(), float(loss), {k: p.grad for k, p in m.named_parameters() if p.grad is not None},
                m.all_users_emb.clone()))
(u0, l0, g0, a0), (u1, l1, g1, a1) = res
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
worst = max(rel(g1[k], g0[k]) for k in g0)
ok = rel(u1, u0) <= 1e-6 and abs(l1 - l0) <= 1e-6 * abs(l0) and rel(a1, a0) <= 1e-6 and worst <= 1e-5
print(f"[rank {rank}/{world}] sharded vs unsharded: out {rel(u1, u0):.2e} loss {l0:.7f}/{l1:.7f} all_E {rel(a1, a0):.2e} "
      f"worst grad {worst:.2e} -> {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
