"""score_topk (fused U·I^T + per-row top-k, demo.py:234-235 / experiment.py:104,109) against torch.mm + torch.topk at
Gowalla scale.  python tools/topk_time.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import seoul_tourism_recommendation_ngcf_b200 as pkg

dev = torch.device("cuda:0")
n_items, D = 40981, 256
I = torch.randn(n_items, D, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def t(fn, reps=10):
    for _ in range(2): fn()
    ev = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); ev.append((e0, e1))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3

for n_users, k in ((30, 100), (30, 20), (1024, 20), (29858, 20)):
    U = torch.randn(n_users, D, device=dev)
    ours = t(lambda: pkg.score_topk(U, I, k))
    ref = t(lambda: torch.topk(U @ I.T, k))
    v0, i0 = pkg.score_topk(U, I, k)
    v1, i1 = torch.topk(U @ I.T, k)
    same = float((i0 == i1).float().mean())
    print(f"users {n_users:6d} x items {n_items} D {D} k {k:3d}: score_topk {ours:9.1f} us   torch mm+topk {ref:9.1f} us   "
          f"same indices {same:.4f}  (score matrix {n_users * n_items * 4 / 1e6:.0f} MB)", flush=True)
