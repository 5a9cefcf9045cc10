#!/usr/bin/env python
"""Small driver for ncu: builds the bench workload and launches the propagation SpMM a few times.
  python tools/prof_spmm.py [shape] [reps] [drop_p]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from seoul_tourism_recommendation_ngcf_b200.plan import LaplacianPlan, spmm  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "gowalla"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
L, batches, info = bench.make_workload(shape)
dev = torch.device("cuda:0")
plan = LaplacianPlan(L, dev)
d = info["emb"]
X = torch.randn(plan.N, d, device=dev)
Y = torch.empty(plan.N, d, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
print("n_hub", plan.fwd.n_hub, "n_chunks", plan.fwd.n_chunks, "symmetric", plan.symmetric)
for r in range(reps):
    flush.zero_()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    spmm(plan.fwd, None, X, d, out=Y, drop_p=drop, seed=1, layer=0)
    e1.record()
    spmm(plan.fwd, None, X, d, out=Y, drop_p=drop, seed=1, layer=0)
    e2.record()
    torch.cuda.synchronize()
    print(f"rep {r}: cold {e0.elapsed_time(e1) * 1e3:.1f} us, warm {e1.elapsed_time(e2) * 1e3:.1f} us")
