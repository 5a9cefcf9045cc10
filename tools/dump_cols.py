"""Writes the column stream of the Gowalla-shaped bench graph's SpMM (execution order of the plan) as int32, for
tools/l2_gather_bench: python tools/dump_cols.py gpurun_out/cols.bin"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seoul_tourism_recommendation_ngcf_b200.plan import LaplacianPlan
L, _, info = bench.make_workload(sys.argv[2] if len(sys.argv) > 2 else "gowalla")
plan = LaplacianPlan(L, torch.device("cuda:0"))
cols = (plan.fwd.colidx & ((1 << 27) - 1)).to(torch.int32).cpu().numpy()
cols.tofile(sys.argv[1])
print("wrote", cols.size, "columns; max", cols.max())
