"""Per-CTA phase times of the row pass of one SpMM at a named shape (globaltimer stamps: start, tile staged, rows
gathered) and how many CTAs each SM had in flight.  python tools/spmm_timeline.py [shape] [layer|-1 = no dropout]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from seoul_tourism_recommendation_ngcf_b200 import _lib
from seoul_tourism_recommendation_ngcf_b200.plan import LaplacianPlan, node_dropout_compact, spmm

shape = sys.argv[1] if len(sys.argv) > 1 else "gowalla"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L, batches, info = bench.make_workload(shape)
dev = torch.device("cuda:0")
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
raw.ngcf_debug_spmm_timeline.argtypes = [C.c_void_p]
plan = LaplacianPlan(L, dev)
N, d = plan.N, info["emb"]
X = torch.randn(N, d, device=dev); Y = torch.empty(N, d, device=dev)
comp = None
if layer >= 0:
    cl, _ = node_dropout_compact(plan.fwd, 0.3, 1, None, 3, as_L=True, as_Lt=False)
    comp = cl[layer]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    spmm(plan.fwd, None, X, d, out=Y, compact=comp)
n_t, n_c = int(plan.fwd.tiles.shape[0]), int(plan.fwd.chunk_tiles.shape[0])
buf = torch.zeros((n_t + n_c) * 4, dtype=torch.int64, device=dev)
flush.zero_(); torch.cuda.synchronize()
raw.ngcf_debug_spmm_timeline(buf.data_ptr())
spmm(plan.fwd, None, X, d, out=Y, compact=comp)
torch.cuda.synchronize()
raw.ngcf_debug_spmm_timeline(None)
t_all = buf.cpu().numpy().reshape(-1, 4).astype(np.int64)
t0 = t_all[:, 0].min()
print(f"one launch: {n_c} chunk-tile CTAs + {n_t} row-tile CTAs, span {(t_all[:, 2].max() - t0) / 1e3:.1f} us")
for name, t in (("chunk tiles", t_all[:n_c]), ("row tiles", t_all[n_c:])):
    if not len(t):
        continue
    start, staged, done = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0
    print(f"{name}: staging mean {np.mean(staged - start):.0f} ns (p90 {np.percentile(staged - start, 90):.0f}), "
          f"stream+write mean {np.mean(done - staged):.0f} ns (p90 {np.percentile(done - staged, 90):.0f}), "
          f"life mean {np.mean(done - start):.0f} ns; first start {start.min() / 1e3:.1f} us, last done {done.max() / 1e3:.1f} us")
start, done = t_all[:, 0] - t0, t_all[:, 2] - t0
print(f"avg resident CTAs per SM over the span: {np.sum(done - start) / (done.max() * 148):.2f}")
for q in (0.25, 0.5, 0.75, 0.9, 1.0):
    print(f"  {int(q * 100):3d}% of CTAs started by {np.quantile(start, q) / 1e3:6.1f} us, done by {np.quantile(done, q) / 1e3:6.1f} us")

# ---- the compaction pass (row launch) ----
raw.ngcf_debug_compact_timeline.argtypes = [C.c_void_p]
buf2 = torch.zeros(n_t * 4, dtype=torch.int64, device=dev)
for _ in range(2):
    node_dropout_compact(plan.fwd, 0.3, 1, None, 3, as_L=True, as_Lt=True)
flush.zero_(); torch.cuda.synchronize()
raw.ngcf_debug_compact_timeline(buf2.data_ptr())
node_dropout_compact(plan.fwd, 0.3, 1, None, 3, as_L=True, as_Lt=True)
torch.cuda.synchronize()
raw.ngcf_debug_compact_timeline(None)
t = buf2.cpu().numpy().reshape(-1, 4).astype(np.int64)
t0 = t[:, 0].min()
start, loaded, done = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0
print(f"compaction, row launch: {n_t} CTAs, span {done.max() / 1e3:.1f} us; load+decide mean {np.mean(loaded - start):.0f} ns, "
      f"scan+write mean {np.mean(done - loaded):.0f} ns, life mean {np.mean(done - start):.0f} ns; "
      f"resident CTAs per SM {np.sum(done - start) / (done.max() * 148):.1f}")
