"""Time of one row-block exchange (ngcf_push_rows over symmetric memory) against ncclAllGather, under torchrun:
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/push_bench.py [rows] [d]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from seoul_tourism_recommendation_ngcf_b200.sharded import PeerExchange, RowShards, all_gather_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 70839
d = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sh = RowShards(N, world, rank)
x = PeerExchange(None, dev)
M = x.matrix("m", sh.N_pad, d)
M[sh.r0:sh.r0 + sh.rows] = rank + 1
torch.cuda.synchronize(); dist.barrier()
x.push("m", sh.r0, sh.rows)
torch.cuda.synchronize(); dist.barrier()
want = torch.arange(1, world + 1, device=dev, dtype=torch.float32).repeat_interleave(sh.rows)
assert torch.equal(M[:, 0], want) and torch.equal(M[:, d - 1], want), "exchange delivered wrong rows"
full, mine = torch.empty(sh.N_pad, d, device=dev), torch.ones(sh.rows, d, device=dev)

def timed(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    ev = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); ev.append((e0, e1))
    torch.cuda.synchronize()
    t = torch.tensor([statistics.median(a.elapsed_time(b) for a, b in ev)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3

t_push = timed(lambda: x.push("m", sh.r0, sh.rows))
t_nccl = timed(lambda: all_gather_rows(full, mine, None))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10): x.push("m", sh.r0, sh.rows)
t_graph = timed(g.replay, reps=10) / 10
if rank == 0:
    mb = sh.N_pad * d * 4 / 1e6
    print(f"world {world}: matrix {mb:.1f} MB ({sh.rows} rows x {d} per rank): push {t_push:.1f} us (in a graph, back to back: "
          f"{t_graph:.1f} us), ncclAllGather {t_nccl:.1f} us; bytes received per rank {(world - 1) / world * mb:.1f} MB -> "
          f"{(world - 1) / world * mb / t_graph * 1e3 / 1e3:.0f} GB/s per direction in the graph", flush=True)
dist.barrier()
os._exit(0)
