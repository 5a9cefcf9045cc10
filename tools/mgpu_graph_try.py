"""Experiment: whole-step CUDA graph (GraphedStep) over the row-sharded path, NCCL collectives captured.
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/mgpu_graph_try.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import seoul_tourism_recommendation_ngcf_b200 as pkg
from seoul_tourism_recommendation_ngcf_b200 import laplacian, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
def log(*a): print(f"[rank {rank}] {time.time() % 1000:8.2f}", *a, flush=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_user, n_item, n_edges, emb, K = synth.SHAPES["gowalla"]
B = 1024
u, i, r = synth.powerlaw_bipartite(n_user, n_item, n_edges, seed=0)
L = laplacian.laplacian_coo(u, i, r, n_user, n_item)
nd = synth.num_dict_for(n_user, n_item)
torch.manual_seed(0)
m = pkg.NGCF(emb, [emb] * K, 0.3, [0.1] * K, 1.0, [L, L], nd, B, dev).to(dev)
m.shard()
m.train()
crit = pkg.BPR(0.025, B)
batches = [{k: torch.from_numpy(v) for k, v in synth.random_batch(n_user, n_item, B, seed=1 + j).items()} for j in range(4)]
dbatches = [{k: (v if k == "year" else v.to(dev)) for k, v in b.items()} for b in batches]
step = pkg.GraphedStep(m, crit, B, node_flag=True)
log("capturing")
loss = step(dbatches[0])
torch.cuda.synchronize()
log("captured + first replay, loss", float(loss))
for j in range(3):
    l = float(step(dbatches[j % 4]))
log("replays ok, loss", l)
dist.barrier(); torch.cuda.synchronize()
t0 = time.time()
n = 50
for j in range(n):
    step(dbatches[j % 4])
torch.cuda.synchronize()
log(f"graphed sharded step: {(time.time() - t0) / n * 1e3:.3f} ms/step (wall, {world} GPUs)")
dist.barrier()
torch.cuda.synchronize()
log("done")
sys.stdout.flush()
os._exit(0)      # destroy_process_group() blocks while captured graphs hold NCCL kernels
