"""Host-side cost of the eager drop-in step (the path main.py takes): cProfile over 200 steps.  python tools/host_profile.py"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import seoul_tourism_recommendation_ngcf_b200 as pkg

L, batches, info = bench.make_workload("gowalla")
dev = torch.device("cuda:0")
model = bench.make_model(info, L, dev).to(dev).train()
crit = pkg.BPR(bench.WEIGHT_DECAY, bench.BATCH)
db = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in batches]
for b in db:
    b["year"] = b["year"].cpu()

def step(b):
    model.zero_grad(set_to_none=True)
    u, p, n = model(year=b["year"], u_id=b["u_id"], age=b["age"], sex=b["sex"], month=b["month"], day=b["day"],
                    dow=b["dow"], pos_item=b["pos_item"], neg_item=b["neg_item"], node_flag=True)
    loss = crit(u, p, n)
    loss.backward()
    return loss

for j in range(10):
    step(db[j % len(db)])
torch.cuda.synchronize()
t0 = time.time()
for j in range(200):
    step(db[j % len(db)])
t_issue = time.time() - t0
torch.cuda.synchronize()
t_all = time.time() - t0
print(f"200 eager steps: host issue {t_issue / 200 * 1e3:.3f} ms/step, wall {t_all / 200 * 1e3:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for j in range(200):
    step(db[j % len(db)])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("cumulative")
st.print_stats(28)
