"""Where the end-to-end step time goes (host enqueue vs device): python scripts/e2e_breakdown.py [c2-dense128|c2-natural]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from mobgt_b200 import collator, model as M, synth

workload = sys.argv[1] if len(sys.argv) > 1 else "c2-dense128"
dev = torch.device("cuda")
world = synth.make_world("c2", seed=1)
items = bench.make_workload(workload, world, 256, 0)
latlon = torch.from_numpy(world.latlon).to(dev)
torch.manual_seed(1)
model = M.Graphormer(dataset_name="toyotagraph", world=world, **bench.HP).to(dev).train()
params = list(model.parameters())
flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
off = 0
for p in params:
    p.grad = flat[off:off + p.numel()].view_as(p)
    off += p.numel()
opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=0.01, fused=True)


def collate():
    return collator.collate_packed(items, world, latlon, 512, 20, 1024, device=dev)


def step(b):
    flat.zero_()
    loss = model.training_step(b)
    loss.backward()
    opt.step()
    return loss


def wall(fn, n=10, sync=True):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        t1 = time.perf_counter()
        if sync:
            torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t0))
    a = np.array(ts[2:]) * 1e3
    return a[:, 0].mean(), a[:, 1].mean()


b = collate()
for _ in range(3):
    step(b)
c_enq, c_tot = wall(collate)
s_enq, s_tot = wall(lambda: step(b))
print(f"collate : host enqueue {c_enq:6.2f} ms, to device-idle {c_tot:6.2f} ms")
print(f"step    : host enqueue {s_enq:6.2f} ms, to device-idle {s_tot:6.2f} ms")
# pieces of the collate host time
import cProfile, pstats, io
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    collate()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue()[:3500])
