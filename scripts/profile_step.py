"""Kernel-level breakdown of one training step (torch.profiler): top CUDA kernels by device time + host wall time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mobgt_b200 import collator, model as M, synth

workload = sys.argv[1] if len(sys.argv) > 1 else "c2-dense128"
world = synth.make_world("c2", seed=1)
items = bench.make_workload(workload, world, 256, 0)
torch.manual_seed(1)
model = M.Graphormer(dataset_name="toyotagraph", world=world, **bench.HP).cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=0.01, fused=True)
b = collator.collate_packed(items, world, None, 512, 20, 1024)


def step():
    opt.zero_grad(set_to_none=False)
    loss = model.training_step(b)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host launch time/step {1e3*(t1-t0)/5:.2f} ms ; wall/step {1e3*(t2-t0)/5:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
