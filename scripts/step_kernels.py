"""Every kernel of one eager training step through trainer.Trainer (the product loop), device time per step, grouped.
python scripts/step_kernels.py [workload] [--steps=3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from mobgt_b200 import collator, model as M, synth, trainer

args = [a for a in sys.argv[1:] if not a.startswith("--")]
workload = args[0] if args else "c2-dense128"
steps = int([a.split("=")[1] for a in sys.argv if a.startswith("--steps=")][0]) if any(a.startswith("--steps=") for a in sys.argv) else 3
cfg, ds = bench.TRAIN_WORKLOADS[workload][:2]
world = bench.make_world_for(workload)
items = bench.make_workload(workload, world, 256, 0)
torch.manual_seed(1)
model = M.Graphormer(dataset_name=ds, world=world, **bench.HP).cuda().train()
tr = trainer.Trainer(model, "cuda", 1, cuda_graph=False)
b = collator.collate_packed(items, world, None, 512, 20, 1024)
for _ in range(3):
    tr.train_step(b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        tr.train_step(b)
    torch.cuda.synchronize()


def group(name):
    if "mobgt::" in name:
        k = name.split("mobgt::")[1].split("(")[0].split("<")[0]
        return "libmobgt", k
    if "nvjet" in name or "cutlass" in name or "gemm" in name.lower() or "cublas" in name.lower() or "gemv" in name.lower():
        return "library GEMM", name[:60]
    if "nccl" in name.lower():
        return "nccl", name[:60]
    if "Memcpy" in name or "Memset" in name:
        return "memcpy/memset", name[:60]
    return "torch elementwise / reduce / other", name[:400]


rows, aten = {}, {}
for e in prof.key_averages():
    t = getattr(e, "self_device_time_total", None)
    if t is None:
        t = e.self_cuda_time_total
    if t <= 0 or e.device_type.name != "CUDA":
        if e.key.startswith("aten::") and t > 0:
            aten[e.key] = (t / steps, e.count / steps)
        continue
    g, k = group(e.key)
    r = rows.setdefault((g, k), [0.0, 0])
    r[0] += t / steps
    r[1] += e.count / steps
tot = sum(v[0] for v in rows.values())
print(f"workload {workload}: {tot / 1e3:.3f} ms of kernel time per eager step, {sum(v[1] for v in rows.values()):.0f} launches")
by_group = {}
for (g, k), (t, c) in rows.items():
    by_group.setdefault(g, []).append((t, c, k))
for g, lst in sorted(by_group.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    gt = sum(x[0] for x in lst)
    print(f"\n== {g}: {gt / 1e3:.3f} ms ({100 * gt / tot:.1f} %), {sum(x[1] for x in lst):.0f} launches")
    for t, c, k in sorted(lst, reverse=True):
        print(f"  {t:9.1f} us  x{c:6.1f}  {t / max(c, 1e-9):8.1f} us each  {k}")

print("\n== aten ops by self device time (which torch calls launch the non-libmobgt kernels)")
for k, (t, c) in sorted(aten.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {t:9.1f} us  x{c:6.1f}  {k}")
