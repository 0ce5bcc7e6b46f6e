"""One c5-shard call of the fused head (for ncu): python scripts/k5_one.py [M] [V] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mobgt_b200 import ops
M, V, k = (int(a) for a in (sys.argv[1:4] + ["4096", "125000", "10"][len(sys.argv) - 1:]))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(5)
z = torch.randn(M, 320, device=dev, generator=g).to(torch.bfloat16)
W = (torch.randn(V, 320, device=dev, generator=g) * 0.02).to(torch.bfloat16)
b = torch.randn(V, device=dev, generator=g) * 0.1
tgt = torch.randint(0, V, (M,), device=dev, generator=g).int()
st = ops.head_target_logit(z, W, b, tgt)
r = ops.head_topk_local(z, W, b, tgt, k, st=st)
torch.cuda.synchronize()
print("ok", r["idx"][0, :5].tolist())
