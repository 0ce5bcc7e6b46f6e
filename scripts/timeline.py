"""Pipeline timeline of one CTA of the attention backward (debug): python scripts/timeline.py [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mobgt_b200 import _C, collator, model as M, ops, synth

workload = sys.argv[1] if len(sys.argv) > 1 else "c2-dense128"
world = synth.make_world("c2", seed=1)
items = bench.make_workload(workload, world, 256, 0)
b = collator.collate_packed(items, world, None, 512, 20, 1024)
ntok = int(b.tok_pos.numel())
T = b.N + 1
bias = torch.randn(b.B, 8, T, ops.bias_pitch(T), device="cuda").to(torch.bfloat16)
qkv = torch.randn(ntok, 576, device="cuda").to(torch.bfloat16)
out, lse = ops.attn_fwd_raw(qkv, bias, b)
dout = torch.randn(ntok, 192, device="cuda").to(torch.bfloat16)
dbias = torch.zeros(bias.shape, dtype=torch.bfloat16, device="cuda")
tl = torch.zeros(256, dtype=torch.int64, device="cuda")
dp = float(os.environ.get("DROP", "0.1"))
for name, fn in (("bwd", lambda: ops.attn_bwd_raw(qkv, bias, out, dout, lse, b, dbias, 2, drop_p=dp, seed=77)),
                 ("fwd", lambda: ops.attn_fwd_raw(qkv, bias, b, drop_p=dp, seed=77))):
    fn()
    torch.cuda.synchronize()
    tl.zero_()
    _C.call("mobgt_debug_set_timeline", tl.data_ptr())
    fn()
    torch.cuda.synchronize()
    _C.call("mobgt_debug_set_timeline", None)
    t = tl.cpu().tolist()
    t0 = t[0]
    print(f"== {name}: cycles since CTA start (slot: cycles)")
    print("  head:", {k: t[k] - t0 for k in range(1, 8) if t[k]})
    for it in range(0, 30):
        row = t[8 + 8 * it: 16 + 8 * it]
        if any(row):
            print(f"  iter {it}:", [(v - t0 if v else None) for v in row])
