"""Small instances of every libmobgt kernel family, for compute-sanitizer (scripts/sanitize.sh): K1 (cluster / DSMEM column
broadcast), K2 fwd / bwd, K3 fwd / bwd (general kernel incl. fold path, single-box two-CTA kernel), K3 in fp32 mode, K4, K5 (mbarrier ring),
K6, K7, K8, K9, K10 (TMA ring + TMEM double buffering), one training step of a 2-layer model."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from mobgt_b200 import collator, model as M, ops, synth
from mobgt_b200.optim import FlatAdamW

which = set(sys.argv[1:]) or {"k1", "k2", "k3", "k3f", "k5", "k7", "k8", "k9", "k10", "step"}
dev = torch.device("cuda")
w = synth.make_world("tiny", seed=1)
sizes = (12, 3, 129, 128, 40, 1, 257)            # fold tails at 129 (T = 130: no), 128 (T = 129), 256 + 1
items = []
for k, n in enumerate((12, 3, 128, 40, 1)):
    items += synth.make_items(w, 1, 512, seed=5 + k, n_fixed=min(n, 90), start=k)
if "k1" in which:
    big = synth.make_world("c1", seed=1)
    its = []
    for k, n in enumerate((5, 33, 128, 200, 300, 512)):     # cluster sizes 1, 2, 4, 8
        its += synth.make_items(big, 1, 512, seed=9 + k, n_fixed=n, start=k)
    b = collator.collator_toyota(its, max_node=512, multi_hop_max_dist=7, rel_pos_max=1024, world=big)
    torch.cuda.synchronize()
    print("k1 ok", int(b.maxdist.max()))
b = collator.collator_toyota(items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=w)
g = torch.Generator().manual_seed(0)
tabs = [torch.randn(512, 8, generator=g) * 0.3, torch.randn(64, 8, generator=g) * 0.3, torch.randn(128, 8, generator=g) * 0.3,
        torch.randn(128 * 64, generator=g) * 0.3, torch.randn(8, generator=g) * 0.3]
for t in tabs[:3]:
    t[0] = 0
cu = [t.to(dev).contiguous() for t in tabs]
bias = ops.bias_fwd_raw(b, *cu)
ntok = int(b.tok_pos.numel())
if "k2" in which:
    planes = torch.randn((2,) + tuple(bias.shape), device=dev).to(torch.bfloat16)
    ops.bias_bwd_raw(b, planes, cu[2], cu[3], 64)
    torch.cuda.synchronize()
    print("k2 ok")
if "k3" in which:
    big = synth.make_world("c1", seed=1)
    for ns in ((128, 5, 60), (256, 7), (12, 90, 128, 1)):
        its = []
        for k, n in enumerate(ns):
            its += synth.make_items(big, 1, 512, seed=20 + k, n_fixed=n, start=k)
        bb = collator.collator_toyota(its, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=big)
        bs = ops.bias_fwd_raw(bb, *cu)
        nt = int(bb.tok_pos.numel())
        qkv = torch.randn(nt, 576, device=dev).to(torch.bfloat16)
        dout = torch.randn(nt, 192, device=dev).to(torch.bfloat16)
        for p in (0.0, 0.1):
            out, lse = ops.attn_fwd_raw(qkv, bs, bb, drop_p=p, seed=77)
            pl = torch.zeros_like(bs)
            ops.attn_bwd_raw(qkv, bs, out, dout, lse, bb, pl, 2, drop_p=p, seed=77)            # single-box kernel when N <= 128
            d32 = torch.zeros(bs.shape, dtype=torch.float32, device=dev)
            ops.attn_bwd_raw(qkv, bs, out, dout, lse, bb, d32, 0, drop_p=p, seed=77)           # general kernel
        torch.cuda.synchronize()
    print("k3 ok")
if "k3f" in which:          # fp32-mode attention (csrc/k3_attn_f32.cu): dynamic shared memory sized by the largest graph
    big = synth.make_world("c1", seed=1)
    its = []
    for k, n in enumerate((70, 5, 1, 33)):
        its += synth.make_items(big, 1, 512, seed=40 + k, n_fixed=n, start=k)
    bb = collator.collator_toyota(its, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=big)
    bs = ops.bias_fwd_raw(bb, *cu, out_dtype=torch.float32)
    nt = int(bb.tok_pos.numel())
    qkv = torch.randn(nt, 576, device=dev)
    dout = torch.randn(nt, 192, device=dev)
    for p in (0.0, 0.1):
        out, lse = ops.attn_f32_fwd_raw(qkv, bs, bb, drop_p=p, seed=78)
        d32 = torch.zeros_like(bs)
        ops.attn_f32_bwd_raw(qkv, bs, out, dout, lse, bb, d32, 1, drop_p=p, seed=78)
    torch.cuda.synchronize()
    print("k3f ok")
if "k5" in which:
    z = torch.randn(200, 320, device=dev).to(torch.bfloat16)
    W = (torch.randn(3001, 320, device=dev) * 0.05).to(torch.bfloat16)
    t = torch.randint(0, 3001, (200,), device=dev).int()
    r = ops.head_topk_local(z, W, torch.randn(3001, device=dev), t, 10)
    torch.cuda.synchronize()
    print("k5 ok", int(r["cnt"].max()))
if "k7" in which:
    x = torch.randn(37, 3001, device=dev, requires_grad=True)
    t = torch.randint(0, 3001, (37,), device=dev)
    (ops.log_softmax_nll_loss(x, t) + ops.gradient_tail_loss(x, t, 0.2)).backward()
    torch.cuda.synchronize()
    print("k7 ok")
if "k10" in which:
    a = torch.randn(700, 192, device=dev).to(torch.bfloat16)
    w1 = (torch.randn(256, 192, device=dev) * 0.07).to(torch.bfloat16)
    bb1 = torch.randn(256, device=dev)
    h = ops.gemm_bf16(a, w1, bb1, mode=1)
    dy = torch.randn(700, 192, device=dev).to(torch.bfloat16)
    w2t = (torch.randn(256, 192, device=dev) * 0.07).to(torch.bfloat16)
    ops.gemm_bf16(dy, w2t, bb1, mode=2, a2=a, w2=w1, want_colsum=True)
    torch.cuda.synchronize()
    print("k10 ok", float(h.float().abs().max()))
if "step" in which or "k8" in which or "k9" in which:
    hp = dict(n_layers=2, num_heads=8, hidden_dim=128, dropout_rate=0.1, intput_dropout_rate=0.1, weight_decay=0.01, ffn_dim=256,
              dataset_name="toyotagraph", warmup_updates=10, tot_updates=100, peak_lr=2e-4, end_lr=1e-9, edge_type="multi_hop",
              multi_hop_max_dist=20, attention_dropout_rate=0.1)
    torch.manual_seed(1)
    m = M.Graphormer(world=w, **hp).to(dev).train()
    (opt,), _ = m.configure_optimizers()
    assert isinstance(opt, FlatAdamW)
    for _ in range(2):
        opt.zero_grad()
        loss = m.training_step(b)
        loss.backward()
        opt.step()
    with torch.no_grad():
        m.eval()
        r = m.test_step(b)
    torch.cuda.synchronize()
    print("step ok", float(loss), int(r["rank"].max()))
