"""GPU: K5 — head GEMM on tcgen05 fused with per-row top-k / rank (csrc/k5_head.cu)."""
import numpy as np
import pytest
import torch

import model_oracle as mo

pytestmark = pytest.mark.gpu


def case(M, V, K=320, seed=0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, K, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(V, generator=g) * 0.1
    target = torch.randint(0, V, (M,), generator=g).int()
    return z, W, bias, target


@pytest.mark.parametrize("M,V,k", [(256, 60001, 20), (100, 3680, 10), (7, 130, 5), (300, 12345, 20)])
def test_head_topk_rank_vs_logits(lib_built, M, V, k):
    from mobgt_b200 import ops
    z, W, bias, target = case(M, V)
    out = ops.head_topk_local(z.cuda(), W.cuda(), bias.cuda(), target.cuda(), k, dump_logits=True)
    torch.cuda.synchronize()
    logits = out["logits"].cpu()
    ref = z.float() @ W.float().t() + bias
    # (1) the GEMM: fp32 accumulation of exact bf16 products
    assert (logits - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
    # (2) s_t comes out of the same MMA arithmetic: bitwise equal to the dumped logit
    st = logits.gather(1, target.long().view(-1, 1)).view(-1)
    assert torch.equal(out["st"].cpu(), st)
    # (3) top-k indices and ranks are exact w.r.t. the logits the kernel itself produced
    tv, ti = logits.topk(k, dim=1)
    assert torch.equal(out["idx"].cpu().long(), ti)
    assert torch.equal(out["val"].cpu(), tv)
    idx = torch.arange(V).view(1, -1)
    rank = (logits > st.view(-1, 1)).sum(1) + ((logits == st.view(-1, 1)) & (idx < target.long().view(-1, 1))).sum(1)
    assert torch.equal(out["cnt"].cpu().long(), rank)
    # (4) metrics from the rank == the reference's get_acc / MRR_metric on the same logits
    t = target.long().clone()
    if M > 50:
        t[40] = 0                                        # the reference's `break` quirk
        rank = (logits > logits.gather(1, t.view(-1, 1))).sum(1)
    m = ops.metrics_from_rank(rank.cuda(), t.cuda())
    acc, ndcg = mo.get_acc(t, logits)
    assert m["acc1"] == acc[2, 0] and m["acc5"] == acc[1, 0] and m["acc10"] == acc[0, 0] and m["acc20"] == acc[3, 0]
    assert abs(m["ndcg10"] - ndcg[0, 0]) < 1e-9 and abs(m["ndcg5"] - ndcg[1, 0]) < 1e-9
    assert abs(m["mrr"] - mo.mrr_metric(t, logits)) < 1e-9


def test_head_ties_prefer_lower_index(lib_built):
    from mobgt_b200 import ops
    M, V, K, k = 130, 1000, 64, 10
    z = torch.zeros(M, K).to(torch.bfloat16)
    z[:, 0] = 1
    W = torch.zeros(V, K).to(torch.bfloat16)
    W[:, 0] = torch.tensor([float((i * 7) % 13) for i in range(V)]).to(torch.bfloat16)     # many exact ties
    target = torch.randint(0, V, (M,)).int()
    out = ops.head_topk_local(z.cuda(), W.cuda(), None, target.cuda(), k, dump_logits=True)
    logits = out["logits"].cpu()
    order = torch.argsort(-logits, dim=1, stable=True)[:, :k]                              # ties -> lower index first
    assert torch.equal(out["idx"].cpu().long(), order)
    st = logits.gather(1, target.long().view(-1, 1))
    idx = torch.arange(V).view(1, -1)
    rank = (logits > st).sum(1) + ((logits == st) & (idx < target.long().view(-1, 1))).sum(1)
    assert torch.equal(out["cnt"].cpu().long(), rank)


def test_vocab_sharded_merge_on_one_gpu(lib_built):
    """Two vocabulary shards evaluated separately (as two ranks would) and merged == the unsharded result."""
    from mobgt_b200 import ops
    M, V, k = 200, 20001, 10
    z, W, bias, target = case(M, V, seed=3)
    zc, Wc, bc, tc = z.cuda(), W.cuda(), bias.cuda(), target.cuda()
    full = ops.head_topk_local(zc, Wc, bc, tc, k)
    cut = 9000
    shards = [(0, Wc[:cut].contiguous(), bc[:cut].contiguous()), (cut, Wc[cut:].contiguous(), bc[cut:].contiguous())]
    st = torch.stack([ops.head_target_logit(zc, w, b, tc, off) for off, w, b in shards]).max(0).values   # all-reduce MAX
    assert torch.equal(st, full["st"])
    loc = [ops.head_topk_local(zc, w, b, tc, k, off, st=st) for off, w, b in shards]
    val, idx = ops.topk_merge_lists(torch.stack([l["val"] for l in loc], 1), torch.stack([l["idx"] for l in loc], 1))
    assert torch.equal(idx, full["idx"]) and torch.equal(val, full["val"])
    assert torch.equal(loc[0]["cnt"] + loc[1]["cnt"], full["cnt"])                                      # all-reduce SUM
