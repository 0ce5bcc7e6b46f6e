"""Host-side multi-GPU plumbing of the MobGT hot path: one process per GPU, torch.distributed.

The path shards naturally, so there are exactly three exchange steps and nothing else:
  * training    — trajectory graphs are independent: rank r takes graphs g = r (mod world) (`shard_graphs`); the only
                  collective is ONE all-reduce of the flat fp32 gradient buffer per step (`FlatGrads`), which replaces the
                  reference's Lightning-DDP bucketed all-reduce (entry.py:141, README.md:62 `--accelerator ddp`);
  * eval head   — out_proj is sharded by vocabulary rows (`shard_vocab`); per row of z the ranks exchange the target logit
                  (all-reduce MAX), the per-shard top-k lists (all-gather) and the partial rank counts (all-reduce SUM)
                  — `sharded_head_topk`;
  * metrics     — integer hit counts and float64 NDCG/MRR sums are summed (`reduce_metric_sums`), replacing the
                  reference's `sync_dist=True` mean-of-means (model_fqandtoyo.py:1524,1589).
Backend: NCCL over NVLink 5 / NVSwitch on the GPU box; the same code runs over gloo on CPU tensors in
tests/test_parallel_gloo.py (world_size 2), where the three local kernels are replaced by a torch checker.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def is_rank0(group=None):
    return _world(group)[1] == 0


def shard_graphs(num_graphs, rank, world, order=None, pad=True):
    """Indices of the graphs rank `rank` owns: positions rank, rank + world, ... of `order` (default 0..num_graphs-1).
    pad=True (torch's DistributedSampler(drop_last=False), what Lightning's DDP gives the reference, entry.py:141): the order
    is extended by wrapping around to a multiple of `world`, so EVERY rank gets the same number of graphs — hence the same
    number of batches and of gradient all-reduces per epoch (unequal counts would dead-lock the collective).  pad=False
    (evaluation): no graph is counted twice; ranks may differ by one graph, which is harmless because the metric exchange
    (`reduce_metric_sums`) happens once per epoch."""
    order = list(range(num_graphs)) if order is None else list(order)
    if pad and world > 1 and len(order) % world and len(order) > 0:
        need = world - len(order) % world
        order = order + [order[i % len(order)] for i in range(need)]
    return order[rank::world]


def shard_vocab(V, rank, world):
    """Row range [offset, offset + size) of out_proj.weight owned by `rank`; the last shard takes the remainder."""
    per = (V + world - 1) // world
    off = min(V, rank * per)
    return off, max(0, min(V, off + per) - off)


class FlatGrads:
    """All parameter gradients as views of ONE flat fp32 buffer, so the data-parallel exchange is a single all-reduce."""

    def __init__(self, params, device=None):
        self.params = [p for p in params]
        device = device if device is not None else self.params[0].device
        # every gradient starts on a 16-byte boundary (kernels may use 128-bit accesses); the padding stays zero
        self.flat = torch.zeros(sum((p.numel() + 3) // 4 * 4 for p in self.params), dtype=torch.float32, device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += (p.numel() + 3) // 4 * 4

    def zero_(self):
        self.flat.zero_()
        if self.flat.is_cuda:
            from . import ops
            ops.grads_zeroed(self.flat)      # the Linear backwards may now WRITE their gradients (ops._claim)

    def all_reduce_mean(self, group=None):
        world, _ = _world(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)
        return self.flat


def sharded_head_topk(kernels, z, W_shard, bias_shard, target, k, vocab_offset, group=None):
    """Vocabulary-parallel evaluation head.  `kernels` supplies the three local ops (mobgt_b200.ops on the GPU):
        head_target_logit(z, W, bias, target, vocab_offset) -> st [M]  (-inf where this shard does not own the target)
        head_topk_local(z, W, bias, target, k, vocab_offset, st=st) -> dict(val [M,k], idx [M,k] global ids, cnt [M])
        topk_merge_lists(val [M,S,k], idx [M,S,k]) -> (val [M,k], idx [M,k])   (ties -> lower index)
    Returns dict(val, idx, rank, st): rank = 0-based rank of each row's target over the WHOLE vocabulary."""
    world, _ = _world(group)
    st = kernels.head_target_logit(z, W_shard, bias_shard, target, vocab_offset)
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.MAX, group=group)
    loc = kernels.head_topk_local(z, W_shard, bias_shard, target, k, vocab_offset, st=st)
    if world == 1:
        return dict(val=loc["val"], idx=loc["idx"], rank=loc["cnt"], st=st)
    M = z.shape[0]
    gv = torch.empty(world * M, k, dtype=loc["val"].dtype, device=z.device)      # rank-major concatenation along dim 0
    gi = torch.empty(world * M, k, dtype=loc["idx"].dtype, device=z.device)
    dist.all_gather_into_tensor(gv, loc["val"].contiguous(), group=group)
    dist.all_gather_into_tensor(gi, loc["idx"].contiguous(), group=group)
    gv, gi = gv.view(world, M, k), gi.view(world, M, k)
    cnt = loc["cnt"].clone()
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    val, idx = kernels.topk_merge_lists(gv.permute(1, 0, 2).contiguous(), gi.permute(1, 0, 2).contiguous())
    return dict(val=val, idx=idx, rank=cnt, st=st)


def vocab_parallel_eval_head(kernels, z, W, bias, target, k, group=None):
    """The evaluation head with out_proj sharded by vocabulary rows over the ranks of `group` (BASELINE
    configs[4]).  Every rank brings the z rows / targets of ITS graphs; the rows of all ranks are all-gathered (padded to the
    largest per-rank count), every rank scores them against its row range of W (replicated weights: a slice, no copy of the
    rest), and the top-k lists / rank counts are merged by `sharded_head_topk`.  Returns this rank's rows:
    dict(idx [Br,k] global ids, val [Br,k], rank [Br])."""
    world, rank = _world(group)
    Br, K = z.shape
    V = W.shape[0]
    off, size = shard_vocab(V, rank, world)
    Ws, bs = W[off:off + size].contiguous(), (bias[off:off + size].contiguous() if bias is not None else None)
    if world == 1:
        r = sharded_head_topk(kernels, z, Ws, bs, target, k, off, group)
        return dict(idx=r["idx"], val=r["val"], rank=r["rank"])
    nb = torch.tensor([Br], dtype=torch.int64, device=z.device)
    dist.all_reduce(nb, op=dist.ReduceOp.MAX, group=group)
    Bm = int(nb.item())
    zp = torch.zeros(Bm, K, dtype=z.dtype, device=z.device)
    tp = torch.zeros(Bm, dtype=target.dtype, device=z.device)
    zp[:Br], tp[:Br] = z, target
    zg = torch.empty(world * Bm, K, dtype=z.dtype, device=z.device)
    tg = torch.empty(world * Bm, dtype=target.dtype, device=z.device)
    dist.all_gather_into_tensor(zg, zp, group=group)
    dist.all_gather_into_tensor(tg, tp, group=group)
    r = sharded_head_topk(kernels, zg, Ws, bs, tg, k, off, group)
    lo = rank * Bm
    return dict(idx=r["idx"][lo:lo + Br], val=r["val"][lo:lo + Br], rank=r["rank"][lo:lo + Br])


def reduce_metric_sums(sums, n_samples, group=None, device=None):
    """sums: dict name -> float (per-rank SUMS over samples, as get_acc / MRR_metric return them).  Returns the global
    dict of sums and the global sample count (float64 all-reduce SUM)."""
    world, _ = _world(group)
    keys = sorted(sums)
    t = torch.tensor([float(sums[k]) for k in keys] + [float(n_samples)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: float(v) for k, v in zip(keys, t[:-1])}, int(t[-1].item())
