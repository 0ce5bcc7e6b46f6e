"""CPU, world_size 2 over gloo: the host-side multi-GPU logic of mobgt_b200/parallel.py.

The three local kernels of the sharded evaluation head are replaced by a plain-torch checker, so what is tested is the
collective choreography (MAX of target logits -> local lists/counts -> all-gather -> SUM -> merge), the graph / vocabulary
sharding arithmetic, the flat-gradient all-reduce and the metric-sum reduction.  The CUDA kernels behind the same
interface are tested on the GPU in tests/test_k5_head.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class TorchHeadChecker:
    """Reference semantics of the three K5 entry points on explicit logits (ties -> lower global index)."""

    @staticmethod
    def _logits(z, W, bias):
        s = z.float() @ W.float().t()
        return s + bias if bias is not None else s

    @classmethod
    def head_target_logit(cls, z, W, bias, target, vocab_offset=0):
        s = cls._logits(z, W, bias)
        loc = target.long() - vocab_offset
        own = (loc >= 0) & (loc < W.shape[0])
        st = torch.full((z.shape[0],), float("-inf"))
        st[own] = s[own].gather(1, loc[own].view(-1, 1)).view(-1)
        return st

    @classmethod
    def head_topk_local(cls, z, W, bias, target, k, vocab_offset=0, st=None):
        s = cls._logits(z, W, bias)
        V = W.shape[0]
        gidx = torch.arange(V).view(1, -1) + vocab_offset
        order = torch.argsort(-s, dim=1, stable=True)[:, :k]
        cnt = (s > st.view(-1, 1)).sum(1) + ((s == st.view(-1, 1)) & (gidx < target.long().view(-1, 1))).sum(1)
        return dict(val=s.gather(1, order), idx=(order + vocab_offset).int(), cnt=cnt.int(), st=st)

    @staticmethod
    def topk_merge_lists(val, idx):
        M, S, k = val.shape
        v, i = val.reshape(M, S * k), idx.reshape(M, S * k).long()
        key = torch.argsort(i, dim=1, stable=True)                      # lower index first ...
        v, i = v.gather(1, key), i.gather(1, key)
        o = torch.argsort(-v, dim=1, stable=True)[:, :k]                # ... among equal values
        return v.gather(1, o), i.gather(1, o).int()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mobgt_b200 import parallel
        from mobgt_b200.ops import metrics_from_rank
        out = {}
        # ---- vocabulary-sharded head == unsharded head
        g = torch.Generator().manual_seed(3)
        M, V, K, k = 37, 1001, 32, 10
        z = torch.randn(M, K, generator=g)
        W = torch.randn(V, K, generator=g).round(decimals=1)            # coarse values -> exact ties across shards
        bias = torch.randn(V, generator=g).round(decimals=1)
        target = torch.randint(0, V, (M,), generator=g).int()
        off, size = parallel.shard_vocab(V, rank, world)
        r = parallel.sharded_head_topk(TorchHeadChecker, z, W[off:off + size], bias[off:off + size], target, k, off)
        st_full = TorchHeadChecker.head_target_logit(z, W, bias, target, 0)
        full = TorchHeadChecker.head_topk_local(z, W, bias, target, k, 0, st=st_full)
        out["head"] = bool(torch.equal(r["idx"], full["idx"]) and torch.equal(r["val"], full["val"]) and
                           torch.equal(r["rank"], full["cnt"]) and torch.equal(r["st"], st_full))
        # ---- metric sums: each rank evaluates its own rows, the reduced sums equal the single-process sums
        rows = parallel.shard_graphs(M, rank, world, pad=False)        # evaluation: no graph counted twice
        mine = metrics_from_rank(full["cnt"][rows], target[rows].long() + 1)
        tot, n = parallel.reduce_metric_sums(mine, len(rows))
        ref = metrics_from_rank(full["cnt"], target.long() + 1)
        out["metrics"] = bool(n == M and all(abs(tot[key] - ref[key]) < 1e-9 for key in ref))
        # ---- flat gradient all-reduce: mean over ranks, every parameter's .grad is a view of the flat buffer
        torch.manual_seed(0)
        lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        fg = parallel.FlatGrads(lin.parameters())
        x = torch.full((4, 5), float(rank + 1))
        fg.zero_()
        lin(x).sum().backward()
        local = fg.flat.clone()
        fg.all_reduce_mean()
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        out["grads"] = bool(torch.allclose(fg.flat, sum(both) / world) and
                            all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in lin.parameters()) and
                            torch.equal(lin[0].weight.grad.reshape(-1), fg.flat[:35]))
        # ---- sharding arithmetic
        cover = sorted(sum((parallel.shard_graphs(11, r_, world, pad=False) for r_ in range(world)), []))
        padded = [parallel.shard_graphs(11, r_, world) for r_ in range(world)]       # training: equal counts on every rank
        spans = [parallel.shard_vocab(1001, r_, world) for r_ in range(world)]
        out["shards"] = bool(cover == list(range(11)) and len({len(p_) for p_ in padded}) == 1 and spans[0][0] == 0 and sum(s for _, s in spans) == 1001 and
                             all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1)))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    res = [q.get(timeout=10) for _ in range(world)]
    for rank, out in res:
        assert all(out.values()), (rank, out)


def test_single_process_paths():
    from mobgt_b200 import parallel
    assert parallel.shard_vocab(10, 0, 1) == (0, 10)
    assert parallel.shard_vocab(10, 3, 4) == (9, 1)
    assert parallel.shard_vocab(2, 3, 4) == (2, 0)
    sums, n = parallel.reduce_metric_sums({"acc1": 3.0, "mrr": 1.5}, 7)
    assert sums == {"acc1": 3.0, "mrr": 1.5} and n == 7
