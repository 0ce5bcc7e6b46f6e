"""CPU: host logic added in round 2 — equal-length graph shards, packing validation, hop stride, dense -> packed conversion."""
import numpy as np
import pytest
import torch

import model_oracle as mo


def test_shard_graphs_equal_lengths_and_coverage():
    from mobgt_b200 import parallel
    for n, world in ((2049, 2), (10, 4), (7, 8), (16, 4), (1, 2)):
        order = list(np.random.default_rng(n).permutation(n))
        shards = [parallel.shard_graphs(n, r, world, order=order, pad=True) for r in range(world)]
        assert len({len(s) for s in shards}) == 1, (n, world)                   # same number of batches / all-reduces per rank
        assert set(i for s in shards for i in s) == set(range(n))               # every graph is seen
        assert sum(len(s) for s in shards) - n < world                           # at most world-1 repeats (DistributedSampler)
        ev = [parallel.shard_graphs(n, r, world, order=order, pad=False) for r in range(world)]
        assert sorted(i for s in ev for i in s) == list(range(n))               # evaluation: no graph counted twice


def test_hop_stride():
    from mobgt_b200.algos import hop_stride
    assert [hop_stride(d) for d in (1, 4, 5, 8, 20, 32)] == [4, 4, 8, 8, 20, 32]
    for bad in (0, 33, -1):
        with pytest.raises(ValueError):
            hop_stride(bad)


def test_pack_host_rejects_out_of_table_indices():
    """The reference raises IndexError when edge_input / degree indices leave the 128-row embedding tables; pack_host does too
    (instead of wrapping in uint8 or reading out of bounds on the device)."""
    from mobgt_b200 import collator, synth
    w = synth.make_world("tiny", seed=1)
    items = synth.make_items(w, 2, 12, seed=1)
    collator.pack_host(items)                                                    # fine
    bad = synth.make_items(w, 2, 12, seed=1)
    bad[0].edge_attr = bad[0].edge_attr.copy()
    bad[0].edge_attr[0] = 125                                                    # 125 + 3 = 128
    with pytest.raises(IndexError):
        collator.pack_host(bad)
    # a hub with 127 out-going edges: degree + 1 = 128
    n = 130
    src = np.zeros(127, np.int64)
    dst = np.arange(1, 128, dtype=np.int64)
    hub = synth.Data(idx=0, x=np.arange(1, n + 1).reshape(-1, 1) % w.P + 1, edge_index=np.stack([src, dst]),
                     edge_attr=np.ones(127, np.int64), y=np.array([1]), time=np.zeros((n, 1), np.int64),
                     time_normal=np.zeros((n, 1), np.float32), user=np.array([[0]]), cat=np.ones((n, 1), np.int64))
    with pytest.raises(IndexError):
        collator.pack_host([hub])


@pytest.mark.parametrize("dk", [20, 5])
def test_from_dense_matches_oracle_collation(dk):
    """Batch1.from_dense on a reference-shaped dense batch (the oracle's collator, pinned to the reference's) gives the packed
    fields pack_host + K1 produce: checked here against the oracle's own per-graph arrays, on the CPU."""
    from mobgt_b200 import collator, synth
    w = synth.make_world("tiny", seed=1)
    items = synth.make_items(w, 5, 12, seed=3)
    pre = [mo.preprocess_item(it, hop_cap=20) for it in items]
    ob = mo.collate(pre, w, multi_hop_max_dist=dk, rel_pos_max=1024)
    b = collator.Batch1.from_dense(ob, multi_hop_max_dist=dk, device="cpu")
    hp = collator.pack_host(items)
    assert b.B == 5 and b.N == int(hp.N) and b.dk == dk and b.hops % 4 == 0 and b.hops >= dk
    assert np.array_equal(b.n_host, hp.ns)
    views = {k: np.frombuffer(hp.buf[off:off + nb].tobytes(), dtype=np.dtype(dt)).reshape(shape)
             for k, (off, shape, dt, nb) in hp.layout.items()}
    for k in ("n", "sq_off", "node_off", "tok_off", "tok_graph", "tok_pos", "feat8", "x_nodes", "slot", "in_deg", "out_deg", "user",
              "y", "node_rows"):
        assert np.array_equal(getattr(b, k).numpy().astype(np.int64), views[k].astype(np.int64)), k
    for g, it in enumerate(pre):
        n = int(hp.ns[g])
        lo, hi = int(b.sq_off[g]), int(b.sq_off[g + 1])
        assert np.array_equal(b.rel_pos16[lo:hi].numpy().reshape(n, n), it.rel_pos.numpy() + 1)
        e = it.edge_input.numpy()[:, :, :dk, 0] + 1
        got = b.edge_in8[lo:hi].numpy().reshape(n, n, b.hops)
        h = min(dk, e.shape[2])
        assert np.array_equal(got[:, :, :h], e[:, :, :h]) and (got[:, :, h:] == 0).all()
        assert int(b.maxdist[g]) == int(it.rel_pos.max())
    # the dense views of the converted batch reproduce the dense batch it came from
    for name in ("x", "rel_pos", "in_degree", "out_degree", "poi_pos", "attn_bias"):
        assert torch.equal(getattr(b, name), getattr(ob, name)), name
    assert torch.equal(b.edge_input, ob.edge_input[:, :, :, :b.edge_input.shape[3]])


def test_pack_host_buckets_pad_only_the_tails():
    """pack_host(bucket=True): per-graph sizes and the real prefix of every array are those of the unpadded pack; padding tokens
    are graph-token rows (pos 0) that no tok_off range covers, padding nodes map to them one to one with padding-row indices."""
    from mobgt_b200 import collator, synth
    w = synth.make_world("tiny", seed=1)
    items = synth.make_items(w, 7, 12, seed=5)
    a, b = collator.pack_host(items), collator.pack_host(items, bucket=True)
    view = lambda hp: {k: np.frombuffer(hp.buf[off:off + nb].tobytes(), dtype=np.dtype(dt)).reshape(shape)
                       for k, (off, shape, dt, nb) in hp.layout.items()}
    va, vb = view(a), view(b)
    assert b.padded and not a.padded and b.B == a.B and np.array_equal(a.ns, b.ns)
    ntok = len(va["tok_pos"])
    assert len(vb["tok_pos"]) % 512 == 0 and len(vb["tok_pos"]) >= ntok
    assert b.N >= a.N and (b.N & (b.N - 1)) == 0 and (b.cells & (b.cells - 1)) == 0 and b.cells >= a.cells
    for k in ("n", "sq_off", "node_off", "tok_off", "user", "y", "idx"):
        assert np.array_equal(va[k], vb[k]), k
    for k in ("tok_graph", "tok_pos", "x_nodes", "slot", "in_deg", "out_deg", "node_rows", "time_normal_nodes"):
        assert np.array_equal(va[k], vb[k][:len(va[k])]), k
    assert np.array_equal(va["feat8"], vb["feat8"][:a.cells]) and not vb["feat8"][a.cells:].any()
    pad = len(vb["tok_pos"]) - ntok
    assert len(vb["x_nodes"]) - len(va["x_nodes"]) == pad                       # padding nodes == padding tokens
    assert not vb["tok_pos"][ntok:].any() and not vb["in_deg"][len(va["in_deg"]):].any()
    assert np.array_equal(vb["node_rows"][len(va["node_rows"]):], ntok + np.arange(pad))
    assert int(vb["tok_off"][-1]) == ntok                                         # no graph owns a padding token


def test_gradient_claims_write_once_then_accumulate():
    """ops._claim — the rule behind "gradients are written, not accumulated": a parameter whose gradient view lies in a flat
    buffer the trainer has just zeroed (ops.grads_zeroed) may be WRITTEN once; a second gradient before the next zeroing, a
    buffer that was never announced, a misaligned view or non-adjacent parameters fall back to accumulation.  Pure host logic
    (pointer arithmetic on the gradient views): runs on CPU tensors."""
    from mobgt_b200 import ops
    q, k, v = (torch.nn.Parameter(torch.zeros(8, 4)) for _ in range(3))
    b = torch.nn.Parameter(torch.zeros(8))
    flat = torch.zeros(8 * 4 * 3 + 8 + 4)
    off = 0
    for p_ in (q, k, v, b):
        p_.grad = flat[off:off + p_.numel()].view_as(p_)
        off += p_.numel()
    assert ops._claim((q,)) is None or ops._zero_epoch > 0          # never announced: no claim unless another test zeroed THIS range
    ops.grads_zeroed(flat)
    t = ops._claim((q, k, v))                                       # q | k | v back to back -> one [24, 4] view of the buffer
    assert t is not None and tuple(t.shape) == (24, 4) and t.data_ptr() == q.grad.data_ptr()
    t.fill_(1.0)
    assert float(q.grad.sum() + k.grad.sum() + v.grad.sum()) == 96.0 and float(b.grad.abs().sum()) == 0.0
    assert ops._claim((q,)) is None and ops._claim((k, v)) is None  # already written since the zeroing -> accumulate
    assert ops._claim((b,)) is not None                             # independent parameter
    ops.grads_zeroed(flat)
    assert ops._claim((q, v)) is None                               # not adjacent
    assert ops._claim((q, k)) is not None                           # (does not consume v)
    assert ops._claim((v,)) is not None
    other = torch.nn.Parameter(torch.zeros(4))
    other.grad = torch.zeros(5)[1:]                                 # 4-byte aligned only, and outside every announced buffer
    assert ops._claim((other,)) is None
    detached = torch.zeros(4, requires_grad=True) * 2               # not a leaf
    assert ops._claim((detached,)) is None


def test_head_split_is_even_and_bounded():
    from mobgt_b200 import ops
    for M, V in ((256, 60001), (4096, 125000), (4096, 1000000), (1, 97), (1024, 125000), (130, 3680)):
        ns = ops.head_split(M, V)
        assert ns % 2 == 0 and 2 <= ns <= 160, (M, V, ns)


def test_flat_parameter_order_groups_qkv_and_puts_the_head_last():
    """Graphormer._flat_param_order (the layout of the flat parameter / gradient buffers): every parameter exactly once, the
    q | k | v weights (and biases) of a layer adjacent — one GEMM / column-sum output writes all three gradients — and
    out_proj.weight last, so that what remains after its early all-reduce is one contiguous range."""
    from mobgt_b200 import model, synth
    w = synth.make_world("tiny", seed=1)
    m = model.Graphormer(dataset_name="toyotagraph", world=w, n_layers=2, num_heads=8, hidden_dim=128, dropout_rate=0.1,
                         intput_dropout_rate=0.1, weight_decay=0.01, ffn_dim=256, warmup_updates=10, tot_updates=100, peak_lr=2e-4,
                         end_lr=1e-9, edge_type="multi_hop", multi_hop_max_dist=20, attention_dropout_rate=0.1)
    order = m._flat_param_order()
    assert sorted(id(p) for p in order) == sorted(id(p) for p in m.parameters())
    pos = {id(p): i for i, p in enumerate(order)}
    for layer in m.layers:
        a = layer.self_attention
        assert pos[id(a.linear_k.weight)] == pos[id(a.linear_q.weight)] + 1 and pos[id(a.linear_v.weight)] == pos[id(a.linear_q.weight)] + 2
        assert pos[id(a.linear_k.bias)] == pos[id(a.linear_q.bias)] + 1 and pos[id(a.linear_v.bias)] == pos[id(a.linear_q.bias)] + 2
    assert order[-1] is m.out_proj.weight
