"""GPU: end-to-end parity of the product `Graphormer` (packed, libmobgt kernels, bf16 GEMMs) against the oracle
restatement of model_fqandtoyo.py (padded, CPU fp32) on the same seeded inputs and the same weights.

Tolerance: 2e-2 relative (north_star, bf16 mode) on logits, loss and gradients; dropout off on both sides."""
import numpy as np
import pytest
import torch

import model_oracle as mo

pytestmark = pytest.mark.gpu
HP = dict(n_layers=2, num_heads=8, hidden_dim=128, dropout_rate=0.0, intput_dropout_rate=0.0, weight_decay=0.01, ffn_dim=256,
          warmup_updates=10, tot_updates=100, peak_lr=2e-4, end_lr=1e-9, edge_type="multi_hop", multi_hop_max_dist=20,
          attention_dropout_rate=0.0)


def build(dataset_name, cfg="tiny", B=6, cap=12, seed=1, n_fixed=None):
    from mobgt_b200 import collator, model, synth
    w = synth.make_world(cfg, seed=seed, dataset_name=dataset_name)
    items = synth.make_items(w, B, cap, seed=seed, n_fixed=n_fixed)
    torch.manual_seed(seed)
    om = mo.Graphormer(w, n_layers=HP["n_layers"], ffn_dim=HP["ffn_dim"], dataset_name=dataset_name).eval()
    # non-trivial values in the zero-initialised / tiny tables so that every term is exercised
    with torch.no_grad():
        for emb in (om.edge_encoder, om.rel_pos_encoder, om.poi_pos_encoder):
            emb.weight.mul_(0.3)
            emb.weight[0].zero_()
        om.edge_dis_encoder.weight.mul_(0.3)
    pm = model.Graphormer(dataset_name=dataset_name, world=w, **HP).cuda().eval()
    missing, unexpected = pm.load_state_dict(om.state_dict(), strict=False)
    assert not missing, missing
    ob = mo.collate([mo.preprocess_item(it, hop_cap=20) for it in items], w, multi_hop_max_dist=20, rel_pos_max=1024)
    pb = collator.collator_toyota(items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=w)
    return w, om, pm, ob, pb


def rel_err(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


@pytest.mark.parametrize("dataset_name", ["toyotagraph", "foursquaregraph", "gowalla_nevda"])
def test_forward_logits_match_oracle(lib_built, dataset_name):
    w, om, pm, ob, pb = build(dataset_name)
    with torch.no_grad():
        ref = om(ob)
        got = pm(pb)
    assert got[0].shape == ref[0].shape and got[1].shape == ref[1].shape
    assert rel_err(got[0].float().cpu(), ref[0]) <= 2e-2
    assert rel_err(got[1].float().cpu(), ref[1]) <= 2e-2
    assert torch.equal(pm.cat_target.cpu(), om.cat_target)


@pytest.mark.parametrize("dataset_name,cfg,B,cap", [("toyotagraph", "tiny", 6, 12), ("gowalla_nevda", "c1", 5, 40)])
def test_loss_and_gradients_match_oracle(lib_built, dataset_name, cfg, B, cap):
    w, om, pm, ob, pb = build(dataset_name, cfg, B, cap)
    om.train()
    pm.train()                    # dropout rates are 0; GCN dropout is p=0.3 in train mode -> keep GCNs in eval
    pm.pos_embed.p = 0.0          # LearnablePositionalEncoding's hard-coded Dropout(0.1) (model_fqandtoyo.py:334) off too
    for m in (om, pm):
        m.poi_distance_model.eval()
        m.poi_cat_model.eval()
    lref = om.training_loss(ob)
    lref.backward()
    lgot = pm.training_step(pb)
    lgot.backward()
    assert abs(lgot.item() - lref.item()) <= 2e-2 * abs(lref.item())
    ref_g = {k: p.grad for k, p in om.named_parameters() if p.grad is not None}
    gmax = max(r.abs().max().item() for r in ref_g.values())
    bad = []
    for k, p in pm.named_parameters():
        if k not in ref_g:
            continue
        g = p.grad
        r = ref_g[k]
        if g is None:             # e.g. fre_embed_model: only its (zero, gradient-free) padding row is ever read
            assert r.abs().max().item() == 0.0, k
            continue
        scale = r.abs().max().item()
        if scale < 1e-6 * gmax:   # mathematically-zero gradients (softmax is invariant to the key bias): rounding noise only
            assert g.abs().max().item() < 1e-3 * gmax, k
            continue
        g = g.float().cpu()
        normwise = (g - r).norm().item() / r.norm().item()
        elem = (g - r).abs().max().item() / scale
        # bf16 mode: 2e-2 per op (north_star); through 2 encoder layers of bf16 GEMMs the whole-tensor error stays
        # below 5e-2 normwise and no element is off by more than 12 % of the tensor's largest entry
        if normwise > 5e-2 or elem > 0.12:
            bad.append((k, normwise, elem))
    assert not bad, bad


def test_metrics_match_reference_semantics(lib_built):
    from mobgt_b200 import metrics
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(64, 500, generator=g)
    target = torch.randint(1, 500, (64,), generator=g)
    target[40] = 0                                   # reference breaks the batch loop here (model_fqandtoyo.py:88-89)
    acc, ndcg = metrics.get_acc(target.cuda(), scores.cuda())
    racc, rndcg = mo.get_acc(target, scores)
    assert np.allclose(acc, racc) and np.allclose(ndcg, rndcg)
    t2 = torch.randint(0, 500, (64,), generator=g)
    assert abs(metrics.MRR_metric(t2.cuda(), scores.cuda()) - mo.mrr_metric(t2, scores)) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("split_head", [False, True])
def test_cuda_graph_step_matches_eager(lib_built, split_head):
    """graphs.GraphedTrainStep: replaying the captured forward + backward on a freshly loaded batch gives the same loss and
    the same gradients as the eager step (eval mode: no dropout, so the comparison is exact up to atomics-free determinism).
    split_head: the step captured as TWO graphs cut behind the head backward (what the data-parallel trainer replays around
    the early all-reduce of the out_proj bucket)."""
    import torch
    from mobgt_b200 import collator, graphs, model as M, ops, synth
    world = synth.make_world("tiny", seed=1)
    dev = torch.device("cuda")
    torch.manual_seed(3)
    model = M.Graphormer(dataset_name="toyotagraph", world=world, n_layers=2, num_heads=8, hidden_dim=128, dropout_rate=0.1,
                         intput_dropout_rate=0.1, weight_decay=0.01, ffn_dim=256, warmup_updates=10, tot_updates=100, peak_lr=2e-4,
                         end_lr=1e-9, edge_type="multi_hop", multi_hop_max_dist=20, attention_dropout_rate=0.1).to(dev).eval()
    params = list(model.parameters())
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    latlon = torch.from_numpy(world.latlon).to(dev)
    mk = lambda seed: collator.collate_packed(synth.make_items(world, 6, 12, seed=seed, cfg_id=2, n_fixed=12), world, latlon, 512, 20, 1024)
    b0, b1 = mk(1), mk(2)
    try:
        g = graphs.GraphedTrainStep(model, flat, b0, split_head=split_head)
    except graphs.GraphCaptureError as e:
        pytest.skip(f"not capturable here: {e}")
    for b in (b1, mk(1)):
        g.load(b)
        loss_g = float(g.run())
        grad_g = flat.clone()
        flat.zero_()
        ops.grads_zeroed(flat)       # same gradient-delivery path as the captured step (fp32 weight gradients written in place)
        loss_e = model.training_step(b)
        loss_e.backward()
        assert abs(loss_g - float(loss_e)) <= 1e-5 * max(1.0, abs(float(loss_e)))
        scale = flat.abs().max().item()
        assert (grad_g - flat).abs().max().item() <= 1e-4 * scale + 1e-7
    with pytest.raises(graphs.ShapeMismatch):
        g.load(collator.collate_packed(synth.make_items(world, 6, 12, seed=5, cfg_id=2, n_fixed=9), world, latlon, 512, 20, 1024))
