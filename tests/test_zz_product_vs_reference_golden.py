"""GPU: the PRODUCT path (packed collator + libmobgt kernels, bf16 GEMMs) against the outputs of the REAL reference frozen in
tests/golden/model_golden_<dataset>.npz (tests/golden/make_model_golden.py: unmodified wrapper / collator / model_fqandtoyo run
on the CPU in the build container).  Same seeded items, same per-name seeded weights.  Integer fields bit-exact; logits within
the north_star's bf16 tolerance (2e-2 relative).  (Runs last: the file name sorts after the kernel-level parity tests.)"""
import importlib.util
import os

import numpy as np
import pytest
import torch

import model_oracle as mo

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DATASETS = ("toyotagraph", "gowalla_nevda", "foursquaregraph", "toyotagraph_n40", "toyotagraph_n128", "gowalla_nevda_n256")      # golden cases


def _gen():
    spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(HERE, "golden", "make_model_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("dataset_name", DATASETS)
def test_product_matches_reference_golden(lib_built, dataset_name):
    from mobgt_b200 import collator, model
    g = _gen()
    gold = np.load(os.path.join(HERE, "golden", f"model_golden_{dataset_name}.npz"))
    world, items = g.make_world_and_items(dataset_name)
    # weights: the oracle module is only the carrier of the per-name seeded values (its state_dict keys are the reference's)
    om = mo.Graphormer(world, n_layers=g.HP["n_layers"], ffn_dim=g.HP["ffn_dim"], dataset_name=g.CASES[dataset_name][0]).eval()
    with torch.no_grad():
        for name, p in om.named_parameters():
            p.copy_(g.golden_weights(name, tuple(p.shape)))
    pm = model.Graphormer(dataset_name=g.CASES[dataset_name][0], world=world, **g.HP).cuda().eval()
    missing, _ = pm.load_state_dict(om.state_dict(), strict=False)
    assert not missing, missing
    pb = collator.collator_toyota(items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=world)
    # ---- K1 + collation against the reference's wrapper.preprocess_item + collator_*: bit for bit
    for name in ("x", "rel_pos", "edge_input", "in_degree", "out_degree", "y", "user"):
        ref = torch.from_numpy(gold["f_" + name])
        got = getattr(pb, name).cpu()
        assert tuple(got.shape) == tuple(ref.shape), (name, got.shape, ref.shape)
        assert torch.equal(got.long(), ref.long()), name
    assert torch.equal(pb.attn_bias.cpu(), torch.from_numpy(gold["f_attn_bias"]))
    # ---- forward against the reference's Graphormer.forward
    with torch.no_grad():
        poi, cat = pm(pb)
    rp, rc = torch.from_numpy(gold["poi_logits"]), torch.from_numpy(gold["cat_logits"])
    assert tuple(poi.shape) == tuple(rp.shape) and tuple(cat.shape) == tuple(rc.shape)
    assert (poi.float().cpu() - rp).abs().max().item() <= 2e-2 * max(1.0, rp.abs().max().item())
    assert (cat.float().cpu() - rc).abs().max().item() <= 2e-2 * max(1.0, rc.abs().max().item())
    assert torch.equal(pm.cat_target.cpu().view(-1).long(), torch.from_numpy(gold["cat_target"]).long())


def test_attention_mixed_batch_with_single_token_tails(lib_built):
    """K3 forward + backward on ONE batch that mixes graphs of 128 m + 1 tokens (single-token tail path: 128 and 256 nodes)
    with ordinary sizes (1, 5, 60, 127, 129 nodes), dropout on and off, against torch autograd with the same mask."""
    from mobgt_b200 import collator, ops, synth
    from test_k2_k3_k4 import tables, torch_attention_diff
    w = synth.make_world("c1", seed=1)
    items = []
    for k, n in enumerate((128, 5, 256, 60, 1, 127, 128, 129)):
        items += synth.make_items(w, 1, 512, seed=40 + k, n_fixed=n)
    b = collator.collator_toyota(items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=w)
    B = len(items)
    R, Pp, E, W, t = tables(seed=9)
    cu = [x.cuda().contiguous() for x in (R, Pp, E, W.view(-1), t.view(-1))]
    bias = ops.bias_fwd_raw(b, *cu, out_dtype=torch.bfloat16)
    ntok = int(b.tok_pos.numel())
    gen = torch.Generator().manual_seed(33)
    qkv = (torch.randn(ntok, 3 * 192, generator=gen) * 1.2).to(torch.bfloat16)
    dout = (torch.randn(ntok, 192, generator=gen)).to(torch.bfloat16)
    tok_off = b.tok_off.cpu().numpy()
    for p, seed in ((0.0, 0), (0.1, 0x1234567890ABCDEF)):
        out, lse = ops.attn_fwd_raw(qkv.cuda(), bias, b, drop_p=p, seed=seed)
        dbias = torch.full(bias.shape, float("nan"), dtype=torch.float32, device="cuda")
        dqkv = ops.attn_bwd_raw(qkv.cuda(), bias, out, dout.cuda(), lse, b, dbias, 0, drop_p=p, seed=seed)
        torch.cuda.synchronize()
        q32 = qkv.float().requires_grad_(True)
        b32 = bias.float().cpu().requires_grad_(True)
        ref, _ = torch_attention_diff(q32, b32, tok_off, drop=(p, seed) if p > 0 else None)
        err = (out.float().cpu() - ref.detach()).abs().max().item()
        assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"p={p}: attention out max err {err}"
        (ref * dout.float()).sum().backward()
        gq = q32.grad
        err = (dqkv.float().cpu() - gq).abs().max().item()
        assert err <= 2e-2 * max(1.0, gq.abs().max().item()), f"p={p}: dqkv max err {err}"
        for g in range(B):
            Tg = int(tok_off[g + 1] - tok_off[g])
            gb = b32.grad[g, :, :Tg, :Tg]
            got = dbias[g, :, :Tg, :Tg].cpu()
            assert torch.isfinite(got).all(), (p, g)
            assert (got - gb).abs().max().item() <= 2e-2 * max(1.0, gb.abs().max().item()), (p, g)
