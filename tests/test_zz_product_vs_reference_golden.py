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
DATASETS = ("toyotagraph", "gowalla_nevda", "foursquaregraph", "toyotagraph_n40")      # golden cases


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
