"""GPU: the REAL Gowalla-Nevada data set of the reference (tests/golden/gowalla_nevda_real.npz, frozen through the reference's
own dataset code: tests/golden/make_gowalla_real.py) through the whole path — collation + K1, both precisions of the model
against the oracle, and `entry` training / evaluating on the real splits."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
NPZ = os.path.join(HERE, "golden", "gowalla_nevda_real.npz")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def real():
    from mobgt_b200 import owndata
    return owndata.unpack_dataset(np.load(NPZ))


@pytest.mark.parametrize("precision,tol", [(16, 2e-2), (32, 1e-5)])
def test_real_batch_fields_and_logits_match_oracle(lib_built, real, precision, tol):
    """A batch of real trajectories (the first train graphs of <= 96 nodes, real POI / category graphs): the collated integer
    fields (rel_pos, edge_input from K1; degrees, x) equal the oracle's bit for bit, both heads' logits are within the mode's
    tolerance (bf16: 2e-2, fp32: 1e-5) of the fp32 oracle, and the training loss likewise."""
    import model_oracle as mo
    from mobgt_b200 import collator, model
    world, splits = real
    items = [it for it in splits["train"][:80] if len(it.x) <= 96][:32]
    assert max(len(it.x) for it in items) > 40
    hp = dict(n_layers=6, num_heads=8, hidden_dim=128, dropout_rate=0.0, intput_dropout_rate=0.0, weight_decay=0.01, ffn_dim=1024,
              warmup_updates=10, tot_updates=100, peak_lr=2e-4, end_lr=1e-9, edge_type="multi_hop", multi_hop_max_dist=20,
              attention_dropout_rate=0.0)
    torch.manual_seed(5)
    om = mo.Graphormer(world, n_layers=6, ffn_dim=1024, dataset_name="gowalla_nevda", multi_hop_max_dist=20).eval()
    with torch.no_grad():
        for emb in (om.edge_encoder, om.rel_pos_encoder, om.poi_pos_encoder):
            emb.weight.mul_(0.3)
            emb.weight[0].zero_()
        om.edge_dis_encoder.weight.mul_(0.3)
    pm = model.Graphormer(dataset_name="gowalla_nevda", world=world, precision=precision, **hp).cuda().eval()
    missing, _ = pm.load_state_dict(om.state_dict(), strict=False)
    assert not missing, missing
    ob = mo.collate([mo.preprocess_item(it, hop_cap=20) for it in items], world, multi_hop_max_dist=20, rel_pos_max=1024)
    pb = collator.collator_gowalla(items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=world)
    for name in ("x", "rel_pos", "edge_input", "in_degree", "out_degree", "y", "user"):
        a, r = getattr(pb, name), getattr(ob, name)
        assert tuple(a.shape) == tuple(r.shape), (name, a.shape, r.shape)
        assert torch.equal(a.cpu().long(), r.long()), name
    with torch.no_grad():
        ref = om(ob)
        got = pm(pb)
    for a, r in zip(got, ref):
        assert (a.float().cpu() - r).abs().max().item() <= tol * max(1.0, r.abs().max().item())
    for m_ in (om, pm):
        m_.train()
        m_.poi_distance_model.eval()
        m_.poi_cat_model.eval()
    pm.pos_embed.p = 0.0
    lref, lgot = om.training_loss(ob).item(), pm.training_step(pb).item()
    assert abs(lgot - lref) <= tol * abs(lref), (lgot, lref)
    model.ops.enable_tf32(True)


def test_entry_trains_and_evaluates_on_the_real_data_set(lib_built, real, tmp_path, capsys):
    """`entry --data_npz tests/golden/gowalla_nevda_real.npz`: four optimizer steps over real 256-trajectory batches (graphs up
    to 512 nodes: T = 513), a checkpoint, then the evaluation pass over the WHOLE real test split prints the reference's metric
    lines; every test trajectory of <= 512 nodes is counted once."""
    from mobgt_b200 import entry
    world, splits = real
    args = ["--dataset_name", "gowalla_nevda", "--gpus", "1", "--precision", "16", "--batch_size", "256", "--hidden_dim", "128",
            "--num_heads", "8", "--n_layers", "6", "--ffn_dim", "1024", "--dropout_rate", "0.1", "--intput_dropout_rate", "0.1",
            "--attention_dropout_rate", "0.1", "--weight_decay", "0.01", "--peak_lr", "2e-4", "--end_lr", "1e-9", "--edge_type",
            "multi_hop", "--warmup_updates", "40000", "--tot_updates", "400000", "--seed", "1", "--max_epochs", "1",
            "--multi_hop_max_dist", "20", "--data_npz", NPZ, "--limit_train_steps", "4", "--default_root_dir", str(tmp_path)]
    r = entry.cli_main(args)
    out = capsys.readouterr().out
    assert r["steps"] == 4 and np.isfinite(r["loss"])
    assert "4970 train / 1899 test trajectories, 3679 POIs, 253 categories" in out
    assert "ACC @1:" in out and "MRR:" in out
    assert r["metrics"]["n"] == sum(len(it.x) <= 512 for it in splits["test"])
