"""CPU: oracle/model_oracle.py (the restatement every GPU parity test is measured against) reproduces the outputs of the REAL
reference — `wrapper.preprocess_item`, `collator.collator_toyota` and `model_fqandtoyo.Graphormer.forward` + losses, executed
unmodified in the build container by tests/golden/make_model_golden.py (import stubs for pytorch_lightning / torch_geometric /
ogb only) and frozen in tests/golden/model_golden_<dataset>.npz for the three constructor / forward / loss branches of the live
model (toyotagraph, gowalla_nevda, foursquaregraph).  This pins the model half of the oracle to the reference itself."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import model_oracle as mo

HERE = os.path.dirname(os.path.abspath(__file__))


def _gen():
    spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(HERE, "golden", "make_model_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


DATASETS = ("toyotagraph", "gowalla_nevda", "foursquaregraph", "toyotagraph_n40", "toyotagraph_n128", "gowalla_nevda_n256")      # golden cases


def _setup(dataset_name):
    g = _gen()
    gold = np.load(os.path.join(HERE, "golden", f"model_golden_{dataset_name}.npz"))
    world, items = g.make_world_and_items(dataset_name)
    ob = mo.collate([mo.preprocess_item(it, hop_cap=20) for it in items], world, multi_hop_max_dist=20, rel_pos_max=1024)
    return g, gold, world, ob


@pytest.mark.parametrize("dataset_name", DATASETS)
def test_oracle_collator_fields_equal_reference_collator(dataset_name):
    """wrapper.py:25-102 + collator.py:310-458 / 460-608 / 610-748 executed by the reference == the oracle's preprocess_item + collate, field by
    field, bit for bit (poi_pos excepted: the reference bins a distance pickle that is not shipped)."""
    g, gold, world, ob = _setup(dataset_name)
    for name in ("x", "rel_pos", "edge_input", "in_degree", "out_degree", "y", "user"):
        ref = torch.from_numpy(gold["f_" + name])
        got = getattr(ob, name)
        assert tuple(got.shape) == tuple(ref.shape), (name, got.shape, ref.shape)
        assert torch.equal(got.long(), ref.long()), name
    assert torch.equal(ob.attn_bias, torch.from_numpy(gold["f_attn_bias"]))          # 0 / -inf pattern
    assert torch.equal(ob.time_normal.float(), torch.from_numpy(gold["f_time_normal"]).float())


@pytest.mark.parametrize("dataset_name", DATASETS)
def test_oracle_forward_and_loss_equal_reference_model(dataset_name):
    """model_fqandtoyo.py:1123-1432 (eval mode) and the losses :545-550 / :1446-1471 of each dataset branch: the oracle with the
    same parameters (filled per name from the shared seeded generator) gives the reference's outputs within 1e-5 relative."""
    g, gold, world, ob = _setup(dataset_name)
    om = mo.Graphormer(world, n_layers=g.HP["n_layers"], ffn_dim=g.HP["ffn_dim"], dataset_name=g.CASES[dataset_name][0]).eval()
    with torch.no_grad():
        for name, p in om.named_parameters():
            p.copy_(g.golden_weights(name, tuple(p.shape)))
        poi, cat = om(ob)
        loss = om.training_loss(ob)
    rp, rc = torch.from_numpy(gold["poi_logits"]), torch.from_numpy(gold["cat_logits"])
    assert tuple(poi.shape) == tuple(rp.shape) and tuple(cat.shape) == tuple(rc.shape)
    assert (poi - rp).abs().max().item() <= 1e-5 * max(1.0, rp.abs().max().item()), (poi - rp).abs().max().item()
    assert (cat - rc).abs().max().item() <= 1e-5 * max(1.0, rc.abs().max().item()), (cat - rc).abs().max().item()
    assert torch.equal(om.cat_target.view(-1).long(), torch.from_numpy(gold["cat_target"]).long())
    assert abs(float(loss) - float(gold["loss"][0])) <= 1e-5 * abs(float(gold["loss"][0]))
    assert torch.equal(poi.argsort(dim=1, descending=True)[:, :10], rp.argsort(dim=1, descending=True)[:, :10])


@pytest.mark.parametrize("dataset_name", DATASETS)
def test_oracle_gradients_equal_reference_backward(dataset_name):
    """loss.backward() through the reference's own forward graph == through the oracle's: the norm of every parameter gradient,
    and the full gradients of the tables K2 / K4 backward produce (rel_pos / edge / edge_dis / poi_pos encoders, virtual
    distance, graph token) plus one encoder weight."""
    g, gold, world, ob = _setup(dataset_name)
    om = mo.Graphormer(world, n_layers=g.HP["n_layers"], ffn_dim=g.HP["ffn_dim"], dataset_name=g.CASES[dataset_name][0]).eval()
    with torch.no_grad():
        for name, p in om.named_parameters():
            p.copy_(g.golden_weights(name, tuple(p.shape)))
    # the live model rounds the edge encoding through fp16 in the bias build (model_fqandtoyo.py:1178-1198, also on the CPU):
    # the oracle's 'ref_half' bias mode restates exactly that; its 'fp32' mode (model.py:157-190) is what the kernels target
    om.training_loss(ob, bias_mode="ref_half").backward()
    grads = {n: p.grad for n, p in om.named_parameters() if p.grad is not None}
    names = [str(n) for n in gold["grad_names"]]
    assert len(names) > 40
    top = float(gold["grad_norms"].max())
    for n, ref_norm in zip(names, gold["grad_norms"]):
        assert n in grads, n
        got = float(grads[n].double().norm())
        # (some gradients are mathematically zero — e.g. linear_k.bias: softmax ignores a per-row constant — hence the atol)
        assert abs(got - ref_norm) <= 2e-3 * ref_norm + 1e-6 * top, (n, got, ref_norm)
    for n in g.GRAD_FULL:
        ref = torch.from_numpy(gold["g_" + n])
        # fp16 round trips amplify (n = 256: 65 k pairs summed behind an fp16 bmm -> 1.2e-3 on a gradient of 5e-5)
        tol = 2e-3 if n in ("edge_encoder.weight", "edge_dis_encoder.weight") else 1e-5
        assert (grads[n] - ref).abs().max().item() <= tol * ref.abs().max().item(), n


@pytest.mark.parametrize("tag", ["a", "b"])
def test_metrics_equal_reference(tag):
    """get_acc / MRR_metric (model_fqandtoyo.py:48-90, 122-131) run by the reference on seeded logits == the oracle's restatement
    and the product's device-side mirror (mobgt_b200.metrics), including the `break` at the first zero target (case b)."""
    from mobgt_b200 import metrics as pm
    gold = np.load(os.path.join(HERE, "golden", "metrics_golden.npz"))
    scores, y = torch.from_numpy(gold["scores"]), torch.from_numpy(gold["y_" + tag])
    for impl in (mo, pm):
        acc, ndcg = impl.get_acc(y, scores)
        assert np.array_equal(np.asarray(acc), gold["acc_" + tag]), impl.__name__
        assert np.allclose(np.asarray(ndcg), gold["ndcg_" + tag], rtol=1e-12, atol=0), impl.__name__
        mrr = getattr(impl, "MRR_metric", None) or impl.mrr_metric
        assert abs(float(mrr(y, scores)) - float(gold["mrr_" + tag][0])) <= 1e-9 * float(gold["mrr_" + tag][0])


def test_lr_schedule_equals_reference_get_lr():
    """mobgt_b200.lr.PolynomialDecayLR stepped like configure_optimizers drives it (model_fqandtoyo.py:1599-1616) follows the
    values of the reference's own `get_lr` (lr.py:18-32) per `_step_count`."""
    from mobgt_b200.lr import PolynomialDecayLR
    gold = np.load(os.path.join(HERE, "golden", "metrics_golden.npz"))["lr_by_step_count"]
    opt = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=2e-4)
    sch = PolynomialDecayLR(opt, warmup_updates=10, tot_updates=25, lr=2e-4, end_lr=1e-9, power=1.0)
    got = [opt.param_groups[0]["lr"]]                       # _step_count == 1 after construction
    for _ in range(len(gold) - 1):
        opt.step()
        sch.step()
        got.append(opt.param_groups[0]["lr"])
    assert np.allclose(np.array(got), gold, rtol=1e-12, atol=0)



@pytest.mark.parametrize("dataset_name", DATASETS)
def test_product_host_packing_equals_reference_collator(dataset_name):
    """The PRODUCT's host half of collation (mobgt_b200.collator.pack_host, numpy, runs in the DataLoader workers) against the
    fields the reference's wrapper.preprocess_item + collator_* produced: node ids, degrees (+1 shift), users, targets, time."""
    from mobgt_b200 import collator
    g = _gen()
    gold = np.load(os.path.join(HERE, "golden", f"model_golden_{dataset_name}.npz"))
    world, items = g.make_world_and_items(dataset_name)
    hp = collator.pack_host(items)

    def field(name):
        off, shape, dts, nbytes = hp.layout[name]
        return hp.buf[off:off + nbytes].view(np.dtype(dts)).reshape(shape)

    ns, no = field("n"), field("node_off")
    B, N = len(ns), int(ns.max())
    assert gold["f_x"].shape[:2] == (B, N)

    def padded(flat, dtype):
        out = np.zeros((B, N), dtype)
        for gi in range(B):
            out[gi, :ns[gi]] = flat[no[gi]:no[gi + 1]]
        return out

    assert np.array_equal(padded(field("x_nodes"), np.int64), gold["f_x"][:, :, 0])            # 0 = pad (collator.py:29-37)
    assert np.array_equal(padded(field("in_deg"), np.int64), gold["f_in_degree"])              # degree + 1, 0 = pad
    assert np.array_equal(padded(field("out_deg"), np.int64), gold["f_out_degree"])
    assert np.array_equal(field("user").reshape(-1), gold["f_user"].reshape(-1))               # uid + 1 (wrapper.py:39)
    assert np.array_equal(field("y").reshape(-1), gold["f_y"].reshape(-1))
    assert np.array_equal(padded(field("time_normal_nodes"), np.float32), gold["f_time_normal"][:, :, 0].astype(np.float32))
    # the u8 edge-type plane K1 consumes == attn_edge_type of wrapper.py:49-53 (count + 2 on edges), checked through rel_pos:
    # an edge (distance 1 -> rel_pos 2 after the collator's +1) exists exactly where the plane is non-zero, off the diagonal
    sq, feat = field("sq_off"), field("feat8")
    for gi in range(B):
        n = int(ns[gi])
        plane = feat[sq[gi]:sq[gi + 1]].reshape(n, n)
        edge = (plane != 0) & ~np.eye(n, dtype=bool)
        assert np.array_equal(edge, (gold["f_rel_pos"][gi, :n, :n] == 2) & ~np.eye(n, dtype=bool))


def test_model_flags_equal_reference():
    """Graphormer.add_model_specific_args (model_fqandtoyo.py:1618-1641): same flag names, defaults and types as the reference's
    own static method produced in the build container."""
    import argparse
    import json
    from mobgt_b200.model import Graphormer
    gold = json.loads(str(np.load(os.path.join(HERE, "golden", "metrics_golden.npz"))["model_flags_json"]))
    ap = Graphormer.add_model_specific_args(argparse.ArgumentParser())
    mine = {a.dest: [a.default, type(a.default).__name__, a.type.__name__ if a.type else None]
            for a in ap._actions if a.dest != "help"}
    assert json.loads(json.dumps(mine, sort_keys=True)) == gold


@pytest.mark.parametrize("dataset_name", DATASETS[:3])
def test_product_state_dict_keys_load_from_reference(dataset_name):
    """Every parameter of the PRODUCT model exists in the reference model's state_dict under the same name with the same shape
    (so a reference checkpoint loads into it; the reference has more entries — modules its live forward never calls)."""
    import json
    from mobgt_b200 import model
    g = _gen()
    gold = np.load(os.path.join(HERE, "golden", f"model_golden_{dataset_name}.npz"))
    ref_shapes = json.loads(str(gold["state_shapes_json"]))
    world, _ = g.make_world_and_items(dataset_name)
    pm = model.Graphormer(dataset_name=g.CASES[dataset_name][0], world=world, **g.HP)        # CPU construction: no kernels involved
    mine = {k: list(v.shape) for k, v in pm.named_parameters()}
    assert len(mine) > 60
    for k, shp in mine.items():
        assert k in ref_shapes, k
        assert ref_shapes[k] == shp, (k, ref_shapes[k], shp)


@pytest.mark.parametrize("dataset_name", DATASETS)
def test_eval_targets_match_reference_steps(dataset_name):
    """y_true of validation_step / test_step (model_fqandtoyo.py:1484-1496, 1530-1544): `y - 1` for the datasets in the
    reference's list, plain `y` for toyotagraph — frozen from the reference's own steps."""
    from mobgt_b200 import model
    g = _gen()
    gold = np.load(os.path.join(HERE, "golden", f"model_golden_{dataset_name}.npz"))
    world, _ = g.make_world_and_items(dataset_name)
    pm = model.Graphormer(dataset_name=g.CASES[dataset_name][0], world=world, **g.HP)        # CPU construction: no kernels involved

    class B:
        y = torch.from_numpy(gold["f_y"])

    assert torch.equal(pm.eval_targets(B), torch.from_numpy(gold["y_true_test"]))
    assert torch.equal(pm.eval_targets(B), torch.from_numpy(gold["y_true_val"]))


# ------------------------------------------------------------------------------------------------------- real data
def test_oracle_equals_reference_on_real_data():
    """The REAL Gowalla-Nevada data set of the reference (tests/golden/make_model_golden_real.py: the unmodified reference
    constructor on the archive's own Graph_*.csv — dense calculate_laplacian_matrix, dense GCN products — its preprocess_item,
    collator_gowalla, forward, GradientTailLoss and backward on real train trajectories): the oracle on the world that
    mobgt_b200.owndata builds from the same files (CSR adjacency; here through the committed packed fixture) reproduces the
    collated fields bit for bit, the logits and the loss within 1e-5, and the gradient norms."""
    from mobgt_b200 import owndata
    g = _gen()
    spec = importlib.util.spec_from_file_location("make_model_golden_real", os.path.join(HERE, "golden", "make_model_golden_real.py"))
    gr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gr)
    gold = np.load(os.path.join(HERE, "golden", "model_golden_gowalla_real.npz"))
    world, splits = owndata.unpack_dataset(np.load(os.path.join(HERE, "golden", "gowalla_nevda_real.npz")))
    items = gr.real_items(splits)
    assert [it.idx for it in items] == gold["item_idx"].tolist()
    ob = mo.collate([mo.preprocess_item(it, hop_cap=20) for it in items], world, multi_hop_max_dist=20, rel_pos_max=1024)
    for name in ("x", "rel_pos", "edge_input", "in_degree", "out_degree", "y", "user"):
        ref = torch.from_numpy(gold["f_" + name])
        got = getattr(ob, name)
        assert tuple(got.shape) == tuple(ref.shape), (name, got.shape, ref.shape)
        assert torch.equal(got.long(), ref.long()), name
    assert torch.equal(ob.attn_bias, torch.from_numpy(gold["f_attn_bias"]))
    om = mo.Graphormer(world, n_layers=g.HP["n_layers"], ffn_dim=g.HP["ffn_dim"], dataset_name="gowalla_nevda").eval()
    with torch.no_grad():
        for name, p in om.named_parameters():
            p.copy_(g.golden_weights(name, tuple(p.shape)))
        poi, cat = om(ob)
        loss = om.training_loss(ob)
    rp, rc = torch.from_numpy(gold["poi_logits"]), torch.from_numpy(gold["cat_logits"])
    assert tuple(poi.shape) == tuple(rp.shape) and tuple(cat.shape) == tuple(rc.shape)
    assert (poi - rp).abs().max().item() <= 1e-5 * max(1.0, rp.abs().max().item()), (poi - rp).abs().max().item()
    assert (cat - rc).abs().max().item() <= 1e-5 * max(1.0, rc.abs().max().item()), (cat - rc).abs().max().item()
    assert abs(float(loss) - float(gold["loss"][0])) <= 1e-5 * abs(float(gold["loss"][0]))
    om.training_loss(ob, bias_mode="ref_half").backward()
    grads = {n: p.grad for n, p in om.named_parameters() if p.grad is not None}
    names = [str(n) for n in gold["grad_names"]]
    assert len(names) > 40
    top = float(gold["grad_norms"].max())
    for n, ref_norm in zip(names, gold["grad_norms"]):
        got = float(grads[n].double().norm())
        assert abs(got - ref_norm) <= 2e-3 * ref_norm + 1e-6 * top, (n, got, ref_norm)
    for n in g.GRAD_FULL:
        ref = torch.from_numpy(gold["g_" + n])
        tol = 2e-3 if n in ("edge_encoder.weight", "edge_dis_encoder.weight") else 1e-5
        assert (grads[n] - ref).abs().max().item() <= tol * ref.abs().max().item(), n
