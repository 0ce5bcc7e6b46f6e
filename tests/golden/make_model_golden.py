"""Golden vectors for the MODEL half of the path, produced by the REAL reference code (run in the build container only).

    python tests/golden/make_model_golden.py          # writes tests/golden/model_golden_<dataset>.npz for the three POI datasets

The reference's `graphormer/model_fqandtoyo.py` cannot be imported as shipped: it needs pytorch_lightning, torch_geometric, ogb
(not installed, no network) and its dataset files.  None of those take part in the arithmetic of the path, so this script
  * registers import stubs for them (a LightningModule that is a plain nn.Module, empty torch_geometric / ogb name spaces),
  * writes the dataset files the constructor reads (`../dataset/<name>/raw/Graph_{cat,dist,adj,poi}.csv`,
    `../dataset/poi_data/{toyota,gowalla,tky}_distance.pkl`; model_fqandtoyo.py:636-1027) from the seeded synthetic `tiny` world
    (mobgt_b200/synth.py) into a temporary directory,
  * imports the UNMODIFIED reference modules from /root/reference/graphormer (model_fqandtoyo, modelGNN, collator, wrapper;
    `algos` is the compiled reference in oracle/_ref), instantiates `Graphormer(dataset_name=...)` on the CPU for each of
    toyotagraph, gowalla_nevda and foursquaregraph (the three constructor / forward / loss branches of the live model),
  * fills every parameter from a per-name seeded generator (`golden_weights`, shared with the test),
  * runs the reference `preprocess_item` + `collator_{toyota,gowalla,foursquare}` and the reference `forward` / losses on a
    seeded batch,
and stores the outputs.  tests/test_oracle_model_golden.py then checks that oracle/model_oracle.py — the restatement every
GPU parity test is measured against — reproduces them: that pins the oracle's model half to the reference itself.

Only `poi_pos` is taken from the oracle collator: the reference bins a POI distance pickle that is not shipped
(`poi_data/` is missing from the repository), `PoiWorld.poi_pos_bins` is the documented stand-in.
"""
import os
import pickle
import sys
import tempfile
import types
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/graphormer"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

HP = dict(n_layers=2, num_heads=8, hidden_dim=128, dropout_rate=0.0, intput_dropout_rate=0.0, weight_decay=0.01, ffn_dim=256,
          warmup_updates=10, tot_updates=100, peak_lr=2e-4, end_lr=1e-9, edge_type="multi_hop",
          multi_hop_max_dist=20, attention_dropout_rate=0.0)
B, CAP, SEED = 5, 12, 11
# dataset -> (users the reference hard-codes, distance pickle, collator)   model_fqandtoyo.py:721/852/981, collator.py:429/579/722
DATASETS = {"toyotagraph": (995, "toyota_distance.pkl", "collator_toyota"),
            "gowalla_nevda": (1080, "gowalla_distance.pkl", "collator_gowalla"),
            "foursquaregraph": (1080, "tky_distance.pkl", "collator_foursquare")}
# parameters whose full gradient is stored (the outputs of K2 / K4 backward and one encoder weight); all others: the norm
GRAD_FULL = ("rel_pos_encoder.weight", "edge_encoder.weight", "edge_dis_encoder.weight", "graph_token_virtual_distance.weight",
             "poi_pos_encoder.weight", "graph_token.weight", "layers.0.self_attention.linear_q.weight", "final_ln.weight")
PAD0 = ("edge_encoder.weight", "rel_pos_encoder.weight", "in_degree_encoder.weight", "out_degree_encoder.weight",
        "fre_embed_model.weight", "poi_pos_encoder.weight")          # nn.Embedding(padding_idx=0) tables: row 0 stays zero


def golden_weights(name, shape):
    """Deterministic value of parameter `name` (shared by the generator and the test): N(0, 0.1), LayerNorm weights around 1."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    w = (rng.standard_normal(shape) * 0.1).astype(np.float32)
    if name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "final_ln.weight":
        w = w + 1.0
    if name in PAD0:
        w[0] = 0.0
    return torch.from_numpy(w)


def fill(model):
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(golden_weights(name, tuple(p.shape)))


# golden cases: name -> (dataset branch, graphs, node cap, item seed).  `toyotagraph_n40`: longer trajectories (up to 40 nodes:
# walks longer than multi_hop_max_dist, the clamp of model_fqandtoyo.py:1168-1176, more unreachable pairs)
CASES = {"toyotagraph": ("toyotagraph", B, CAP, SEED), "gowalla_nevda": ("gowalla_nevda", B, CAP, SEED),
         "foursquaregraph": ("foursquaregraph", B, CAP, SEED), "toyotagraph_n40": ("toyotagraph", 4, 40, 12),
         # BASELINE shapes: graphs AT the node cap of configs[1] (128 nodes -> T = 129 = 128 + 1: the single-token tail of K3)
         # and of configs[3] (256 nodes -> T = 257: two key blocks, the fold path), each next to a short graph in the same batch
         "toyotagraph_n128": ("toyotagraph", [(128, 1, 13), (9, 1, 14), (128, 1, 15)], 128, None),
         "gowalla_nevda_n256": ("gowalla_nevda", [(256, 1, 16), (31, 1, 17)], 256, None)}
MID_WORLD = dict(P=400, C=11, U=13, cap=256, batch=4)          # a world with enough POIs for 256 distinct nodes per graph


def make_world_and_items(case="toyotagraph"):
    from mobgt_b200 import synth
    dataset_name, nb, cap, seed = CASES[case]
    if isinstance(nb, list):          # [(exact node count, graphs, item seed), ...]
        world = synth.make_world(dict(MID_WORLD, dataset_name=dataset_name), seed=1, U=DATASETS[dataset_name][0])
        items = []
        for n_fixed, cnt, sd in nb:
            items += synth.make_items(world, cnt, cap, seed=sd, n_fixed=n_fixed, start=len(items))
        return world, items
    world = synth.make_world("tiny", seed=1, U=DATASETS[dataset_name][0], dataset_name=dataset_name)   # hard-coded user counts
    items = synth.make_items(world, nb, cap, seed=seed)
    return world, items


def _dense_raw(csr, n):
    """binary raw adjacency (no self loops) whose hat_rw_normd_lap_mat is the world's CSR"""
    crow, col, _ = csr
    a = np.zeros((n, n), np.int64)
    rows = np.repeat(np.arange(n), np.diff(crow))
    a[rows, col] = 1
    a[np.arange(n), np.arange(n)] = 0
    return a


def write_dataset(world, root, dataset_name):
    raw = os.path.join(root, "dataset", dataset_name, "raw")
    os.makedirs(raw)
    os.makedirs(os.path.join(root, "dataset", "poi_data"), exist_ok=True)
    os.makedirs(os.path.join(root, "graphormer"), exist_ok=True)
    P, C = world.P, world.C

    def csv(name, a):
        with open(os.path.join(raw, name), "w") as f:
            f.write(",".join(str(i) for i in range(a.shape[1])) + "\n")      # pd.read_csv consumes one header line
            for r in a:
                f.write(",".join(str(int(v)) for v in r) + "\n")

    csv("Graph_cat.csv", _dense_raw(world.C_A, C))
    csv("Graph_dist.csv", _dense_raw(world.D_A, P))
    csv("Graph_adj.csv", _dense_raw(world.D_A, P))
    with open(os.path.join(raw, "Graph_poi.csv"), "w") as f:      # columns read positionally (:930-941) and by name (:1106-1117)
        f.write("POI ID,check_freq,latitude,longitude,cat\n")
        for i in range(P):
            f.write(f"{i + 1},{int(world.check_freq[i])},{float(world.latlon[i, 0])!r},{float(world.latlon[i, 1])!r},"
                    f"{int(world.cat_of_poi[i])}\n")
    d = np.zeros((P + 1, P + 1), np.float64)
    ll = world.latlon.astype(np.float64)
    d[1:, 1:] = np.sqrt(((ll[:, None, :] - ll[None, :, :]) ** 2).sum(-1))
    pickle.dump(d, open(os.path.join(root, "dataset", "poi_data", DATASETS[dataset_name][1]), "wb"))


def install_stubs():
    import torch.nn as nn

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        @property
        def device(self):
            return torch.device("cpu")

        def log(self, *a, **k):
            pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    dummy = type("Dummy", (), {})
    mod("pytorch_lightning", LightningModule=LightningModule, LightningDataModule=dummy)
    tg = mod("torch_geometric")
    tg.nn = mod("torch_geometric.nn", GCNConv=dummy, GATConv=dummy, MessagePassing=dummy)
    tg.utils = mod("torch_geometric.utils", to_undirected=None, add_self_loops=None, degree=None)
    tg.datasets = mod("torch_geometric.datasets")
    tg.data = mod("torch_geometric.data", Data=dummy, InMemoryDataset=dummy, Dataset=dummy)
    ogb = mod("ogb")
    ogb.graphproppred = mod("ogb.graphproppred", PygGraphPropPredDataset=dummy, Evaluator=dummy)
    ogb.lsc = mod("ogb.lsc")
    mod("ogb.lsc.pcqm4m_pyg", PygPCQM4MDataset=dummy)
    mod("ogb.utils")
    mod("owndata", Foursquare=dummy, FoursquareGraph=dummy, ToyotaGraph=dummy)
    # data.py:165-168 for toyotagraph: NLLLoss(ignore_index=0); evaluator / metric are only read by the Lightning hooks
    mod("data", get_dataset=lambda *a, **k: dict(evaluator=None, metric="NLLLoss", loss_fn=torch.nn.NLLLoss(ignore_index=0),
                                                 num_class=1))
    mod("wandb")
    import build_ref
    algos = build_ref.load()
    assert algos is not None, "oracle/_ref (compiled algos.pyx) must be built first: python oracle/build_ref.py"
    sys.modules["algos"] = algos          # wrapper.py:12-15 would rebuild algos.pyx through pyximport


class RawItem:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def to_ref_item(it):
    """a synthetic raw item with torch fields, as owndata.py:340-349 stores them in a PyG Data object"""
    t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt)
    return RawItem(idx=int(it.idx), x=t(it.x, torch.long), edge_index=t(it.edge_index, torch.long), edge_attr=t(it.edge_attr, torch.long),
                   y=t(it.y, torch.long), time=t(it.time, torch.long), time_normal=t(it.time_normal, torch.float32),
                   user=t(it.user, torch.long), cat=t(it.cat, torch.long))


def run_dataset(case, tmp, ref_model, ref_collator, ref_wrapper, mo):
    import copy
    dataset_name = CASES[case][0]
    world, items = make_world_and_items(case)
    import shutil
    shutil.rmtree(os.path.join(tmp, "dataset", dataset_name), ignore_errors=True)     # the world differs between cases
    write_dataset(world, tmp, dataset_name)
    torch.manual_seed(0)
    rm = ref_model.Graphormer(dataset_name=dataset_name, **HP).eval()
    # the shipped pickle is missing; the table only has to cover the bins the stand-in produces
    rm.poi_pos_encoder = torch.nn.Embedding(world.num_bins, HP["num_heads"], padding_idx=0)
    fill(rm)

    # ---- reference preprocessing + collation (wrapper.py:25-102, collator.py:310-748)
    ref_items = [ref_wrapper.preprocess_item(to_ref_item(it)) for it in items]
    rb = getattr(ref_collator, DATASETS[dataset_name][2])(ref_items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024)
    ob = mo.collate([mo.preprocess_item(it, hop_cap=20) for it in items], world, multi_hop_max_dist=20, rel_pos_max=1024)
    rb.poi_pos = ob.poi_pos.clone()       # stand-in binning (see the module docstring)

    # ---- reference forward (model_fqandtoyo.py:1123-1432), eval mode: [poi logits (log_softmax for toyotagraph), cat logits]
    with torch.no_grad():
        out = rm(copy.deepcopy(rb))
    poi, cat = out[0].detach(), out[1].detach()
    cat_target = rm.cat_target.clone().view(-1).long()
    # the evaluation targets of the reference's own steps (model_fqandtoyo.py:1484-1496, 1530-1544)
    with torch.no_grad():
        y_true_test = rm.test_step(copy.deepcopy(rb), 0)["y_true"].view(-1).long()
        y_true_val = rm.validation_step(copy.deepcopy(rb), 0)["y_true"].view(-1).long()

    def gtl(inputs, targets, alpha):      # GradientTailLoss :545-550 with its `.to("cuda")` dropped
        one_hot = torch.zeros_like(inputs)
        one_hot.scatter_(1, targets[:len(inputs)].view(-1, 1), 1)
        prob = torch.sigmoid(inputs)
        return (-alpha * (1 - prob) * one_hot * torch.log(prob) - (1 - one_hot) * prob * torch.log(1 - prob)).mean()

    if dataset_name == "toyotagraph":     # :1464-1471
        loss = gtl(cat, cat_target, 0.1) + torch.nn.NLLLoss(ignore_index=0)(poi, rb.y)
    else:                                 # :1446-1460
        loss = gtl(poi, rb.y - 1, 0.2)
    # ---- reference backward: autograd through the reference's own forward graph (the loss formula above, differentiable)
    rm.zero_grad()
    out_g = rm(copy.deepcopy(rb))
    ct = rm.cat_target.clone().view(-1).long()
    loss_g = (gtl(out_g[1], ct, 0.1) + torch.nn.NLLLoss(ignore_index=0)(out_g[0], rb.y)) if dataset_name == "toyotagraph" \
        else gtl(out_g[0], rb.y - 1, 0.2)
    loss_g.backward()
    gnames = sorted(n for n, p_ in rm.named_parameters() if p_.grad is not None and float(p_.grad.abs().sum()) > 0)
    gnorm = np.array([float(dict(rm.named_parameters())[n].grad.double().norm()) for n in gnames], np.float64)
    full = {n: dict(rm.named_parameters())[n].grad.numpy().copy() for n in GRAD_FULL}
    fields = dict(x=rb.x, rel_pos=rb.rel_pos, edge_input=rb.edge_input, attn_bias=rb.attn_bias, in_degree=rb.in_degree,
                  out_degree=rb.out_degree, y=rb.y, user=rb.user, time_normal=rb.time_normal)
    path = os.path.join(HERE, f"model_golden_{case}.npz")
    np.savez_compressed(path, poi_logits=poi.numpy(), cat_logits=cat.numpy(), cat_target=cat_target.numpy(),
                        loss=np.array([float(loss)], np.float64), grad_names=np.array(gnames), grad_norms=gnorm,
                        y_true_test=y_true_test.numpy(), y_true_val=y_true_val.numpy(),
                        state_shapes_json=np.array(__import__("json").dumps({k: list(v.shape) for k, v in rm.state_dict().items()},
                                                                            sort_keys=True)),
                        **{"g_" + k: v for k, v in full.items()}, **{"f_" + k: v.numpy() for k, v in fields.items()})
    print("wrote", path, "poi", tuple(poi.shape), "cat", tuple(cat.shape), "loss", float(loss))


def main():
    tmp = tempfile.mkdtemp(prefix="mobgt_ref_")
    os.makedirs(os.path.join(tmp, "graphormer"))
    install_stubs()
    sys.path.insert(0, REF)
    os.chdir(os.path.join(tmp, "graphormer"))
    import model_fqandtoyo as ref_model
    import collator as ref_collator
    import wrapper as ref_wrapper
    import model_oracle as mo
    for case in CASES:
        run_dataset(case, tmp, ref_model, ref_collator, ref_wrapper, mo)
    # ---- metrics (model_fqandtoyo.py:48-90 get_acc, :122-131 MRR_metric) on seeded logits; two cases: no zero target, and a
    #      zero target in the middle of the batch (the reference's loop BREAKS there, :88-89)
    g = torch.Generator().manual_seed(3)
    scores = torch.randn(64, 98, generator=g)
    out = {}
    for tag, zero_at in (("a", None), ("b", 40)):
        y = torch.randint(1, 98, (64,), generator=g)
        y[::7] = scores.argmax(1)[::7].clamp(min=1)          # some top-1 hits
        y[3::9] = scores.topk(8, 1)[1][3::9, 6].clamp(min=1)   # some top-10 hits
        if zero_at is not None:
            y[zero_at] = 0
        acc, ndcg = ref_model.get_acc(y, scores)
        out.update({f"y_{tag}": y.numpy(), f"acc_{tag}": acc, f"ndcg_{tag}": ndcg,
                    f"mrr_{tag}": np.array([ref_model.MRR_metric(y, scores)], np.float64)})
    # ---- learning-rate schedule: lr.py:18-32 `get_lr`, the reference's own method, evaluated for _step_count = 1..40 on a
    #      plain attribute holder (the class itself cannot be constructed under torch 2.11: it passes `verbose` to
    #      _LRScheduler.__init__, lr.py:15, which no longer exists)
    import lr as ref_lr
    holder = types.SimpleNamespace(warmup_updates=10, tot_updates=25, lr=2e-4, end_lr=1e-9, power=1.0,
                                   optimizer=types.SimpleNamespace(param_groups=[{}]), _step_count=0)
    lrs = []
    for c in range(1, 41):
        holder._step_count = c
        lrs.append(ref_lr.PolynomialDecayLR.get_lr(holder)[0])
    out["lr_by_step_count"] = np.array(lrs, np.float64)
    # ---- the model's command-line flags (model_fqandtoyo.py:1618-1641): name -> (default, type)
    import argparse
    import json
    ap = ref_model.Graphormer.add_model_specific_args(argparse.ArgumentParser())
    flags = {a.dest: [a.default, type(a.default).__name__, a.type.__name__ if a.type else None]
             for a in ap._actions if a.dest != "help"}
    out["model_flags_json"] = np.array(json.dumps(flags, sort_keys=True))
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), scores=scores.numpy(), **out)
    print("wrote metrics_golden.npz", {k: np.asarray(v).reshape(-1)[:4] for k, v in out.items() if not k.startswith("y_")})


if __name__ == "__main__":
    main()
