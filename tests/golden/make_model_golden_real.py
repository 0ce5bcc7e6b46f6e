"""Golden vectors of the MODEL half on the REAL Gowalla-Nevada data set (build container only; ~2 minutes).

The UNMODIFIED reference — `wrapper.preprocess_item`, `collator.collator_gowalla`, `model_fqandtoyo.Graphormer`
(dataset_name="gowalla_nevda") — is run on the files of the archive the reference ships: its constructor reads the real
`../dataset/gowalla_nevda/raw/Graph_{cat,dist,adj,poi}.csv` (3 679 POIs, 253 categories; `calculate_laplacian_matrix` on the
dense matrices, dense `torch.mm` GCN tables), the batch is the first real train trajectories of <= 64 nodes in the reference's
queue order (`owndata.GowallaGraph.process`, pinned by make_gowalla_real.py).  Forward, GradientTailLoss and its gradients are
frozen in tests/golden/model_golden_gowalla_real.npz; tests/test_oracle_model_golden.py::test_oracle_equals_reference_on_real_data
checks the oracle (on the world of mobgt_b200.owndata, i.e. the CSR adjacency) against them, and tests/test_real_data_gpu.py
compares the product with that oracle on real batches — so the chain reference -> oracle -> product is closed on real data.

As in make_model_golden.py, only `poi_pos` is the documented stand-in (the reference's distance pickle is not shipped; a
lat / lon distance matrix is written in its place so that the reference collator runs, and the bins of PoiWorld.poi_pos_bins
replace the collator's).
"""
import copy
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
N_ITEMS, MAX_NODES = 10, 64


def real_items(splits):
    """the batch of the golden: the first train trajectories of <= MAX_NODES nodes, a 40+-node one among them"""
    items = [it for it in splits["train"][:200] if len(it.x) <= MAX_NODES]
    head = items[:N_ITEMS - 1]
    big = next(it for it in items if len(it.x) >= 40 and all(it is not h for h in head))
    return head + [big]


def main():
    import _gowalla
    import make_model_golden as mg
    from mobgt_b200 import owndata
    tmp = tempfile.mkdtemp(prefix="mobgt_ref_real_")
    raw = os.path.join(tmp, "dataset", "gowalla_nevda", "raw")
    os.makedirs(raw)
    os.makedirs(os.path.join(tmp, "dataset", "poi_data"))
    os.makedirs(os.path.join(tmp, "graphormer"))
    for name, blob in _gowalla.unpack().items():
        if name.endswith((".csv", ".pickle", ".pkl")):
            with open(os.path.join(raw, name), "wb") as f:
                f.write(blob)
    world = owndata.load_world(raw, "gowalla_nevda")
    splits = {"train": owndata.load_items(raw, "train")}
    items = real_items(splits)
    ll = world.latlon.astype(np.float64)
    d = np.zeros((world.P + 1, world.P + 1), np.float64)
    d[1:, 1:] = np.sqrt(((ll[:, None, :] - ll[None, :, :]) ** 2).sum(-1))
    pickle.dump(d, open(os.path.join(tmp, "dataset", "poi_data", "gowalla_distance.pkl"), "wb"))
    del d
    mg.install_stubs()
    sys.path.insert(0, mg.REF)
    os.chdir(os.path.join(tmp, "graphormer"))
    import model_fqandtoyo as ref_model
    import collator as ref_collator
    import wrapper as ref_wrapper
    import model_oracle as mo
    torch.manual_seed(0)
    rm = ref_model.Graphormer(dataset_name="gowalla_nevda", **mg.HP).eval()      # reads the REAL Graph_*.csv
    rm.poi_pos_encoder = torch.nn.Embedding(world.num_bins, mg.HP["num_heads"], padding_idx=0)
    mg.fill(rm)
    ref_items = [ref_wrapper.preprocess_item(mg.to_ref_item(it)) for it in items]
    rb = ref_collator.collator_gowalla(ref_items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024)
    ob = mo.collate([mo.preprocess_item(it, hop_cap=20) for it in items], world, multi_hop_max_dist=20, rel_pos_max=1024)
    rb.poi_pos = ob.poi_pos.clone()
    with torch.no_grad():
        out = rm(copy.deepcopy(rb))
    poi, cat = out[0].detach(), out[1].detach()

    def gtl(inputs, targets, alpha):      # GradientTailLoss :545-550 with its `.to("cuda")` dropped
        one_hot = torch.zeros_like(inputs)
        one_hot.scatter_(1, targets[:len(inputs)].view(-1, 1), 1)
        prob = torch.sigmoid(inputs)
        return (-alpha * (1 - prob) * one_hot * torch.log(prob) - (1 - one_hot) * prob * torch.log(1 - prob)).mean()

    rm.zero_grad()
    out_g = rm(copy.deepcopy(rb))
    loss = gtl(out_g[0], rb.y - 1, 0.2)                                          # :1446-1460
    loss.backward()
    params = dict(rm.named_parameters())
    gnames = sorted(n for n, p_ in params.items() if p_.grad is not None and float(p_.grad.abs().sum()) > 0)
    gnorm = np.array([float(params[n].grad.double().norm()) for n in gnames], np.float64)
    full = {n: params[n].grad.numpy().copy() for n in mg.GRAD_FULL}
    # the reference's dense \hat A of the REAL distance graph, as its constructor built it: a digest and the row sums
    import hashlib
    da = rm.D_A.numpy()
    fields = dict(x=rb.x, rel_pos=rb.rel_pos, edge_input=rb.edge_input, attn_bias=rb.attn_bias, in_degree=rb.in_degree,
                  out_degree=rb.out_degree, y=rb.y, user=rb.user, time_normal=rb.time_normal)
    path = os.path.join(HERE, "model_golden_gowalla_real.npz")
    np.savez_compressed(path, poi_logits=poi.numpy(), cat_logits=cat.numpy(), loss=np.array([float(loss)], np.float64),
                        grad_names=np.array(gnames), grad_norms=gnorm, item_idx=np.array([it.idx for it in items], np.int64),
                        ref_D_A_sha256=np.array(hashlib.sha256(np.ascontiguousarray(da, np.float32).tobytes()).hexdigest()),
                        **{"g_" + k: v for k, v in full.items()}, **{"f_" + k: v.numpy() for k, v in fields.items()})
    print("wrote", path, "items", [len(it.x) for it in items], "poi", tuple(poi.shape), "loss", float(loss), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
