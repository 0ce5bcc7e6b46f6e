"""The reference's ON-DISK dataset format -> the items and tables this package trains on.

A MobGT dataset directory (`../dataset/<name>/raw/`, e.g. the `gowalla_nevda.7z` the reference ships) holds

    train.pickle / test.pickle      {user: {session: {'node_name', 'edge_type', 'target', 'time', 'time_normal', 'user', 'cat', ...}}}
    train_idx.pkl / test_idx.pkl    {user: [session, ...]}          (which sessions of a user are in the split)
    Graph_poi.csv                   POI ID, check_freq, lat, lon, cat, cat_freq
    Graph_dist.csv / Graph_cat.csv  the global POI-distance and category graphs (dense adjacency, one header row)

`load_items` is `owndata.GowallaGraph.process` / `FoursquareGraph.process` (owndata.py:288-373, 375-460: the same body) without
torch_geometric: the per-user queue of `generate_queue` (owndata.py:60-85, `np.random.seed(1)`), then one raw item per
(user, session) with the fields `wrapper.preprocess_item` reads.  `load_world` is the table part of `Graphormer.__init__`
(model_fqandtoyo.py:647-700 Gowalla, :791-832 Foursquare, same statements): the row-normalised \\hat A = (D + I)^-1 (A + I) of
`calculate_laplacian_matrix(..., 'hat_rw_normd_lap_mat')` (:458-488) — kept as CSR, never as a dense P x P matrix — the POI
feature matrix X = [check_freq | one-hot category | lat | lon] and the one-hot category features C_X.

Not in the archive, hence not read: the P x P distance pickle the reference's collator bins into `poi_pos`
(collator.py:428-437, `poi_data/…_distance.pkl`); `PoiWorld.poi_pos_bins` / `mobgt_poi_pos` bin the lat / lon distance instead
(the stand-in every round has documented).  pandas is only needed here.
"""
import os
import pickle
from collections import deque

import numpy as np

from .synth import Data, PoiWorld

# `self.num_users` of the reference's per-dataset constructor branches (model_fqandtoyo.py:721-723, 852, 981)
NUM_USERS = {"gowalla_nevda": 1080, "gowalla_7day": 937, "foursquaregraph": 1080, "toyotagraph": 995}


def hat_rw_normd_csr(adj):
    """calculate_laplacian_matrix(adj, 'hat_rw_normd_lap_mat') (model_fqandtoyo.py:458-488) as CSR (crow, col, val f32):
    (D + I)^-1 (A + I) with D = diag(row sums of A).  Same float64 arithmetic as the reference's `inv(D + I) @ (A + I)` — a
    reciprocal per row, one product per entry — then its `.to(torch.float)`."""
    a = np.asarray(adj, np.float64)
    n = a.shape[0]
    assert a.shape == (n, n)
    inv = 1.0 / (a.sum(1) + 1.0)
    w = a + np.eye(n)
    rows, cols = np.nonzero(w)
    val = (inv[rows] * w[rows, cols]).astype(np.float32)
    crow = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=crow[1:])
    return crow, cols.astype(np.int64), val


def load_world(raw_dir, dataset_name, num_bins=64):
    """The dataset tables of model_fqandtoyo.py:647-700 / :791-832 from `raw_dir` -> PoiWorld."""
    import pandas as pd
    raw_c = pd.read_csv(os.path.join(raw_dir, "Graph_cat.csv")).to_numpy()
    raw_d = pd.read_csv(os.path.join(raw_dir, "Graph_dist.csv")).to_numpy()
    raw_x = pd.read_csv(os.path.join(raw_dir, "Graph_poi.csv")).to_numpy()
    P = raw_x.shape[0]
    cat = raw_x[:, 4].astype(np.int64)
    cats = np.unique(cat)                                  # OneHotEncoder: sorted distinct categories
    C = len(cats)
    if not np.array_equal(cats, np.arange(1, C + 1)):
        raise ValueError("Graph_poi.csv: the category ids must be 1..C (the reference indexes category tables with cat - 1)")
    if raw_c.shape != (C, C) or raw_d.shape != (P, P):
        raise ValueError(f"graph sizes {raw_c.shape} / {raw_d.shape} do not match {C} categories / {P} POIs")
    X = np.zeros((P, 3 + C), np.float32)
    X[:, 0] = raw_x[:, 1]                                  # check_freq
    X[np.arange(P), cat] = 1.0                             # one-hot category in columns 1..C
    X[:, C + 1] = raw_x[:, 2]
    X[:, C + 2] = raw_x[:, 3]
    latlon = np.ascontiguousarray(raw_x[:, 2:4], np.float32)
    span = latlon.max(0) - latlon.min(0)
    if dataset_name not in NUM_USERS:
        raise ValueError(f"dataset_name={dataset_name!r}: the reference sizes its user table per data set (model_fqandtoyo.py:721-723, "
                         f"852, 981); known: {sorted(NUM_USERS)}")
    return PoiWorld(P=P, C=C, U=NUM_USERS[dataset_name], dataset_name=dataset_name, cat_of_poi=cat, latlon=latlon,
                    check_freq=raw_x[:, 1].astype(np.int64), num_bins=num_bins, dist_max=float(np.sqrt((span ** 2).sum())) + 1e-6,
                    D_A=hat_rw_normd_csr(raw_d), C_A=hat_rw_normd_csr(raw_c), X=X, C_X=np.eye(C, dtype=np.float32))


def generate_queue(split_idx, mode, seed=1):
    """owndata.py:60-85: the (user, session) order of a split.  'normal' (test): users in dict order, sessions in list order.
    'random' (train): rounds of — shuffle the users with the legacy global generator seeded with `seed`, take the next session
    of the first int(0.01 * users) + 1 users that still have one — until every user's queue is empty."""
    users = list(split_idx.keys())
    out = []
    if mode == "normal":
        for u in users:
            out += [(u, s) for s in split_idx[u]]
        return out
    rng = np.random.RandomState(seed)                      # np.random.seed(seed) ; np.random.shuffle(user)
    left = {u: deque(split_idx[u]) for u in users}
    stop = int(0.01 * len(users))
    while any(len(q) for q in left.values()):
        rng.shuffle(users)
        for j, u in enumerate(users):
            if left[u]:
                out.append((u, left[u].popleft()))
            if j >= stop:
                break
    return out


def _np(v, dtype):
    return np.asarray(v.numpy() if hasattr(v, "numpy") else v).astype(dtype)


def item_from_session(mol, idx):
    """One trajectory-graph dict of train.pickle -> the raw item of owndata.py:430-444 (numpy instead of torch tensors)."""
    adj = _np(mol["edge_type"], np.int64)
    src, dst = np.nonzero(adj)                             # adj.nonzero(as_tuple=False).t(): row-major order
    return Data(idx=idx, x=_np(mol["node_name"], np.int64).reshape(-1, 1), edge_index=np.stack([src, dst]).astype(np.int64),
                edge_attr=adj[src, dst], y=_np(mol["target"], np.int64).reshape(-1),
                time=_np(mol["time"], np.int64).reshape(-1, 1), time_normal=_np(mol["time_normal"], np.float32).reshape(-1, 1),
                user=_np(mol["user"], np.int64).reshape(-1, 1), cat=_np(mol["cat"], np.int64).reshape(-1, 1))


def load_items(raw_dir, split, seed=1):
    """`GowallaGraph(root, split=...)` / `FoursquareGraph`: the items of a split in the reference's order."""
    assert split in ("train", "test")
    with open(os.path.join(raw_dir, f"{split}.pickle"), "rb") as f:
        mols = pickle.load(f)
    with open(os.path.join(raw_dir, f"{split}_idx.pkl"), "rb") as f:
        split_idx = pickle.load(f)
    queue = generate_queue(split_idx, "random" if split == "train" else "normal", seed)
    return [item_from_session(mols[u][s], k) for k, (u, s) in enumerate(queue)]


# ---- compact form of a dataset (what tests/golden/gowalla_nevda_real.npz holds: the archive's pickles are 58 MB of torch
# tensors, its graph CSVs 108 MB of text; the same content is ~2 MB as packed arrays)
def pack_dataset(world, splits):
    """world + {split: items} -> dict of numpy arrays (np.savez_compressed-able)."""
    out = {"poi": np.stack([world.check_freq.astype(np.float64), world.latlon[:, 0].astype(np.float64),
                            world.latlon[:, 1].astype(np.float64), world.cat_of_poi.astype(np.float64)], 1),
           "meta": np.array([world.P, world.C, world.U, world.num_bins], np.int64), "dist_max": np.array([world.dist_max]),
           "dataset_name": np.array(world.dataset_name)}
    for name, csr in (("D_A", world.D_A), ("C_A", world.C_A)):
        crow, col, val = csr
        out[name + "_crow"], out[name + "_col"], out[name + "_val"] = crow.astype(np.int64), col.astype(np.int32), val
    for split, items in splits.items():
        n = np.array([len(it.x) for it in items], np.int32)
        e = np.array([it.edge_index.shape[1] for it in items], np.int32)
        cat = lambda f, dt: np.concatenate([np.asarray(getattr(it, f)).reshape(-1) for it in items]).astype(dt)
        out.update({f"{split}_n": n, f"{split}_e": e, f"{split}_x": cat("x", np.int32), f"{split}_time": cat("time", np.int16),
                    f"{split}_tn": cat("time_normal", np.float32), f"{split}_cat": cat("cat", np.int16),
                    f"{split}_y": cat("y", np.int32), f"{split}_user": cat("user", np.int32),
                    f"{split}_src": np.concatenate([it.edge_index[0] for it in items]).astype(np.int16),
                    f"{split}_dst": np.concatenate([it.edge_index[1] for it in items]).astype(np.int16),
                    f"{split}_ea": cat("edge_attr", np.int16)})
    return out


def unpack_dataset(z):
    """Inverse of pack_dataset: npz / dict -> (PoiWorld, {split: items})."""
    P, C, U, num_bins = (int(v) for v in z["meta"])
    poi = z["poi"]
    cat = poi[:, 3].astype(np.int64)
    X = np.zeros((P, 3 + C), np.float32)
    X[:, 0] = poi[:, 0]
    X[np.arange(P), cat] = 1.0
    X[:, C + 1] = poi[:, 1]
    X[:, C + 2] = poi[:, 2]
    csr = lambda n: (z[n + "_crow"].astype(np.int64), z[n + "_col"].astype(np.int64), z[n + "_val"].astype(np.float32))
    world = PoiWorld(P=P, C=C, U=U, dataset_name=str(z["dataset_name"]), cat_of_poi=cat,
                     latlon=np.ascontiguousarray(poi[:, 1:3], np.float32), check_freq=poi[:, 0].astype(np.int64), num_bins=num_bins,
                     dist_max=float(z["dist_max"][0]), D_A=csr("D_A"), C_A=csr("C_A"), X=X, C_X=np.eye(C, dtype=np.float32))
    splits = {}
    for split in ("train", "test"):
        if f"{split}_n" not in z:
            continue
        n, e = z[f"{split}_n"], z[f"{split}_e"]
        no, eo = np.concatenate([[0], np.cumsum(n)]), np.concatenate([[0], np.cumsum(e)])
        items = []
        for k in range(len(n)):
            a, b, c, d = no[k], no[k + 1], eo[k], eo[k + 1]
            items.append(Data(idx=k, x=z[f"{split}_x"][a:b].astype(np.int64).reshape(-1, 1),
                              edge_index=np.stack([z[f"{split}_src"][c:d], z[f"{split}_dst"][c:d]]).astype(np.int64),
                              edge_attr=z[f"{split}_ea"][c:d].astype(np.int64), y=z[f"{split}_y"][k:k + 1].astype(np.int64),
                              time=z[f"{split}_time"][a:b].astype(np.int64).reshape(-1, 1),
                              time_normal=z[f"{split}_tn"][a:b].reshape(-1, 1),
                              user=z[f"{split}_user"][k:k + 1].astype(np.int64).reshape(-1, 1),
                              cat=z[f"{split}_cat"][a:b].astype(np.int64).reshape(-1, 1)))
        splits[split] = items
    return world, splits
