"""Seeded synthetic MobGT-shaped data.

There is no network and the reference's `poi_data` blob is missing, so every workload in
BASELINE.json is generated here: trajectory graphs shaped like gen_pickles.py:735-833 builds them
(nodes = unique POIs ordered by last occurrence, edge_type[i,j] = number of consecutive i->j
transitions, self-loops possible), a POI table (category, lat/lon), and the global POI / category
graphs the GCN tables are computed from (model_fqandtoyo.py:791-832).

Node-count law: the empirical histogram of the real Gowalla-Nevada train split
(data/gowalla_node_hist.json, regenerated from the shipped archive), clipped to a cap; or
`n_fixed` for the stress variant (every graph at the cap).
"""
import json
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# BASELINE.json configs: P POIs, C categories, U users, node cap
CONFIGS = {
    "c1": dict(P=7856, C=300, U=1080, cap=32, batch=16, dataset_name="foursquaregraph"),
    "c2": dict(P=60000, C=300, U=995, cap=128, batch=256, dataset_name="toyotagraph"),
    "c3": dict(P=60000, C=300, U=995, cap=512, batch=100000, dataset_name="toyotagraph"),
    "c4": dict(P=3679, C=253, U=1080, cap=256, batch=256, dataset_name="gowalla_nevda"),
    "tiny": dict(P=97, C=11, U=13, cap=12, batch=4, dataset_name="toyotagraph"),
}


class Data:
    """Minimal stand-in for torch_geometric.data.Data (owndata.py:340-349 fills these fields)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def node_count_sampler(rng, G, cap, n_fixed=None):
    if n_fixed is not None:
        return np.full(G, int(n_fixed), np.int32)
    hist = np.array(json.load(open(os.path.join(HERE, "data", "gowalla_node_hist.json")))["hist"], np.float64)
    hist[0] = 0
    ns = rng.choice(len(hist), size=G, p=hist / hist.sum())
    return np.clip(ns, 1, cap).astype(np.int32)


def zipf_table(P, s=1.1):
    w = 1.0 / np.arange(1, P + 1, dtype=np.float64) ** s
    return np.cumsum(w / w.sum())


def gen_trajectory(rng, n, P, cdf, revisit=0.3, max_count=124, max_deg=126):
    """One trajectory graph with exactly n unique POIs.  Returns (node_name[n] 1-based POI ids,
    edge_type[n,n] transition counts)."""
    while True:
        seq, seen = [], {}
        while len(seen) < n:
            if seq and rng.random() < revisit:
                v = seq[int(rng.integers(len(seq)))]
            else:
                v = int(np.searchsorted(cdf, rng.random())) + 1
                v = min(v, P)
                if v in seen and len(seen) + 1 <= P:
                    # draw a fresh POI so that the node count is reached in bounded time
                    v = int(rng.integers(1, P + 1))
            seen[v] = len(seq)
            seq.append(v)
        for _ in range(int(rng.geometric(0.5)) - 1):   # a few trailing revisits
            v = seq[int(rng.integers(len(seq)))]
            seen[v] = len(seq)
            seq.append(v)
        # nodes ordered by last occurrence (gen_pickles.py:789-791 drop_duplicates(keep='last'))
        order = sorted(seen, key=lambda v: seen[v])
        if len(order) != n:
            continue
        idx = {v: i for i, v in enumerate(order)}
        et = np.zeros((n, n), np.int64)
        for a, b in zip(seq[:-1], seq[1:]):
            et[idx[a], idx[b]] += 1
        adj = et > 0
        if et.max() <= max_count and adj.sum(0).max(initial=0) <= max_deg and adj.sum(1).max(initial=0) <= max_deg:
            return np.array(order, np.int64), et


@dataclass
class PoiWorld:
    """The dataset-level tables a MobGT model is built from (model_fqandtoyo.py:791-900)."""
    P: int
    C: int
    U: int
    dataset_name: str
    cat_of_poi: np.ndarray        # [P] int64, 1..C   (Graph_poi.csv 'cat' column)
    latlon: np.ndarray            # [P,2] float32
    check_freq: np.ndarray        # [P] int64
    num_bins: int = 64
    dist_max: float = 1.0
    # global graphs in CSR (row-normalised \hat A = D^-1 (A + I), calculate_laplacian_matrix 'hat_rw_normd_lap_mat')
    D_A: tuple = field(default=None, repr=False)   # (crow[P+1], col[nnz], val[nnz])
    C_A: tuple = field(default=None, repr=False)
    X: np.ndarray = field(default=None, repr=False)    # [P, 3+C] float32 (check_freq, one-hot cat, lat, lon)
    C_X: np.ndarray = field(default=None, repr=False)  # [C, C] one-hot

    def poi_pos_bins(self, ids_a, ids_b):
        """Distance bin in 1..num_bins-1 between 1-based POI ids (collator.py:428-437 np.digitize stand-in)."""
        a = self.latlon[np.asarray(ids_a) - 1]
        b = self.latlon[np.asarray(ids_b) - 1]
        d = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))
        return 1 + np.minimum((d / self.dist_max * (self.num_bins - 2)).astype(np.int64), self.num_bins - 2)


def _row_norm_csr(P, rows, cols):
    r"""\hat A = (D+I)^-1 (A+I) as CSR from an edge list (no self loops in input)."""
    rows = np.concatenate([rows, np.arange(P)])
    cols = np.concatenate([cols, np.arange(P)])
    key = rows.astype(np.int64) * P + cols
    key = np.unique(key)
    rows, cols = key // P, key % P
    deg = np.bincount(rows, minlength=P).astype(np.float32)
    val = (1.0 / deg[rows]).astype(np.float32)
    crow = np.zeros(P + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=P), out=crow[1:])
    return crow, cols.astype(np.int64), val


def make_world(cfg="c2", seed=1, knn=16, **over):
    spec = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    spec.update(over)
    P, C, U = spec["P"], spec["C"], spec["U"]
    rng = np.random.default_rng(np.random.SeedSequence([seed, 7, P, C]))
    cat = rng.integers(1, C + 1, size=P).astype(np.int64)
    cat[:C] = np.arange(1, C + 1)          # every category occurs (OneHotEncoder width == C)
    latlon = rng.random((P, 2)).astype(np.float32)
    freq = np.maximum(1, (200.0 / np.arange(1, P + 1) ** 0.6)).astype(np.int64)
    # distance graph: each POI linked to `knn` random near-ish POIs (symmetric), category graph: random 20 %
    k = min(knn, P - 1)
    src = np.repeat(np.arange(P), k)
    dst = (src + rng.integers(1, max(2, min(P, 64 * k)), size=src.shape)) % P
    rows = np.concatenate([src, dst])
    cols = np.concatenate([dst, src])
    keep = rows != cols
    D_A = _row_norm_csr(P, rows[keep], cols[keep])
    cm = rng.random((C, C)) < 0.2
    cm = np.triu(cm, 1)
    r, c = np.nonzero(cm | cm.T)
    C_A = _row_norm_csr(C, r, c)
    X = np.zeros((P, 3 + C), np.float32)
    X[:, 0] = freq
    X[np.arange(P), cat] = 1.0            # columns 1..C one-hot (model_fqandtoyo.py:821-825)
    X[:, C + 1] = latlon[:, 0]
    X[:, C + 2] = latlon[:, 1]
    C_X = np.eye(C, dtype=np.float32)
    return PoiWorld(P=P, C=C, U=U, dataset_name=spec["dataset_name"], cat_of_poi=cat, latlon=latlon,
                    check_freq=freq, num_bins=64, dist_max=float(np.sqrt(2.0)) + 1e-6, D_A=D_A, C_A=C_A, X=X, C_X=C_X)


def make_items(world, G, cap, seed=1, cfg_id=2, n_fixed=None, start=0):
    """G raw dataset items (what owndata.py:340-349 stores), seeded per graph as
    SeedSequence([seed, cfg_id, g]) so any subset is reproducible."""
    rng0 = np.random.default_rng(np.random.SeedSequence([seed, cfg_id, 999983]))
    ns = node_count_sampler(rng0, start + G, cap, n_fixed)[start:]
    cdf = zipf_table(world.P)
    items = []
    for g in range(G):
        rng = np.random.default_rng(np.random.SeedSequence([seed, cfg_id, start + g]))
        n = int(min(ns[g], world.P))
        names, et = gen_trajectory(rng, n, world.P, cdf)
        src, dst = np.nonzero(et)
        slot = rng.integers(0, 48, size=n).astype(np.int64)
        y = int(min(world.P, np.searchsorted(cdf, rng.random()) + 1))
        items.append(Data(
            idx=start + g,
            x=names.reshape(-1, 1),
            edge_index=np.stack([src, dst]).astype(np.int64),
            edge_attr=et[src, dst].astype(np.int64),
            y=np.array([y], np.int64),
            time=slot.reshape(-1, 1),
            time_normal=(slot / 48.0).astype(np.float32).reshape(-1, 1),
            user=np.array([[int(rng.integers(0, world.U))]], np.int64),
            cat=world.cat_of_poi[names - 1].reshape(-1, 1),
        ))
    return items
