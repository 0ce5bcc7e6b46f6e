"""Shared test helpers: golden-fixture access and the reference-style driving of algos (wrapper.py:42-60)."""
import hashlib

import numpy as np

HOPS = 20


def golden_graph(gd, idx):
    """-> (name, n, src, dst, cnt) of fixture graph idx."""
    a, b = int(gd["eoff"][idx]), int(gd["eoff"][idx + 1])
    return (str(gd["names"][idx]), int(gd["n"][idx]), gd["esrc"][a:b].astype(np.int64),
            gd["edst"][a:b].astype(np.int64), gd["ecnt"][a:b].astype(np.int64))


def dense_inputs(n, src, dst, cnt):
    """adjacency (bool [n,n]) and attn_edge_type (int64 [n,n,1], count+2 on edges) as wrapper.py:42-53."""
    adj = np.zeros((n, n), bool)
    ef = np.zeros((n, n, 1), np.int64)
    adj[src, dst] = True
    ef[src, dst, 0] = cnt + 2
    return adj, ef


def run_algos(algos, n, src, dst, cnt, hop_cap=None):
    """Drive an `algos`-shaped module like wrapper.py:55-60 and reduce to (M16, path16, e20 int8, max_dist)."""
    adj, ef = dense_inputs(n, src, dst, cnt)
    M, path = algos.floyd_warshall(adj)
    md = int(M.max()) if n > 0 else 0
    if hop_cap is None:
        e = algos.gen_edge_input(md, path, ef)
    else:
        e = algos.gen_edge_input(md, path, ef, hop_cap=hop_cap)
    e20 = np.full((n, n, HOPS), -1, np.int8)
    h = min(HOPS, e.shape[2])
    e20[:, :, :h] = e[:, :, :h, 0].astype(np.int8)
    return M.astype(np.int16), path.astype(np.int16), e20, md


def digest(M, path, e20):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(M).tobytes())
    h.update(np.ascontiguousarray(path).tobytes())
    h.update(np.ascontiguousarray(e20).tobytes())
    return np.frombuffer(h.digest()[:16], np.uint8)
