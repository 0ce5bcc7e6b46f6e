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


# ---- numpy restatement of libmobgt's attention-dropout keep mask (mobgt_b200/csrc/common.cuh: attn_drop_*)
def _lowbias32(z):
    z = z.astype(np.uint32)
    z ^= z >> np.uint32(16)
    z = (z * np.uint32(0x7FEB352D)).astype(np.uint32)
    z ^= z >> np.uint32(15)
    z = (z * np.uint32(0x846CA68B)).astype(np.uint32)
    z ^= z >> np.uint32(16)
    return z


def attn_drop_keep(plane, T, p, seed):
    """keep mask bool [T, T] of attention plane `plane` (= graph * H + head) for dropout rate p and the 64-bit call seed."""
    th16 = min(int(p * 65536.0 + 0.5), 65535) if p > 0 else 0
    lo, hi = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        rows = np.arange(T, dtype=np.uint32)
        rowkey = _lowbias32(np.uint32(plane) * np.uint32(1024) + rows + lo) ^ hi                   # [T]
        nch = (T + 7) // 8
        chunk = np.arange(nch, dtype=np.uint32)
        w0 = _lowbias32(rowkey[:, None] + chunk[None, :] * np.uint32(0x9E3779B1))                   # [T, nch]
        words = [w0]
        for c in (0xC2B2AE35, 0x27D4EB2F, 0x165667B1):
            w = (w0 * np.uint32(c)).astype(np.uint32)
            words.append(w ^ (w >> np.uint32(15)))
        fields = np.stack([f for w in words for f in (w & np.uint32(0xFFFF), w >> np.uint32(16))], axis=-1)   # [T, nch, 8]
    keep = (fields >= th16).reshape(T, nch * 8)[:, :T]
    return keep, 65536.0 / (65536 - th16)
