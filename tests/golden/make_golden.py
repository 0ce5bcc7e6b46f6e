"""Generate tests/golden/algos_golden.npz from the REAL reference (oracle/_ref).

Run once in the build container (needs /root/reference):
    python tests/golden/make_golden.py
Every output below is produced by the unmodified, Cython-compiled
/root/reference/graphormer/algos.pyx driven exactly like wrapper.py:42-60:
    adj[src,dst] = True ; ef[src,dst,0] = count + 2
    M, path = algos.floyd_warshall(adj)
    e = algos.gen_edge_input(int(M.max()), path, ef)
and then reduced to:  M (int16), path (int16), e[:, :, :20, 0] padded with -1
to 20 hops (int8).  Small graphs keep the full arrays; every graph keeps a
16-byte sha256 digest of M||path||e20.

Content:
  * hand KATs (SURVEY.md §8c A-E + hub-0, self-loop, short-hop-axis cases)
  * random digraphs (several densities)
  * ALL real Gowalla-Nevada train trajectories with n <= 256 (edge lists only)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import build_ref  # noqa: E402

HOPS = 20


def run_ref(algos, n, src, dst, cnt):
    adj = np.zeros((n, n), bool)
    ef = np.zeros((n, n, 1), np.int64)
    adj[src, dst] = True
    ef[src, dst, 0] = cnt + 2
    M, path = algos.floyd_warshall(adj)
    md = int(M.max()) if n > 0 else 0
    e = algos.gen_edge_input(md, path, ef)
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


def main():
    build_ref.build()
    algos = build_ref.load()
    graphs = []   # (name, n, src, dst, cnt)

    def add(name, n, edges):
        e = np.array(edges, np.int64).reshape(-1, 3)
        graphs.append((name, n, e[:, 0], e[:, 1], e[:, 2]))

    # ---- hand KATs (counts chosen so that feat = count+2 matches the survey's values)
    add("kat_a", 5, [(0, 1, 1), (1, 2, 1), (2, 3, 1), (3, 1, 1), (2, 2, 1), (3, 4, 1)])
    add("kat_b_node0", 3, [(1, 0, 5), (0, 2, 7)])
    add("kat_c_cycle", 3, [(0, 1, 3), (1, 2, 3), (2, 0, 3)])
    add("kat_d_chain512", 512, [(i, i + 1, 1) for i in range(511)])
    add("kat_e_single", 1, [])
    add("single_selfloop", 1, [(0, 0, 4)])
    add("two_isolated", 2, [])
    add("hub0_star", 9, [(i, 0, i) for i in range(1, 9)] + [(0, i, 9 - i) for i in range(1, 9)])
    add("selfloops_chain", 6, [(i, i, 2) for i in range(6)] + [(i, i + 1, 1) for i in range(5)])
    add("bidir_ring40", 40, [(i, (i + 1) % 40, 1) for i in range(40)] + [((i + 1) % 40, i, 2) for i in range(40)])
    add("complete7", 7, [(i, j, 1 + (i * 7 + j) % 5) for i in range(7) for j in range(7) if i != j])
    add("chain25_long_hops", 25, [(i, i + 1, 1 + i % 3) for i in range(24)])
    add("rev_chain30", 30, [(i + 1, i, 1) for i in range(29)])
    rng = np.random.default_rng(20261017)
    for t in range(160):
        n = int(rng.integers(2, 70))
        dens = float(rng.choice([0.02, 0.05, 0.1, 0.25, 0.6]))
        a = rng.random((n, n)) < dens
        s, d = np.nonzero(a)
        c = rng.integers(1, 48, size=len(s))
        graphs.append((f"rand_{t}_n{n}_d{dens}", n, s, d, c))
    for t, n in enumerate([100, 129, 160, 200, 256, 300]):
        a = rng.random((n, n)) < (2.5 / n)
        s, d = np.nonzero(a)
        c = rng.integers(1, 20, size=len(s))
        graphs.append((f"sparse_{t}_n{n}", n, s, d, c))

    # ---- real Gowalla-Nevada trajectories
    import _gowalla
    parts = _gowalla.unpack()
    tr = _gowalla.sessions(parts["train.pickle"])
    for gi, s in enumerate(tr):
        n = int(s["num_node"])
        if n > 256:
            continue
        et = s["edge_type"].numpy()
        src, dst = np.nonzero(et)
        graphs.append((f"gowalla_train_{gi}", n, src, dst, et[src, dst]))

    names, ns, eoff, esrc, edst, ecnt, digs, mds = [], [], [0], [], [], [], [], []
    full = {}
    for name, n, s, d, c in graphs:
        M, path, e20, md = run_ref(algos, n, s, d, c)
        names.append(name)
        ns.append(n)
        esrc.append(np.asarray(s, np.int16))
        edst.append(np.asarray(d, np.int16))
        ecnt.append(np.asarray(c, np.int16))
        eoff.append(eoff[-1] + len(s))
        digs.append(digest(M, path, e20))
        mds.append(md)
        if n <= 12 and not name.startswith("gowalla") or name in ("kat_a", "kat_b_node0", "kat_c_cycle"):
            full[f"full_{name}_M"] = M
            full[f"full_{name}_path"] = path
            full[f"full_{name}_e20"] = e20
    # spot values of the 512-chain (SURVEY.md §8c KAT-D)
    M, path, e20, md = run_ref(algos, *graphs[3][1:])
    assert M[0, 509] == 509 and M[0, 510] == 510 and M[1, 511] == 510 and path[0, 509] == 508 and path[0, 510] == 510
    out = os.path.join(HERE, "algos_golden.npz")
    np.savez_compressed(
        out, names=np.array(names), n=np.array(ns, np.int32), eoff=np.array(eoff, np.int64),
        esrc=np.concatenate(esrc), edst=np.concatenate(edst), ecnt=np.concatenate(ecnt),
        digest=np.stack(digs), max_dist=np.array(mds, np.int32), **full)
    print("wrote", out, os.path.getsize(out), "bytes;", len(names), "graphs;", len(full) // 3, "with full arrays")


if __name__ == "__main__":
    main()
