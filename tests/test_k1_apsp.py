"""GPU: K1 (batched APSP + path-edge extraction, csrc/k1_apsp.cu) is bit-exact against the reference.

Checker = the golden vectors produced by the compiled reference algos.pyx, and the C oracle
(oracle/algos_oracle.c) on fresh seeded inputs.  All calls go through the C-ABI (mobgt_b200._C)."""
import numpy as np
import pytest
import torch

import algos_oracle
from helpers import dense_inputs, digest, golden_graph, run_algos

pytestmark = pytest.mark.gpu
HOPS = 20


def pack(graphs):
    """graphs: list of (n, src, dst, cnt) -> feat u8 packed, n, sq_off (numpy)."""
    from mobgt_b200.algos import pack_graphs
    n, sq, no = pack_graphs([g[0] for g in graphs])
    feat = np.zeros(int(sq[-1]), np.uint8)
    for gi, (nn, s, d, c) in enumerate(graphs):
        blk = feat[sq[gi]:sq[gi + 1]].reshape(nn, nn)
        blk[s, d] = c + 2
    return feat, n, sq, no


def run_gpu(graphs, shift=0, want_path=True, hops=HOPS):
    from mobgt_b200.algos import apsp_edge_input_packed
    feat, n, sq, no = pack(graphs)
    out = apsp_edge_input_packed(torch.from_numpy(feat).cuda(), torch.from_numpy(n).cuda(),
                                 torch.from_numpy(sq).cuda(), n, hops=hops, shift=shift, want_path=want_path)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}, n, sq


def unpack(res, n, sq, gi, hops=HOPS):
    nn = int(n[gi])
    a, b = int(sq[gi]), int(sq[gi + 1])
    M = res["dist"][a:b].reshape(nn, nn)
    P = res["path"][a:b].reshape(nn, nn) if res["path"] is not None else None
    e = res["edge_in"][a:b].reshape(nn, nn, hops).view(np.int8)
    return M, P, e


def test_golden_all(lib_built, golden):
    """Every fixture graph (KATs + random + 4 959 real Gowalla graphs) in ONE batched launch set."""
    G = len(golden["n"])
    graphs = [golden_graph(golden, i)[1:] for i in range(G)]
    res, n, sq = run_gpu(graphs)
    bad = []
    for gi in range(G):
        M, P, e = unpack(res, n, sq, gi)
        md = int(golden["max_dist"][gi])
        e20 = e.copy()
        if md < HOPS:
            assert (e20[:, :, md:] == -1).all()
        if not (digest(M, P, e20) == golden["digest"][gi]).all() or int(res["maxdist"][gi]) != md:
            bad.append((gi, str(golden["names"][gi])))
    assert not bad, f"{len(bad)} of {G} graphs differ from the reference: {bad[:8]}"


def test_kats_by_value(lib_built, golden):
    names = [str(x) for x in golden["names"]]
    for name in ("kat_a", "kat_b_node0", "kat_c_cycle", "hub0_star", "kat_e_single"):
        gi = names.index(name)
        res, n, sq = run_gpu([golden_graph(golden, gi)[1:]])
        M, P, e = unpack(res, n, sq, 0)
        assert (M == golden[f"full_{name}_M"]).all()
        assert (P == golden[f"full_{name}_path"]).all()
        assert (e == golden[f"full_{name}_e20"]).all()


@pytest.mark.parametrize("seed,nlo,nhi,dens", [(1, 1, 33, 0.2), (2, 30, 130, 0.03), (3, 120, 260, 0.012), (4, 400, 513, 0.004)])
def test_random_vs_oracle(lib_built, seed, nlo, nhi, dens):
    rng = np.random.default_rng(seed)
    graphs = []
    for _ in range(24 if nhi < 300 else 6):
        n = int(rng.integers(nlo, nhi))
        a = rng.random((n, n)) < dens * rng.choice([0.5, 1.0, 3.0])
        s, d = np.nonzero(a)
        graphs.append((n, s, d, rng.integers(1, 100, size=len(s))))
    res, n, sq = run_gpu(graphs)
    for gi, g in enumerate(graphs):
        Mo, Po, eo, md = run_algos(algos_oracle, *g, hop_cap=HOPS)
        M, P, e = unpack(res, n, sq, gi)
        assert (M == Mo).all(), f"dist differs graph {gi} n={g[0]}"
        assert (P == Po).all(), f"path differs graph {gi} n={g[0]}"
        assert (e == eo).all(), f"edge_input differs graph {gi} n={g[0]}"
        assert int(res["maxdist"][gi]) == md


def test_chain_512_sentinel(lib_built):
    """Directed chain of 512 nodes: true distances 510/511 are reported unreachable (SURVEY.md KAT-D)."""
    n = 512
    g = (n, np.arange(n - 1), np.arange(1, n), np.ones(n - 1, np.int64))
    res, nn, sq = run_gpu([g])
    M, P, e = unpack(res, nn, sq, 0)
    assert M[0, 509] == 509 and M[0, 510] == 510 and M[1, 511] == 510
    assert P[0, 509] == 508 and P[0, 510] == 510
    Mo, Po, eo, _ = run_algos(algos_oracle, *g, hop_cap=HOPS)
    assert (M == Mo).all() and (P == Po).all() and (e == eo).all()


@pytest.mark.parametrize("n", [511, 512])
def test_node_510_is_a_real_intermediate(lib_built, n):
    """n > 510: node 510's index collides with the reference's 510 sentinel, so pairs whose LAST improving intermediate
    is node 510 are skipped by gen_edge_input (algos.pyx:87-88) while walks of other pairs still pass through them.
    Node 510 is made a hub so that many shortest paths use it."""
    rng = np.random.default_rng(510 + n)
    a = rng.random((n, n)) < 0.003
    a[510, rng.random(n) < 0.3] = True
    a[rng.random(n) < 0.3, 510] = True
    s, d = np.nonzero(a)
    g = (n, s, d, rng.integers(1, 100, size=len(s)))
    res, nn, sq = run_gpu([g, g])           # twice: the second copy runs as another cluster of the same launch
    Mo, Po, eo, md = run_algos(algos_oracle, *g, hop_cap=HOPS)
    assert ((Po == 510) & (Mo < 510)).sum() > 100, "fixture must exercise the path == 510 collision"
    for gi in range(2):
        M, P, e = unpack(res, nn, sq, gi)
        assert (M == Mo).all() and (P == Po).all() and (e == eo).all()
        assert int(res["maxdist"][gi]) == md


def test_shift_and_no_path_variant(lib_built):
    rng = np.random.default_rng(9)
    graphs = []
    for _ in range(20):
        n = int(rng.integers(1, 90))
        a = rng.random((n, n)) < 0.06
        s, d = np.nonzero(a)
        graphs.append((n, s, d, rng.integers(1, 60, size=len(s))))
    raw, n, sq = run_gpu(graphs, shift=0, want_path=True)
    sh, _, _ = run_gpu(graphs, shift=1, want_path=False)
    assert sh["path"] is None
    assert (sh["dist"] == raw["dist"] + 1).all()
    assert (sh["edge_in"] == (raw["edge_in"].astype(np.int16) + 1).astype(np.uint8)).all()   # 255 -> 0
    assert (sh["maxdist"] == raw["maxdist"]).all()


def test_dropin_mirror_matches_reference_call_surface(lib_built):
    """mobgt_b200.algos.{floyd_warshall, gen_edge_input} used exactly like wrapper.py:55-60."""
    from mobgt_b200 import algos
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 40, 131):
        a = rng.random((n, n)) < 0.08
        s, d = np.nonzero(a)
        adj, ef = dense_inputs(n, s, d, rng.integers(1, 40, size=len(s)))
        M, path = algos.floyd_warshall(adj)
        Mo, po = algos_oracle.floyd_warshall(adj)
        assert M.dtype == np.int64 and path.dtype == np.int64
        assert (M == Mo).all() and (path == po).all()
        md = int(np.amax(M))
        e = algos.gen_edge_input(md, path, ef)
        eo = algos_oracle.gen_edge_input(md, po, ef)
        assert e.dtype == np.float32 and e.shape == eo.shape == (n, n, md, 1)
        assert (e == eo).all()
    with pytest.raises(AssertionError):
        algos.floyd_warshall(np.zeros((3, 4), bool))          # algos.pyx:12 `assert nrows == ncols`


def test_degrees(lib_built):
    from mobgt_b200 import _C
    rng = np.random.default_rng(4)
    graphs = []
    for _ in range(12):
        n = int(rng.integers(1, 70))
        a = rng.random((n, n)) < 0.1
        s, d = np.nonzero(a)
        graphs.append((n, s, d, rng.integers(1, 9, size=len(s))))
    feat, n, sq, no = pack(graphs)
    f = torch.from_numpy(feat).cuda()
    ind = torch.empty(int(no[-1]), dtype=torch.int16, device="cuda")
    outd = torch.empty_like(ind)
    n_d, sq_d, no_d = torch.from_numpy(n).cuda(), torch.from_numpy(sq).cuda(), torch.from_numpy(no).cuda()
    _C.call("mobgt_degrees", _C.ptr(f), _C.ptr(n_d), _C.ptr(sq_d), _C.ptr(no_d), len(graphs), 1, _C.ptr(ind),
            _C.ptr(outd), _C.stream_ptr())
    torch.cuda.synchronize()
    for gi, (nn, s, d, c) in enumerate(graphs):
        adj, _ = dense_inputs(nn, s, d, c)
        assert (ind[no[gi]:no[gi + 1]].cpu().numpy() == adj.sum(1) + 1).all()     # wrapper.py:97 (+1 collator.py:12)
        assert (outd[no[gi]:no[gi + 1]].cpu().numpy() == adj.sum(0) + 1).all()    # wrapper.py:98


def test_stress_n512_roundtrip_properties(lib_built):
    """Full-size (c3 stress) size-independent properties: M symmetric for symmetric graphs, triangle
    inequality on finite entries, hop count == M where node 0 is not involved (node-0-free graphs)."""
    rng = np.random.default_rng(77)
    n = 512
    a = rng.random((n, n)) < 0.006
    a = a | a.T
    for v in (0, 510):                   # isolate node 0 (path == 0 quirk) and node 510 (its index collides with the
        a[v, :] = False                  # reference's 510 sentinel, algos.pyx:87-88): then hop count must equal distance
        a[:, v] = False
    s, d = np.nonzero(a)
    res, nn, sq = run_gpu([(n, s, d, np.ones(len(s), np.int64))])
    M, P, e = unpack(res, nn, sq, 0)
    M = M.astype(np.int64)
    assert (M == M.T).all()
    fin = M < 510
    hops = (e != -1).sum(-1)
    assert (hops[fin] == np.minimum(M[fin], HOPS)).all()
    k = rng.integers(0, n, size=64)
    for kk in k:
        assert (M <= np.minimum(510, M[:, [kk]] + M[[kk], :])).all()
