"""CPU: pin the oracle (oracle/algos_oracle.c) to the reference's own outputs.

The golden fixture (tests/golden/algos_golden.npz) was produced by the unmodified, Cython-compiled
/root/reference/graphormer/algos.pyx (tests/golden/make_golden.py).  The reference ships no tests of
its own, so these vectors are the pin."""
import numpy as np
import pytest

import algos_oracle
from helpers import digest, golden_graph, run_algos


def test_kats_full_arrays(golden):
    names = [str(x) for x in golden["names"]]
    checked = 0
    for key in golden.files:
        if not key.endswith("_M") or not key.startswith("full_"):
            continue
        name = key[len("full_"):-2]
        idx = names.index(name)
        _, n, s, d, c = golden_graph(golden, idx)
        M, path, e20, md = run_algos(algos_oracle, n, s, d, c, hop_cap=20)
        assert (M == golden[f"full_{name}_M"]).all(), name
        assert (path == golden[f"full_{name}_path"]).all(), name
        assert (e20 == golden[f"full_{name}_e20"]).all(), name
        assert md == int(golden["max_dist"][idx])
        checked += 1
    assert checked >= 30


def test_survey_kats_by_value(golden):
    """SURVEY.md §8c KAT-A / KAT-B / KAT-C literal values."""
    a = golden["full_kat_a_M"]
    assert a.tolist() == [[0, 1, 2, 3, 4], [510, 0, 1, 2, 3], [510, 2, 0, 1, 2], [510, 1, 2, 0, 1], [510, 510, 510, 510, 0]]
    assert golden["full_kat_a_path"].tolist() == [[0, 0, 1, 2, 3], [510, 0, 0, 2, 3], [510, 3, 0, 0, 3], [510, 0, 1, 0, 0],
                                                  [510, 510, 510, 510, 0]]
    assert golden["full_kat_a_e20"][0, 3, :5].tolist() == [3, 3, 3, -1, -1]
    assert golden["full_kat_b_node0_e20"][1, 2, :3].tolist() == [0, -1, -1]      # node-0 quirk
    assert golden["full_kat_c_cycle_e20"][2, 1, :2].tolist() == [0, -1]


def test_all_digests(golden):
    """Every fixture graph (hand KATs, random digraphs, all 4 959 real Gowalla-Nevada train graphs n<=256)."""
    G = len(golden["n"])
    bad = []
    for idx in range(G):
        _, n, s, d, c = golden_graph(golden, idx)
        M, path, e20, md = run_algos(algos_oracle, n, s, d, c, hop_cap=20)
        if not (digest(M, path, e20) == golden["digest"][idx]).all() or md != int(golden["max_dist"][idx]):
            bad.append(idx)
    assert not bad, f"{len(bad)} of {G} fixture graphs differ from the reference: {bad[:10]}"


def test_oracle_vs_compiled_reference_random():
    """When oracle/_ref (the really compiled reference) is present, fuzz the oracle against it."""
    import build_ref
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    for _ in range(60):
        n = int(rng.integers(1, 30))
        adj = rng.random((n, n)) < rng.choice([0.05, 0.15, 0.4])
        ef = np.zeros((n, n, 1), np.int64)
        ef[adj] = rng.integers(3, 50, size=(int(adj.sum()), 1))
        M, p = ref.floyd_warshall(adj)
        M2, p2 = algos_oracle.floyd_warshall(adj)
        assert (M == M2).all() and (p == p2).all()
        e = ref.gen_edge_input(int(M.max()), p, ef)
        e2 = algos_oracle.gen_edge_input(int(M.max()), p2, ef)
        assert e.shape == e2.shape and (e == e2).all()


def test_oracle_vs_compiled_reference_node_510_collision():
    """n = 512 with node 510 as a hub: reachable pairs whose last improving intermediate is node 510 carry path == 510 and
    are skipped by the reference's gen_edge_input (algos.pyx:87-88).  The oracle must reproduce that."""
    import build_ref
    ref = build_ref.load()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(1022)
    n = 512
    adj = rng.random((n, n)) < 0.003
    adj[510, rng.random(n) < 0.3] = True
    adj[rng.random(n) < 0.3, 510] = True
    ef = np.zeros((n, n, 1), np.int64)
    ef[adj] = rng.integers(3, 50, size=(int(adj.sum()), 1))
    M, p = ref.floyd_warshall(adj)
    M2, p2 = algos_oracle.floyd_warshall(adj)
    assert (M == M2).all() and (p == p2).all()
    assert ((p == 510) & (M < 510)).sum() > 100
    md = int(M[M < 510].max())              # hop axis just long enough for every walk (keeps the reference's temp small)
    e = ref.gen_edge_input(md, p, ef)
    e2 = algos_oracle.gen_edge_input(md, p2, ef)
    assert e.shape == e2.shape and (e == e2).all()


def test_fw_invariants_property():
    """SURVEY.md §4 property checks on the oracle."""
    rng = np.random.default_rng(11)
    for _ in range(40):
        n = int(rng.integers(1, 50))
        adj = rng.random((n, n)) < 0.1
        M, p = algos_oracle.floyd_warshall(adj)
        assert (np.diag(M) == 0).all() and M.max() <= 510
        assert ((p == 510) == (M == 510)).all()
        ef = np.zeros((n, n, 1), np.int64)
        ef[adj] = 5
        md = int(M.max())
        e = algos_oracle.gen_edge_input(md, p, ef)[..., 0]
        if e.shape[-1] == 0:
            continue
        hops = (e != -1).sum(-1)
        assert (hops <= M).all()                      # the node-0 quirk can only shorten the walk
        # -1 only after the walk: once a hop is -1 all later hops are -1
        first_neg = np.where((e == -1).any(-1), (e == -1).argmax(-1), e.shape[-1])
        assert (first_neg == hops).all()
