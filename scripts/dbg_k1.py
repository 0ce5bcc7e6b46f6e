import sys, numpy as np, torch
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import algos_oracle
from helpers import run_algos
from test_k1_apsp import run_gpu, unpack
for n in (64, 200, 256, 257, 400, 500, 511, 512):
    for want_path in (True, False):
        g = (n, np.arange(n - 1), np.arange(1, n), np.ones(n - 1, np.int64))
        res, nn, sq = run_gpu([g], want_path=want_path)
        M, P, e = unpack(res, nn, sq, 0)
        Mo, Po, eo, md = run_algos(algos_oracle, *g, hop_cap=20)
        msg = []
        for name, a, b in (("M", M, Mo), ("P", P, Po), ("e", e, eo)):
            if a is None: continue
            bad = np.argwhere(a != b)
            if len(bad):
                i = tuple(bad[0]); msg.append(f"{name}: {len(bad)} bad, first {i} got {a[i]} exp {b[i]}; last {tuple(bad[-1])}")
        print(n, want_path, "OK" if not msg else msg, "maxdist", res["maxdist"], md)
