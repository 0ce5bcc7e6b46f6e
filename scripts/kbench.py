"""Per-kernel device times at the bench workload's shapes (no training loop): python scripts/kbench.py [workload] [--k1]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from mobgt_b200 import collator, model as M, synth
from mobgt_b200.algos import apsp_edge_input_packed, pack_graphs

args = [a for a in sys.argv[1:] if not a.startswith("--")]
iters = int([a.split("=")[1] for a in sys.argv if a.startswith("--iters=")][0]) if any(a.startswith("--iters=") for a in sys.argv) else 8
workload = args[0] if args else "c2-dense128"
pk = bench.peaks()
world = synth.make_world("c2", seed=1)
items = bench.make_workload(workload, world, 256, 0)
torch.manual_seed(1)
model = M.Graphormer(dataset_name="toyotagraph", world=world, **bench.HP).cuda().train()
t0 = time.perf_counter()
b = collator.collate_packed(items, world, None, 512, 20, 1024)
torch.cuda.synchronize()
t1 = time.perf_counter()
b = collator.collate_packed(items, world, None, 512, 20, 1024)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"collate wall: first {1e3 * (t1 - t0):.1f} ms, second {1e3 * (t2 - t1):.1f} ms")
rep = bench.kernel_report(model, b, pk, iters=iters)
for k, v in rep.items():
    print(f"{k:22s} {v['ms'] * 1e3:9.1f} us  {v['gbs']:8.1f} GB/s  hbm {100 * v['frac_hbm']:5.1f}%"
          + (f"  {v['tflops']:7.1f} TF/s tc {100 * v['frac_tc']:4.1f}%" if "tflops" in v else ""))

if "--k1" in sys.argv:
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(0)
    for n, G in ((8, 4096), (32, 2048), (64, 1024), (128, 256), (128, 1024), (256, 256), (512, 64)):
        its = synth.make_items(world, G, n, seed=3, cfg_id=3, n_fixed=n)
        ns = np.full(G, n, np.int32)
        nn, sq, no = pack_graphs(ns)
        feat = np.zeros(int(sq[-1]), np.uint8)
        for g, it in enumerate(its):
            ei = np.asarray(it.edge_index)
            feat[sq[g] + ei[0] * n + ei[1]] = np.asarray(it.edge_attr).reshape(-1) + 2
        fd, nd, sd = torch.from_numpy(feat).cuda(), torch.from_numpy(nn).cuda(), torch.from_numpy(sq).cuda()
        for edges in (True, False):
            ms = bench.time_kernel(lambda: apsp_edge_input_packed(fd, nd, sd, nn, 20, 1, want_edges=edges), flush, iters=4)
            cells = int(sq[-1])
            by = cells * (1 + 2 + (20 if edges else 0))
            print(f"k1 n={n:4d} G={G:5d} edges={int(edges)}: {ms * 1e3:9.1f} us  {G / ms * 1e3:10.0f} graphs/s  {by / ms / 1e6:8.1f} GB/s "
                  f"({100 * by / ms / 1e6 / pk['hbm']:.1f}% hbm)  {2.0 * n ** 3 * G / ms / 1e9:7.2f} Tminplus/s")
