import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mobgt_b200 import synth
from mobgt_b200.algos import apsp_edge_input_packed, pack_graphs
world = synth.make_world("tiny", seed=1)
n, G = int(sys.argv[1]), int(sys.argv[2])
its = synth.make_items(world, G, n, seed=3, cfg_id=3, n_fixed=n)
print("item n:", sorted(set(len(np.asarray(it.x)) for it in its)))
ns = np.full(G, n, np.int32)
nn, sq, no = pack_graphs(ns)
feat = np.zeros(int(sq[-1]), np.uint8)
for g, it in enumerate(its):
    ei = np.asarray(it.edge_index)
    feat[sq[g] + ei[0] * n + ei[1]] = np.asarray(it.edge_attr).reshape(-1) + 2
fd, nd, sd = torch.from_numpy(feat).cuda(), torch.from_numpy(nn).cuda(), torch.from_numpy(sq).cuda()
for edges in (True, False):
    r = apsp_edge_input_packed(fd, nd, sd, nn, 20, 1, want_edges=edges)
    torch.cuda.synchronize()
    print("ok", edges, int(r["maxdist"].max()))
