"""Build container only: how does the oracle "port" that bench.py times as the CPU baseline compare with the REAL reference on the
same CPU?  Runs the unmodified reference (wrapper.preprocess_item + collator_toyota + Graphormer fwd + loss + bwd, behind the import
stubs of tests/golden/make_model_golden.py) and the oracle on the same seeded batch of the `tiny` world and prints both times.

    python scripts/ref_vs_port_cpu.py [graphs=16] [cap=32] [layers=6] [ffn=1024]
"""
import copy
import importlib.util
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("g", os.path.join(ROOT, "tests", "golden", "make_model_golden.py"))
g = importlib.util.module_from_spec(spec)
spec.loader.exec_module(g)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 32
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ffn = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
g.CASES["bench"] = ("toyotagraph", B, cap, 21)
hp = dict(g.HP, n_layers=layers, ffn_dim=ffn)
world, items = g.make_world_and_items("bench")
tmp = tempfile.mkdtemp(prefix="mobgt_ref_")
g.write_dataset(world, tmp, "toyotagraph")
g.install_stubs()
sys.path.insert(0, g.REF)
os.chdir(os.path.join(tmp, "graphormer"))
import model_fqandtoyo as ref_model
import collator as ref_collator
import wrapper as ref_wrapper
import model_oracle as mo

torch.manual_seed(0)
rm = ref_model.Graphormer(dataset_name="toyotagraph", **hp).train()
rm.poi_pos_encoder = torch.nn.Embedding(world.num_bins, 8, padding_idx=0)
om = mo.Graphormer(world, n_layers=layers, ffn_dim=ffn, dataset_name="toyotagraph", dropout_rate=0.1, intput_dropout_rate=0.1,
                   attention_dropout_rate=0.1, pos_dropout=0.1).train()


def gtl(inputs, targets, alpha):
    one_hot = torch.zeros_like(inputs)
    one_hot.scatter_(1, targets[:len(inputs)].view(-1, 1), 1)
    prob = torch.sigmoid(inputs)
    return (-alpha * (1 - prob) * one_hot * torch.log(prob) - (1 - one_hot) * prob * torch.log(1 - prob)).mean()


def ref_step():
    its = [ref_wrapper.preprocess_item(g.to_ref_item(it)) for it in items]
    rb = ref_collator.collator_toyota(its, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024)
    rb.poi_pos = rb.poi_pos.clamp(max=world.num_bins - 1)
    out = rm(rb)
    loss = gtl(out[1], rm.cat_target.view(-1).long(), 0.1) + torch.nn.NLLLoss(ignore_index=0)(out[0], rb.y)
    rm.zero_grad()
    loss.backward()


def port_step():
    ob = mo.collate([mo.preprocess_item(it, hop_cap=None) for it in items], world, multi_hop_max_dist=20, rel_pos_max=1024)
    loss = om.training_loss(ob)
    om.zero_grad()
    loss.backward()


for name, fn in (("reference", ref_step), ("oracle port", port_step)):
    fn()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n
    print(f"{name:12s}: {dt * 1e3:8.1f} ms / step of {B} graphs (cap {cap}, {layers} layers, ffn {ffn}, {torch.get_num_threads()} threads)"
          f" = {B / dt:7.1f} graphs/s")
