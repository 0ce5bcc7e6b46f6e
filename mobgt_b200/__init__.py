"""mobgt_b200 — B200-native (sm_100a) implementation of the MobGT hot path behind the reference's Python surface.

    algos      drop-in for graphormer/algos.pyx (floyd_warshall, gen_edge_input) + the batched packed form   (K1)
    collator   Batch1 / collator_{foursquare,gowalla,toyota}                                                (packing, K1, poi_pos)
    model      Graphormer: forward(batched_data) -> [poi_logits, cat_logits], training/test steps            (K2, K3, K4, K5)
    entry      the reference's entry.py flag surface + a torch.distributed training loop
    ops        autograd wrappers of the libmobgt kernels; _C: the ctypes binding of include/mobgt.h
"""
__version__ = "0.1.0"
