"""Acc@k / NDCG@k / MRR with the reference's semantics (model_fqandtoyo.py:48-90, 122-131), computed on the device.

`get_acc` keeps the reference's quirks: hits need target > 0, and the batch loop BREAKS at the first row whose target is 0
(:88-89).  `MRR_metric` replaces the reference's full descending argsort per row on the CPU by a rank count:
rank0 = #(s > s_t) + #(s == s_t and idx < t) (ties broken towards the lower index; the reference leaves ties to
np.argsort).  Return types match the reference: numpy (4,1) arrays / a Python float."""
import numpy as np
import torch


def rank_of_target(scores, target):
    s_t = scores.gather(1, target.view(-1, 1))
    idx = torch.arange(scores.shape[1], device=scores.device).view(1, -1)
    return ((scores > s_t).sum(1) + ((scores == s_t) & (idx < target.view(-1, 1))).sum(1))


def get_acc(target, scores):
    """-> (acc[4,1], ndcg[4,1]) in the reference's slot order: [top10, top5, top1, top20]."""
    target = target.view(-1)
    _, idxx = scores.topk(20, 1)
    zero = (target == 0).nonzero()
    stop = int(zero[0]) if len(zero) else len(target)          # `break` at the first target == 0
    acc = np.zeros((4, 1))
    ndcg = np.zeros((4, 1))
    if stop == 0:
        return acc, ndcg
    t, p = target[:stop], idxx[:stop]
    hit = (p == t.view(-1, 1)) & (t.view(-1, 1) > 0)
    pos = hit.float().argmax(1)                                 # index of the target inside the top-20 list
    found = hit.any(1)
    for slot, k in ((3, 20), (0, 10), (1, 5), (2, 1)):
        ok = found & (pos < k)
        acc[slot] = float(ok.sum())
        ndcg[slot] = float((1.0 / torch.log2(pos[ok].double() + 2)).sum())
    return acc, ndcg


def MRR_metric(target, scores):
    r = rank_of_target(scores, target.view(-1))
    return float((1.0 / (r.double() + 1)).sum())
