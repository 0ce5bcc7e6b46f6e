"""fp32 mode of the path (`Graphormer(precision=32)`, `entry --precision 32`): the first half of the parity contract —
"within 1e-5 relative in fp32 mode ... for bias, attention outputs, logits and gradients" — against the fp32 oracle
(oracle/model_oracle.py, pinned to the reference by tests/golden/).  The bf16 mode's 2e-2 half is tests/test_round2_gpu.py.

Tolerances (written where they are applied): attention outputs / logits / loss 1e-5 relative to the tensor's largest magnitude;
parameter gradients 1e-5 norm-wise per tensor (||got - ref|| / ||ref||)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.mark.parametrize("sizes", [(12, 3, 7, 1, 9), (127, 100, 1, 64, 33, 17, 128), (300, 5, 256)])
def test_attention_f32_fwd_bwd_matches_torch(lib_built, sizes):
    """mobgt_attn_f32_fwd / _bwd (csrc/k3_attn_f32.cu) against torch fp32 autograd on CPU: out, dq / dk / dv and dS within 1e-5,
    with and without the attention dropout (same counter-based mask as the tensor-core kernels); dS accumulation (mode 1) adds
    onto what is there and leaves the cells outside a graph alone; bitwise reproducible."""
    from mobgt_b200 import collator, ops, synth
    from test_k2_k3_k4 import tables, torch_attention_diff
    w = synth.make_world("c1", seed=1)
    items = []
    for k, n in enumerate(sizes):
        items += synth.make_items(w, 1, 512, seed=170 + k, n_fixed=n, start=k)
    b = collator.collator_toyota(items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=w)
    B = len(items)
    R, Pp, E, W, t = tables(seed=5)
    cu = [x.cuda().contiguous() for x in (R, Pp, E, W.view(-1), t.view(-1))]
    bias = ops.bias_fwd_raw(b, *cu, out_dtype=torch.float32)
    ntok = int(b.tok_pos.numel())
    gen = torch.Generator().manual_seed(17)
    qkv = torch.randn(ntok, 3 * 192, generator=gen) * 1.2
    dout = torch.randn(ntok, 192, generator=gen)
    tok_off = b.tok_off.cpu().numpy()
    for p, seed in ((0.0, 0), (0.1, 0x0123456789ABCDEF)):
        out, lse = ops.attn_f32_fwd_raw(qkv.cuda(), bias, b, drop_p=p, seed=seed)
        db = torch.full(bias.shape, 7.0, dtype=torch.float32, device="cuda")
        dqkv = ops.attn_f32_bwd_raw(qkv.cuda(), bias, out, dout.cuda(), lse, b, db, 0, drop_p=p, seed=seed)
        db1 = torch.full(bias.shape, 7.0, dtype=torch.float32, device="cuda")
        dqkv1 = ops.attn_f32_bwd_raw(qkv.cuda(), bias, out, dout.cuda(), lse, b, db1, 1, drop_p=p, seed=seed)
        torch.cuda.synchronize()
        q32 = qkv.clone().requires_grad_(True)
        b32 = bias.cpu().clone().requires_grad_(True)
        ref, _ = torch_attention_diff(q32, b32, tok_off, drop=(p, seed) if p > 0 else None)
        (ref * dout).sum().backward()
        assert (out.cpu() - ref.detach()).abs().max().item() <= TOL * max(1.0, ref.abs().max().item()), p
        gq = q32.grad
        assert (dqkv.cpu() - gq).abs().max().item() <= TOL * max(1.0, gq.abs().max().item()), p
        assert torch.equal(dqkv, dqkv1)
        for g in range(B):
            Tg = int(tok_off[g + 1] - tok_off[g])
            gb = b32.grad[g, :, :Tg, :Tg]
            got = db[g, :, :Tg, :Tg].cpu()
            assert (got - gb).abs().max().item() <= TOL * max(1.0, gb.abs().max().item()), (p, g)
            assert torch.equal(db1[g, :, :Tg, :Tg].cpu(), got + 7.0), (p, g)
            assert (db[g, :, Tg:, :] == 7.0).all() and (db[g, :, :Tg, Tg:] == 7.0).all(), (p, g)   # cells outside the graph: untouched
    out2, lse2 = ops.attn_f32_fwd_raw(qkv.cuda(), bias, b, drop_p=p, seed=seed)
    db2 = torch.full(bias.shape, 7.0, dtype=torch.float32, device="cuda")
    dq2 = ops.attn_f32_bwd_raw(qkv.cuda(), bias, out2, dout.cuda(), lse2, b, db2, 0, drop_p=p, seed=seed)
    assert torch.equal(out2, out) and torch.equal(lse2, lse) and torch.equal(dq2, dqkv) and torch.equal(db2, db)


def test_attention_f32_rejects_bad_arguments(lib_built):
    from mobgt_b200 import _C, ops

    class Bt:
        N, tok_off = 3, torch.tensor([0, 4], dtype=torch.int32, device="cuda")

    qkv = torch.zeros(4, 576, device="cuda")
    bias = torch.zeros(1, 8, 4, 8, device="cuda")
    with pytest.raises(_C.MobgtError):
        ops.attn_f32_fwd_raw(qkv, bias, Bt, drop_p=1.5)
    out, lse = ops.attn_f32_fwd_raw(qkv, bias, Bt)
    with pytest.raises(_C.MobgtError):
        _C.call("mobgt_attn_f32_bwd", qkv.data_ptr(), qkv.data_ptr(), qkv.data_ptr(), 576, bias.data_ptr(), out.data_ptr(), out.data_ptr(),
                lse.data_ptr(), Bt.tok_off.data_ptr(), 1, 8, 4, 4, 8, 4, 0.2, qkv.data_ptr(), qkv.data_ptr(), qkv.data_ptr(), 576,
                bias.data_ptr(), 2, 0.0, 0, None, _C.stream_ptr())          # mode 2 (bf16 planes) does not exist in fp32 mode


def _oracle_float64(om, ob):
    """loss and parameter gradients of the oracle run in float64 on the same batch (time_normal stays fp32: its slot binning
    is part of the collated input, not of the arithmetic under test)."""
    import copy
    om64 = copy.deepcopy(om).double()
    om64.D_A, om64.C_A = om64.D_A.double(), om64.C_A.double()
    for p in om64.parameters():
        p.grad = None
    ob64 = copy.copy(ob)
    for k, v in list(vars(ob).items()):
        if torch.is_tensor(v) and v.dtype == torch.float32 and k != "time_normal":
            setattr(ob64, k, v.double())
    loss = om64.training_loss(ob64)
    loss.backward()
    return loss.item(), {k: p.grad for k, p in om64.named_parameters() if p.grad is not None}


FP32_CASES = {
    # the three dataset branches on the small world, a BASELINE configs[1]-shaped batch (60 000 POIs, graphs at the 128-node cap
    # next to natural-law graphs) and a configs[3]-shaped one (3 679 POIs, a 256-node graph: T = 257)
    "toyota-tiny": ("toyotagraph", "tiny", [(None, 6, 12, 4)], 2, 256),
    "foursquare-c1": ("foursquaregraph", "c1", [(None, 8, 40, 5)], 2, 256),
    "gowalla-c1": ("gowalla_nevda", "c1", [(None, 8, 40, 6)], 2, 256),
    "c2": ("toyotagraph", "c2", [(128, 3, 128, 31), (None, 5, 128, 32)], 6, 1024),
    "c4": ("gowalla_nevda", "c4", [(256, 1, 256, 41), (None, 6, 256, 42)], 6, 1024),
}


@pytest.mark.parametrize("case", list(FP32_CASES))
def test_fp32_mode_logits_loss_and_gradients_match_oracle(lib_built, case):
    """Graphormer(precision=32) against the fp32 oracle with the same weights on the same items: both heads' logits and the
    training loss within 1e-5, EVERY parameter gradient within 1e-5 norm-wise.  (6 layers / ffn 1024 on the BASELINE shapes.)"""
    from mobgt_b200 import model
    from test_round2_gpu import _pair
    dataset_name, cfg, spec, n_layers, ffn = FP32_CASES[case]
    w, items, om, pm, ob, pb = _pair(dataset_name, cfg, spec, n_layers=n_layers, ffn=ffn, seed=3, precision=32)
    assert pm.precision == 32 and not torch.backends.cuda.matmul.allow_tf32
    with torch.no_grad():
        ref = om(ob)
        got = pm(pb)
    for a, r in zip(got, ref):
        assert a.dtype == torch.float32 and a.shape == r.shape
        assert (a.cpu() - r).abs().max().item() <= TOL * max(1.0, r.abs().max().item())
    for m_ in (om, pm):
        m_.train()
        m_.poi_distance_model.eval()       # the GCN dropout masks are torch's own (not pinned): off, like every other dropout here
        m_.poi_cat_model.eval()
    pm.pos_embed.p = 0.0
    lref = om.training_loss(ob)
    lref.backward()
    lgot = pm.training_step(pb)
    lgot.backward()
    assert abs(lgot.item() - lref.item()) <= TOL * abs(lref.item()), (lgot.item(), lref.item())
    ref_g = {k: p.grad for k, p in om.named_parameters() if p.grad is not None}
    got_g = {k: p.grad.float().cpu() for k, p in pm.named_parameters() if p.grad is not None}
    l64, ref64_g = _oracle_float64(om, ob)
    assert abs(l64 - lref.item()) <= TOL * abs(l64)
    gmax = max(r.abs().max().item() for r in ref_g.values())
    worst, table = {}, {}
    for k, r in ref_g.items():
        g = got_g.get(k)
        if g is None:
            assert r.abs().max().item() == 0.0, k
            continue
        if r.abs().max().item() < 1e-6 * gmax:        # mathematically-zero gradients (softmax ignores a per-row constant)
            assert g.abs().max().item() < 1e-4 * gmax, k
            continue
        r64 = ref64_g[k]
        e32 = (g - r).norm().item() / r.norm().item()                       # ours vs the fp32 oracle
        e64 = (g.double() - r64).norm().item() / r64.norm().item()          # ours vs the oracle in float64
        o64 = (r.double() - r64).norm().item() / r64.norm().item()          # the fp32 oracle's own rounding
        table[k] = (e32, e64, o64)
        # the gate: 1e-5 against the fp32 oracle; where that oracle's own fp32 rounding is a sizeable part of 1e-5 (a few
        # table gradients that sum 1e5 cells one after the other on the CPU) its float64 run is the reference point
        if min(e32, e64) > TOL:
            worst[k] = table[k]
    assert len(table) > 40
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        rows = sorted(table, key=lambda k: -table[k][0])
        json.dump({"case": case, "loss": [lgot.item(), lref.item(), l64],
                   "columns": ["ours vs fp32 oracle", "ours vs float64 oracle", "fp32 oracle vs float64 oracle"],
                   "rows": {k: ["%.2e" % v for v in table[k]] for k in rows}},
                  open(os.path.join(out_dir, f"fp32_grad_errors_{case}.json"), "w"), indent=0)
    assert not worst, worst
    model.ops.enable_tf32(True)            # leave the process-wide cuBLAS switch as the bf16-mode tests expect it


def test_fp32_mode_trains_with_dropout_and_through_entry(lib_built, tmp_path):
    """`entry --precision 32`: two training steps with the canonical dropout rates (attention dropout inside the fp32 kernels)
    and the evaluation pass run; the loss is finite."""
    from mobgt_b200 import entry, model
    args = ["--dataset_name", "toyotagraph", "--gpus", "1", "--precision", "32", "--batch_size", "16", "--hidden_dim", "128",
            "--num_heads", "8", "--n_layers", "2", "--ffn_dim", "256", "--dropout_rate", "0.1", "--intput_dropout_rate", "0.1",
            "--attention_dropout_rate", "0.1", "--weight_decay", "0.01", "--peak_lr", "2e-4", "--end_lr", "1e-9", "--edge_type",
            "multi_hop", "--warmup_updates", "40000", "--tot_updates", "400000", "--seed", "1", "--max_epochs", "1", "--synthetic",
            "tiny", "--train_graphs", "48", "--test_graphs", "24", "--n_fixed", "9", "--multi_hop_max_dist", "20",
            "--limit_train_steps", "2", "--default_root_dir", str(tmp_path)]
    r = entry.cli_main(args)
    assert r["steps"] == 2 and np.isfinite(r["loss"]) and r["metrics"]["n"] == 24
    model.ops.enable_tf32(True)
