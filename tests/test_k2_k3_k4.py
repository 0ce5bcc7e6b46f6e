"""GPU: kernel-level parity of K2 (bias build fwd/bwd), K3 (tcgen05 attention fwd) and K4 (embedding gather / sum /
deterministic segmented scatter-add) against the oracle restatement (oracle/model_oracle.py) and plain torch fp32."""
import numpy as np
import pytest
import torch

import model_oracle as mo
from helpers import attn_drop_keep

pytestmark = pytest.mark.gpu


def make_case(cfg="tiny", B=6, cap=12, seed=1, n_fixed=None, **world_kw):
    from mobgt_b200 import collator, synth
    w = synth.make_world(cfg, seed=seed, **world_kw)
    items = synth.make_items(w, B, cap, seed=seed, n_fixed=n_fixed)
    ob = mo.collate([mo.preprocess_item(it, hop_cap=20) for it in items], w, multi_hop_max_dist=20, rel_pos_max=1024)
    b = collator.collator_toyota(items, max_node=512, multi_hop_max_dist=20, rel_pos_max=1024, world=w)
    return w, items, ob, b


def tables(H=8, bins=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    R = torch.randn(512, H, generator=g) * 0.3
    Pp = torch.randn(bins, H, generator=g) * 0.3
    E = torch.randn(128, H, generator=g) * 0.3
    W = torch.randn(128 * H * H, 1, generator=g) * 0.3
    t = torch.randn(1, H, generator=g) * 0.3
    R[0] = 0
    Pp[0] = 0
    E[0] = 0
    return R, Pp, E, W, t


def oracle_bias(ob, R, Pp, E, W, t, H=8):
    class M:
        pass
    m = mo.Graphormer.__new__(mo.Graphormer)
    torch.nn.Module.__init__(m)
    m.num_heads, m.multi_hop_max_dist = H, 20
    m.rel_pos_encoder = torch.nn.Embedding.from_pretrained(R.clone(), freeze=False, padding_idx=0)
    m.poi_pos_encoder = torch.nn.Embedding.from_pretrained(Pp.clone(), freeze=False, padding_idx=0)
    m.edge_encoder = torch.nn.Embedding.from_pretrained(E.clone(), freeze=False, padding_idx=0)
    m.edge_dis_encoder = torch.nn.Embedding.from_pretrained(W.clone(), freeze=False)
    m.graph_token_virtual_distance = torch.nn.Embedding.from_pretrained(t.clone(), freeze=False)
    return m, m.attn_bias_build(ob, "fp32")


@pytest.mark.parametrize("cfg,B,cap,nfix", [("tiny", 6, 12, None), ("c1", 5, 40, None), ("c1", 3, 70, 70)])
def test_collated_fields_match_oracle_collator(lib_built, cfg, B, cap, nfix):
    """Batch1's reference-shaped views == the oracle's collator (collator.py restatement), bit-exact."""
    w, items, ob, b = make_case(cfg, B, cap, n_fixed=nfix)
    for name in ("x", "rel_pos", "edge_input", "in_degree", "out_degree", "attn_bias", "attn_edge_type", "adj1",
                 "time", "time_normal", "cat", "poi_pos", "user", "y"):
        got = getattr(b, name).cpu()
        exp = getattr(ob, name)
        assert got.shape == exp.shape, (name, got.shape, exp.shape)
        assert got.dtype == exp.dtype, (name, got.dtype, exp.dtype)
        assert torch.equal(got, exp), name
    assert b.spatial_pos is b.rel_pos or torch.equal(b.spatial_pos, b.rel_pos)
    assert len(b) == B


@pytest.mark.parametrize("cfg,B,cap,nfix", [("tiny", 6, 12, None), ("c1", 8, 60, None), ("c1", 4, 128, 128)])
def test_bias_fwd_fp32_and_bf16(lib_built, cfg, B, cap, nfix):
    from mobgt_b200 import ops
    w, items, ob, b = make_case(cfg, B, cap, n_fixed=nfix)
    R, Pp, E, W, t = tables()
    _, ref = oracle_bias(ob, R, Pp, E, W, t)
    ref = ref.detach()
    cu = [x.cuda().contiguous() for x in (R, Pp, E, W.view(-1), t.view(-1))]
    out32 = ops.bias_fwd_raw(b, *cu, out_dtype=torch.float32).cpu()
    out16 = ops.bias_fwd_raw(b, *cu, out_dtype=torch.bfloat16).float().cpu()
    T = b.N + 1
    n = b.n_host
    for g in range(B):
        Tg = int(n[g]) + 1
        r = ref[g, :, :Tg, :Tg]
        assert torch.isfinite(r).all()
        # fp32 mode: 1e-5 relative (north_star tolerance)
        assert torch.allclose(out32[g, :, :Tg, :Tg], r, rtol=1e-5, atol=1e-6), g
        # bf16 mode: 2e-2 relative
        assert torch.allclose(out16[g, :, :Tg, :Tg], r, rtol=2e-2, atol=2e-2), g
        # padding columns of real rows are -inf in the reference; the kernels mask them from seqlens instead
        assert (ref[g, :, :Tg, Tg:] == float("-inf")).all()


def test_bias_bwd_matches_autograd_of_oracle(lib_built):
    from mobgt_b200 import ops
    w, items, ob, b = make_case("c1", 6, 50)
    R, Pp, E, W, t = tables(seed=3)
    m, ref = oracle_bias(ob, R, Pp, E, W, t)
    B, H, T = ref.shape[0], 8, b.N + 1
    Tp = ops.bias_pitch(T)
    gen = torch.Generator().manual_seed(5)
    dB = torch.zeros(B, H, T, Tp)
    for g in range(B):
        Tg = int(b.n_host[g]) + 1
        dB[g, :, :Tg, :Tg] = torch.randn(H, Tg, Tg, generator=gen)
    finite = torch.where(torch.isfinite(ref), ref, torch.zeros_like(ref))
    (finite * dB[..., :T]).sum().backward()
    dR, dP, dE, dW, dt = ops.bias_bwd_raw(b, dB.cuda(), E.cuda().contiguous(), W.view(-1).cuda().contiguous(), Pp.shape[0])
    torch.cuda.synchronize()

    def close(a, bb, name):
        a, bb = a.cpu(), bb
        scale = bb.abs().max().item() + 1e-6
        assert (a - bb).abs().max().item() <= 2e-5 * scale + 1e-5, (name, (a - bb).abs().max().item(), scale)

    close(dR, m.rel_pos_encoder.weight.grad, "dR")
    close(dP, m.poi_pos_encoder.weight.grad, "dPpos")
    close(dE, m.edge_encoder.weight.grad, "dE")
    close(dW.view(-1, 1), m.edge_dis_encoder.weight.grad, "dW")
    close(dt.view(1, -1), m.graph_token_virtual_distance.weight.grad, "dtvd")
    # per-layer bf16 planes (mobgt_attn_bwd mode 2) summed inside the kernel == the f32 path on the summed planes
    planes = torch.stack([(dB * c).to(torch.bfloat16) for c in (0.5, 0.25, 1.0)]).cuda()
    ref5 = ops.bias_bwd_raw(b, planes.float().sum(0).contiguous(), E.cuda().contiguous(), W.view(-1).cuda().contiguous(), Pp.shape[0])
    got5 = ops.bias_bwd_raw(b, planes, E.cuda().contiguous(), W.view(-1).cuda().contiguous(), Pp.shape[0])
    for a5, b5 in zip(got5, ref5):
        assert torch.allclose(a5, b5, rtol=1e-5, atol=1e-5 * (b5.abs().max().item() + 1e-6))
    # bitwise reproducible: the histograms are reduced in a fixed order and the rare path (walk bytes that deviate from the
    # expected walk, rel_pos keys outside the plan) accumulates in 64-bit fixed point, where the order of the atomics is immaterial
    for _ in range(3):
        again = ops.bias_bwd_raw(b, planes, E.cuda().contiguous(), W.view(-1).cuda().contiguous(), Pp.shape[0])
        for a5, b5 in zip(again, got5):
            assert torch.equal(a5, b5)
    again = ops.bias_bwd_raw(b, dB.cuda(), E.cuda().contiguous(), W.view(-1).cuda().contiguous(), Pp.shape[0])
    for a5, b5 in zip(again, (dR, dP, dE, dW, dt)):
        assert torch.equal(a5, b5)


def torch_attention(qkv, bias, tok_off, H=8, d=24):
    """fp32 reference of model_fqandtoyo.py:1693-1706 per packed graph (bias already restricted to real columns)."""
    D = H * d
    out = torch.zeros(qkv.shape[0], D)
    lse = torch.zeros(qkv.shape[0], H)
    for g in range(len(tok_off) - 1):
        a, e = int(tok_off[g]), int(tok_off[g + 1])
        T = e - a
        q = qkv[a:e, :D].view(T, H, d).transpose(0, 1).float()
        k = qkv[a:e, D:2 * D].view(T, H, d).transpose(0, 1).float()
        v = qkv[a:e, 2 * D:].view(T, H, d).transpose(0, 1).float()
        s = (q * d ** -0.5) @ k.transpose(1, 2) + bias[g, :, :T, :T].float()
        lse[a:e] = torch.logsumexp(s, dim=-1).t()
        out[a:e] = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(T, D)
    return out, lse


@pytest.mark.parametrize("cfg,B,cap,nfix", [("tiny", 6, 12, None), ("c1", 8, 60, None), ("c1", 5, 128, 128),
                                            ("c1", 3, 300, 300), ("c1", 2, 512, 512)])
def test_attention_fwd_matches_torch(lib_built, cfg, B, cap, nfix):
    from mobgt_b200 import ops
    w, items, ob, b = make_case(cfg, B, cap, n_fixed=nfix)
    R, Pp, E, W, t = tables(seed=7)
    cu = [x.cuda().contiguous() for x in (R, Pp, E, W.view(-1), t.view(-1))]
    bias = ops.bias_fwd_raw(b, *cu, out_dtype=torch.bfloat16)
    ntok = int(b.tok_pos.numel())
    g = torch.Generator().manual_seed(11)
    qkv = (torch.randn(ntok, 3 * 192, generator=g) * 1.5).to(torch.bfloat16)
    out, lse = ops.attn_fwd_raw(qkv.cuda(), bias, b)
    torch.cuda.synchronize()
    ref, ref_lse = torch_attention(qkv, bias.cpu(), b.tok_off.cpu().numpy())
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"attention out max err {err}"
    assert (lse.cpu() - ref_lse).abs().max().item() <= 2e-2


def test_embed_gather_sum_and_segment_sum(lib_built):
    from mobgt_b200 import ops
    w, items, ob, b = make_case("c1", 7, 45)
    g = torch.Generator().manual_seed(2)
    P, C = w.P, w.C
    Gd = torch.randn(P, 128, generator=g)
    Tm = torch.randn(48, 32, generator=g)
    Gc = torch.randn(C, 32, generator=g)
    cop = torch.from_numpy(w.cat_of_poi).int().cuda()
    out = ops.embed_gather_raw(b, cop, Gd.cuda(), Tm.cuda(), Gc.cuda(), torch.float32).cpu()
    x = b.x_nodes.cpu().long()
    exp = torch.cat([Gd[x - 1], Tm[b.slot.cpu().long()], Gc[torch.from_numpy(w.cat_of_poi)[x - 1] - 1]], 1)
    assert torch.equal(out, exp)
    # sum
    Din = torch.randn(128, 192, generator=g)
    Dout = torch.randn(128, 192, generator=g)
    pe = torch.randn(2000, 192, generator=g)
    gt = torch.randn(1, 192, generator=g)
    nf = torch.randn(x.numel(), 192, generator=g)
    tok = ops.embed_sum_raw(b, nf.cuda(), Din.cuda(), Dout.cuda(), pe.cuda(), gt.view(-1).cuda()).cpu()
    tp, tg = b.tok_pos.cpu().long(), b.tok_graph.cpu().long()
    node_rows = (tp > 0).nonzero().view(-1)
    exp_tok = torch.zeros_like(tok)
    exp_tok[tp == 0] = gt + pe[0]
    exp_tok[node_rows] = ((nf + Din[b.in_deg.cpu().long()]) + Dout[b.out_deg.cpu().long()]) + pe[tp[node_rows]]
    assert torch.allclose(tok, exp_tok, rtol=1e-6, atol=1e-6)
    # deterministic segmented sum: == index_add in fp64, and bitwise reproducible
    src = torch.randn(tok.shape[0], 192, generator=g).cuda()
    for keys, nk in ((b.tok_pos, 2000), (b.tok_graph, b.B), (torch.zeros_like(b.tok_pos), 3)):
        plan = ops.sort_plan(keys)
        t1 = ops.segment_sum_raw(src, 0, 192, plan, nk)
        t2 = ops.segment_sum_raw(src, 0, 192, plan, nk)
        assert torch.equal(t1, t2)
        ref = torch.zeros(nk, 192, dtype=torch.float64).index_add_(0, keys.cpu().long(), src.cpu().double())
        assert torch.allclose(t1.cpu().double(), ref, rtol=1e-5, atol=1e-4)
    plan = ops.sort_plan(b.tok_pos)
    part = ops.segment_sum_raw(src, 64, 32, plan, 2000)
    ref = torch.zeros(2000, 32, dtype=torch.float64).index_add_(0, b.tok_pos.cpu().long(), src[:, 64:96].cpu().double())
    assert torch.allclose(part.cpu().double(), ref, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("cfg,B,cap,nfix", [("tiny", 6, 12, None), ("c1", 8, 60, None), ("c1", 5, 128, 128),
                                            ("c1", 3, 300, 300), ("c1", 2, 512, 512)])
def test_attention_bwd_matches_torch_autograd(lib_built, cfg, B, cap, nfix):
    from mobgt_b200 import ops
    w, items, ob, b = make_case(cfg, B, cap, n_fixed=nfix)
    R, Pp, E, W, t = tables(seed=9)
    cu = [x.cuda().contiguous() for x in (R, Pp, E, W.view(-1), t.view(-1))]
    bias = ops.bias_fwd_raw(b, *cu, out_dtype=torch.bfloat16)
    ntok = int(b.tok_pos.numel())
    gen = torch.Generator().manual_seed(21)
    qkv = (torch.randn(ntok, 3 * 192, generator=gen) * 1.2).to(torch.bfloat16)
    dout = (torch.randn(ntok, 192, generator=gen)).to(torch.bfloat16)
    out, lse = ops.attn_fwd_raw(qkv.cuda(), bias, b)
    dbias = torch.full(bias.shape, float("nan"), dtype=torch.float32, device="cuda")
    dqkv = ops.attn_bwd_raw(qkv.cuda(), bias, out, dout.cuda(), lse, b, dbias, 0)
    dbias2 = dbias.clone()
    ops.attn_bwd_raw(qkv.cuda(), bias, out, dout.cuda(), lse, b, dbias2, 1)      # accumulate: 2x
    plane = torch.full(bias.shape, float("nan"), dtype=torch.bfloat16, device="cuda")
    dqkv3 = ops.attn_bwd_raw(qkv.cuda(), bias, out, dout.cuda(), lse, b, plane, 2)
    torch.cuda.synchronize()
    # fp32 autograd reference on the same bf16-rounded inputs
    tok_off = b.tok_off.cpu().numpy()
    q32 = qkv.float().requires_grad_(True)
    b32 = bias.float().cpu().requires_grad_(True)
    ref, _ = torch_attention_diff(q32, b32, tok_off)
    (ref * dout.float()).sum().backward()
    gq = q32.grad
    scale = gq.abs().max().item()
    err = (dqkv.float().cpu() - gq).abs().max().item()
    assert err <= 2e-2 * max(1.0, scale), f"dqkv max err {err} (scale {scale})"
    for g in range(B):
        Tg = int(tok_off[g + 1] - tok_off[g])
        gb = b32.grad[g, :, :Tg, :Tg]
        got = dbias[g, :, :Tg, :Tg].cpu()
        assert torch.isfinite(got).all()
        assert (got - gb).abs().max().item() <= 2e-2 * max(1.0, gb.abs().max().item())
        assert torch.allclose(dbias2[g, :, :Tg, :Tg].cpu(), 2 * got, rtol=1e-6, atol=1e-7)
        # mode 2 (training path): the same dS, rounded once to bf16, TMA-stored from the MMA operand tile
        got16 = plane[g, :, :Tg, :Tg].float().cpu()
        assert torch.equal(got16, got.to(torch.bfloat16).float())
    assert torch.equal(dqkv3.cpu(), dqkv.cpu())


def torch_attention_diff(qkv, bias, tok_off, H=8, d=24, drop=None):
    """drop = (p, seed): apply libmobgt's counter-based keep mask (tests/helpers.attn_drop_keep) like nn.Dropout would."""
    D = H * d
    outs = []
    for g in range(len(tok_off) - 1):
        a, e = int(tok_off[g]), int(tok_off[g + 1])
        T = e - a
        q = qkv[a:e, :D].view(T, H, d).transpose(0, 1)
        k = qkv[a:e, D:2 * D].view(T, H, d).transpose(0, 1)
        v = qkv[a:e, 2 * D:].view(T, H, d).transpose(0, 1)
        s = (q * d ** -0.5) @ k.transpose(1, 2) + bias[g, :, :T, :T]
        pr = torch.softmax(s, -1)
        if drop is not None:
            masks = [attn_drop_keep(g * H + h, T, drop[0], drop[1]) for h in range(H)]
            keep = torch.from_numpy(np.stack([m[0] for m in masks]))
            pr = pr * keep.to(pr.dtype) * masks[0][1]                     # model_fqandtoyo.py:1704
        outs.append((pr @ v).transpose(0, 1).reshape(T, D))
    return torch.cat(outs), None


@pytest.mark.parametrize("cfg,B,cap,nfix", [("tiny", 6, 12, None), ("c1", 5, 128, 128), ("c1", 2, 300, 300)])
def test_attention_dropout_fwd_bwd_matches_torch_with_same_mask(lib_built, cfg, B, cap, nfix):
    """Training-mode attention dropout (model_fqandtoyo.py:1674, 1704): forward and backward regenerate the same counter-based
    mask; against torch autograd with that mask restated in numpy."""
    from mobgt_b200 import ops
    w, items, ob, b = make_case(cfg, B, cap, n_fixed=nfix)
    R, Pp, E, W, t = tables(seed=9)
    cu = [x.cuda().contiguous() for x in (R, Pp, E, W.view(-1), t.view(-1))]
    bias = ops.bias_fwd_raw(b, *cu, out_dtype=torch.bfloat16)
    ntok = int(b.tok_pos.numel())
    gen = torch.Generator().manual_seed(31)
    qkv = (torch.randn(ntok, 3 * 192, generator=gen) * 1.2).to(torch.bfloat16)
    dout = (torch.randn(ntok, 192, generator=gen)).to(torch.bfloat16)
    p, seed = 0.1, 0x9E3779B97F4A7C15
    out, lse = ops.attn_fwd_raw(qkv.cuda(), bias, b, drop_p=p, seed=seed)
    out0, lse0 = ops.attn_fwd_raw(qkv.cuda(), bias, b)
    out_b, _ = ops.attn_fwd_raw(qkv.cuda(), bias, b, drop_p=p, seed=seed + 1)
    dbias = torch.full(bias.shape, float("nan"), dtype=torch.float32, device="cuda")
    dqkv = ops.attn_bwd_raw(qkv.cuda(), bias, out, dout.cuda(), lse, b, dbias, 0, drop_p=p, seed=seed)
    torch.cuda.synchronize()
    assert torch.equal(lse, lse0)                              # the softmax denominator is taken before the mask
    assert not torch.equal(out, out0) and not torch.equal(out, out_b)
    tok_off = b.tok_off.cpu().numpy()
    q32 = qkv.float().requires_grad_(True)
    b32 = bias.float().cpu().requires_grad_(True)
    ref, _ = torch_attention_diff(q32, b32, tok_off, drop=(p, seed))
    err = (out.float().cpu() - ref.detach()).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), f"attention out max err {err}"
    (ref * dout.float()).sum().backward()
    gq = q32.grad
    scale = gq.abs().max().item()
    err = (dqkv.float().cpu() - gq).abs().max().item()
    assert err <= 2e-2 * max(1.0, scale), f"dqkv max err {err} (scale {scale})"
    for g in range(B):
        Tg = int(tok_off[g + 1] - tok_off[g])
        gb = b32.grad[g, :, :Tg, :Tg]
        got = dbias[g, :, :Tg, :Tg].cpu()
        assert torch.isfinite(got).all()
        assert (got - gb).abs().max().item() <= 2e-2 * max(1.0, gb.abs().max().item())


# ---------------------------------------------------------------------------------------------------- K6
@pytest.mark.parametrize("N,D", [(1000, 192), (257, 320), (33, 64)])
def test_layernorm_fwd_bwd_vs_torch(lib_built, N, D):
    """K6 LayerNorm == torch.nn.functional.layer_norm (fp32), forward and autograd backward, incl. the two-output form."""
    from mobgt_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(N, D, generator=g) * 2 + 0.5).cuda()
    ln = torch.nn.LayerNorm(D).cuda()
    with torch.no_grad():
        ln.weight.copy_(torch.randn(D, generator=g).cuda() * 0.5 + 1)
        ln.bias.copy_(torch.randn(D, generator=g).cuda() * 0.1)
    dy32 = torch.randn(N, D, generator=g).cuda()
    dy16 = torch.randn(N, D, generator=g).cuda().to(torch.bfloat16)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), ln.weight, ln.bias, ln.eps)
    (ref * (dy32 + dy16.float())).sum().backward()
    ref_dx, ref_dg, ref_db = xr.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone()
    ln.zero_grad()
    xo = x.clone().requires_grad_(True)
    out, out16 = ops.layer_norm(xo, ln, "both")
    assert torch.allclose(out, ref.detach(), rtol=1e-5, atol=1e-5)
    assert torch.equal(out16, out.to(torch.bfloat16))
    ((out * dy32).sum() + (out16.float() * dy16.float()).sum()).backward()
    scale = ref_dx.abs().max().item()
    assert (xo.grad - ref_dx).abs().max().item() <= 2e-5 * scale + 1e-5
    assert (ln.weight.grad - ref_dg).abs().max().item() <= 2e-5 * ref_dg.abs().max().item() + 1e-4
    assert (ln.bias.grad - ref_db).abs().max().item() <= 2e-5 * ref_db.abs().max().item() + 1e-4
    # single-output forms
    o32 = ops.layer_norm(x, ln, "f32")
    o16 = ops.layer_norm(x, ln, "bf16")
    assert torch.equal(o32, out.detach()) and torch.equal(o16, out16.detach())


@pytest.mark.parametrize("N,C,dt", [(33024, 192, torch.bfloat16), (5000, 1024, torch.bfloat16), (77, 576, torch.bfloat16), (300, 64, torch.float32)])
def test_colsum_vs_torch(lib_built, N, C, dt):
    from mobgt_b200 import ops
    g = torch.Generator().manual_seed(9)
    src = torch.randn(N, C, generator=g).cuda().to(dt)
    ref = src.double().sum(0)
    got = ops.colsum(src)
    assert (got.double() - ref).abs().max().item() <= 1e-5 * N ** 0.5 * max(1.0, ref.abs().max().item())
    # strided view (a slice of a wider matrix)
    wide = torch.randn(N, 2 * C, generator=g).cuda().to(dt)
    got2 = ops.colsum(wide[:, C:])
    assert (got2.double() - wide[:, C:].double().sum(0)).abs().max().item() <= 1e-5 * N ** 0.5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_add_dropout_layernorm_vs_torch(lib_built, p):
    """s = x + dropout(y); out = LN(s): against torch with the SAME mask (recovered from the kernel's own s output)."""
    from mobgt_b200 import ops
    N, D = 777, 192
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, D, generator=g).cuda()
    y = torch.randn(N, D, generator=g).cuda().to(torch.bfloat16)
    ln = torch.nn.LayerNorm(D).cuda()
    with torch.no_grad():
        ln.weight.copy_(torch.randn(D, generator=g).cuda() * 0.5 + 1)
        ln.bias.copy_(torch.randn(D, generator=g).cuda() * 0.1)
    xo, yo = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    s, out, out16 = ops.add_dropout_layer_norm(xo, yo, ln, p, True, "both", need_s=True)
    keep = (s.detach() != x) | (y.float() == 0)                       # the mask the kernel drew
    if p == 0.0:
        assert keep.all()
    else:
        assert 0.85 < keep.float().mean().item() < 0.95
    scale = 1.0 / (1.0 - p)
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    dropped = (yr * scale * keep).to(torch.bfloat16) if p > 0 else yr
    sr = xr + dropped.float()
    ref = torch.nn.functional.layer_norm(sr, (D,), ln.weight, ln.bias, ln.eps)
    assert torch.allclose(s, sr.detach(), rtol=0, atol=1e-6)
    assert torch.allclose(out, ref.detach(), rtol=1e-5, atol=1e-5)
    dy32, dy16, dse = (torch.randn(N, D, generator=g).cuda() for _ in range(3))
    dy16 = dy16.to(torch.bfloat16)
    ln.zero_grad()
    (ref * (dy32 + dy16.float())).sum().backward(retain_graph=True)
    (sr * dse).sum().backward()
    ref_g = (xr.grad.clone(), yr.grad.float().clone(), ln.weight.grad.clone(), ln.bias.grad.clone())
    ln.zero_grad()
    ((out * dy32).sum() + (out16.float() * dy16.float()).sum() + (s * dse).sum()).backward()
    got = (xo.grad, yo.grad.float(), ln.weight.grad, ln.bias.grad)
    for a, b, tol in zip(got, ref_g, (2e-5, 1e-2, 2e-5, 2e-5)):
        assert (a - b).abs().max().item() <= tol * b.abs().max().item() + 1e-4


@pytest.mark.parametrize("N,C", [(33024, 1024), (777, 256), (5, 8)])
def test_gelu_bwd_colsum_vs_torch(lib_built, N, C):
    """K6: backward of nn.GELU() (exact erf form, model_fqandtoyo.py:1650) fused with the bias-gradient column sum, against torch
    autograd on the same bf16 inputs."""
    from mobgt_b200 import ops
    g = torch.Generator().manual_seed(5)
    h = (torch.randn(N, C, generator=g) * 2.0).to(torch.bfloat16).cuda()
    da = torch.randn(N, C, generator=g).to(torch.bfloat16).cuda()
    hr = h.float().requires_grad_(True)
    ref = torch.nn.functional.gelu(hr)
    dh, db = ops.gelu_bwd_colsum_raw(da, h)
    (ref * da.float()).sum().backward()
    assert (dh.float() - hr.grad).abs().max().item() <= 1e-2 * max(1.0, hr.grad.abs().max().item())
    # the fused column sum is the fp32 sum of the bf16 values that were stored
    want = dh.double().sum(0)
    assert (db.double() - want).abs().max().item() <= 1e-5 * N ** 0.5 * max(1.0, want.abs().max().item())


def test_linear_gelu_fn_matches_unfused(lib_built):
    from mobgt_b200 import ops
    g = torch.Generator().manual_seed(6)
    lin = torch.nn.Linear(192, 1024).cuda()
    x = torch.randn(1000, 192, generator=g).to(torch.bfloat16).cuda()
    dy = torch.randn(1000, 1024, generator=g).to(torch.bfloat16).cuda()
    xa = x.clone().requires_grad_(True)
    ya = ops.linear_gelu_bf16(xa, lin)
    (ya.float() * dy.float()).sum().backward()
    ga = (xa.grad.float().clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    lin.zero_grad()
    xb = x.clone().requires_grad_(True)
    yb = torch.nn.functional.gelu(ops.linear_bf16(xb, lin))
    (yb.float() * dy.float()).sum().backward()
    gb = (xb.grad.float(), lin.weight.grad, lin.bias.grad)
    assert (ya.float() - yb.float()).abs().max().item() <= 2e-2 * max(1.0, yb.float().abs().max().item())
    for a_, b_ in zip(ga, gb):
        assert (a_ - b_).abs().max().item() <= 2e-2 * max(1.0, b_.abs().max().item())
