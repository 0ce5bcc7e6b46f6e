#!/usr/bin/env python
"""bench.py — MobGT training throughput (trajectory graphs / s) on N B200s, next to the reference's CPU path.

    python bench.py --gpus 1 --steps 20 --warmup 5                       # this framework (libmobgt kernels)
    python bench.py --impl reference --gpus 1 --steps K --warmup W       # the reference's CPU implementation of the path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                        # data parallel, one rank per GPU (weak scaling)

Workload (BASELINE.json configs[1]): toyotagraph-shaped synthetic world (60 000 POIs, 300 categories, 995 users),
hidden 128 (+64), 8 heads, 6 layers, ffn 1024, multi_hop_max_dist 20, batch 256 graphs per GPU, bf16 GEMMs/attention.
`--workload c2-dense128` (default) = every graph at the 128-node cap (the stress variant SURVEY.md §8d quotes its byte
counts on); `--workload c2-natural` = node counts drawn from the real Gowalla-Nevada histogram clipped to 128.
One step = forward + loss + backward + (NCCL all-reduce) + AdamW over one batch; nothing is skipped or cached.

value  = graphs/s with the collated batch already resident in HBM;
e2e    = graphs/s through the public API from HOST buffers: raw items -> collator (pack, H2D, K1 APSP + path edges on
         the GPU) -> training step -> loss read back, every step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HP = dict(n_layers=6, num_heads=8, hidden_dim=128, dropout_rate=0.1, intput_dropout_rate=0.1, weight_decay=0.01, ffn_dim=1024,
          warmup_updates=40000, tot_updates=400000, peak_lr=2e-4, end_lr=1e-9, edge_type="multi_hop", multi_hop_max_dist=20,
          attention_dropout_rate=0.1)      # README.md:62 canonical flags


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tc=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        rows = [r for r in rows if len(r) >= 7]
        if not rows:
            return out
        sm = [float(r[0]) for r in rows]
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = float(rows[0][1])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i].strip().lower().startswith("active") for r in rows)]
        out["samples"] = len(rows)
        return out


def make_workload(workload, world, B, rank, seed=1):
    from mobgt_b200 import synth
    n_fixed = 128 if workload == "c2-dense128" else None
    return synth.make_items(world, B, 128, seed=seed, cfg_id=2, n_fixed=n_fixed, start=rank * B)


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_rate(workload, steps, warmup, sample=None, verbose=False):
    """The reference's own CPU implementation of the path on the host cores: compiled algos.pyx (oracle/_ref) when present
    (else the C port), the collator and model_fqandtoyo restatement (oracle/model_oracle.py), forward + loss + backward +
    AdamW, torch threads = all cores.  Each step processes a bounded SAMPLE of the workload's batch."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import algos_oracle
    import build_ref
    import model_oracle as mo
    from mobgt_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref_algos = build_ref.load()
    world = synth.make_world("c2", seed=1)
    sample = sample or 4
    items = make_workload(workload, world, sample, 0)
    torch.manual_seed(1)
    model = mo.Graphormer(world, n_layers=HP["n_layers"], ffn_dim=HP["ffn_dim"], dropout_rate=HP["dropout_rate"],
                          intput_dropout_rate=HP["intput_dropout_rate"], attention_dropout_rate=HP["attention_dropout_rate"],
                          pos_dropout=0.1).train()
    opt = torch.optim.AdamW(model.parameters(), lr=HP["peak_lr"], weight_decay=HP["weight_decay"])

    def preprocess(it):
        if ref_algos is None:
            return mo.preprocess_item(it, hop_cap=None)
        saved = (algos_oracle.floyd_warshall, algos_oracle.gen_edge_input)
        algos_oracle.floyd_warshall = ref_algos.floyd_warshall
        algos_oracle.gen_edge_input = lambda md, p, ef, hop_cap=None: ref_algos.gen_edge_input(md, p, ef)
        try:
            return mo.preprocess_item(it, hop_cap=None)          # full wrapper.py:55-60 cost, 510-deep hop axis included
        finally:
            algos_oracle.floyd_warshall, algos_oracle.gen_edge_input = saved

    def step():
        b = mo.collate([preprocess(it) for it in items], world, multi_hop_max_dist=20, rel_pos_max=1024)
        loss = model.training_loss(b)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dict(value=sample * steps / dt, ms_per_step=1e3 * dt / steps, cores=cores, sample_graphs=sample,
                kind="reference+port" if ref_algos is not None else "port")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # exactly K timed steps after W warm-up steps, each a bounded 4-graph sample (~1 s on 16 cores): K = 20, W = 5 is ~25 s
    steps, warm = max(1, args.steps), max(1, args.warmup)
    r = cpu_reference_rate(args.workload, steps, warm)
    sample = (f"{r['sample_graphs']} graphs/step of the {args.workload} batch, {steps} timed steps after {warm} warm-up; "
              f"compiled algos.pyx (oracle/_ref) + CPU restatement of collator/model_fqandtoyo, fwd+bwd+AdamW")
    line = {"metric": "train_graphs_per_sec", "value": r["value"], "unit": "graphs/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": args.workload, "world": "toyotagraph-shaped P=60000 C=300 U=995", "hidden": 128, "layers": 6,
                       "heads": 8, "ffn": 1024, "multi_hop_max_dist": 20, "graphs_per_step": r["sample_graphs"]},
            "cpu_baseline": {"value": r["value"], "unit": "graphs/s", "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ kernel micro-timing
def time_kernel(fn, flush, iters=8, warm=2):
    """CUDA-event time of fn() on the current stream, L2 flushed before every launch; returns mean ms."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(warm):
        fn()
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def kernel_report(model, batch, pk, iters=8, head_c5=True):
    """Per-kernel device time at the workload's shapes + algorithmic bytes -> roofline."""
    from mobgt_b200 import ops
    from mobgt_b200.algos import apsp_edge_input_packed
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, H, T = batch.B, 8, batch.N + 1
    Tp = ops.bias_pitch(T)
    n = batch.n_host.astype(np.int64)
    cells, pairs_t, ntok = int((n * n).sum()), int(((n + 1) ** 2).sum()), int(batch.tok_pos.numel())
    hops = batch.hops
    tabs = [t.detach().float().contiguous() for t in (model.rel_pos_encoder.weight, model.poi_pos_encoder.weight,
                                                       model.edge_encoder.weight, model.edge_dis_encoder.weight.view(-1),
                                                       model.graph_token_virtual_distance.weight.view(-1))]
    bias = ops.bias_fwd_raw(batch, *tabs)
    qkv = torch.randn(ntok, 576, device=dev).to(torch.bfloat16)
    dp, ds = HP["attention_dropout_rate"], 0x5DEECE66D1234567      # K3 is timed as the training step runs it: dropout on
    out, lse = ops.attn_fwd_raw(qkv, bias, batch, drop_p=dp, seed=ds)
    dout = torch.randn(ntok, 192, device=dev).to(torch.bfloat16)
    n_layers = len(model.layers)
    planes = torch.zeros((n_layers,) + tuple(bias.shape), dtype=torch.bfloat16, device=dev)   # per-layer dS planes (bf16)
    rep = {}
    tk = lambda fn: time_kernel(fn, flush, iters=iters, warm=min(2, iters))

    def add(name, ms, nbytes, per_step, flops=None):
        rep[name] = dict(ms=ms, gbs=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / pk["hbm"], bytes=nbytes, launches_per_step=per_step)
        if flops:
            rep[name]["tflops"] = flops / ms / 1e9
            rep[name]["frac_tc"] = flops / ms / 1e9 / pk["tc"]

    add("k1_apsp_edge_input", tk(lambda: apsp_edge_input_packed(batch.feat8, batch.n, batch.sq_off, batch.n_host, hops, 1)),
        cells * (1 + 2 + hops), 0)
    add("k2_bias_fwd", tk(lambda: ops.bias_fwd_raw(batch, *tabs)), cells * (4 + hops) + H * pairs_t * 2, 1)
    add("k2_bias_bwd", tk(lambda: ops.bias_bwd_raw(batch, planes, tabs[2], tabs[3], tabs[1].shape[0])),
        cells * (4 + hops) + n_layers * H * pairs_t * 2, 1)
    fl = 4.0 * H * pairs_t * 24
    add("k3_attn_fwd", tk(lambda: ops.attn_fwd_raw(qkv, bias, batch, drop_p=dp, seed=ds)), 4 * ntok * 192 * 2 + H * pairs_t * 2, 6, fl)
    add("k3_attn_bwd", tk(lambda: ops.attn_bwd_raw(qkv, bias, out, dout, lse, batch, planes[0], 2, drop_p=dp, seed=ds)),
        8 * ntok * 192 * 2 + H * pairs_t * 2 + H * pairs_t * 2, 6, 2.5 * fl)
    Gd, Gc = model.gcn_tables()
    Gd, Gc = Gd.detach(), Gc.detach()
    Tm = model.time_embed_model_48.weight.detach()
    nn_ = ntok - B
    add("k4_embed_gather", tk(lambda: ops.embed_gather_raw(batch, model.cat_of_poi, Gd, Tm, Gc)),
        nn_ * (192 * 4 + 192 * 2 + 12), 1)
    nf = torch.randn(nn_, 192, device=dev).to(torch.bfloat16)
    pe, gt = model.pos_embed.pe.detach(), model.graph_token.weight.detach().view(-1)
    Din, Dout = model.in_degree_encoder.weight.detach(), model.out_degree_encoder.weight.detach()
    add("k4_embed_sum", tk(lambda: ops.embed_sum_raw(batch, nf, Din, Dout, pe, gt)),
        nn_ * (192 * 2 + 3 * 192 * 4 + 16) + ntok * 192 * 2, 1)
    g = torch.randn(ntok, 192, device=dev).to(torch.bfloat16)
    plan = ops.sort_plan(batch.tok_pos)
    add("k4_segment_sum", tk(lambda: ops.segment_sum_raw(g, 0, 192, plan, 2000)), ntok * (192 * 2 + 8), 6)
    # K6 — encoder row ops (first slice): LayerNorm fwd / bwd at [ntok, 192], bias-gradient column sums
    ln = model.layers[0].ffn_norm2
    xs = torch.randn(ntok, 192, device=dev)
    dy32, dy16 = torch.randn(ntok, 192, device=dev), torch.randn(ntok, 192, device=dev).to(torch.bfloat16)
    add("k6_layernorm_fwd", tk(lambda: ops.layer_norm(xs, ln, "both")), ntok * 192 * (4 + 4 + 2), 12)
    # backward through the C-ABI directly (an autograd.grad call would time Python / autograd dispatch, not the kernel)
    from mobgt_b200 import _C as C_
    mean_, rstd_ = torch.empty(ntok, device=dev), torch.empty(ntok, device=dev)
    o32_, o16_ = torch.empty(ntok, 192, device=dev), torch.empty(ntok, 192, device=dev, dtype=torch.bfloat16)
    gam, bet = ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous()
    C_.call("mobgt_layernorm_fwd", C_.ptr(xs), C_.ptr(gam), C_.ptr(bet), float(ln.eps), ntok, 192, C_.ptr(o32_), C_.ptr(o16_),
            C_.ptr(mean_), C_.ptr(rstd_), C_.stream_ptr())
    dx_, dg_, db_ = torch.empty_like(xs), torch.empty(192, device=dev), torch.empty(192, device=dev)
    wsb = int(C_.lib().mobgt_layernorm_bwd_workspace_bytes(192))
    ws_ = torch.empty(wsb, dtype=torch.uint8, device=dev)
    add("k6_layernorm_bwd", tk(lambda: C_.call("mobgt_layernorm_bwd", C_.ptr(dy32), C_.ptr(dy16), C_.ptr(xs), C_.ptr(gam), C_.ptr(mean_),
                                               C_.ptr(rstd_), ntok, 192, C_.ptr(dx_), C_.ptr(dg_), C_.ptr(db_), C_.ptr(ws_), wsb,
                                               C_.stream_ptr())),
        ntok * 192 * (4 + 2 + 4 + 4), 12)
    wide = torch.randn(ntok, 1024, device=dev).to(torch.bfloat16)
    wide2 = torch.randn(ntok, 1024, device=dev).to(torch.bfloat16)
    add("k6_gelu_bwd_colsum", tk(lambda: ops.gelu_bwd_colsum_raw(wide2, wide)), ntok * 1024 * (2 + 2 + 2), 6)
    del wide2
    add("k6_colsum_1024", tk(lambda: ops.colsum(wide)), ntok * 1024 * 2, 0)
    add("k6_colsum_192", tk(lambda: ops.colsum(dy16)), ntok * 192 * 2, 12)
    del wide
    # K5 — the evaluation head (not part of the training step): c2 shape, and one GPU's shard of the c5 shape
    for name, Mh, Vh in (("k5_head_c2", 256, 60001),) + ((("k5_head_c5_shard", 4096, 125000),) if head_c5 else ()):
        g_ = torch.Generator(device=dev).manual_seed(5)
        z = torch.randn(Mh, 320, device=dev, generator=g_).to(torch.bfloat16)
        Wh = (torch.randn(Vh, 320, device=dev, generator=g_) * 0.02).to(torch.bfloat16)
        tgt = torch.randint(0, Vh, (Mh,), device=dev, generator=g_).int()
        st = ops.head_target_logit(z, Wh, None, tgt)
        add(name, tk(lambda: ops.head_topk_local(z, Wh, None, tgt, 10, st=st)), Vh * 320 * 2 + Mh * 320 * 2 + Mh * 10 * 8, 0,
            2.0 * Mh * 320 * Vh)
        del z, Wh
    del flush
    return rep


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch.distributed as dist
    from mobgt_b200 import _C, collator, graphs, model as M, synth
    rank, world_size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    _C.require_cuda()
    pk = peaks()
    world = synth.make_world("c2", seed=1)
    B = args.batch
    items = make_workload(args.workload, world, B, rank)
    torch.manual_seed(1)
    model = M.Graphormer(dataset_name="toyotagraph", world=world, **HP).to(dev).train()
    params = [p for p in model.parameters()]
    # one flat fp32 gradient buffer: p.grad are views, so the data-parallel exchange is a single NCCL all-reduce
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    opt = torch.optim.AdamW(params, lr=HP["peak_lr"], weight_decay=HP["weight_decay"], fused=True)
    from mobgt_b200.lr import PolynomialDecayLR
    sched = PolynomialDecayLR(opt, HP["warmup_updates"], HP["tot_updates"], HP["peak_lr"], HP["end_lr"], 1.0)
    latlon = torch.from_numpy(world.latlon).to(dev)

    def collate():
        return collator.collate_packed(items, world, latlon, 512, 20, 1024, device=dev)

    graphed = {"g": None}

    def train_step(b):
        g = graphed["g"]
        loss = None
        if g is not None:
            try:
                g.load(b)
                loss = g.run()
            except graphs.ShapeMismatch:
                loss = None
        if loss is None:
            graphed["eager_steps"] = graphed.get("eager_steps", 0) + 1
            flat.zero_()
            loss = model.training_step(b)
            loss.backward()
        if world_size > 1:
            dist.all_reduce(flat)
            flat.div_(world_size)
        opt.step()
        sched.step()
        return loss

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finalize=None):
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        if finalize is not None:
            finalize()
        b_.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b_)], device=dev)
        if world_size > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    batch = collate()
    for _ in range(args.warmup):
        train_step(batch)
    # fixed-shape workloads: capture forward + backward in a CUDA graph (mobgt_b200/graphs.py); anything not capturable, and every
    # batch whose shapes differ from the captured ones, runs eagerly
    graph_note = "off"
    if args.cuda_graph and args.workload == "c2-dense128":
        try:
            graphed["g"] = graphs.GraphedTrainStep(model, flat, batch)
            graph_note = "fwd+bwd captured"
            for _ in range(2):
                train_step(batch)
        except graphs.GraphCaptureError as e:
            graph_note = f"eager ({str(e)[:160]})"
            print(f"[bench] {graph_note}", file=sys.stderr)
    n0 = _C.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(lambda: train_step(batch), args.steps)
    launches = (_C.launch_count() - n0) / args.steps        # libmobgt launches issued from the host in the timed region ...
    if graphed["g"] is not None:
        launches += graphed["g"].launches                    # ... plus the libmobgt kernel nodes of the graph replayed every step
    # end to end: host items -> collate (H2D + K1) -> step -> loss on the host, every step
    losses = []

    def endless():
        while True:
            yield items

    # one-batch-ahead loader (the reference's DataLoader-worker role, data.py:255-267): the numpy packing of the raw items runs
    # in 2 worker processes that hand over one pinned byte buffer per batch; the H2D copy, the sort plans, K1 and poi_pos of
    # batch i+1 are issued right after the kernels of step i have been enqueued.  Every timed step still contains exactly one
    # collate (pack + pinned H2D + K1 + poi_pos), one training step and one D2H read of the loss.
    # side-stream collation: K1 / poi_pos / the sort plans of batch i+1 run next to the head of training step i.  The loader
    # gates them behind the previous step (PackedLoader.current records the gate), so they never run next to the NCCL
    # all-reduce: N = 2 e2e 8.96 ms with the gate, 11.0 ms without it, 9.76 ms on the training stream
    # (profiles/r04z_bench_n2*.json)
    side = not args.no_side_stream
    loader = collator.PackedLoader(endless(), num_workers=args.loader_workers, side_stream=side, world=world,
                                   latlon_dev=latlon,
                                   multi_hop_max_dist=20, rel_pos_max=1024, device=dev)

    # the loss of every step is read back to the host (async D2H into pinned memory + event); the host consumes it one step
    # late, so the read-back does not drain the GPU queue between steps.  The last read happens inside the timed region.
    loss_pin = torch.empty(2, dtype=torch.float32).pin_memory()
    state = dict(pending=None, k=0)

    def e2e_flush():
        if state["pending"] is not None:
            ev, slot = state["pending"]
            ev.synchronize()
            losses.append(float(loss_pin[slot]))
            state["pending"] = None

    trace = [] if os.environ.get("MOBGT_E2E_TRACE") else None

    def e2e_step():
        if trace is not None:
            trace.append(time.perf_counter())
        b = loader.current()
        loss = train_step(b)
        slot = state["k"] & 1
        loss_pin[slot:slot + 1].copy_(loss.detach().view(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        loader.advance()
        e2e_flush()                       # the PREVIOUS step's loss (its event has long completed)
        state["pending"], state["k"] = (ev, slot), state["k"] + 1
        return b

    # warm-up: the caching allocator's pool of collation buffers (two batches in flight + blocks waiting for the consumer
    # stream) and the loader workers reach their steady state
    for _ in range(max(6, args.warmup)):
        bb = e2e_step()
    e2e_flush()
    e2e_steps = max(4, args.steps)
    ms_e2e = timed(e2e_step, e2e_steps, finalize=e2e_flush)
    clocks = sampler.stop() if sampler else None
    if trace is not None and rank == 0:
        dt = np.diff(np.array(trace[-e2e_steps:])) * 1e3
        print("[bench] e2e host step intervals (ms): " + " ".join(f"{x:.1f}" for x in dt), file=sys.stderr)
        print(f"[bench] eager (non-graph) steps so far: {graphed.get('eager_steps', 0)}", file=sys.stderr)
    h2d = int(bb.h2d_bytes)
    if rank == 0:
        graphs = B * world_size
        line = {"metric": "train_graphs_per_sec", "value": graphs * args.steps / (ms / 1e3), "unit": "graphs/s",
                "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": args.workload, "world": "toyotagraph-shaped P=60000 C=300 U=995", "hidden": 128, "layers": 6,
                           "heads": 8, "ffn": 1024, "multi_hop_max_dist": 20, "graphs_per_gpu": B,
                           "tokens_per_gpu": int(batch.tok_pos.numel()), "parallelism": f"dp{world_size}", "cuda_graph": graph_note,
                           "e2e_collate_stream": "side" if side else "training",
                           "l2": "per-step working set (activations + bias planes, > 1 GB) exceeds the 126 MB L2"},
                "e2e": {"value": graphs * e2e_steps / (ms_e2e / 1e3), "unit": "graphs/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / e2e_steps},
                "gpu_launches": launches, "clocks": clocks, "final_loss": losses[-1] if losses else None}
        if world_size == 1 and not args.no_kernel_report:
            rep = kernel_report(model, batch, pk)
            top = max(rep, key=lambda k: rep[k]["ms"] * rep[k]["launches_per_step"])
            r = rep[top]
            tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram bytes / launch from the committed ncu capture
            traffic = json.load(open(tp)).get(top) if (os.path.exists(tp) and args.workload == "c2-dense128") else None
            line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": r["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                                "frac": r["frac_hbm"], "traffic": traffic, "peak_source": pk["src"],
                                "share_of_step": r["ms"] * r["launches_per_step"] / (ms / args.steps)}
            line["kernels"] = {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in rep.items()}
        if world_size == 1 and not args.no_cpu_baseline:
            r = cpu_reference_rate(args.workload, steps=4, warmup=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": "graphs/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['sample_graphs']} graphs/step of the same batch, 4 timed steps after 1 warm-up; "
                                              f"compiled algos.pyx (oracle/_ref) when present + CPU restatement of "
                                              f"collator/model_fqandtoyo (fwd+bwd+AdamW), {r['cores']} torch threads"}
        emit(line)
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2-dense128", choices=["c2-dense128", "c2-natural"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--loader-workers", type=int, default=2, help="DataLoader worker processes packing raw items (e2e path)")
    ap.add_argument("--no-side-stream", action="store_true", help="e2e: collate on the training stream (no overlap)")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false", help="run every step eagerly")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-report", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    # stdout carries exactly ONE JSON line: everything libraries print to fd 1 meanwhile (the NCCL version banner, ...) goes to
    # stderr; the line is written to the saved descriptor at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
