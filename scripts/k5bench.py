"""K5 alone: device time of the fused head (mode 1 + merge, and mode 0) at the c2 / c5-shard shapes: python scripts/k5bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mobgt_b200 import ops

pk = bench.peaks()
_share = torch.zeros(32 * 8192, dtype=torch.int32, device='cuda')


def share_zero():
    _share.zero_()
    return _share.data_ptr()


dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
from mobgt_b200 import _C as _C0
for cl, (name, M, V, k) in [(c, t) for c in (1, 0) for t in (("c2", 256, 60001, 10), ("c5_shard", 4096, 125000, 10), ("c5_shard_k20", 4096, 125000, 20), ("mid", 1024, 125000, 10))]:
    _C0.call("mobgt_debug_head_cluster", cl)          # 1: adjacent row tiles paired into 2-CTA clusters (W stages multicast), 0: off
    name = f"{name}{'' if cl else ' (no cluster)'}"
    g = torch.Generator(device=dev).manual_seed(5)
    z = torch.randn(M, 320, device=dev, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, 320, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    b = torch.randn(V, device=dev, generator=g) * 0.1
    tgt = torch.randint(0, V, (M,), device=dev, generator=g).int()
    st = ops.head_target_logit(z, W, b, tgt)
    t1 = bench.time_kernel(lambda: ops.head_topk_local(z, W, b, tgt, k, st=st), flush, iters=6)
    t0 = bench.time_kernel(lambda: ops.head_target_logit(z, W, b, tgt), flush, iters=6)
    # the two launches of mode 1 apart, on preallocated buffers (no allocator / host time between the events)
    from mobgt_b200 import _C
    ns = ops.head_split(M, V)
    tv = torch.empty(M, ns, k, dtype=torch.float32, device=dev); ti = torch.empty(M, ns, k, dtype=torch.int32, device=dev)
    cg = torch.empty(M, ns, dtype=torch.int32, device=dev); ce = torch.empty(M, ns, dtype=torch.int32, device=dev)
    ov = torch.empty(M, k, dtype=torch.float32, device=dev); oi = torch.empty(M, k, dtype=torch.int32, device=dev)
    rk = torch.empty(M, dtype=torch.int32, device=dev)
    sp = _C.stream_ptr()
    th = bench.time_kernel(lambda: _C.call("mobgt_head_topk", _C.ptr(z), _C.ptr(W), _C.ptr(b), _C.ptr(tgt), M, V, 320, 0, k, ns, 1,
                                           _C.ptr(st), _C.ptr(tv), _C.ptr(ti), _C.ptr(cg), _C.ptr(ce), None, share_zero(), sp), flush, iters=6)
    tm = bench.time_kernel(lambda: _C.call("mobgt_topk_merge", _C.ptr(tv), _C.ptr(ti), _C.ptr(cg), _C.ptr(ce), M, ns, k,
                                           _C.ptr(ov), _C.ptr(oi), _C.ptr(rk), sp), flush, iters=6)
    fl = 2.0 * M * 320 * V
    print(f"{name:14s} M={M} V={V} k={k} nsplit={ops.head_split(M, V)}: mode1+merge {t1*1e3:8.1f} us = {fl/t1/1e9:7.1f} TF/s "
          f"({100*fl/t1/1e9/pk['tc']:.1f}% of bf16 peak) ; mode0 {t0*1e3:7.1f} us ; head kernel alone {th*1e3:8.1f} us = {fl/th/1e9:7.1f} TF/s ({100*fl/th/1e9/pk['tc']:.1f}%) ; merge alone {tm*1e3:6.1f} us")

_C0.call("mobgt_debug_head_cluster", 1)
if "--dbg" in sys.argv:
    # timing experiments on the c5 shard shape: which part of the head kernel bounds it (results of dbg runs are NOT valid top-k)
    M, V, k = 4096, 125000, 10
    g = torch.Generator(device=dev).manual_seed(5)
    z = torch.randn(M, 320, device=dev, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, 320, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    b = torch.randn(V, device=dev, generator=g) * 0.1
    tgt = torch.randint(0, V, (M,), device=dev, generator=g).int()
    st = ops.head_target_logit(z, W, b, tgt)
    ns = ops.head_split(M, V)
    tv = torch.empty(M, ns, k, dtype=torch.float32, device=dev); ti = torch.empty(M, ns, k, dtype=torch.int32, device=dev)
    cg = torch.empty(M, ns, dtype=torch.int32, device=dev); ce = torch.empty(M, ns, dtype=torch.int32, device=dev)
    fl = 2.0 * M * 320 * V
    for name, mode in (("full, spin 24ns", 1), ("full, spin 8ns", 1 | (1 << 12)), ("full, spin 64ns", 1 | (8 << 12)), ("full, spin 120ns", 1 | (15 << 12)),
                       ("no harvest", 1 | (1 << 8)), ("ld + release only", 1 | (2 << 8))):
        for nsp in (ns, 2 * ns if 2 * ns <= 64 else ns):
            tvx = torch.empty(M, nsp, k, dtype=torch.float32, device=dev); tix = torch.empty(M, nsp, k, dtype=torch.int32, device=dev)
            cgx = torch.empty(M, nsp, dtype=torch.int32, device=dev); cex = torch.empty(M, nsp, dtype=torch.int32, device=dev)
            t = bench.time_kernel(lambda: _C.call("mobgt_head_topk", _C.ptr(z), _C.ptr(W), _C.ptr(b), _C.ptr(tgt), M, V, 320, 0, k, nsp, mode,
                                                  _C.ptr(st), _C.ptr(tvx), _C.ptr(tix), _C.ptr(cgx), _C.ptr(cex), None, share_zero(), sp), flush, iters=6)
            print(f"dbg {name:20s} nsplit={nsp:3d}: {t*1e3:8.1f} us = {fl/t/1e9:7.1f} TF/s ({100*fl/t/1e9/pk['tc']:.1f}%)")

if "--timeline" in sys.argv:
    M, V, k = 4096, 125000, 10
    g = torch.Generator(device=dev).manual_seed(5)
    z = torch.randn(M, 320, device=dev, generator=g).to(torch.bfloat16)
    W = (torch.randn(V, 320, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    b = torch.randn(V, device=dev, generator=g) * 0.1
    tgt = torch.randint(0, V, (M,), device=dev, generator=g).int()
    st = ops.head_target_logit(z, W, b, tgt)
    ns = ops.head_split(M, V)
    tv = torch.empty(M, ns, k, dtype=torch.float32, device=dev); ti = torch.empty(M, ns, k, dtype=torch.int32, device=dev)
    cg = torch.empty(M, ns, dtype=torch.int32, device=dev); ce = torch.empty(M, ns, dtype=torch.int32, device=dev)
    tl = torch.zeros(256, dtype=torch.int64, device=dev)
    for name, mode in (("full", 1), ("no harvest", 1 | (1 << 8)), ("ld + release only", 1 | (2 << 8))):
        tl.zero_()
        _C.call("mobgt_debug_set_timeline", tl.data_ptr())
        _C.call("mobgt_head_topk", _C.ptr(z), _C.ptr(W), _C.ptr(b), _C.ptr(tgt), M, V, 320, 0, k, ns, mode,
                _C.ptr(st), _C.ptr(tv), _C.ptr(ti), _C.ptr(cg), _C.ptr(ce), None, share_zero(), sp)
        torch.cuda.synchronize()
        _C.call("mobgt_debug_set_timeline", None)
        t = tl.cpu().tolist()
        t0 = t[16]
        print(f"== timeline {name}: cycles since the MMA warp started tile 0 (CTA (1,1))")
        for i in range(14):
            m = [t[16 + 4 * i + j] - t0 if t[16 + 4 * i + j] else None for j in range(3)]
            e = [t[80 + 4 * i + j] - t0 if t[80 + 4 * i + j] else None for j in range(3)]
            print(f"  tile {i:2d}: mma wait-acc {m[0]} got-acc {m[1]} last-kblock-issued {m[2]} | epi acc-full {e[0]} released {e[1]} done {e[2]}")
