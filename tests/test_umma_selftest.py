"""GPU: the tcgen05 / TMA building blocks (csrc/umma.cuh) in isolation — descriptor conventions for
K-major and MN-major no-swizzle operands loaded by 3-D TMA, accumulators read back from TMEM."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("N,K", [(16, 16), (32, 32), (144, 32), (128, 128), (256, 64)])
def test_umma_matches_torch(lib_built, a_mn, b_mn, N, K):
    from mobgt_b200 import _C
    _C.require_cuda()
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K + a_mn * 7 + b_mn * 13)
    A = torch.randn(128, K, device="cuda", generator=g).to(torch.bfloat16)       # logical [M,K]
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)         # logical [N,K]
    Ag = A.t().contiguous() if a_mn else A.contiguous()
    Bg = B.t().contiguous() if b_mn else B.contiguous()
    out = torch.full((128, N), float("nan"), device="cuda")
    _C.call("mobgt_selftest_umma", _C.ptr(Ag), _C.ptr(Bg), N, K, a_mn, b_mn, _C.ptr(out), _C.stream_ptr())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (out - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("N,K", [(128, 64), (128, 128), (256, 64), (64, 256)])
def test_umma_sw128_kmajor(lib_built, N, K):
    """K-major operands in the 128-byte-swizzle layout (2-D TMA boxes {64, rows}, make_smem_desc_sw128) — what K5 uses."""
    from mobgt_b200 import _C
    _C.require_cuda()
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K + 5)
    A = torch.randn(128, K, device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.full((128, N), float("nan"), device="cuda")
    _C.call("mobgt_selftest_umma", _C.ptr(A), _C.ptr(B), N, K, 2, 2, _C.ptr(out), _C.stream_ptr())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (out - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
