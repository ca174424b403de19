"""GPU: flash attention kernel (head dims 72->80, 32, 16; ragged key counts) against fp32 softmax attention."""
import pytest
import torch

from tpxl_b200 import _lib
from gpu_util import rel_l2, st

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,Nq,Nk,Dh,DhP", [(2, 16, 2048, 2048, 72, 80), (1, 16, 2048, 1370, 72, 80), (2, 4, 300, 200, 72, 80), (3, 8, 64, 64, 32, 32),
                                             (1, 4, 130, 1, 16, 16), (1, 16, 256, 77, 24, 32)])
def test_attention(B, H, Nq, Nk, Dh, DhP):
    g = torch.Generator(device="cuda").manual_seed(Nq + Nk)
    def mk(n):
        t = torch.zeros(B, H, n, DhP, dtype=torch.float16, device="cuda")
        t[..., :Dh] = torch.randn(B, H, n, Dh, generator=g, device="cuda").half()
        return t
    q, k, v = mk(Nq), mk(Nk), mk(Nk)
    out = torch.empty(B, Nq, H * Dh, dtype=torch.float16, device="cuda")
    scale = Dh ** -0.5
    _lib.check(_lib.lib().tpx_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Nq, Nk, Dh, DhP, scale, st()))
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    ref = torch.matmul(torch.softmax(s, -1), v.float())[..., :Dh].permute(0, 2, 1, 3).reshape(B, Nq, H * Dh)
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref) < 2e-3
    assert (out.float() - ref).abs().max() < 2e-2


def test_attention_uniform_when_keys_identical():
    """The identity the null-conditioning collapse relies on: identical keys -> output == v (exactly, in fp16)."""
    B, H, Nq, Nk, Dh, DhP = 1, 2, 128, 1370, 72, 80
    q = torch.zeros(B, H, Nq, DhP, dtype=torch.float16, device="cuda")
    q[..., :Dh] = torch.randn(B, H, Nq, Dh, device="cuda").half()
    krow = torch.zeros(B, H, 1, DhP, dtype=torch.float16, device="cuda")
    krow[..., :Dh] = torch.randn(B, H, 1, Dh, device="cuda").half()
    vrow = torch.zeros_like(krow)
    vrow[..., :Dh] = torch.randn(B, H, 1, Dh, device="cuda").half()
    k, v = krow.expand(B, H, Nk, DhP).contiguous(), vrow.expand(B, H, Nk, DhP).contiguous()
    out = torch.empty(B, Nq, H * Dh, dtype=torch.float16, device="cuda")
    _lib.check(_lib.lib().tpx_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Nq, Nk, Dh, DhP, Dh ** -0.5, st()))
    ref = vrow[..., :Dh].permute(0, 2, 1, 3).reshape(B, 1, H * Dh).expand(B, Nq, H * Dh)
    torch.cuda.synchronize()
    assert (out.float() - ref.float()).abs().max() <= 1e-3 * ref.float().abs().max()


@pytest.mark.parametrize("B,H,Nq,Nk,Dh", [(2, 16, 2048, 2048, 72), (1, 16, 2048, 1370, 72), (1, 2, 256, 128, 72), (2, 3, 300, 77, 72), (1, 4, 128, 1000, 72),
                                           (1, 1, 513, 257, 72), (1, 12, 1374, 1374, 64), (2, 3, 200, 333, 64)])
def test_attention_tcgen05(B, H, Nq, Nk, Dh):
    """tcgen05 path (Dh 72 or 64 -> 80-wide tiles): V given transposed [B,H,80,NkPad]; ragged query and key counts; the DINOv2 shape
    (12 heads x 64, 1374 tokens)."""
    DhP = 80
    NkPad = (Nk + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(Nq * 7 + Nk)
    def mk(n):
        t = torch.zeros(B, H, n, DhP, dtype=torch.float16, device="cuda")
        t[..., :Dh] = torch.randn(B, H, n, Dh, generator=g, device="cuda").half()
        return t
    q, k, v = mk(Nq), mk(Nk), mk(Nk)
    vT = torch.zeros(B, H, DhP, NkPad, dtype=torch.float16, device="cuda")
    vT[..., :Nk] = v.transpose(-1, -2)
    out = torch.full((B, Nq, H * Dh), 9.0, dtype=torch.float16, device="cuda")
    scale = Dh ** -0.5
    _lib.check(_lib.lib().tpx_attention_tc(q.data_ptr(), k.data_ptr(), vT.data_ptr(), out.data_ptr(), B, H, Nq, Nk, NkPad, Dh, scale, st()))
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    ref = torch.matmul(torch.softmax(s, -1), v.float())[..., :Dh].permute(0, 2, 1, 3).reshape(B, Nq, H * Dh)
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref) < 2e-3
    assert (out.float() - ref).abs().max() < 2e-2

