"""GPU: the tcgen05/TMA GEMM and its fused epilogues through the C ABI, against fp32 torch math."""
import pytest
import torch

from tpxl_b200 import _lib
from gpu_util import dev, linear, linear_ref, rel_l2, st

pytestmark = pytest.mark.gpu


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).half()


@pytest.mark.parametrize("M,N,K,tile", [(128, 128, 64, 128), (256, 256, 128, 128), (512, 384, 256, 192), (512, 512, 1152, 256),
                                        (300, 136, 1152, 128), (4096, 1152, 1152, 0), (1370, 2304, 768, 0), (4096, 1152, 4608, 0)])
def test_linear_bias(M, N, K, tile):
    A, W, b = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, seed=3)
    out = linear(A, W, b, tile_n=tile)
    ref = linear_ref(A, W, b)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-3, (M, N, K, tile)
    assert (out.float() - ref.float()).abs().max() < 0.05


def test_linear_identity_exact():
    """A @ I^T must reproduce A bit-exactly: catches swizzle / descriptor / row-mapping errors with an exact check."""
    M, K = 256, 128
    A = _rand(M, K, seed=4)
    W = torch.eye(K, device="cuda").half()
    out = linear(A, W, None, tile_n=128)
    torch.cuda.synchronize()
    assert torch.equal(out, A)


def test_linear_gelu_and_post_scale():
    A, W, b = _rand(512, 256, seed=5), _rand(768, 256, scale=1 / 16, seed=6), _rand(768, seed=7)
    assert rel_l2(linear(A, W, b, act=1), linear_ref(A, W, b, act=1)) < 2e-3
    s = 72 ** -0.5
    assert rel_l2(linear(A, W, b, post_scale=s), linear_ref(A, W, b, post_scale=s)) < 2e-3


def test_linear_heads_layout_and_padding():
    B, N, H, Dh, DhP, K = 2, 200, 16, 72, 80, 256
    D = H * Dh
    A, W, b = _rand(B * N, K, seed=8), _rand(3 * D, K, scale=1 / 16, seed=9), _rand(3 * D, seed=10)
    outs = [torch.full((B, H, N, DhP), 7.0, dtype=torch.float16, device="cuda") for _ in range(3)]
    _lib.check(_lib.lib().tpx_linear_heads(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                           B * N, 3 * D, K, D, H, Dh, DhP, N, 1.0, 0, -1, 0, st()))
    ref = linear_ref(A, W, b).reshape(B, N, 3, H, Dh)
    torch.cuda.synchronize()
    for w in range(3):
        got = outs[w]
        assert torch.all(got[..., Dh:] == 0), "zero padding of the head dim"
        assert rel_l2(got[..., :Dh], ref[:, :, w].permute(0, 2, 1, 3)) < 2e-3


def test_linear_gated_residual():
    B, N, D, K = 2, 256, 384, 256
    A, W, b = _rand(B * N, K, seed=11), _rand(D, K, scale=1 / 16, seed=12), _rand(D, seed=13)
    gate = _rand(1, 5 * D, seed=14)          # one modulation row shared by both sequences (CFG: same t)
    x = torch.randn(B * N, D, device="cuda")
    x0 = x.clone()
    _lib.check(_lib.lib().tpx_linear_gated(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), gate[:, 2 * D:].data_ptr(), 5 * D, 1, N, x.data_ptr(), D,
                                           B * N, D, K, 0, st()))
    ref = x0 + (gate[0, 2 * D:3 * D].float() * linear_ref(A, W, b).float()).half().float()
    torch.cuda.synchronize()
    assert (x - ref).abs().max() < 2e-2 and rel_l2(x, ref) < 1e-3


def test_linear_gated_ragged_rows_and_per_row_gate():
    """Sequence length that is not a multiple of the 128-row tile: the gate is looked up per row, the last tile is partial,
    and two launches accumulate (the residual update is an L2-side add)."""
    B, N, D, K = 3, 200, 256, 128
    A, W, b = _rand(B * N, K, seed=51), _rand(D, K, scale=1 / 11, seed=52), _rand(D, seed=53)
    gate = _rand(B, D, seed=54)
    x = torch.randn(B * N, D, device="cuda")
    x0 = x.clone()
    for _ in range(2):
        _lib.check(_lib.lib().tpx_linear_gated(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), gate.data_ptr(), D, B, N, x.data_ptr(), D, B * N, D, K, 0, st()))
    upd = (gate.float().repeat_interleave(N, 0) * linear_ref(A, W, b).float()).half().float()
    ref = (x0 + upd) + upd
    torch.cuda.synchronize()
    # (one fp16 ulp where the tensor core's fp32 summation order rounds acc+bias the other way than torch's)
    assert rel_l2(x, ref) < 1e-3 and (x - ref).abs().max() < 4e-2, float((x - ref).abs().max())


@pytest.mark.parametrize("tile", [0, 144, 192])
def test_linear_heads_transposed_v(tile):
    """V stored as [b, head, DhP, tokens] for the tcgen05 attention's PV operand; zero rows for d >= Dh.  Tiles 144 / 192 take
    the bulk-tensor-store epilogue (24-column groups), 0 picks a width that takes the per-thread one."""
    B, N, H, Dh, DhP, K = 2, 256, 16, 72, 80, 256
    D = H * Dh
    A, W, b = _rand(B * N, K, seed=18), _rand(3 * D, K, scale=1 / 16, seed=19), _rand(3 * D, seed=20)
    q = torch.full((B, H, N, DhP), 7.0, dtype=torch.float16, device="cuda")
    k = torch.full((B, H, N, DhP), 7.0, dtype=torch.float16, device="cuda")
    vT = torch.full((B, H, DhP, N), 7.0, dtype=torch.float16, device="cuda")
    _lib.check(_lib.lib().tpx_linear_heads(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), q.data_ptr(), k.data_ptr(), vT.data_ptr(),
                                           B * N, 3 * D, K, D, H, Dh, DhP, N, 1.0, tile, 2, N, st()))
    ref = linear_ref(A, W, b).reshape(B, N, 3, H, Dh)
    torch.cuda.synchronize()
    assert torch.all(vT[:, :, Dh:, :] == 0) and torch.all(k[..., Dh:] == 0) and torch.all(q[..., Dh:] == 0)
    assert rel_l2(vT[:, :, :Dh, :], ref[:, :, 2].permute(0, 2, 3, 1)) < 2e-3
    assert rel_l2(q[..., :Dh], ref[:, :, 0].permute(0, 2, 1, 3)) < 2e-3
    assert rel_l2(k[..., :Dh], ref[:, :, 1].permute(0, 2, 1, 3)) < 2e-3


def test_linear_heads_bulk_store_q_scaled():
    """The N = 1152 head projection as the DiT issues it (144-wide tiles, post-scale re-rounding of q, M = 2 sequences)."""
    B, N, H, Dh, DhP, K = 2, 384, 16, 72, 80, 1152
    D = H * Dh
    A, W, b = _rand(B * N, K, seed=61), _rand(D, K, scale=K ** -0.5, seed=62), _rand(D, seed=63)
    q = torch.full((B, H, N, DhP), 7.0, dtype=torch.float16, device="cuda")
    s = Dh ** -0.5
    _lib.check(_lib.lib().tpx_linear_heads(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), q.data_ptr(), None, None, B * N, D, K, D, H, Dh, DhP, N, s, 144, -1, 0,
                                           st()))
    ref = linear_ref(A, W, b, post_scale=s).reshape(B, N, H, Dh)
    torch.cuda.synchronize()
    assert torch.all(q[..., Dh:] == 0)
    assert rel_l2(q[..., :Dh], ref.permute(0, 2, 1, 3)) < 2e-3


@pytest.mark.parametrize("M,N,K,tile", [(256, 128, 64, -128), (512, 256, 256, -256), (4096, 1152, 1152, -128), (4096, 3456, 1152, -192), (4096, 4608, 1152, -256),
                                        (4096, 1152, 4608, -128), (2048, 1152, 1152, -128), (1370, 2304, 768, -256), (300, 136, 1152, -128)])
def test_linear_two_cta_pairs(M, N, K, tile):
    """cta_group::2 kernel (negative tile_n selects it): 256-row pair tiles, B split across the two CTAs of a cluster."""
    A, W, b = _rand(M, K, seed=31), _rand(N, K, scale=K ** -0.5, seed=32), _rand(N, seed=33)
    out = linear(A, W, b, tile_n=tile)
    ref = linear_ref(A, W, b)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-3, (M, N, K, tile)
    assert rel_l2(linear(A, W, b, act=1, tile_n=tile), linear_ref(A, W, b, act=1)) < 2e-3


def test_two_cta_gated_and_heads_epilogues():
    B, N, H, Dh, DhP, K = 2, 512, 16, 72, 80, 256
    D = H * Dh
    A, W, b = _rand(B * N, K, seed=41), _rand(3 * D, K, scale=1 / 16, seed=42), _rand(3 * D, seed=43)
    q = torch.zeros(B, H, N, DhP, dtype=torch.float16, device="cuda")
    k = torch.zeros_like(q)
    vT = torch.zeros(B, H, DhP, N, dtype=torch.float16, device="cuda")
    _lib.check(_lib.lib().tpx_linear_heads(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), q.data_ptr(), k.data_ptr(), vT.data_ptr(),
                                           B * N, 3 * D, K, D, H, Dh, DhP, N, 1.0, -192, 2, N, st()))
    ref = linear_ref(A, W, b).reshape(B, N, 3, H, Dh)
    torch.cuda.synchronize()
    assert rel_l2(q[..., :Dh], ref[:, :, 0].permute(0, 2, 1, 3)) < 2e-3 and rel_l2(vT[:, :, :Dh, :], ref[:, :, 2].permute(0, 2, 3, 1)) < 2e-3
    Wp, bp = _rand(D, K, scale=1 / 16, seed=44), _rand(D, seed=45)
    gate = _rand(1, D, seed=46)
    x = torch.randn(B * N, D, device="cuda")
    x0 = x.clone()
    _lib.check(_lib.lib().tpx_linear_gated(A.data_ptr(), K, Wp.data_ptr(), bp.data_ptr(), gate.data_ptr(), D, 1, N, x.data_ptr(), D, B * N, D, K, -128, st()))
    refx = x0 + (gate[0].float() * linear_ref(A, Wp, bp).float()).half().float()
    torch.cuda.synchronize()
    assert rel_l2(x, refx) < 1e-3
