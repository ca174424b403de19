"""-m gpu: ray-march preview (tpx_raymarch_preview / tpxl_b200.RayMarcher) against the oracle's kernel-shaped restatement of
compute_raydirs + mvpraymarch("fixedorder").  Tolerance: fp32 throughout; the kernel uses the fast-math exp / pow the reference is
compiled with and fused multiply-adds, the oracle plain torch math, and a grazing ray can enter a box one step earlier or later:
relative L2 over the image <= 2e-3, and no pixel off by more than 2 % of the value range."""
import pytest
import torch

import oracle
import tpxl_b200

pytestmark = pytest.mark.gpu


def _scene(K, S, seed, rot=False):
    g = torch.Generator().manual_seed(seed)
    pos = (torch.rand(1, K, 3, generator=g) - 0.5) * 1.0
    scale = 1.0 / (0.12 + 0.1 * torch.rand(1, K, 1, generator=g)).repeat(1, 1, 3)          # primscale = inverse half-size
    if rot:
        q = torch.randn(1, K, 4, generator=g)
        q = q / q.norm(dim=-1, keepdim=True)
        w, x, y, z = q.unbind(-1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(1, K, 3, 3)
    else:
        R = torch.eye(3)[None, None].repeat(1, K, 1, 1)
    rgba = torch.rand(1, K, 4, S, S, S, generator=g)
    rgba[:, :, :3] *= 255.0
    rgba[:, :, 3] *= 60.0
    return rgba, pos, R, scale


def _camera(H, W, volradius):
    RT = torch.tensor([[[1.0, 0, 0, 0], [0, -1.0, 0, 0], [0, 0, -1.0, 3.0 * volradius]]])
    Kc = torch.tensor([[[2.4 * W, 0, W / 2], [0, 2.4 * W, H / 2], [0, 0, 1.0]]])
    return Kc, RT


@pytest.mark.parametrize("K,S,H,W,rot", [(24, 4, 32, 24, False), (40, 8, 33, 21, True), (600, 4, 16, 16, False)])
def test_preview_matches_oracle(K, S, H, W, rot):
    """Small scenes (coarse step so the oracle's python loop is quick); ragged image sizes exercise the border rays; K = 600 overlapping boxes
    push a warp's hit list past the reference's cap of 512."""
    volradius, dt = 100.0, 1.0
    rgba, pos, R, scale = _scene(K, S, 5 * K + H, rot)
    if K == 600:
        pos = pos * 0.2
    Kc, RT = _camera(H, W, volradius)
    rm = tpxl_b200.RayMarcher(H, W, volradius, dt=dt).cuda()
    out = rm(rgba.cuda(), (pos * volradius).cuda(), R.cuda(), scale.cuda(), Kc.cuda(), RT.cuda())["rgba_image"].cpu()
    ref = oracle.raymarch.ray_marcher_forward(rgba, pos * volradius, R, scale, Kc, RT, H, W, volradius, dt)
    assert out.shape == (1, 4, H, W)
    rel = float((out - ref).norm() / ref.norm())
    worst = float(((out - ref).abs() / torch.tensor([255.0, 255.0, 255.0, 1.0]).view(1, 4, 1, 1)).max())
    print(f"K={K} S={S} {H}x{W}: rel-L2 {rel:.2e}, worst pixel {worst:.2e}, coverage {float((ref[:, 3] > 0).float().mean()):.2f}, max alpha {float(ref[:, 3].max()):.2f}")
    assert float(ref[:, 3].max()) > 0.5 and float((ref[:, 3] > 0).float().mean()) > 0.05    # the scene is actually rendered
    assert rel < 2e-3 and worst < 2e-2


def test_preview_full_size_properties():
    """The shipped size (2048 primitives of 8^3, 518 x 518, volradius 10000, dt 1 as configs/inference_dit.yml): size-independent
    properties — empty space stays exactly zero, alpha never exceeds 1, an opaque primitive saturates, the image is deterministic."""
    K, S, H, W, volradius = 2048, 8, 518, 518, 10000.0
    g = torch.Generator().manual_seed(3)
    d = torch.randn(1, K, 3, generator=g)
    pos = d / d.norm(dim=-1, keepdim=True) * 0.5
    scale = torch.full((1, K, 3), 1.0 / 0.05)
    rgba = torch.rand(1, K, 4, S, S, S, generator=g) * 255.0
    rgba[:, :, 3] = 255.0 * torch.exp(-(torch.rand(1, K, S, S, S, generator=g) * 0.02 / 0.005) ** 2)
    R = torch.eye(3)[None, None].repeat(1, K, 1, 1)
    RT = torch.tensor([[[1.0, 0, 0, 0], [0, -1.0, 0, 0], [0, 0, -1.0, 5 * volradius]]])
    Kc = torch.tensor([[[2084.9526697685183 * H / 1024, 0, 512.0 * H / 1024], [0, 2084.9526697685183 * W / 1024, 512.0 * W / 1024], [0, 0, 1.0]]])
    rm = tpxl_b200.RayMarcher(H, W, volradius).cuda()
    a = rm(rgba.cuda(), (pos * volradius).cuda(), R.cuda(), scale.cuda(), Kc.cuda(), RT.cuda())["rgba_image"]
    b = rm(rgba.cuda(), (pos * volradius).cuda(), R.cuda(), scale.cuda(), Kc.cuda(), RT.cuda())["rgba_image"]
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    alpha = a[0, 3]
    assert float(alpha.max()) <= 1.0 + 1e-5 and float(alpha.min()) >= 0.0
    assert float(alpha[:8, :8].abs().max()) == 0.0                       # the image corner looks past the object
    assert float((alpha > 0.99).float().mean()) > 0.05                   # the opaque shell saturates
    assert float(a[0, :3].max()) <= 255.0 * 1.0001
