"""Time the ray-march preview at the shipped size (2048 primitives of 8^3 on a shell, 518 x 518, volradius 10000, dt 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpxl_b200

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
args = [t.cuda() for t in (rgba, pos * volradius, R, scale, Kc, RT)]
for _ in range(2): out = rm(*args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): out = rm(*args)
b.record(); torch.cuda.synchronize()
al = out["rgba_image"][0, 3]
print(f"ray-march preview {H}x{W}, {K} prims: {a.elapsed_time(b)/3:.2f} ms  (covered pixels {float((al > 0).float().mean()):.2f}, saturated {float((al > 0.99).float().mean()):.2f})")
