#!/usr/bin/env python
"""Time the PrimSDF point query at the reference's mesh-extraction size: 256^3 grid points against 2048 primitives
(inference.py:108-116 walks them in 8192-point chunks through a dense weight matrix).  Prints ms and points/s; also times
a dense [points x prims] torch formulation of the same query on the GPU for a 64-chunk sample as the stock-code comparison
(restated here; tools do not use oracle/)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tpxl_b200  # noqa: E402


def dense_query(x, srt, feat, S=8, C=6):
    """Covered points only (the bulk of the cost): dense weight matrix, gather of the covering (point, prim) pairs, grid_sample."""
    import torch.nn.functional as F
    local = (x[:, None, :] - srt[None, :, 1:4]) / srt[None, :, 0:1]
    w = F.relu(1 - local.abs().amax(-1))
    w = w / (w.sum(-1, keepdim=True) + 1e-6)
    ib, ip = torch.where(w > 0)
    samp = F.grid_sample(feat[ip].reshape(-1, C, S, S, S), local[ib, ip].reshape(-1, 1, 1, 1, 3), mode="bilinear", padding_mode="zeros",
                         align_corners=True).reshape(-1, C)
    out = torch.zeros(x.shape[0], C, device=x.device)
    out.index_add_(0, ib, samp * w[ib, ip][:, None])
    return out


def main():
    K, S, G = 2048, 8, int(os.environ.get("GRID", 256))
    g = torch.Generator().manual_seed(0)
    # a plausible object: primitives on a noisy sphere shell, scales ~ what the released VAE emits
    d = torch.randn(K, 3, generator=g)
    pos = d / d.norm(dim=1, keepdim=True) * (0.55 + 0.1 * torch.rand(K, 1, generator=g))
    srt = torch.cat([0.03 + 0.05 * torch.rand(K, 1, generator=g), pos], 1)
    m = tpxl_b200.PrimSDF(num_prims=K, dim_feat=6, prim_shape=S).eval()
    m.srt_param.data = srt
    m.feat_param.data = torch.randn(K, 6 * S ** 3, generator=g)
    m = m.cuda()
    lin = torch.linspace(-1, 1, G, device="cuda")
    pts = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3).contiguous()
    for _ in range(2):
        out = m(pts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = m(pts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    cov = float((out["tex"].sum(1) != 0).float().mean())
    print(f"PrimSDF.forward (grid-binned query, grid cached): {pts.shape[0]} points x {K} prims  {ms:.2f} ms  {pts.shape[0] / ms / 1e6:.2f} Gpoints/s  covered {cov:.3f}")
    from tpxl_b200 import _lib
    lib = _lib.lib()
    srt_c, feat_c = m.srt_param.data.float().contiguous(), m.feat_param.data.float().contiguous()
    o2 = torch.empty(pts.shape[0], 6, device="cuda")
    ws = m._grid[2]
    hdr = ws[:40].view(torch.int32).cpu()
    e0.record()
    for _ in range(5):
        lib.tpx_primsdf_grid_build(srt_c.data_ptr(), K, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    print(f"grid build: {e0.elapsed_time(e1) / 5:.3f} ms  (cover entries {int(hdr[5])}, nearest-candidate entries {int(hdr[6])}, overflow {int(hdr[7])})")
    e0.record()
    lib.tpx_primsdf_query(pts.data_ptr(), srt_c.data_ptr(), feat_c.data_ptr(), pts.shape[0], K, S, 6, 1, o2.data_ptr(), _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1)
    same = bool(torch.equal(o2[:, 0:1], out["sdf"]) and torch.equal(o2[:, 1:4], out["tex"]))
    print(f"exhaustive tpx_primsdf_query: {ms2:.2f} ms ({pts.shape[0] * K / ms2 / 1e9:.2f} T box-tests/s); identical output: {same}")
    # stock formulation (dense weight matrix per 8192-point chunk) on the same GPU, 64 chunks
    srt_d, feat_d = m.srt_param.data, m.feat_param.data
    chunks = pts[: 64 * 8192].split(8192)
    dense_query(chunks[0], srt_d, feat_d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in chunks:
        dense_query(c, srt_d, feat_d)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    print(f"dense torch formulation on GPU: {dt / 64:.3f} ms per 8192-point chunk -> {dt / 64 * pts.shape[0] / 8192:.0f} ms for the grid")


if __name__ == "__main__":
    main()
