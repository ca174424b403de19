#!/usr/bin/env python
"""Golden vector for the ray-march oracle, produced by RUNNING THE REFERENCE'S OWN pure-PyTorch ray-marcher.

Run in the build container only (needs /root/reference):

    python tests/golden/make_raymarch_golden.py

The reference's CUDA ray-marcher (dva/mvp/extensions/mvpraymarch, an sm_70 torch extension) cannot be built or run here.  Its
gradcheck script, however, carries a plain PyTorch implementation of the same forward pass as the arm the CUDA kernel is checked
against (mvpraymarch.py, the block under "# python raymarching implementation" inside ``gradcheck``: every ray visits every
primitive at every step, ``grid_sample`` trilinear lookup, box test, fade, additive alpha with saturation).  That block is not a
function and the module imports the compiled extension, so it cannot be imported; this script reads the block's lines from the
reference file AT RUN TIME (nothing of it is stored in this repository), drops the device moves / timers, and executes it on the CPU
over a small seeded scene.  Inputs and the resulting image go to tests/golden/raymarch_ref.npz; tests/test_oracle_golden.py holds
``oracle.raymarch.raymarch_dense`` to it.
"""
import os
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_FILE = "/root/reference/dva/mvp/extensions/mvpraymarch/mvpraymarch.py"
OUT = os.path.join(ROOT, "tests", "golden", "raymarch_ref.npz")


def reference_block() -> str:
    lines = open(REF_FILE).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "# python raymarching implementation" in l)
    end = next(i for i in range(start, len(lines)) if lines[i].strip().startswith("print(rayrgba["))
    keep = []
    for l in lines[start:end]:
        s = l.strip()
        if s.startswith("torch.cuda.synchronize") or s.startswith("time0 ="):
            continue
        keep.append(l.replace('.to("cuda")', ""))
    return textwrap.dedent("\n".join(keep))


def scene(seed=7, N=1, H=12, W=12, K=6, M=8):
    g = torch.Generator().manual_seed(seed)
    focal, princ = W * 2.0, (W * 0.5, H * 0.5)
    py, px = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    d = torch.stack([(px - princ[0]) / focal, (py - princ[1]) / focal, torch.ones_like(px)], -1)
    raydir = (d / d.norm(dim=-1, keepdim=True))[None].repeat(N, 1, 1, 1).contiguous()
    raypos = torch.tensor([0.05, -0.03, -4.0])[None, None, None, :].repeat(N, H, W, 1).contiguous()
    tminmax = torch.stack([2.6 + 0.3 * torch.rand(N, H, W, generator=g), 5.2 + 0.4 * torch.rand(N, H, W, generator=g)], -1)
    template = F.softplus(torch.randn(N, K, 4, M, M, M, generator=g) * 1.2)          # channels-first, as the reference block samples it
    primpos = torch.randn(N, K, 3, generator=g) * 0.45
    q, _ = torch.linalg.qr(torch.randn(N, K, 3, 3, generator=g))
    primrot = q.contiguous()
    primscale = 1.4 + 0.8 * torch.rand(N, K, 3, generator=g)
    return dict(N=N, H=H, W=W, K=K, raypos=raypos, raydir=raydir, tminmax=tminmax, stepsize=0.21, template=template, primpos=primpos, primrot=primrot,
                primscale=primscale, fadescale=8.0, fadeexp=8.0)


def main():
    sc = scene()
    ns = dict(sc)
    ns.update(torch=torch, F=F, dowarp=False, accum=0, warp=None)
    exec(compile(reference_block(), REF_FILE + ":<python raymarching block>", "exec"), ns)
    out = ns["rayrgba"].detach()
    assert out.shape == (sc["N"], sc["H"], sc["W"], 4)
    cov = float((out[..., 3] > 0).float().mean())
    sat = float((out[..., 3] >= 1.0 - 1e-6).float().mean())
    print(f"reference torch ray-marcher: steps {ns['step']}, covered pixels {cov:.2f}, saturated {sat:.2f}, alpha max {float(out[..., 3].max()):.3f}")
    assert cov > 0.3, "scene too empty to pin anything"
    np.savez_compressed(
        OUT,
        raypos=sc["raypos"].numpy(), raydir=sc["raydir"].numpy(), tminmax=sc["tminmax"].numpy(), stepsize=np.float32(sc["stepsize"]),
        template_chlast=sc["template"].permute(0, 1, 3, 4, 5, 2).contiguous().numpy(),   # [N,K,D,H,W,4]: the layout the oracle / the kernel take
        primpos=sc["primpos"].numpy(), primrot=sc["primrot"].numpy(), primscale=sc["primscale"].numpy(),
        fadescale=np.float32(sc["fadescale"]), fadeexp=np.float32(sc["fadeexp"]), rayrgba=out.numpy())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
