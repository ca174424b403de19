#!/usr/bin/env python
"""Golden vectors for the sample -> decode glue, produced by RUNNING THE REFERENCE'S OWN STATEMENTS.

Run in the build container only (needs /root/reference):

    python tests/golden/make_inference_glue_golden.py

``inference.py`` cannot be imported here (rembg, nvdiffrast, ... at module level), and the statements in question are not a function:
they are the body of the preview branch of its sampling loop (latent de-normalisation, 0:4 | 4:68 slicing, per-sample ``vae.decode``,
inverse feature normalisation, channel-major packing, concat).  This script reads exactly that statement block from the reference file
AT RUN TIME (nothing of it is stored in this repository), and executes it on the CPU with a deterministic, index-revealing stand-in
for ``vae.decode`` — for both settings of ``perchannel_norm`` and two values of ``latent_nf``.  Inputs and the resulting
``recon_param`` go to tests/golden/inference_glue.npz; tests/test_oracle_golden.py holds ``oracle.vae.inference_glue`` to them bit for bit,
and tests/test_gpu_pipeline.py holds the CUDA kernels (tpx_latent_split / tpx_primvolume_pack) to that oracle function.
"""
import os
import textwrap
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_FILE = "/root/reference/inference.py"
OUT = os.path.join(ROOT, "tests", "golden", "inference_glue.npz")


def reference_block() -> str:
    lines = open(REF_FILE).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip().startswith('recon_param = samples["sample"].reshape('))
    end = next(i for i in range(start, len(lines)) if lines[i].strip().startswith("recon_param = torch.concat([recon_srt_param, recon_feat_param]"))
    return textwrap.dedent("\n".join(lines[start:end + 1]))


def fake_decode(z):
    """index-revealing stand-in for vae.decode: [n,1,4,4,4] -> [n,6,8,8,8] (the same formula as tests/test_gpu_pipeline.py)"""
    n = z.shape[0]
    base = torch.arange(n * 6 * 512, dtype=torch.float32).reshape(n, 6, 8, 8, 8)
    return base * 1e-3 - 3.0 + z.reshape(n, -1).sum(1).reshape(n, 1, 1, 1, 1)


def main():
    block = compile(reference_block(), REF_FILE + ":<preview glue block>", "exec")
    g = torch.Generator().manual_seed(5)
    bs, prims = 2, 4
    sample = torch.randn(bs, prims * 68, generator=g) * 1.7
    mean = torch.randn(1, 1, 68, generator=g) * 0.3
    std = torch.rand(1, 1, 68, generator=g) + 0.5
    out = {"sample": sample.numpy(), "latent_mean": mean.numpy(), "latent_std": std.numpy()}
    for perchannel in (True, False):
        for nf in (1.0, 0.7):
            ns = dict(torch=torch, samples={"sample": sample.clone()}, inf_bs=bs, perchannel_norm=perchannel, latent_std=std, latent_mean=mean,
                      latent=torch.empty(1, prims, 1, 4, 4, 4), vae=types.SimpleNamespace(decode=fake_decode),
                      config=types.SimpleNamespace(model=types.SimpleNamespace(num_prims=prims, latent_nf=nf)))
            exec(block, ns)
            rp = ns["recon_param"]
            assert rp.shape == (bs, prims, 4 + 6 * 512)
            out[f"recon_param_pc{int(perchannel)}_nf{nf}"] = rp.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", [k for k in out if k.startswith("recon")])


if __name__ == "__main__":
    main()
