#!/usr/bin/env python
"""Recipe that stages the UNMODIFIED reference modules of the hot path under ``oracle/_ref/`` (TEST INFRASTRUCTURE).

    python oracle/stage_ref.py            # in the build container, where /root/reference exists

``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the staged
files travel to the GPU box with the snapshot, exactly like the built ``.so``.  There they let the ``-m gpu`` parity
tests and ``bench.py --impl reference`` execute the reference's own ``DiT`` / ``create_diffusion`` / ``VAE`` — on the
GPU under ``torch.autocast('cuda', fp16)`` (the contract north_star's tolerance is stated against,
models/dit_crossattn.py:197) and on the host cores in fp32 (the timing baseline).

What is staged (byte-identical copies, sha256 recorded in ``oracle/_ref/MANIFEST.json``):
    models/__init__.py, models/{dit_crossattn,attention,utils,vae3d_dib}.py, models/diffusion/*.py, utils/typing.py
Nothing else of the reference is needed on this path.  The one un-vendored dependency, ``xformers.ops``
(models/attention.py:17; unpinned, README.md:67), is NOT staged or copied from anywhere: ``oracle/refmods.py``
restates its contract with ``torch.nn.functional.scaled_dot_product_attention``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
DEFAULT_SRC = "/root/reference"
FILES = [
    "models/__init__.py", "models/dit_crossattn.py", "models/attention.py", "models/utils.py", "models/vae3d_dib.py",
    "models/diffusion/__init__.py", "models/diffusion/diffusion_utils.py", "models/diffusion/gaussian_diffusion.py",
    "models/diffusion/respace.py", "models/diffusion/timestep_sampler.py", "utils/typing.py",
]


def stage(src: str = DEFAULT_SRC, dest: str = DEST) -> dict:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} not found: the reference only exists in the build container")
    manifest = {"source": src, "files": {}}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest["files"][rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def staged(dest: str = DEST) -> bool:
    return all(os.path.exists(os.path.join(dest, rel)) for rel in FILES)


if __name__ == "__main__":
    m = stage(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_SRC)
    print(f"staged {len(m['files'])} reference files under {DEST}")
