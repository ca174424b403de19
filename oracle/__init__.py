"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU (device-agnostic torch/numpy) restatement of the 3DTopia-XL denoising hot path:
DiT.forward / forward_with_cfg, the DDIM/DDPM sampler and VAE.decode.  Every function cites the
reference file:line it restates.  It is the *checker*, never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
    ``bench.py`` may import it;
  * the shipped package (``3dtopia-xl_b200/``) never imports it and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against fixtures produced by importing the reference's own Python modules in the
build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; the attention core of
the un-vendored, unpinned ``xformers`` dependency is restated as softmax(QK^T * Dh^-1/2) V in fp32).
"""
from . import dit, diffusion, vae, primsdf, dinov2, raymarch  # noqa: F401
