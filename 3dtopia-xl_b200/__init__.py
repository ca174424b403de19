"""B200-native denoising hot path for 3DTopia-XL: DiT step (CFG x2) + DDIM/DDPM update + VAE decode.

Python host layer = the reference's own interfaces (``DiT``, ``create_diffusion``, ``VAE``) calling the
sm_100a kernels in ``lib/libtpx_b200.so`` through the C ABI of ``include/tpx.h``.  See DESIGN.md and
INTEGRATION.md.  There is no CPU or other-architecture fallback anywhere in this package.
"""
import sys
import types

from . import synth  # noqa: F401
from . import _lib  # noqa: F401
from . import diffusion, dinov2, dit, pipeline, primsdf, ray_marcher, shard, vae  # noqa: F401
from .diffusion import SpacedDiffusion, create_diffusion  # noqa: F401
from .dit import DiT  # noqa: F401
from .pipeline import LatentCodec, PrimXPipeline  # noqa: F401
from .primsdf import PrimSDF  # noqa: F401
from .ray_marcher import RayMarcher  # noqa: F401
from .vae import VAE  # noqa: F401

__version__ = "0.1.0"


def install() -> None:
    """Make the reference's import paths resolve to this implementation, so ``inference.py`` / ``app.py`` run
    unchanged:  ``models.dit_crossattn`` -> dit, ``models.vae3d_dib`` -> vae, ``models.diffusion`` -> diffusion,
    ``models.primsdf`` -> primsdf, ``models.conditioner.image_dinov2`` -> dinov2, ``dva.ray_marcher`` -> ray_marcher.
    (Equivalent to editing ``class_name`` in configs/inference_dit.yml:32,53; see INTEGRATION.md.)  If the reference's
    ``models`` package is importable it is imported first so its other members (conditioner, primsdf) keep working."""
    try:
        import models as ref_models  # the reference checkout, when it is on sys.path
    except Exception:
        ref_models = types.ModuleType("models")
        ref_models.__path__ = []
        sys.modules["models"] = ref_models
    for name, mod in (("dit_crossattn", dit), ("vae3d_dib", vae), ("diffusion", diffusion), ("primsdf", primsdf)):
        sys.modules["models." + name] = mod
        setattr(ref_models, name, mod)
    # the image conditioner's encoder (configs/inference_dit.yml:49, ``models.conditioner.image_dinov2.Dinov2Wrapper``): the mirror runs
    # on the tcgen05 GEMM / attention kernels with fp16 tensor-core inputs (9e-4 relative L2 from the reference's fp32 encoder)
    sys.modules["models.conditioner.image_dinov2"] = dinov2
    try:
        import models.conditioner as ref_cond
        setattr(ref_cond, "image_dinov2", dinov2)
    except Exception:
        pass
    # the preview renderer: ``from dva.ray_marcher import RayMarcher`` (inference.py:12) resolves here, so the reference's sm_70
    # ray-march extensions (dva/mvp/extensions/*) need not be built at all
    sys.modules["dva.ray_marcher"] = ray_marcher
    try:
        import dva as ref_dva
        setattr(ref_dva, "ray_marcher", ray_marcher)
    except Exception:
        pass
