"""B200-native denoising hot path for 3DTopia-XL (DiT + DDIM/CFG + VAE decode).  See DESIGN.md."""
from . import synth  # noqa: F401

__version__ = "0.1.0"
