"""Oracle restatement of the image conditioner's encoder (TEST INFRASTRUCTURE — see oracle/__init__.py): SURVEY §8f-2, the next
row after the PrimSDF query.  No CUDA path exists for it yet; this module and its fixture are the parity anchor for one.

Restates, for the released configuration (configs/inference_dit.yml:48-51: ``dinov2_vitb14_reg``, frozen, no modulation):
  * Dinov2Wrapper.forward            models/conditioner/image_dinov2.py:44-61   ([N,H,W,3] in 0..255 -> /255 -> Resize(518, bicubic)
                                                                                 -> Normalize -> ViT -> cat[cls, patch tokens])
  * DinoVisionTransformer            models/conditioner/dinov2/models/vision_transformer.py:208-265 (prepare_tokens_with_masks,
                                      forward_features; 4 register tokens inserted after the class token, LayerNorm eps 1e-6)
  * Block / Attention / Mlp / LayerScale   dinov2/layers/block.py:87-93, attention.py:56-69, mlp.py:33-39, layer_scale.py:26-27
                                      (pre-norm; 12 heads x 64; GELU (erf); per-channel LayerScale on both branches)
  * PatchEmbed                       dinov2/layers/patch_embed.py:66-80 (14x14 stride-14 convolution, no norm)
Square 518 x 518 inputs only (what inference.py feeds, configs/inference_dit.yml:18-19): then the position embedding is used
as stored (vision_transformer.py:191-192) and the Resize is the identity.  Other square sizes go through the same bicubic
antialiased resize torchvision applies to tensors.  The arithmetic is fp32 like the reference (it runs outside autocast,
inference.py:317).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

IMG, PATCH, DIM, DEPTH, HEADS, N_REG = 518, 14, 768, 12, 12, 4
MEAN = (0.48145466, 0.4578275, 0.40821073)          # image_dinov2.py:21 (the CLIP statistics, as the reference uses them)
STD = (0.26862954, 0.26130258, 0.27577711)


def shapes(depth: int = DEPTH, dim: int = DIM) -> "OrderedDict[str, Tuple[int, ...]]":
    """state_dict keys / shapes of Dinov2Wrapper('dinov2_vitb14_reg') in the reference's order."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    m = "model."
    s[m + "cls_token"] = (1, 1, dim)
    s[m + "pos_embed"] = (1, (IMG // PATCH) ** 2 + 1, dim)
    s[m + "register_tokens"] = (1, N_REG, dim)
    s[m + "patch_embed.proj.weight"], s[m + "patch_embed.proj.bias"] = (dim, 3, PATCH, PATCH), (dim,)
    for i in range(depth):
        b = f"{m}blocks.{i}."
        s[b + "norm1.weight"], s[b + "norm1.bias"] = (dim,), (dim,)
        s[b + "attn.qkv.weight"], s[b + "attn.qkv.bias"] = (3 * dim, dim), (3 * dim,)
        s[b + "attn.proj.weight"], s[b + "attn.proj.bias"] = (dim, dim), (dim,)
        s[b + "ls1.gamma"] = (dim,)
        s[b + "norm2.weight"], s[b + "norm2.bias"] = (dim,), (dim,)
        s[b + "mlp.fc1.weight"], s[b + "mlp.fc1.bias"] = (4 * dim, dim), (4 * dim,)
        s[b + "mlp.fc2.weight"], s[b + "mlp.fc2.bias"] = (dim, 4 * dim), (dim,)
        s[b + "ls2.gamma"] = (dim,)
    s[m + "norm.weight"], s[m + "norm.bias"] = (dim,), (dim,)
    return s


def synth_weights(seed: int, depth: int = DEPTH, dim: int = DIM) -> "OrderedDict[str, torch.Tensor]":
    """Random weights with the magnitudes of a trained ViT (LayerScale ~0.3, LayerNorm gains ~1, Glorot-scaled matrices), from
    one numpy RandomState so the fixture generator and the tests rebuild the same tensors."""
    rs = np.random.RandomState(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shp in shapes(depth, dim).items():
        a = rs.standard_normal(size=shp).astype(np.float32)
        if name.endswith("gamma"):
            a = 0.3 + 0.1 * a
        elif "norm" in name and name.endswith("weight"):
            a = 1.0 + 0.1 * a
        elif name.endswith("bias"):
            a = 0.02 * a
        elif len(shp) >= 2 and "token" not in name and "pos_embed" not in name:
            fan_in, fan_out = int(np.prod(shp[1:])), shp[0]
            a = a * np.float32(np.sqrt(2.0 / (fan_in + fan_out)))
        else:
            a = 0.2 * a                                   # class / register tokens, position embedding
        sd[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def preprocess(image: torch.Tensor) -> torch.Tensor:
    """[N,H,W,3] in 0..255 -> normalised [N,3,518,518]   (image_dinov2.py:48-50 with the Compose of :19-22)."""
    assert image.shape[-1] == 3 and image.shape[1] == image.shape[2], "square RGB images"
    x = image.permute(0, 3, 1, 2) / 255.0
    if x.shape[-1] != IMG:
        x = F.interpolate(x, size=(IMG, IMG), mode="bicubic", align_corners=False, antialias=True)   # torchvision Resize on tensors
    mean = torch.tensor(MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


def block(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, heads: int = HEADS) -> torch.Tensor:
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])                    # softmax(q k^T / sqrt(64)) v
    a = F.linear(a.transpose(1, 2).reshape(B, N, C), sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x = x + a * sd[p + "ls1.gamma"]
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + h * sd[p + "ls2.gamma"]


def forward(sd: Dict[str, torch.Tensor], image: torch.Tensor, depth: int = DEPTH, return_blocks: bool = False):
    """Dinov2Wrapper.forward: [N,518,518,3] -> [N, 1 + 37*37, 768] = cat[class token, patch tokens] after the final LayerNorm."""
    m = "model."
    x = preprocess(image.float())
    x = F.conv2d(x, sd[m + "patch_embed.proj.weight"], sd[m + "patch_embed.proj.bias"], stride=PATCH).flatten(2).transpose(1, 2)
    x = torch.cat((sd[m + "cls_token"].expand(x.shape[0], -1, -1), x), dim=1) + sd[m + "pos_embed"]
    x = torch.cat((x[:, :1], sd[m + "register_tokens"].expand(x.shape[0], -1, -1), x[:, 1:]), dim=1)
    outs = []
    for i in range(depth):
        x = block(sd, f"{m}blocks.{i}.", x)
        if return_blocks:
            outs.append(x)
    x = F.layer_norm(x, (x.shape[-1],), sd[m + "norm.weight"], sd[m + "norm.bias"], 1e-6)
    ret = torch.cat((x[:, :1], x[:, 1 + N_REG:]), dim=1)
    return (ret, outs) if return_blocks else ret
