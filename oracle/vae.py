"""Oracle restatement of the 3D-conv VAE decoder (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates, over a reference-keyed ``state_dict`` (decoder.* and post_quant_conv.* keys):
  * VAE.decode              /root/reference/models/vae3d_dib.py:437-440
  * Decoder.forward         vae3d_dib.py:369-387
  * MidBlock._forward       vae3d_dib.py:220-226
  * UpBlock._forward        vae3d_dib.py:259-267
  * ResnetBlock.forward     vae3d_dib.py:128-145   ((x + shortcut(res)) * skip_scale, skip_scale = sqrt(0.5) :338)
  * VolumeAttention.forward vae3d_dib.py:34-48     (GroupNorm -> tokens -> MemEffAttention(8 heads, no qkv bias) -> (x+res)*skip)

Precision policies: ``fp32`` (as the reference invokes it, inference.py:339) and ``fp16`` (the
reference module under torch.autocast(fp16): convs / linears / attention round to fp16, GroupNorm and
SiLU run in fp32 on the fp16-valued input).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .dit import Policy, attention_core

Tensor = torch.Tensor
SKIP = math.sqrt(0.5)


def _conv(pol: Policy, x, w, b, **kw):
    if not pol.amp:
        return F.conv3d(x, w, b, **kw)
    return pol.rnd(F.conv3d(pol.rnd(x), pol.rnd(w), None if b is None else pol.rnd(b), **kw))


def _convT(pol: Policy, x, w, b, **kw):
    if not pol.amp:
        return F.conv_transpose3d(x, w, b, **kw)
    return pol.rnd(F.conv_transpose3d(pol.rnd(x), pol.rnd(w), None if b is None else pol.rnd(b), **kw))


def _gn(sd, pre: str, x: Tensor) -> Tensor:
    C = x.shape[1]
    return F.group_norm(x, min(32, C), sd[pre + "weight"], sd[pre + "bias"], eps=1e-5)


def resnet_block(sd, pre: str, x: Tensor, pol: Policy) -> Tensor:
    res = x
    h = F.silu(_gn(sd, pre + "norm1.", x))
    h = _conv(pol, h, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    h = F.silu(_gn(sd, pre + "norm2.", h))
    h = _conv(pol, h, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    if pre + "shortcut.weight" in sd:
        res = _conv(pol, res, sd[pre + "shortcut.weight"], sd[pre + "shortcut.bias"])
    return pol.rnd((h + res) * SKIP)


def volume_attention(sd, pre: str, x: Tensor, pol: Policy, heads: int = 8) -> Tensor:
    B, C, H, W, D = x.shape
    res = x
    h = _gn(sd, pre + "norm.", x)
    h = h.permute(0, 2, 3, 4, 1).reshape(B, -1, C)
    qkv = pol.linear(h, sd[pre + "attn.qkv.weight"], sd.get(pre + "attn.qkv.bias")).reshape(B, -1, 3, heads, C // heads)
    q, k, v = torch.unbind(qkv, 2)
    o = attention_core(q, k, v, pol).reshape(B, -1, C)
    o = pol.linear(o, sd[pre + "attn.proj.weight"], sd.get(pre + "attn.proj.bias"))
    o = o.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)
    return pol.rnd((o + res) * SKIP)


def decode(sd: Dict[str, Tensor], z: Tensor, precision: str = "fp32", layers_per_block: int = 2,
           n_up: int = 2, stages: dict | None = None) -> Tensor:
    """VAE.decode: z [P,1,4,4,4] -> [P,6,8,8,8].  ``stages`` (optional dict) receives intermediates."""
    pol = Policy(precision)
    sd = {k: v.float() for k, v in sd.items()}
    x = _conv(pol, z.float(), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    d = "decoder."
    x = _conv(pol, x, sd[d + "conv_in.weight"], sd[d + "conv_in.bias"], padding=1)
    if stages is not None:
        stages["conv_in"] = x
    x = resnet_block(sd, d + "mid_block.nets.0.", x, pol)
    if d + "mid_block.attns.0.norm.weight" in sd:
        x = volume_attention(sd, d + "mid_block.attns.0.", x, pol)
    if stages is not None:
        stages["mid_attn"] = x
    x = resnet_block(sd, d + "mid_block.nets.1.", x, pol)
    if stages is not None:
        stages["mid"] = x
    for u in range(n_up):
        for j in range(layers_per_block):
            x = resnet_block(sd, f"{d}up_blocks.{u}.nets.{j}.", x, pol)
        if f"{d}up_blocks.{u}.upsample.weight" in sd:
            x = _convT(pol, x, sd[f"{d}up_blocks.{u}.upsample.weight"], sd[f"{d}up_blocks.{u}.upsample.bias"], stride=2)
        if stages is not None:
            stages[f"up{u}"] = x
    x = F.silu(_gn(sd, d + "norm_out.", x))
    x = _convT(pol, x, sd[d + "conv_out.weight"], sd[d + "conv_out.bias"], stride=1, padding=1)
    return x


def denormalise_latents(sample: Tensor, latent_mean: Tensor, latent_std: Tensor, latent_nf: float = 1.0):
    """inference.py:328-332 — returns (srt [B,N,4], feat [B,N,64]); slicing 0:4 | 4:68 is the index contract."""
    p = sample / latent_nf * latent_std + latent_mean
    return p[:, :, 0:4], p[:, :, 4:]


def pack_decoded(decoded: Tensor, bs: int, num_prims: int) -> Tensor:
    """inference.py:345-347 — ch0 /= 5, ch1: -> (x+1)/2, then channel-major flatten [bs, prims, 6*512]."""
    d = decoded.clone()
    d[:, 0:1] /= 5.0
    d[:, 1:] = (d[:, 1:] + 1) / 2.0
    return d.reshape(bs, num_prims, -1)


def inference_glue(sample: Tensor, decode, latent_mean: Tensor, latent_std: Tensor, latent_nf: float, perchannel_norm: bool) -> Tensor:
    """inference.py:328-348, statement by statement (variable names kept), with `decode` standing in for vae.decode:
    latent de-normalisation, 0:4 | 4:68 slicing, per-sample decode, inverse feature normalisation (sdf / 5, (rgb, mat + 1) / 2 — and, without
    per-channel statistics, srt scale / 10 + 0.05 and the decoder input / latent_nf), channel-major packing, concat -> [bs, prims, 4 + 6*512].
    Pinned bit for bit to the reference's own statements executed by tests/golden/make_inference_glue_golden.py (inference_glue.npz)."""
    inf_bs, num_prims = sample.shape[0], sample.shape[1]
    latent = torch.empty(1, num_prims, 1, 4, 4, 4)
    recon_param = sample.reshape(inf_bs, num_prims, -1)
    if perchannel_norm:
        recon_param = recon_param / latent_nf * latent_std + latent_mean
    recon_srt_param = recon_param[:, :, 0:4]
    recon_feat_param = recon_param[:, :, 4:]
    recon_feat_param_list = []
    for inf_bidx in range(inf_bs):
        if not perchannel_norm:
            decoded = decode(recon_feat_param[inf_bidx, ...].reshape(1 * num_prims, *latent.shape[-4:]) / latent_nf)
        else:
            decoded = decode(recon_feat_param[inf_bidx, ...].reshape(1 * num_prims, *latent.shape[-4:]))
        recon_feat_param_list.append(decoded.detach())
    recon_feat_param = torch.concat(recon_feat_param_list, dim=0)
    if not perchannel_norm:
        recon_srt_param[:, :, 0:1] = (recon_srt_param[:, :, 0:1] / 10) + 0.05
    recon_feat_param[:, 0:1, ...] /= 5.
    recon_feat_param[:, 1:, ...] = (recon_feat_param[:, 1:, ...] + 1) / 2.
    recon_feat_param = recon_feat_param.reshape(inf_bs, num_prims, -1)
    return torch.concat([recon_srt_param, recon_feat_param], dim=-1)

