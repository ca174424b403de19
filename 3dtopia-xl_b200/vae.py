"""Host-side mirror of the reference's VAE interface (decode path), backed by libtpx_b200.

Drop-in for ``models.vae3d_dib.VAE`` (/root/reference/models/vae3d_dib.py:391-453): same constructor kwargs
(configs/inference_dit.yml:31-40), ``load_state_dict(sd['model_state_dict'])`` with the reference key names
(encoder keys are accepted and kept but never uploaded), ``decode(z[P,1,4,4,4]) -> [P,6,8,8,8]`` in z's dtype,
returned as a fresh contiguous tensor (the caller mutates it in place, inference.py:345-346).
``encode`` / ``forward`` are training-time paths that inference never calls: they raise.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .synth import vae_decoder_shapes


class VAE(nn.Module):
    def __init__(self, in_channels=1, latent_channels=16, out_channels=1, down_channels=(16, 32, 64, 128, 256), mid_attention=True,
                 up_channels=(256, 128, 64, 32, 16), layers_per_block=2, skip_scale=None, gradient_checkpointing=False):
        super().__init__()
        up_channels = tuple(up_channels)
        if not (latent_channels == 1 and up_channels == (256, 32) and mid_attention and layers_per_block == 2 and skip_scale is None):
            raise NotImplementedError("the B200 decoder kernel set covers the released configuration (configs/inference_dit.yml:31-40): "
                                      f"latent_channels=1, up_channels=[256,32], mid_attention, layers_per_block=2; got latent={latent_channels}, "
                                      f"up={list(up_channels)}, mid_attention={mid_attention}, layers={layers_per_block}")
        self.in_channels, self.latent_channels, self.out_channels = in_channels, latent_channels, out_channels
        self.up_channels = up_channels
        self._shapes = vae_decoder_shapes(latent_channels=latent_channels, out_channels=out_channels, up_channels=up_channels,
                                          mid_attention=mid_attention, layers_per_block=layers_per_block)
        self._anchor = nn.Parameter(torch.zeros(1), requires_grad=False)
        self._sd: Optional["OrderedDict[str, torch.Tensor]"] = None
        self._extra = OrderedDict()   # encoder.* / quant_conv.* tensors, kept only so state_dict() round-trips
        self._handle = None
        self._handle_device = None
        self._ws = {}

    def state_dict(self, *args, **kwargs):
        if self._sd is None:
            raise _lib.TpxError("VAE has no parameters yet: load_state_dict first (random initialisation of a decoder is not an inference path)")
        out = OrderedDict(self._extra)
        out.update(self._sd)
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        missing = [k for k in self._shapes if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._shapes and not k.startswith(("encoder.", "quant_conv."))]
        errs = [f"size mismatch for {k}: copying a param with shape {tuple(state_dict[k].shape)}, expected {tuple(s)}"
                for k, s in self._shapes.items() if k in state_dict and tuple(state_dict[k].shape) != tuple(s)]
        if strict and (missing or unexpected):
            errs.append(f"Missing key(s): {missing[:5]}; unexpected key(s): {unexpected[:5]}")
        if errs:
            raise RuntimeError("Error(s) in loading state_dict for VAE:\n\t" + "\n\t".join(errs))
        self._sd = OrderedDict((k, state_dict[k].detach()) for k in self._shapes if k in state_dict)
        self._extra = OrderedDict((k, v) for k, v in state_dict.items() if k.startswith(("encoder.", "quant_conv.")))
        if self._handle is not None:
            self._ingest()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def load_checkpoint(self, path: str, key: Optional[str] = "model_state_dict"):
        """``vae.load_state_dict(torch.load(path, map_location='cpu')['model_state_dict'])`` (inference.py:257-258) from a memory-mapped
        file: only decoder.* / post_quant_conv.* tensors are read and uploaded (the encoder half of the checkpoint is never touched)."""
        try:
            ck = torch.load(path, map_location="cpu", mmap=True, weights_only=True)
        except (RuntimeError, ValueError, TypeError):
            ck = torch.load(path, map_location="cpu", weights_only=True)
        sd = ck[key] if key is not None else ck
        return self.load_state_dict(OrderedDict((k, v) for k, v in sd.items() if not k.startswith(("encoder.", "quant_conv."))))

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        dev = self._anchor.device
        if dev.type == "cuda" and (self._handle is None or self._handle_device != dev):
            self._create_handle(dev)
        return out

    def _create_handle(self, dev):
        lib = _lib.lib()
        self._destroy_handle()
        cfg = _lib.VaeConfig(self.latent_channels, self.out_channels, self.up_channels[0], self.up_channels[1], 8)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.tpx_vae_create(C.byref(cfg), C.byref(h)), "tpx_vae_create")
        self._handle, self._handle_device = h, dev
        self._ws.clear()
        if self._sd is not None:
            self._ingest()

    def _ingest(self):
        lib = _lib.lib()
        dev = self._handle_device
        with torch.cuda.device(dev):
            st = _lib.stream_ptr()
            keep = []
            for k, v in self._sd.items():
                t = v.detach()
                if t.dtype not in (torch.float16, torch.float32):
                    t = t.float()
                t = t.to(dev, non_blocking=True).contiguous()
                keep.append(t)
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.tpx_vae_set_weight(self._handle, k.encode(), t.data_ptr(), _lib.dtype_tag(t), shape, t.dim(), st), f"set_weight({k})")
            _lib.check(lib.tpx_vae_finalize(self._handle, st), "tpx_vae_finalize")
            torch.cuda.current_stream().synchronize()

    def _destroy_handle(self):
        if self.__dict__.get("_handle") is not None:
            try:
                _lib.load_library().tpx_vae_destroy(self._handle)
            except Exception:
                pass
            self.__dict__["_handle"] = None      # not nn.Module.__setattr__: this also runs at interpreter shutdown

    def __del__(self):
        self._destroy_handle()

    def decode(self, x: torch.Tensor) -> torch.Tensor:
        """VAE.decode (vae3d_dib.py:437-440)."""
        if self._handle is None:
            raise _lib.TpxError("VAE is not on a CUDA device: call .to('cuda') first (this implementation has no CPU path)")
        if self._sd is None:
            raise _lib.TpxError("VAE.decode before load_state_dict")
        if x.dim() != 5 or tuple(x.shape[1:]) != (self.latent_channels, 4, 4, 4):
            raise ValueError(f"z must be [P,{self.latent_channels},4,4,4], got {tuple(x.shape)}")
        lib = _lib.lib()
        dev = self._handle_device
        z = x.detach().to(dev)
        if z.dtype not in (torch.float16, torch.float32):
            z = z.float()
        z = z.contiguous()
        P = z.shape[0]
        out = torch.empty(P, self.out_channels, 8, 8, 8, dtype=z.dtype, device=dev)
        if P == 0:
            return out
        with torch.cuda.device(dev):
            nbytes = lib.tpx_vae_workspace_bytes(self._handle, P)
            ws = self._ws.get(P)
            if ws is None:
                self._ws.clear()
                ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
                self._ws[P] = ws
            base = (ws.data_ptr() + 255) & ~255
            _lib.check(lib.tpx_vae_decode(self._handle, z.data_ptr(), _lib.dtype_tag(z), out.data_ptr(), _lib.dtype_tag(out), P, base, nbytes,
                                          _lib.stream_ptr()), "tpx_vae_decode")
        return out

    def encode(self, x):
        raise NotImplementedError("VAE.encode is a training-time path (vae3d_dib.py:431-435); the B200 build covers decode only")

    def forward(self, x, sample=True):
        raise NotImplementedError("VAE.forward is a training-time path (vae3d_dib.py:442-453); the B200 build covers decode only")
