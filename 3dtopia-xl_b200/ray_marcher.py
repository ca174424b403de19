"""Host-side mirror of the reference's preview renderer ``dva.ray_marcher.RayMarcher``, backed by libtpx_b200.

Same constructor kwargs (``RayMarcher(config.image_height, config.image_width, **config.rm)``, inference.py:282-286) and the same
``forward(prim_rgba, prim_pos, prim_rot, prim_scale, K, RT) -> {"rgba_image": [B,4,H,W], "pixel_coords": ...}``, so
``dva.visualize.visualize_primvolume`` / ``visualize_video_primvolume`` run unchanged on top of it.  The reference's forward is
``compute_raydirs`` + ``mvpraymarch`` from an sm_70 torch extension that has to be compiled and imported at module import time; here
one sm_100a kernel (csrc/raymarch.cu) does both, so the extension is not needed at all.
Inference path only: no gradients, ``ray_subsample_factor == 1``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


def convert_camera_parameters(Rt: torch.Tensor, K: torch.Tensor):
    """dva/ray_marcher.py:24-33."""
    R = Rt[:, :3, :3]
    t = -R.permute(0, 2, 1).bmm(Rt[:, :3, 3].unsqueeze(2)).squeeze(2)
    return dict(campos=t, camrot=R, focal=K[:, :2, :2], princpt=K[:, :2, 2])


class RayMarcher(nn.Module):
    def __init__(self, image_height, image_width, volradius, fadescale=8.0, fadeexp=8.0, dt=1.0, ray_subsample_factor=1, accum=2, termthresh=0.99,
                 blocksize=None, with_t_img=True, chlast=False, assets=None):
        super().__init__()
        self.image_height, self.image_width = image_height, image_width
        self.volradius, self.dt = volradius, dt
        self.fadescale, self.fadeexp = fadescale, fadeexp
        self.blocksize = (8, 16) if blocksize is None else blocksize
        self.with_t_img, self.chlast, self.accum, self.termthresh = with_t_img, chlast, accum, termthresh
        self.ray_subsample_factor = ray_subsample_factor
        self.register_buffer("base_pixel_coords", self._pix(image_height, image_width), persistent=False)

    @staticmethod
    def _pix(h, w, device=None):
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device), torch.arange(w, dtype=torch.float32, device=device), indexing="ij")
        return torch.stack([xs, ys], dim=-1)

    def resize(self, h: int, w: int):
        self.image_height, self.image_width = h, w
        self.base_pixel_coords = self._pix(h, w, self.base_pixel_coords.device)

    @torch.no_grad()
    def forward(self, prim_rgba, prim_pos, prim_rot, prim_scale, K, RT, ray_subsample_factor=None):
        if (self.ray_subsample_factor if ray_subsample_factor is None else ray_subsample_factor) != 1:
            raise NotImplementedError("the B200 preview marches every pixel (ray_subsample_factor == 1, as inference.py / app.py use it)")
        if tuple(self.blocksize) != (8, 16):
            raise NotImplementedError("blocksize (8, 16) only: the per-warp hit list of 8 x 4 pixels is part of the reference's result")
        if not prim_rgba.is_cuda:
            raise _lib.TpxError("RayMarcher runs on CUDA only (no CPU path)")
        B, Kp, C, S = prim_rgba.shape[0], prim_rgba.shape[1], prim_rgba.shape[2], prim_rgba.shape[-1]
        if C != 4 or prim_rgba.shape[3] != S or prim_rgba.shape[4] != S:
            raise ValueError(f"prim_rgba must be [B, K, 4, S, S, S], got {tuple(prim_rgba.shape)}")
        dev = prim_rgba.device
        cam = convert_camera_parameters(RT.to(dev).float(), K.to(dev).float())
        f32 = lambda t: t.to(dev).float().contiguous()  # noqa: E731
        tpl = prim_rgba.float().permute(0, 1, 3, 4, 5, 2).contiguous()
        pos = f32(prim_pos / self.volradius)
        rot, scale = f32(prim_rot), f32(prim_scale)
        campos, camrot = f32(cam["campos"]), f32(cam["camrot"])
        focal = f32(torch.diagonal(cam["focal"], dim1=1, dim2=2))
        princpt = f32(cam["princpt"])
        H, W = self.image_height, self.image_width
        out = torch.empty(B, H, W, 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tpx_raymarch_preview(tpl.data_ptr(), pos.data_ptr(), rot.data_ptr(), scale.data_ptr(), campos.data_ptr(), camrot.data_ptr(),
                                                       focal.data_ptr(), princpt.data_ptr(), B, Kp, S, H, W, float(self.volradius),
                                                       float(self.dt / self.volradius), float(self.fadescale), float(self.fadeexp), out.data_ptr(),
                                                       _lib.stream_ptr()), "tpx_raymarch_preview")
        return {"rgba_image": out.permute(0, 3, 1, 2), "pixel_coords": self.base_pixel_coords[None].expand(B, -1, -1, -1).contiguous()}
