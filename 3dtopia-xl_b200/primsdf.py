"""Host-side mirror of the reference's PrimSDF query field, backed by libtpx_b200.

Drop-in for ``models.primsdf.PrimSDF`` (/root/reference/models/primsdf.py) on the inference path: same constructor
kwargs (configs/inference_dit.yml:22-30), the same two parameters ``srt_param [K,4]`` / ``feat_param [K, 6*S^3]`` (so
``load_state_dict`` of ``denoised.pt`` and the ``.data = filtered`` reassignment in inference.py:100-101 work), and
``forward(x[n,3]) -> {"sdf","tex","mat"}``.  The reference evaluates a dense [points x prims] weight matrix per 8192-point
chunk; the kernel visits every primitive per point in registers and samples only the covering ones.
Mesh-based initialisation (``_init_param``) is training-time and absent, as in the released reference (it is ``pass``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class PrimSDF(nn.Module):
    def __init__(self, mesh_obj=None, f_sdf=None, geo_fn=None, asset_list=None, num_prims=1024, dim_feat=6, prim_shape=8, init_scale=0.05,
                 sdf2alpha_var=0.005, auto_scale_init=True, init_sampling="uniform", **_unused):
        super().__init__()
        self.num_prims, self.dim_feat, self.prim_shape = num_prims, dim_feat, prim_shape
        self.sdf2alpha_var = sdf2alpha_var
        self.auto_scale_init, self.init_sampling = auto_scale_init, init_sampling
        self.mesh_obj, self.f_sdf = mesh_obj, f_sdf
        self.srt_param = nn.parameter.Parameter(torch.zeros(num_prims, 1 + 3))
        self.feat_param = nn.parameter.Parameter(torch.zeros(num_prims, dim_feat * prim_shape ** 3))
        s3 = prim_shape ** 3
        self.geo_start_index, self.geo_end_index = 0, s3
        self.tex_start_index, self.tex_end_index = s3, 4 * s3
        self.mat_start_index, self.mat_end_index = 4 * s3, 6 * s3
        self._grid = None          # (key, srt tensor the grid was built from, workspace): rebuilt whenever srt_param changes
        self.grid_entries = 8 << 20   # capacity of the per-cell lists (int32 entries); over-full grids fall back to the exhaustive loop

    def forward(self, x: torch.Tensor):
        if x.dim() != 2 or x.shape[1] != 3:
            raise ValueError(f"x must be [n,3], got {tuple(x.shape)}")
        if not x.is_cuda or not self.srt_param.is_cuda:
            raise _lib.TpxError("PrimSDF query runs on CUDA only (no CPU path): move the module and the points to the GPU")
        lib = _lib.lib()
        xx = x.detach().float().contiguous()
        srt = self.srt_param.detach().float().contiguous()
        feat = self.feat_param.detach().float().contiguous()
        n, K = xx.shape[0], srt.shape[0]
        out = torch.empty(n, self.dim_feat, dtype=torch.float32, device=xx.device)
        with torch.cuda.device(xx.device):
            if K < 1 or K > 4096:      # outside the grid builder's range: exhaustive kernel
                _lib.check(lib.tpx_primsdf_query(xx.data_ptr(), srt.data_ptr(), feat.data_ptr(), n, K, self.prim_shape, self.dim_feat,
                                                 0 if self.training else 1, out.data_ptr(), _lib.stream_ptr()), "tpx_primsdf_query")
            else:
                key = (self.srt_param.data_ptr(), self.srt_param._version, K, str(xx.device))
                if self._grid is None or self._grid[0] != key:
                    nbytes = int(lib.tpx_primsdf_grid_bytes(self.grid_entries))
                    ws = torch.empty(nbytes, dtype=torch.uint8, device=xx.device)
                    _lib.check(lib.tpx_primsdf_grid_build(srt.data_ptr(), K, ws.data_ptr(), nbytes, _lib.stream_ptr()), "tpx_primsdf_grid_build")
                    self._grid = (key, self.srt_param.data, ws)     # holding the tensor keeps its address from being recycled under the key
                ws = self._grid[2]
                _lib.check(lib.tpx_primsdf_query_grid(xx.data_ptr(), srt.data_ptr(), feat.data_ptr(), ws.data_ptr(), ws.numel(), n, K, self.prim_shape,
                                                      self.dim_feat, 0 if self.training else 1, out.data_ptr(), _lib.stream_ptr()), "tpx_primsdf_query_grid")
        return {"sdf": out[:, 0:1], "tex": out[:, 1:4], "mat": out[:, 4:6]}

    def sdf2alpha(self, sdf):
        return torch.exp(-(sdf / self.sdf2alpha_var) ** 2)


def _column_view(param: str, lo, hi):
    """Read-only views of the two parameters under the reference's attribute names (models/primsdf.py:111-137).  ``lo`` / ``hi``
    are column bounds, or names of the per-instance ``*_start_index`` / ``*_end_index`` attributes."""
    def get(self):
        a = getattr(self, lo) if isinstance(lo, str) else lo
        b = getattr(self, hi) if isinstance(hi, str) else hi
        return getattr(self, param)[:, a:b]
    return property(get)


for _name, _param, _lo, _hi in (("pos", "srt_param", 1, 4), ("scale", "srt_param", 0, 1), ("feat", "feat_param", 0, None),
                                ("feat_geo", "feat_param", "geo_start_index", "geo_end_index"),
                                ("feat_tex", "feat_param", "tex_start_index", "tex_end_index"),
                                ("feat_mat", "feat_param", "mat_start_index", "mat_end_index")):
    setattr(PrimSDF, _name, _column_view(_param, _lo, _hi))
