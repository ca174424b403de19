"""CPU: the host side of the DINOv2 encoder mirror (3dtopia-xl_b200/dinov2.py) — state-dict contract, and the
algebra of its launch plan (Normalize folded into the patch embedding, (i, j, c) patch layout, LayerNorm affine expressed as
modulation, LayerScale expressed as the residual gate, token assembly) checked by emulating each kernel's documented contract in
torch and comparing the whole pipeline with the oracle."""
import importlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
import tpxl_b200  # noqa: F401

d = importlib.import_module("tpxl_b200.dinov2")


def _h(t):
    return t.half().float()


def _linear(A, W, b):                       # tpx_linear: fp16 operands, fp32 accumulate, one rounding of acc + bias
    return _h(A.float() @ W.float().t() + b.float())


def _gated(A, W, b, gate, xres):            # tpx_linear_gated: xres(fp32) += h(gate * h(A W^T + b))
    xres += _h(gate.float() * _linear(A, W, b))


def _lnmod(x, shift, scale):                # tpx_ln_modulate: h(LN(x) * h(1 + scale) + shift)
    return _h(F.layer_norm(x, (x.shape[-1],), eps=1e-6) * _h(1 + scale.float()) + shift.float())


def test_state_dict_contract_and_errors():
    m = d.Dinov2Wrapper("dinov2_vitb14_reg", freeze=True)
    assert list(m.state_dict().keys()) == list(oracle.dinov2.shapes().keys())
    sd = oracle.dinov2.synth_weights(3)
    sd["model.mask_token"] = torch.zeros(1, 768)                 # present in hub checkpoints, dropped by the reference loader
    m.load_state_dict(sd)
    bare = {k[len("model."):]: v for k, v in sd.items()}           # dinov2_vitb14_reg4_pretrain.pth has no wrapper prefix
    m.load_state_dict(bare)
    assert torch.equal(m.state_dict()["model.norm.weight"], sd["model.norm.weight"])
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd.items() if "ls1" not in k})
    with pytest.raises(NotImplementedError):
        d.Dinov2Wrapper("dinov2_vitl14_reg")
    with pytest.raises(tpxl_b200._lib.TpxError):
        m(torch.zeros(1, 518, 518, 3))


def test_launch_plan_algebra_matches_oracle():
    torch.set_grad_enabled(False)
    try:
        sd = oracle.dinov2.synth_weights(105)
        m = d.Dinov2Wrapper()
        m.load_state_dict(sd)
        w = m._operands(torch.device("cpu"))
        rs = np.random.RandomState(7)
        img = torch.from_numpy(rs.uniform(0, 255, size=(1, 518, 518, 3)).astype(np.float32))
        n, g, np_, nt = 1, 37, 1369, 1374
        patches = torch.zeros(n * np_, d.KPAD, dtype=torch.float16)
        patches[:, :d.KPATCH] = (img / 255.0).reshape(n, g, 14, g, 14, 3).permute(0, 1, 3, 2, 4, 5).reshape(n * np_, d.KPATCH).half()
        tok = torch.empty(n, nt, 768)
        tok[:, 0], tok[:, 1:5], tok[:, 5:] = w["row0"], w["regs"], w["pos_patches"]
        _gated(patches, w["patch_w"], w["patch_b"], w["ones"], tok[0, 5:])
        xr = tok.view(-1, 768)
        for i in range(d.DEPTH):
            ln = _lnmod(xr, w[f"{i}.ln1.shift"], w[f"{i}.ln1.scale"])
            qkv = _linear(ln, w[f"{i}.attn.qkv.weight"], w[f"{i}.attn.qkv.bias"]).reshape(n, nt, 3, 12, 64).permute(2, 0, 3, 1, 4)
            a = _h(F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])).transpose(1, 2).reshape(n * nt, 768)
            _gated(a, w[f"{i}.attn.proj.weight"], w[f"{i}.attn.proj.bias"], w[f"{i}.ls1"], xr)
            ln = _lnmod(xr, w[f"{i}.ln2.shift"], w[f"{i}.ln2.scale"])
            hid = _h(F.gelu(_linear(ln, w[f"{i}.mlp.fc1.weight"], w[f"{i}.mlp.fc1.bias"])))
            _gated(hid, w[f"{i}.mlp.fc2.weight"], w[f"{i}.mlp.fc2.bias"], w[f"{i}.ls2"], xr)
        o = _lnmod(xr, w["norm.shift"], w["norm.scale"]).view(n, nt, 768)
        out = torch.cat([o[:, :1], o[:, 5:]], 1)
        ref = oracle.dinov2.forward(sd, img)
        rel = float((out - ref).norm() / ref.norm())
        assert rel < 3e-3, rel                                   # fp16-input contract against the fp32 oracle: ~1e-3
    finally:
        torch.set_grad_enabled(True)
