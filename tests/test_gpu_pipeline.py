"""GPU: the sample -> decode glue and the progressive-decode pipeline (§8f-3, config #5) as product code.

a13 / a15 are compared with the reference's own statements (inference.py:328-332, 343-348) executed by eager torch on the same GPU:
the index layout must be bit-exact, and because the kernels repeat the eager ops one rounding at a time the values are too."""
import numpy as np
import pytest
import torch

import oracle
import tpxl_b200
from tpxl_b200 import synth
from gpu_util import rel_l2

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
P = 2048


# inference.py:328-348 restated once, in the oracle, where tests/test_oracle_golden.py pins it bit for bit to the reference's own statements
# (tests/golden/inference_glue.npz, produced by executing that block of /root/reference/inference.py)
_reference_glue = oracle.vae.inference_glue


@pytest.mark.parametrize("perchannel,nf", [(True, 1.0), (True, 0.7), (False, 1.0), (False, 1.3)])
def test_latent_split_and_pack_bit_exact(perchannel, nf):
    g = torch.Generator(device=DEV).manual_seed(5)
    sample = torch.randn(2, 96, 68, generator=g, device=DEV) * 1.7
    table = {}

    def fake_decode(z):        # a deterministic stand-in for vae.decode: records z, returns an index-revealing volume
        table.setdefault("z", []).append(z.clone())
        n = z.shape[0]
        base = torch.arange(n * 6 * 512, device=DEV, dtype=torch.float32).reshape(n, 6, 8, 8, 8)
        return base * 1e-3 - 3.0 + z.reshape(n, -1).sum(1).reshape(n, 1, 1, 1, 1)

    mean = torch.Tensor(synth.LATENT_MEAN)[None, None, :].to(DEV)
    std = torch.Tensor(synth.LATENT_STD)[None, None, :].to(DEV)
    want = _reference_glue(sample.clone(), fake_decode, mean, std, nf, perchannel)
    zs_ref = torch.cat(table["z"], 0)

    codec = tpxl_b200.LatentCodec(synth.LATENT_MEAN if perchannel else None, synth.LATENT_STD if perchannel else None, nf)
    srt, z = codec.split(sample)
    assert z.shape == (2 * 96, 1, 4, 4, 4) and srt.shape == (2, 96, 4)
    assert torch.equal(z, zs_ref)                                   # latents handed to the decoder: bit-exact (indexing AND values)
    table.clear()
    dec = torch.cat([fake_decode(z[:96]), fake_decode(z[96:])], 0)
    got = codec.pack(srt, dec)
    assert got.shape == want.shape == (2, 96, 4 + 6 * 512)
    assert torch.equal(got, want)
    # fp16 voxels (decode of fp16 latents): same layout, fp16 arithmetic of the in-place ops, promoted by the concat
    want16 = _reference_glue(sample.clone(), lambda zz: fake_decode(zz).half(), mean, std, nf, perchannel)
    got16 = codec.pack(srt, dec.half())
    assert got16.dtype == torch.float32 and torch.equal(got16, want16.float())


@pytest.fixture(scope="module")
def pipe():
    cfg = dict(seq_length=P, in_channels=68, condition_channels=768, hidden_size=384, depth=2, num_heads=16, attn_proj_bias=True, cond_drop_prob=0.1)
    m = tpxl_b200.DiT(**cfg)
    m.load_state_dict(synth.synth_state_dict(synth.dit_shapes(**cfg), 91))
    m = m.to(DEV).eval()
    vae = tpxl_b200.VAE(**synth.FULL_VAE)
    vae.load_state_dict(synth.synth_state_dict(synth.vae_decoder_shapes(**synth.FULL_VAE), 92))
    vae = vae.to(DEV)
    return tpxl_b200.PrimXPipeline(m, vae, latent_mean=synth.LATENT_MEAN, latent_std=synth.LATENT_STD, latent_nf=1.0, num_prims=P, cfg_scale=6.0, ddim_steps=25)


def test_progressive_previews_same_with_and_without_overlap(pipe):
    """§8f-3: decoding step k on the side stream while DiT step k+1 runs must not change what is yielded, nor its order."""
    x, y = synth.synth_inputs(1, P, 68, 64, 768, 93)
    x, y = x.to(DEV), y.to(DEV)
    serial = [(o["step"], o["final"], o["sample"].clone(), o["recon_param"].clone()) for o in pipe.sample_progressive(y, x, preview_every=10, overlap=False)]
    over = [(o["step"], o["final"], o["sample"].clone(), o["recon_param"].clone()) for o in pipe.sample_progressive(y, x, preview_every=10, overlap=True)]
    torch.cuda.synchronize()
    assert [s[0] for s in serial] == [s[0] for s in over] == [0, 10, 20, 24]          # inference.py:326: every 10th step and the last
    assert [s[1] for s in over] == [False, False, False, True]
    for a, b in zip(serial, over):
        assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    final = pipe(y, x)
    assert torch.equal(final["recon_param"], serial[-1][3]) and final["srt_param"].shape == (1, P, 4) and final["feat_param"].shape == (1, P, 3072)
    assert torch.isfinite(final["recon_param"]).all()


def test_pipeline_equals_manual_composition_of_the_reference_loop(pipe):
    """The pipeline's output == the reference loop's statements (inference.py:313-348) driven by hand around the same DiT / VAE."""
    x, y = synth.synth_inputs(2, P, 68, 64, 768, 94)
    x, y = x.to(DEV), y.to(DEV)
    d = tpxl_b200.create_diffusion("ddim25", noise_schedule="squaredcos_cap_v2", diffusion_steps=1000, parameterization="v")
    kw = dict(y=y, cfg_scale=6.0, precision_dtype=torch.float16, enable_amp=True)
    with torch.no_grad():
        for samples in d.ddim_sample_loop_progressive(pipe.model.forward_with_cfg, x.shape, x, clip_denoised=False, model_kwargs=kw, progress=False, device=DEV):
            pass
        mean = torch.Tensor(synth.LATENT_MEAN)[None, None, :].to(DEV)
        std = torch.Tensor(synth.LATENT_STD)[None, None, :].to(DEV)
        want = _reference_glue(samples["sample"], pipe.vae.decode, mean, std, 1.0, True)
        got = pipe(y, x)
    assert torch.equal(got["sample"], samples["sample"])
    assert torch.equal(got["recon_param"], want)
