"""Time tpxl_b200.VAE.decode on 2048 primitives (fp16 in/out) and print the per-class device times of one profiled pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import tpxl_b200
from tpxl_b200 import _lib, synth

dev = "cuda:0"
sd = synth.synth_state_dict(synth.vae_decoder_shapes(**synth.FULL_VAE), 103)
vae = tpxl_b200.VAE(**synth.FULL_VAE); vae.load_state_dict(sd); vae = vae.to(dev)
z = torch.randn(2048, 1, 4, 4, 4, device=dev).half()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
with torch.no_grad():
    for _ in range(3): out = vae.decode(z)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): out = vae.decode(z)
    b.record(); torch.cuda.synchronize()
    print(f"decode 2048 primitives: {a.elapsed_time(b)/iters:.3f} ms", flush=True)
    lib = _lib.lib()
    lib.tpx_profile_begin()
    out = vae.decode(z)
    ms = (C.c_float * 8)(); n = (C.c_int64 * 8)()
    lib.tpx_profile_end(ms, n)
    names = ["gemm", "attention", "ln", "gemv", "elementwise", "conv_gemm", "groupnorm", "vae_misc"]
    print({names[i]: (round(ms[i], 3), n[i]) for i in range(8) if n[i]}, flush=True)
