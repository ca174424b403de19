"""Micro-benchmarks of the VAE decoder kernels (device-timed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tpxl_b200 import _lib
from gpu_util import st
lib = _lib.lib()

def timeit(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

P = 2048
for (S, C, Cout, resid) in [(4, 256, 256, False), (4, 256, 256, True), (8, 256, 32, False), (8, 32, 32, False), (8, 32, 32, True)]:
    x = torch.randn(P, S, S, S, C, device="cuda").half()
    w = (torch.randn(Cout, 27 * C, device="cuda") * (27 * C) ** -0.5).half()
    b = torch.randn(Cout, device="cuda").half()
    r = torch.randn(P, S, S, S, Cout, device="cuda").half() if resid else None
    o = torch.empty(P, S, S, S, Cout, device="cuda", dtype=torch.float16)
    ms = timeit(lambda: lib.tpx_conv3d_k3(x.data_ptr(), w.data_ptr(), b.data_ptr(), _lib.ptr(r), 0.7071 if resid else 1.0, o.data_ptr(), P, S, C, Cout, st()))
    fl = 2.0 * P * S ** 3 * Cout * 27 * C
    print(f"conv3 S={S} C={C}->{Cout} resid={resid}: {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TFLOP/s  A-bytes {P*S**3*C*2/1e6:.0f} MB", flush=True)
for (S3, C) in [(64, 256), (512, 256), (512, 32)]:
    x = torch.randn(P, S3, C, device="cuda").half(); g = torch.ones(C, device="cuda").half(); bb = torch.zeros(C, device="cuda").half(); o = torch.empty_like(x)
    ms = timeit(lambda: lib.tpx_groupnorm_silu(x.data_ptr(), g.data_ptr(), bb.data_ptr(), P, S3, C, 32, 1e-5, 1, o.data_ptr(), st()))
    print(f"groupnorm S3={S3} C={C}: {ms*1e3:8.1f} us  {x.numel()*2*2/ms/1e6:8.1f} GB/s (1R+1W)", flush=True)
