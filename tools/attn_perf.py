"""Device-time the tcgen05 attention kernel on the two DiT shapes (self: 2x16x2048x2048, cross: 1x16x2048x1370, Dh 72).
Kernel switches are environment variables read once by the library (TPX_ATT_STALE_MAX, TPX_ATT_STAGGER), so run one
process per setting."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
from tpxl_b200 import _lib  # noqa: E402
from gpu_util import st  # noqa: E402

lib = _lib.lib()
res = []
for (B, H, Nq, Nk, Dh, DhP) in [(2, 16, 2048, 2048, 72, 80), (1, 16, 2048, 1370, 72, 80)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(B, H, Nq, DhP, device="cuda", generator=g).half()
    k = torch.randn(B, H, Nk, DhP, device="cuda", generator=g).half()
    NkPad = (Nk + 7) // 8 * 8
    vT = torch.zeros(B, H, DhP, NkPad, device="cuda", dtype=torch.float16)
    vT[..., :Nk] = torch.randn(B, H, DhP, Nk, device="cuda", generator=g).half()
    o = torch.empty(B, Nq, H * Dh, device="cuda", dtype=torch.float16)
    run = lambda: lib.tpx_attention_tc(q.data_ptr(), k.data_ptr(), vT.data_ptr(), o.data_ptr(), B, H, Nq, Nk, NkPad, Dh, Dh ** -0.5, st())  # noqa: E731
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        run()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / 50 * 1e3
    res.append(f"Nk={Nk}: {us:6.1f} us {4.0 * B * H * Nq * Nk * Dh / us / 1e6:6.1f} TF/s")
print(f"poly={os.environ.get('TPX_ATT_POLY', '0')} stale_max={os.environ.get('TPX_ATT_STALE_MAX', '1')} stagger={os.environ.get('TPX_ATT_STAGGER', '1600')}  " + "   ".join(res), flush=True)
