"""Micro-benchmarks of the hot kernels through the C ABI (device-timed, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpxl_b200 import _lib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import linear, st

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def main():
    lib = _lib.lib()
    for (M, N, K) in [(4096, 1152, 1152), (4096, 3456, 1152), (4096, 4608, 1152), (4096, 1152, 4608), (2048, 1152, 1152)]:
        A = torch.randn(M, K, device="cuda").half(); W = torch.randn(N, K, device="cuda").half() * K ** -0.5; b = torch.randn(N, device="cuda").half()
        for tile in (128, 192, 256, -128, -192, -256):
            for act in (0, 1):
                ms = timeit(lambda: linear(A, W, b, act=act, tile_n=tile))
                print(f"linear M={M} N={N} K={K} tile={tile} act={act}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)
        ms = timeit(lambda: A @ W.t())
        print(f"cublas M={M} N={N} K={K}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)
    # attention
    for (B, H, Nq, Nk, Dh, DhP) in [(2, 16, 2048, 2048, 72, 80), (1, 16, 2048, 1370, 72, 80)]:
        q = torch.randn(B, H, Nq, DhP, device="cuda").half(); k = torch.randn(B, H, Nk, DhP, device="cuda").half(); v = torch.randn(B, H, Nk, DhP, device="cuda").half()
        o = torch.empty(B, Nq, H * Dh, device="cuda", dtype=torch.float16)
        ms = timeit(lambda: lib.tpx_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, Nq, Nk, Dh, DhP, Dh ** -0.5, st()))
        fl = 4.0 * B * H * Nq * Nk * Dh
        print(f"attention B={B} Nq={Nq} Nk={Nk}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s (algorithmic, Dh=72)", flush=True)
        NkPad = (Nk + 7) // 8 * 8
        vT = torch.zeros(B, H, DhP, NkPad, device="cuda", dtype=torch.float16); vT[..., :Nk] = v.transpose(-1, -2)
        ms = timeit(lambda: lib.tpx_attention_tc(q.data_ptr(), k.data_ptr(), vT.data_ptr(), o.data_ptr(), B, H, Nq, Nk, NkPad, Dh, Dh ** -0.5, st()))
        print(f"attention_tc B={B} Nq={Nq} Nk={Nk}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s (algorithmic, Dh=72)", flush=True)
        qq, kk, vv = (t[..., :Dh].contiguous() for t in (q, k, v))
        ms = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv))
        print(f"torch sdpa same shape: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
    # LN
    x = torch.randn(4096, 1152, device="cuda"); mod = torch.randn(1, 3 * 1152, device="cuda").half(); out = torch.empty(4096, 1152, device="cuda", dtype=torch.float16)
    ms = timeit(lambda: lib.tpx_ln_modulate(x.data_ptr(), 4096, 1152, 1e-6, mod.data_ptr(), mod[:, 1152:].data_ptr(), 3456, 2048, 1, out.data_ptr(), None, None, 0, st()))
    print(f"ln_modulate 4096x1152: {ms*1e3:8.1f} us  {(4096*1152*6)/ms/1e6:8.1f} GB/s", flush=True)

main()
