"""Per-role timeline of the tcgen05 GEMM (tpx_debug_gemm_timeline): where do the cycles of one launch go?

For every DiT shape and epilogue the last of a few back-to-back launches is probed; per CTA the kernel reports clock64
stamps and barrier-wait sums of its TMA producer, MMA issuer and epilogue warps.  Printed: averages over CTAs (cycles).
  fill      first full barrier seen by the MMA warp, after griddepcontrol.wait
  mma_span  first MMA issue .. last commit;  ideal = tiles * num_kb * (BK/16) * (BN/2) cycles  (128 x BN x 16 per BN/2 cycles)
  w_full    MMA warp waiting for TMA data   (load-bound share)
  w_tempty  MMA warp waiting for the epilogue to drain an accumulator (epilogue-bound share)
  w_empty   producer waiting for a free stage (MMA-bound share)
  epi_proc  epilogue: tfull seen -> tempty arrive, per tile
  tail      epilogue end - last MMA commit
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tpxl_b200 import _lib
from gpu_util import st


def run(name, M, N, K, kind, tile, iters=6):
    lib = _lib.lib()
    dev = "cuda"
    A = torch.randn(M, K, device=dev).half()
    W = (torch.randn(N, K, device=dev) * K ** -0.5).half()
    b = torch.randn(N, device=dev).half()
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    if kind in ("store", "gelu"):
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        call = lambda: lib.tpx_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, 1 if kind == "gelu" else 0, 1.0, tile, st())
    elif kind == "gated":
        x = torch.randn(M, N, device=dev)
        gate = torch.randn(1, N, device=dev).half()
        call = lambda: lib.tpx_linear_gated(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), gate.data_ptr(), N, 1, 2048, x.data_ptr(), N, M, N, K, tile, st())
    else:  # heads: N = which * 1152
        nwhich = N // 1152
        S = M // 2048
        outs = [torch.empty(S, 16, 2048, 80, dtype=torch.float16, device=dev) for _ in range(3)]
        tw = 2 if nwhich == 3 else -1
        call = lambda: lib.tpx_linear_heads(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), M, N, K,
                                            1152, 16, 72, 80, 2048, 1.0, tile, tw, 2048, st())
    for _ in range(3):
        _lib.check(call(), name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    lib.tpx_debug_gemm_timeline(dbg.data_ptr())
    for _ in range(iters):
        call()
    lib.tpx_debug_gemm_timeline(None)
    torch.cuda.synchronize()
    d = dbg.view(148, 16).cpu().double()
    d = d[d[:, 10] > 0]
    tiles = d[:, 10]
    bn = abs(tile)
    ideal = tiles * (K // 64) * 4 * (bn / 2)
    f = lambda v: f"{v.mean():8.0f}/{v.max():8.0f}"
    wall = (d[:, 12].max() - d[:, 11].min())
    print(f"{name:28s} M={M} N={N} K={K} tile={tile} kind={kind}: {us:6.1f} us (probed wall {wall/1e3:6.1f} us) {2*M*N*K/us/1e6:7.1f} TF/s  CTAs {len(d)} tiles/CTA {tiles.mean():.2f}")
    print(f"   [avg/max cycles]  pdl_wait {f(d[:,0]-d[:,13])}  fill {f(d[:,1]-d[:,0])}  mma_span {f(d[:,2]-d[:,1])}  ideal {f(ideal)}  w_full {f(d[:,3])}  w_tempty {f(d[:,4])}")
    print(f"                     prod w_empty {f(d[:,5])}  epi w_tfull {f(d[:,7])}  epi_proc/tile {f(d[:,8]/tiles)}  tail {f(d[:,9]-d[:,2])}  total {f(d[:,14]-d[:,0])}  clk {((d[:,14]-d[:,0])/(d[:,12]-d[:,11])).mean():.3f} GHz")
    sys.stdout.flush()


def main():
    for (name, M, N, K, kind, tile) in [
        ("fc1 gelu", 4096, 4608, 1152, "gelu", 256),
        ("fc1 store", 4096, 4608, 1152, "store", 256),
        ("fc2 gated", 4096, 1152, 4608, "gated", 128),
        ("fc2 store", 4096, 1152, 4608, "store", 128),
        ("qkv heads", 4096, 3456, 1152, "heads", 192),
        ("qkv store", 4096, 3456, 1152, "store", 192),
        ("proj gated", 4096, 1152, 1152, "gated", 128),
        ("proj store", 4096, 1152, 1152, "store", 128),
        ("to_q heads (M=2048)", 2048, 1152, 1152, "heads", 144),
        ("cproj gated (M=2048)", 2048, 1152, 1152, "gated", 128),
        ("2cta fc1 gelu", 4096, 4608, 1152, "gelu", -256),
        ("2cta fc2 gated", 4096, 1152, 4608, "gated", -128),
        ("2cta qkv heads", 4096, 3456, 1152, "heads", -192),
        ("2cta proj gated", 4096, 1152, 1152, "gated", -128),
        ("2cta to_q heads (M=2048)", 2048, 1152, 1152, "heads", -144),
        ("2cta cproj gated (M=2048)", 2048, 1152, 1152, "gated", -128),
    ]:
        if len(sys.argv) > 1 and sys.argv[1] not in name:
            continue
        run(name, M, N, K, kind, tile)


main()
