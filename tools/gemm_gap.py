"""Kernel-to-kernel gap of back-to-back GEMM launches, per SM (clock64 is continuous per SM across launches):
gap = start of launch i+1 (after griddepcontrol.wait) - end of launch i on the same SM.  Also the host cost per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tpxl_b200 import _lib
from gpu_util import st

def main():
    lib = _lib.lib()
    dev = "cuda"
    for (name, M, N, K, act, tile) in [("fc1 gelu", 4096, 4608, 1152, 1, 256), ("qkv store", 4096, 3456, 1152, 0, 192), ("proj store", 4096, 1152, 1152, 0, 128), ("M2048 store", 2048, 1152, 1152, 0, 128)]:
        A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) * K ** -0.5).half(); b = torch.randn(N, device=dev).half()
        out = torch.empty(M, N, dtype=torch.float16, device=dev)
        call = lambda: lib.tpx_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, act, 1.0, tile, st())
        for _ in range(5): call()
        torch.cuda.synchronize()
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(n): call()
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{name}: device {e0.elapsed_time(e1)/n*1e3:.1f} us/launch, host issue {(t1-t0)/n*1e6:.1f} us/call", flush=True)
        L = 6
        bufs = [torch.zeros(148 * 16, dtype=torch.int64, device=dev) for _ in range(L)]
        for _ in range(3): call()
        for i in range(L):
            lib.tpx_debug_gemm_timeline(bufs[i].data_ptr()); call()
        lib.tpx_debug_gemm_timeline(None)
        torch.cuda.synchronize()
        recs = []
        for bf in bufs:
            d = bf.view(148, 16).cpu()
            d = d[d[:, 10] > 0]
            recs.append({int(r[15]): r for r in d})
        for i in range(1, L):
            gaps, spans, entry = [], [], []
            for sm, r in recs[i].items():
                if sm in recs[i - 1]:
                    p = recs[i - 1][sm]
                    gaps.append(int(r[0] - p[14])); spans.append(int(r[14] - r[0])); entry.append(int(r[13] - p[14]))
            g = torch.tensor(gaps, dtype=torch.float64); s = torch.tensor(spans, dtype=torch.float64); en = torch.tensor(entry, dtype=torch.float64)
            print(f"   launch {i}: same-SM pairs {len(gaps)}  end->next entry (CTA launch+prologue) avg {en.mean():.0f} min {en.min():.0f} max {en.max():.0f} | end->next start avg {g.mean():.0f} min {g.min():.0f} max {g.max():.0f} | span avg {s.mean():.0f} max {s.max():.0f} | period (span+gap) avg {(s+g).mean():.0f}", flush=True)
main()
